// comm_nccl.cu -- the crx_comm of include/crx.h over NCCL, inside libcrx.so (one process per GPU).
//
// The sharded entry points (crx_k_means_sharded, crx_k_means_pp_sharded, crx_pam_lloyds_sharded, ...) call the three
// collectives of a crx_comm.  This file provides them natively: ncclAllReduce / ncclAllGather / ncclBroadcast enqueued on
// the CONTEXT'S stream, so a collective on a device buffer is ordered with the kernels around it and needs no host
// synchronisation (crx_comm.stream_ordered = 1; the engine then skips the synchronisation it makes for callbacks that run
// on another stream).  Host buffers (the few scalars of k-means++) are staged through a small device scratch.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring the copy the process already holds -- the one PyTorch
// loaded): libcrx.so keeps loading on a box without NCCL, and `crx_comm_nccl_create` then fails with an explicit error.
// The 128-byte unique id is made by rank 0 (crx_comm_nccl_unique_id) and handed to the other ranks by whatever bootstrap
// the caller has (torch.distributed broadcast in dist.py, a file, MPI).
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy already in the process (PyTorch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { crx_set_error("libnccl.so.2 cannot be loaded: %s", dlerror()); return nullptr; }
#define BIND(field, sym)                                                                  \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, sym));                     \
    if (!api.field) { crx_set_error("libnccl.so.2 has no symbol %s", sym); return nullptr; }
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(GetErrorString, "ncclGetErrorString");
    BIND(AllReduce, "ncclAllReduce");
    BIND(AllGather, "ncclAllGather");
    BIND(Broadcast, "ncclBroadcast");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetVersion, "ncclGetVersion");
#undef BIND
    api.lib = h;
    return &api;
}

struct NcclComm {
    crx_comm pub;          // first member: a crx_comm* handed out by crx_comm_nccl_create points here
    crx_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    char* scratch = nullptr;      // device staging for host buffers
    size_t scratch_bytes = 0;
    int64_t calls[3] = {0, 0, 0};
};

int nccl_fail(ncclResult_t r, const char* what) {
    NcclApi* a = nccl_api();
    crx_set_error("%s -> %s", what, a ? a->GetErrorString(r) : "NCCL unavailable");
    return 1;
}
#define NCCL_TRY(call)                                            \
    do {                                                          \
        ncclResult_t r__ = (call);                                \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);     \
    } while (0)
#define CUDA_TRY1(call)                                                                                      \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) { crx_set_error("%s -> %s", #call, cudaGetErrorString(e__)); return 1; }     \
    } while (0)

bool nccl_type(int dtype, ncclDataType_t* t, size_t* sz) {
    switch (dtype) {
        case CRX_F32: *t = ncclFloat32; *sz = 4; return true;
        case CRX_F64: *t = ncclFloat64; *sz = 8; return true;
        case CRX_I32: *t = ncclInt32; *sz = 4; return true;
        case CRX_I64: *t = ncclInt64; *sz = 8; return true;
    }
    return false;
}

int ensure_scratch(NcclComm* n, size_t bytes) {
    if (bytes <= n->scratch_bytes) return 0;
    if (n->scratch) CUDA_TRY1(cudaFree(n->scratch));
    n->scratch = nullptr;
    n->scratch_bytes = 0;
    size_t want = std::max<size_t>(bytes, (size_t)1 << 16);
    CUDA_TRY1(cudaMalloc((void**)&n->scratch, want));
    n->scratch_bytes = want;
    return 0;
}

int cb_allreduce(void* user, void* buf, int64_t count, int dtype, int op, int mem) {
    NcclComm* n = static_cast<NcclComm*>(user);
    NcclApi* a = nccl_api();
    ncclDataType_t t;
    size_t sz;
    if (!a || !nccl_type(dtype, &t, &sz)) return 1;
    const ncclRedOp_t o = op == CRX_SUM ? ncclSum : (op == CRX_MAX ? ncclMax : ncclMin);
    cudaStream_t s = n->ctx->stream;
    n->calls[0]++;
    if (mem == CRX_DEVICE) {
        NCCL_TRY(a->AllReduce(buf, buf, (size_t)count, t, o, n->comm, s));
        return 0;
    }
    if (ensure_scratch(n, (size_t)count * sz)) return 1;
    CUDA_TRY1(cudaMemcpyAsync(n->scratch, buf, (size_t)count * sz, cudaMemcpyHostToDevice, s));
    NCCL_TRY(a->AllReduce(n->scratch, n->scratch, (size_t)count, t, o, n->comm, s));
    CUDA_TRY1(cudaMemcpyAsync(buf, n->scratch, (size_t)count * sz, cudaMemcpyDeviceToHost, s));
    CUDA_TRY1(cudaStreamSynchronize(s));
    return 0;
}

int cb_allgather(void* user, const void* send, void* recv, int64_t count, int dtype, int mem) {
    NcclComm* n = static_cast<NcclComm*>(user);
    NcclApi* a = nccl_api();
    ncclDataType_t t;
    size_t sz;
    if (!a || !nccl_type(dtype, &t, &sz)) return 1;
    cudaStream_t s = n->ctx->stream;
    n->calls[1]++;
    if (mem == CRX_DEVICE) {
        NCCL_TRY(a->AllGather(send, recv, (size_t)count, t, n->comm, s));
        return 0;
    }
    const size_t one = (size_t)count * sz, all = one * (size_t)n->pub.world;
    if (ensure_scratch(n, one + all)) return 1;
    CUDA_TRY1(cudaMemcpyAsync(n->scratch, send, one, cudaMemcpyHostToDevice, s));
    NCCL_TRY(a->AllGather(n->scratch, n->scratch + one, (size_t)count, t, n->comm, s));
    CUDA_TRY1(cudaMemcpyAsync(recv, n->scratch + one, all, cudaMemcpyDeviceToHost, s));
    CUDA_TRY1(cudaStreamSynchronize(s));
    return 0;
}

int cb_broadcast(void* user, void* buf, int64_t count, int dtype, int root, int mem) {
    NcclComm* n = static_cast<NcclComm*>(user);
    NcclApi* a = nccl_api();
    ncclDataType_t t;
    size_t sz;
    if (!a || !nccl_type(dtype, &t, &sz)) return 1;
    cudaStream_t s = n->ctx->stream;
    n->calls[2]++;
    if (mem == CRX_DEVICE) {
        NCCL_TRY(a->Broadcast(buf, buf, (size_t)count, t, root, n->comm, s));
        return 0;
    }
    if (ensure_scratch(n, (size_t)count * sz)) return 1;
    if (n->pub.rank == root) CUDA_TRY1(cudaMemcpyAsync(n->scratch, buf, (size_t)count * sz, cudaMemcpyHostToDevice, s));
    NCCL_TRY(a->Broadcast(n->scratch, n->scratch, (size_t)count, t, root, n->comm, s));
    if (n->pub.rank != root) CUDA_TRY1(cudaMemcpyAsync(buf, n->scratch, (size_t)count * sz, cudaMemcpyDeviceToHost, s));
    CUDA_TRY1(cudaStreamSynchronize(s));
    return 0;
}

}  // namespace

extern "C" {

int crx_comm_nccl_version(void) {
    NcclApi* a = nccl_api();
    int v = 0;
    if (!a || a->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int crx_comm_nccl_unique_id(uint8_t id[CRX_NCCL_ID_BYTES]) {
    CRX_REQUIRE(id, "NULL argument");
    static_assert(CRX_NCCL_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");
    NcclApi* a = nccl_api();
    if (!a) return CRX_ERR_COMM;
    ncclUniqueId u;
    ncclResult_t r = a->GetUniqueId(&u);
    if (r != ncclSuccess) { nccl_fail(r, "ncclGetUniqueId"); return CRX_ERR_COMM; }
    memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return CRX_OK;
}

int crx_comm_nccl_create(crx_ctx* c, const uint8_t id[CRX_NCCL_ID_BYTES], int rank, int world, crx_comm** out) {
    CRX_REQUIRE(c && id && out, "NULL argument");
    CRX_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank / world");
    NcclApi* a = nccl_api();
    if (!a) return CRX_ERR_COMM;
    CRX_CUDA(cudaSetDevice(c->device));
    NcclComm* n = new NcclComm();
    n->ctx = c;
    ncclUniqueId u;
    memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
    ncclResult_t r = a->CommInitRank(&n->comm, world, u, rank);
    if (r != ncclSuccess) { nccl_fail(r, "ncclCommInitRank"); delete n; return CRX_ERR_COMM; }
    n->pub.rank = rank;
    n->pub.world = world;
    n->pub.user = n;
    n->pub.allreduce = cb_allreduce;
    n->pub.allgather = cb_allgather;
    n->pub.broadcast = cb_broadcast;
    n->pub.stream_ordered = 1;
    *out = &n->pub;
    return CRX_OK;
}

int crx_comm_nccl_calls(const crx_comm* comm, int64_t out[3]) {
    CRX_REQUIRE(comm && out && comm->allreduce == cb_allreduce, "not a crx_comm_nccl_create communicator");
    const NcclComm* n = static_cast<const NcclComm*>(comm->user);
    for (int i = 0; i < 3; i++) out[i] = n->calls[i];
    return CRX_OK;
}

int crx_comm_nccl_destroy(crx_comm* comm) {
    if (!comm) return CRX_OK;
    CRX_REQUIRE(comm->allreduce == cb_allreduce, "not a crx_comm_nccl_create communicator");
    NcclComm* n = static_cast<NcclComm*>(comm->user);
    NcclApi* a = nccl_api();
    cudaSetDevice(n->ctx->device);
    cudaStreamSynchronize(n->ctx->stream);
    if (a && n->comm) a->CommDestroy(n->comm);
    if (n->scratch) cudaFree(n->scratch);
    delete n;
    return CRX_OK;
}

}  // extern "C"
