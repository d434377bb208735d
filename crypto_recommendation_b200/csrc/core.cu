// core.cu -- context, points upload, per-row norms, pair ops, counters, profiling.
#include <cstdarg>

#include "common.cuh"
#include "tc_scan.cuh"

static thread_local char g_err[1024] = "";

void crx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// Pack caller rows [n][d] into the padded fp32 (and fp64) matrices.  One thread per element of the
// padded row; consecutive threads touch consecutive addresses on both sides.
template <typename T>
__global__ void pack_rows_kernel(const T* __restrict__ src, int64_t n, int d, int ld, float* __restrict__ x32,
                                 double* __restrict__ x64) {
    int64_t total = n * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / ld;
        int c = (int)(i - r * ld);
        T v = c < d ? src[r * d + c] : (T)0;
        x32[i] = (float)v;
        if (x64) x64[i] = (double)v;
    }
}

// Sum of squares of every row, double, index order.  A warp stages 32 rows x 32 columns through
// shared memory so that global reads are coalesced while each lane still walks ITS row in order.
template <typename T>
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const T* __restrict__ x, int64_t n, int d, int ld,
                                                         double* __restrict__ sqn) {
    __shared__ double tile[8][32][17];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * 32;
    if (row0 >= n) return;
    double acc = 0.0;
    for (int c0 = 0; c0 < d; c0 += 16) {
        for (int it = 0; it < 16; it++) {
            int r = it * 2 + (lane >> 4);
            int64_t row = row0 + r;
            int c = c0 + (lane & 15);
            tile[warp][r][lane & 15] = (row < n && c < d) ? (double)x[row * ld + c] : 0.0;
        }
        __syncwarp();
        int lim = min(16, d - c0);
        for (int c = 0; c < lim; c++) {
            double v = tile[warp][lane][c];
            acc = __dadd_rn(acc, __dmul_rn(v, v));
        }
        __syncwarp();
    }
    if (row0 + lane < n) sqn[row0 + lane] = acc;
}

template <typename TA, typename TB>
__global__ void pair_op_kernel(const TA* __restrict__ xa, int lda, const double* __restrict__ na,
                               const int32_t* __restrict__ ia, const TB* __restrict__ xb, int ldb,
                               const double* __restrict__ nb, const int32_t* __restrict__ ib, int64_t m, int d, int op,
                               double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const TA* a = xa + (int64_t)ia[i] * lda;
    const TB* b = xb + (int64_t)ib[i] * ldb;
    double r;
    if (op == 0) r = x87_to_double(dot_x87(a, b, d));   // the long double of cust_vector.hpp:107-121, rounded to double
    else if (op == 1) r = euclid_exact(a, b, d);
    else if (op == 2) r = __dsub_rn(1.0, cos_sim_exact(a, b, d, na[ia[i]], nb[ib[i]]));
    else r = cos_sim_exact(a, b, d, na[ia[i]], nb[ib[i]]);
    out[i] = r;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int crx_version(void) { return 1; }
const char* crx_last_error(void) { return g_err; }

int crx_ctx_create(int device, void* cuda_stream, crx_ctx** out) {
    CRX_REQUIRE(out, "out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        crx_set_error("no CUDA device available (%s): this engine has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return CRX_ERR_CUDA;
    }
    CRX_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    CRX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CRX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        crx_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return CRX_ERR_UNSUPPORTED;
    }
    crx_ctx* c = new crx_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
    else {
        CRX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    CRX_CUDA(cudaMalloc((void**)&c->counters, 8 * sizeof(unsigned long long)));
    CRX_CUDA(cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned long long), c->stream));
    // keep freed stream-ordered allocations in the pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = c;
    return CRX_OK;
}

int crx_ctx_destroy(crx_ctx* c) {
    if (!c) return CRX_OK;
    cudaSetDevice(c->device);
    for (auto& b : c->big_free) cudaFreeAsync(b.first, c->stream);
    c->big_free.clear();
    cudaStreamSynchronize(c->stream);
    for (auto& r : c->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : c->free_events) cudaEventDestroy(e);
    cudaFree(c->counters);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return CRX_OK;
}

int crx_ctx_synchronize(crx_ctx* c) {
    CRX_REQUIRE(c, "ctx is NULL");
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}

int64_t crx_ctx_launch_count(const crx_ctx* c) { return c ? c->launches : 0; }

int crx_ctx_trim(crx_ctx* c) {
    CRX_REQUIRE(c, "ctx is NULL");
    CRX_CUDA(cudaSetDevice(c->device));
    for (auto& b : c->big_free) cudaFreeAsync(b.first, c->stream);   // the context's cache of large temporaries
    c->big_free.clear();
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    cudaMemPool_t pool;
    CRX_CUDA(cudaDeviceGetDefaultMemPool(&pool, c->device));
    CRX_CUDA(cudaMemPoolTrimTo(pool, 0));
    return CRX_OK;
}

int crx_ctx_profile(crx_ctx* c, int enable) {
    CRX_REQUIRE(c, "ctx is NULL");
    c->profiling = enable != 0;
    return CRX_OK;
}

int crx_ctx_profile_reset(crx_ctx* c) {
    CRX_REQUIRE(c, "ctx is NULL");
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    for (auto& r : c->prof) { c->free_events.push_back(r.e0); c->free_events.push_back(r.e1); }
    c->prof.clear();
    return CRX_OK;
}

int crx_ctx_kernel_time(crx_ctx* c, const char* prefix, double* total_ms, int64_t* launches) {
    CRX_REQUIRE(c && prefix, "NULL argument");
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    double t = 0;
    int64_t n = 0;
    size_t pl = strlen(prefix);
    for (auto& r : c->prof) {
        if (strncmp(r.name, prefix, pl) == 0) {
            float ms = 0;
            CRX_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
            t += ms;
            n++;
        }
    }
    if (total_ms) *total_ms = t;
    if (launches) *launches = n;
    return CRX_OK;
}

int crx_ctx_counters(crx_ctx* c, int64_t out[8], int reset) {
    CRX_REQUIRE(c, "ctx is NULL");
    unsigned long long h[8];
    CRX_CUDA(cudaMemcpyAsync(h, c->counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    if (out) for (int i = 0; i < 8; i++) out[i] = (int64_t)h[i];
    if (reset) CRX_CUDA(cudaMemsetAsync(c->counters, 0, sizeof(h), c->stream));
    return CRX_OK;
}

int crx_points_create(crx_ctx* c, const void* data, int dtype, int64_t n, int32_t d, int mem, crx_points** out) {
    CRX_REQUIRE(c && data && out, "NULL argument");
    CRX_REQUIRE(n > 0 && n < (1ll << 31), "n must be in [1, 2^31)");
    CRX_REQUIRE(d > 0 && d <= 512, "d must be in [1, 512]");
    CRX_REQUIRE(dtype == CRX_F32 || dtype == CRX_F64, "dtype");
    CRX_CUDA(cudaSetDevice(c->device));
    crx_points* p = new crx_points();
    p->ctx = c; p->n = n; p->d = d; p->ld = (d + 3) & ~3;
    size_t elems = (size_t)n * p->ld;
    int st = crx_alloc(c, &p->x32, elems);
    if (st == CRX_OK && dtype == CRX_F64) st = crx_alloc(c, &p->x64, elems);
    if (st == CRX_OK) st = crx_alloc(c, &p->sqn, (size_t)n);
    if (st != CRX_OK) {
        crx_points_destroy(p);
        return st;
    }
    size_t esz = dtype == CRX_F32 ? 4 : 8;
    const void* src = data;
    // released on every early return below (CRX_TRY / CRX_CUDA): the half-built object and the staged copy of the caller's rows
    struct Guard {
        crx_ctx* c; crx_points* p; void* staged;
        ~Guard() { if (staged) crx_free(c, staged); if (p) crx_points_destroy(p); }
    } guard{c, p, nullptr};
    if (mem == CRX_HOST) {
        char* st8 = nullptr;
        CRX_TRY(crx_alloc(c, &st8, (size_t)n * d * esz));   // through the context's cache of large blocks: an upload per step reuses it
        guard.staged = st8;
        CRX_CUDA(cudaMemcpyAsync(st8, data, (size_t)n * d * esz, cudaMemcpyHostToDevice, c->stream));
        src = st8;
    }
    int grid = (int)std::min<int64_t>((int64_t)c->sm_count * 16, (int64_t)((elems + 255) / 256));
    {
        CRX_KERNEL(c, "pack_rows");
        if (dtype == CRX_F32) pack_rows_kernel<float><<<grid, 256, 0, c->stream>>>((const float*)src, n, d, p->ld, p->x32, nullptr);
        else pack_rows_kernel<double><<<grid, 256, 0, c->stream>>>((const double*)src, n, d, p->ld, p->x32, p->x64);
    }
    {
        CRX_KERNEL(c, "row_sqnorm");
        int g = (int)((n + 255) / 256);
        if (p->x64) row_sqnorm_kernel<double><<<g, 256, 0, c->stream>>>(p->x64, n, d, p->ld, p->sqn);
        else row_sqnorm_kernel<float><<<g, 256, 0, c->stream>>>(p->x32, n, d, p->ld, p->sqn);
    }
    CRX_CUDA(cudaGetLastError());
    if (mem == CRX_HOST) CRX_CUDA(cudaStreamSynchronize(c->stream));  // caller may reuse its buffer
    guard.p = nullptr;   // built: the caller owns it (the staged copy is still released by the guard, stream-ordered)
    *out = p;
    return CRX_OK;
}

int crx_points_set_ratings(crx_points* p, const uint8_t* unknown, const double* known_mean, int mem) {
    CRX_REQUIRE(p && unknown && known_mean, "NULL argument");
    crx_ctx* c = p->ctx;
    CRX_CUDA(cudaSetDevice(c->device));
    size_t nu = (size_t)p->n * p->d;
    if (!p->unknown) CRX_TRY(crx_alloc(c, &p->unknown, nu));
    if (!p->mean) CRX_TRY(crx_alloc(c, &p->mean, (size_t)p->n));
    cudaMemcpyKind kind = mem == CRX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CRX_CUDA(cudaMemcpyAsync(p->unknown, unknown, nu, kind, c->stream));
    CRX_CUDA(cudaMemcpyAsync(p->mean, known_mean, (size_t)p->n * sizeof(double), kind, c->stream));
    if (mem == CRX_HOST) CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}

int crx_points_destroy(crx_points* p) {
    if (!p) return CRX_OK;
    if (p->tc_l2) { p->tc_l2->free_all(); delete p->tc_l2; p->tc_l2 = nullptr; }
    if (p->ctx) {
        cudaSetDevice(p->ctx->device);
        crx_free(p->ctx, p->x32); crx_free(p->ctx, p->x64); crx_free(p->ctx, p->sqn); crx_free(p->ctx, p->unknown); crx_free(p->ctx, p->mean);
    }
    delete p;
    return CRX_OK;
}

int64_t crx_points_n(const crx_points* p) { return p ? p->n : 0; }
int32_t crx_points_d(const crx_points* p) { return p ? p->d : 0; }

int crx_pair_op(crx_ctx* c, const crx_points* pa, const int32_t* a, const crx_points* pb, const int32_t* b, int64_t m,
                int op, double* out) {
    CRX_REQUIRE(c && pa && pb && a && b && out, "NULL argument");
    CRX_REQUIRE(pa->d == pb->d, "dimension mismatch");
    CRX_REQUIRE(op >= 0 && op <= 3, "op");
    if (m == 0) return CRX_OK;
    CRX_CUDA(cudaSetDevice(c->device));
    IoBuf<int32_t> ia, ib;
    IoBuf<double> o;
    CRX_TRY(ia.bind(c, a, m, CRX_HOST, true));
    CRX_TRY(ib.bind(c, b, m, CRX_HOST, true));
    CRX_TRY(o.bind(c, out, m, CRX_HOST, false));
    int grid = crx_grid(m, 128);
    {
        CRX_KERNEL(c, "pair_op");
        if (pa->x64 && pb->x64)
            pair_op_kernel<double, double><<<grid, 128, 0, c->stream>>>(pa->x64, pa->ld, pa->sqn, ia.dev, pb->x64, pb->ld, pb->sqn, ib.dev, m, pa->d, op, o.dev);
        else if (pa->x64)
            pair_op_kernel<double, float><<<grid, 128, 0, c->stream>>>(pa->x64, pa->ld, pa->sqn, ia.dev, pb->x32, pb->ld, pb->sqn, ib.dev, m, pa->d, op, o.dev);
        else if (pb->x64)
            pair_op_kernel<float, double><<<grid, 128, 0, c->stream>>>(pa->x32, pa->ld, pa->sqn, ia.dev, pb->x64, pb->ld, pb->sqn, ib.dev, m, pa->d, op, o.dev);
        else
            pair_op_kernel<float, float><<<grid, 128, 0, c->stream>>>(pa->x32, pa->ld, pa->sqn, ia.dev, pb->x32, pb->ld, pb->sqn, ib.dev, m, pa->d, op, o.dev);
    }
    CRX_CUDA(cudaGetLastError());
    return o.flush();
}

} // extern "C"
