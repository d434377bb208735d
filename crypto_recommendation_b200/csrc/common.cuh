// common.cuh -- context, handles, launch bookkeeping and the exact FP64 device arithmetic shared by
// every kernel of the engine.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/crx.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
void crx_set_error(const char* fmt, ...);

#define CRX_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            crx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));  \
            return CRX_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define CRX_TRY(call)                  \
    do {                               \
        int s__ = (call);              \
        if (s__ != CRX_OK) return s__; \
    } while (0)

#define CRX_REQUIRE(cond, msg)                                                 \
    do {                                                                       \
        if (!(cond)) {                                                         \
            crx_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
            return CRX_ERR_INVALID;                                            \
        }                                                                      \
    } while (0)
// rows wider than 128 coordinates are accepted by the clustering core only (k_means_pp, lloyds_assignment,
// lloyds_for_remaining, k_means, pair_op); every other entry point says so instead of misbehaving
#define CRX_NARROW(p) CRX_REQUIRE((p)->d <= 128, "this entry point supports d <= 128 (wider rows: k_means_pp, lloyds_assignment, k_means, pair_op)")


// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct crx_prof_rec {
    const char* name;
    cudaEvent_t e0, e1;
};

enum { CRX_CNT_HASH_DD = 0, CRX_CNT_TOPP_RESCAN = 1, CRX_CNT_KPP_NEAR = 2, CRX_CNT_PAM_EXACT = 3, CRX_CNT_LLOYD_EXACT = 4, CRX_CNT_TOPP_TIES = 5, CRX_CNT_TOPP_TIED = 6, CRX_CNT_TOPP_PASS2 = 7 };

struct crx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int64_t launches = 0;
    bool profiling = false;
    std::vector<crx_prof_rec> prof;
    std::vector<cudaEvent_t> free_events;
    unsigned long long* counters = nullptr;  // device, 8 slots
    // Large temporaries (>= CRX_BIG_BYTES: pass masks, collected lists, operands) are kept by the context between calls: the
    // stream-ordered pool splits a freed 20 GB block for the next call's smaller buffers and then has to map fresh memory for
    // the next 20 GB request -- hundreds of milliseconds, at random.  All work of a context is ordered on its one stream, so a
    // block handed back here can be handed out again at once.  crx_ctx_trim / crx_ctx_destroy return them to the driver.
    std::vector<std::pair<void*, size_t>> big_free;       // (block, capacity)
    std::unordered_map<void*, size_t> big_live;           // blocks handed out -> capacity
};
constexpr size_t CRX_BIG_BYTES = (size_t)64 << 20;
// nullptr when the request is small (the caller uses the pool) or no memory is left
static inline void* crx_big_take(crx_ctx* c, size_t bytes, cudaError_t* err) {
    *err = cudaSuccess;
    if (bytes < CRX_BIG_BYTES) return nullptr;
    size_t best = (size_t)-1, cap = 0;
    for (size_t i = 0; i < c->big_free.size(); i++) {
        const size_t b = c->big_free[i].second;
        if (b >= bytes && b <= bytes + bytes / 2 + CRX_BIG_BYTES && (best == (size_t)-1 || b < cap)) { best = i; cap = b; }
    }
    void* p = nullptr;
    if (best != (size_t)-1) {
        p = c->big_free[best].first;
        c->big_free.erase(c->big_free.begin() + best);
    } else {
        *err = cudaMallocAsync(&p, bytes, c->stream);
        if (*err != cudaSuccess) {   // give the cached blocks back and try once more
            for (auto& b : c->big_free) cudaFreeAsync(b.first, c->stream);
            c->big_free.clear();
            cudaGetLastError();
            *err = cudaMallocAsync(&p, bytes, c->stream);
            if (*err != cudaSuccess) return nullptr;
        }
        cap = bytes;
    }
    c->big_live[p] = cap;
    return p;
}
// true when p was a cached block (now back in the cache)
static inline bool crx_big_give(crx_ctx* c, void* p) {
    auto it = c->big_live.find(p);
    if (it == c->big_live.end()) return false;
    c->big_free.emplace_back(p, it->second);
    c->big_live.erase(it);
    return true;
}

struct crx_points {
    crx_ctx* ctx = nullptr;
    int64_t n = 0;
    int d = 0;
    int ld = 0;              // row stride in elements (d rounded up to a multiple of 4; pad = 0)
    float* x32 = nullptr;    // [n][ld] always present (rounded copy when the source was double)
    double* x64 = nullptr;   // [n][ld] only when the source was double
    double* sqn = nullptr;   // [n] sum of squares, double, index order (cust_vector.hpp:148-151)
    uint8_t* unknown = nullptr;  // [n][d]
    double* mean = nullptr;      // [n]
    // lazily built caches of the tensor-core path (tc_scan.cuh); owned by the points object
    mutable double maxabs = -1.0;          // max |x| over all coordinates
    mutable struct TcOperand* tc_l2 = nullptr;   // split-fp16 rows times 2^tc_l2_scale (Euclidean scans)
    mutable int tc_l2_scale = 0;
};

// ------------------------------------------------------------------------------------------------
// launch bookkeeping: counts our own kernels and (optionally) times each with CUDA events
// ------------------------------------------------------------------------------------------------
struct crx_launch_scope {
    crx_ctx* c;
    crx_prof_rec rec;
    bool on;
    crx_launch_scope(crx_ctx* ctx, const char* name) : c(ctx), on(ctx->profiling) {
        c->launches++;
        if (on) {
            rec.name = name;
            auto get = [&](cudaEvent_t& e) {
                if (!c->free_events.empty()) { e = c->free_events.back(); c->free_events.pop_back(); }
                else cudaEventCreate(&e);
            };
            get(rec.e0); get(rec.e1);
            cudaEventRecord(rec.e0, c->stream);
        }
    }
    ~crx_launch_scope() {
        if (on) {
            cudaEventRecord(rec.e1, c->stream);
            c->prof.push_back(rec);
        }
    }
};
#define CRX_CONCAT2(a, b) a##b
#define CRX_CONCAT(a, b) CRX_CONCAT2(a, b)
#define CRX_KERNEL(ctx, name) crx_launch_scope CRX_CONCAT(scope__, __LINE__)(ctx, name)

// stream-ordered temporary device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    cudaStream_t s = nullptr;
    crx_ctx* ctx = nullptr;
    size_t count = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    int alloc(crx_ctx* c, size_t n) {
        release();
        s = c->stream;
        ctx = c;
        count = n;
        if (n == 0) n = 1;
        cudaError_t e = cudaSuccess;
        p = (T*)crx_big_take(c, n * sizeof(T), &e);
        if (!p && e == cudaSuccess) e = cudaMallocAsync((void**)&p, n * sizeof(T), s);
        if (e != cudaSuccess) {
            crx_set_error("cudaMallocAsync(%zu bytes) -> %s", n * sizeof(T), cudaGetErrorString(e));
            p = nullptr;
            return CRX_ERR_NOMEM;
        }
        return CRX_OK;
    }
    void release() {
        if (p && !(ctx && crx_big_give(ctx, p))) cudaFreeAsync(p, s);
        p = nullptr;
    }
    ~DevBuf() { release(); }
    operator T*() const { return p; }
};

// A caller buffer that may live on the host: gives a device pointer, copies in / out on demand.
template <typename T>
struct IoBuf {
    crx_ctx* c = nullptr;
    T* user = nullptr;
    T* dev = nullptr;
    size_t n = 0;
    int mem = CRX_DEVICE;
    DevBuf<T> tmp;
    int bind(crx_ctx* ctx, const T* ptr, size_t count, int where, bool copy_in) {
        c = ctx; user = const_cast<T*>(ptr); n = count; mem = where;
        if (!ptr) { dev = nullptr; return CRX_OK; }
        if (mem == CRX_DEVICE) { dev = user; return CRX_OK; }
        CRX_TRY(tmp.alloc(ctx, count));
        dev = tmp.p;
        if (copy_in) CRX_CUDA(cudaMemcpyAsync(dev, user, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        return CRX_OK;
    }
    // device -> host when the caller's buffer is on the host; synchronises the stream
    int flush() {
        if (!user || mem == CRX_DEVICE) return CRX_OK;
        CRX_CUDA(cudaMemcpyAsync(user, dev, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        return CRX_OK;
    }
};

// long-lived device buffers of handles: stream-ordered pool allocations too (the pool keeps freed memory,
// so rebuilding tables / points every step does not go back to the driver)
template <typename T>
static inline int crx_alloc(crx_ctx* c, T** p, size_t count) {
    cudaError_t e = cudaSuccess;
    *p = (T*)crx_big_take(c, (count ? count : 1) * sizeof(T), &e);
    if (*p) return CRX_OK;
    if (e == cudaSuccess) e = cudaMallocAsync((void**)p, (count ? count : 1) * sizeof(T), c->stream);
    if (e != cudaSuccess) {
        crx_set_error("cudaMallocAsync(%zu bytes) -> %s", count * sizeof(T), cudaGetErrorString(e));
        *p = nullptr;
        return CRX_ERR_NOMEM;
    }
    return CRX_OK;
}
static inline void crx_free(crx_ctx* c, void* p) {
    if (p && !crx_big_give(c, p)) cudaFreeAsync(p, c->stream);
}

static inline int crx_grid(int64_t work, int block) { return (int)((work + block - 1) / block); }

// ------------------------------------------------------------------------------------------------
// exact FP64 device arithmetic
//
// The reference is strict (non-fused) double arithmetic in index order, except its dot product,
// which accumulates double-rounded products in x87 extended precision.  The helpers below
// reproduce the former bit for bit (no FMA contraction: explicit __d*_rn intrinsics) and the
// latter by emulating the 64-bit-mantissa accumulator (x87.cuh: dot_x87 / cos_sim_x87, bit-identical
// to the long double code).  The compensated "dot2" product only decides hash signs / floors that
// fall inside the error bound of the fast projection (hash.cu).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ldv(const float* p, int i) { return (double)p[i]; }
__device__ __forceinline__ double ldv(const double* p, int i) { return p[i]; }

// sum_i x_i^2, double, index order (cust_vector.hpp:148-151: accum = accum + pow(x,2))
template <typename T>
__device__ __forceinline__ double sqnorm_exact(const T* __restrict__ a, int d) {
    double acc = 0.0;
    for (int i = 0; i < d; i++) {
        double v = ldv(a, i);
        acc = __dadd_rn(acc, __dmul_rn(v, v));
    }
    return acc;
}

// squared Euclidean accumulation exactly as cust_vector.hpp:126-136 (before the sqrt)
template <typename TA, typename TB>
__device__ __forceinline__ double euclid_acc_exact(const TA* __restrict__ a, const TB* __restrict__ b, int d) {
    double acc = 0.0;
    for (int i = 0; i < d; i++) {
        double t = __dsub_rn(ldv(a, i), ldv(b, i));
        acc = __dadd_rn(acc, __dmul_rn(t, t));
    }
    return acc;
}
template <typename TA, typename TB>
__device__ __forceinline__ double euclid_exact(const TA* __restrict__ a, const TB* __restrict__ b, int d) {
    return __dsqrt_rn(euclid_acc_exact(a, b, d));
}

#include "x87.cuh"   // two_sum, X87, x87_add, cos_sim_x87 (host-compilable: tests/test_x87_cpu.py checks them against long double)

// error-free product
__device__ __forceinline__ void two_prod(double a, double b, double& p, double& e) {
    p = __dmul_rn(a, b);
    e = __fma_rn(a, b, -p);
}
// compensated dot product: (hi, lo) with hi+lo accurate to ~2^-100 relative to sum|x_i y_i|
template <typename TA, typename TB>
__device__ __forceinline__ void dot2(const TA* __restrict__ a, const TB* __restrict__ b, int d, double& hi, double& lo) {
    double s = 0.0, c = 0.0;
    for (int i = 0; i < d; i++) {
        double p, pe, se;
        two_prod(ldv(a, i), ldv(b, i), p, pe);
        two_sum(s, p, s, se);
        c = __dadd_rn(c, __dadd_rn(pe, se));
    }
    two_sum(s, c, hi, lo);
}
template <typename TA, typename TB>
__device__ __forceinline__ X87 dot_x87(const TA* __restrict__ a, const TB* __restrict__ b, int d) {
    X87 acc = {0.0, 0.0};
    for (int i = 0; i < d; i++) x87_add(acc, __dmul_rn(ldv(a, i), ldv(b, i)));
    return acc;
}
// cosine similarity from a plain double inner product (filters and the batched top-P refine)
__device__ __forceinline__ double cos_sim_from(double ip, double na, double nb) {
    double denom = __dmul_rn(__dsqrt_rn(na), __dsqrt_rn(nb));
    return __ddiv_rn(ip, denom);
}
template <typename TA, typename TB>
__device__ __forceinline__ double cos_sim_exact(const TA* a, const TB* b, int d, double na, double nb) {
    return cos_sim_x87(dot_x87(a, b, d), na, nb);
}
template <typename TA, typename TB>
__device__ __forceinline__ double metric_dist_exact(int metric, const TA* a, const TB* b, int d, double na, double nb) {
    if (metric == CRX_EUCLIDEAN) return euclid_exact(a, b, d);
    return __dsub_rn(1.0, cos_sim_exact(a, b, d, na, nb));
}

// utils.hpp:97-98 for (int, int)
__host__ __device__ __forceinline__ int mod_ii(int x, int n) { return (x % n + n) % n; }

// crypto_rec.hpp:235-277: Lomuto partition (pivot = last, `>=` goes left) over two parallel arrays, explicit
// stack instead of recursion.  Two shortcuts that cannot change the result of the literal algorithm:
//   * a range whose keys are all equal is left untouched by it (every partition swaps elements with
//     themselves and recurses on the prefix) -- skip it instead of spending O(n^2) compares; predicted
//     scores tie massively (every coin unknown to all neighbours predicts exactly the user's mean);
//   * only the first `need` positions are consumed (resize(N), crypto_rec.hpp:322): ranges that start at or
//     beyond `need` never influence them and are skipped (quickselect-style pruning).
template <typename K, typename V>
__host__ __device__ inline void lomuto_desc(K* key, V* val, int n, int need) {
    int stack_lo[64], stack_hi[64];
    int sp = 0;
    stack_lo[0] = 0; stack_hi[0] = n - 1; sp = 1;
    while (sp > 0) {
        sp--;
        int lo = stack_lo[sp], hi = stack_hi[sp];
        while (lo < hi && lo < need) {
            K pivot = key[hi];
            bool all_equal = true;
            for (int j = lo; j < hi; j++)
                if (!(key[j] == pivot)) { all_equal = false; break; }
            if (all_equal) break;
            int i = lo - 1;
            for (int j = lo; j < hi; j++) {
                if (key[j] >= pivot) {
                    i++;
                    K tk = key[i]; key[i] = key[j]; key[j] = tk;
                    V tv = val[i]; val[i] = val[j]; val[j] = tv;
                }
            }
            K tk = key[i + 1]; key[i + 1] = key[hi]; key[hi] = tk;
            V tv = val[i + 1]; val[i + 1] = val[hi]; val[hi] = tv;
            int p = i + 1;
            // push the larger side, continue with the smaller: stack depth <= log2(n) + 1
            int llo = lo, lhi = p - 1, rlo = p + 1, rhi = hi;
            if (lhi - llo < rhi - rlo) {
                if (sp < 64) { stack_lo[sp] = rlo; stack_hi[sp] = rhi; sp++; }
                lo = llo; hi = lhi;
            } else {
                if (sp < 64) { stack_lo[sp] = llo; stack_hi[sp] = lhi; sp++; }
                lo = rlo; hi = rhi;
            }
        }
    }
}
template <typename K, typename V>
__host__ __device__ inline void lomuto_desc(K* key, V* val, int n) { lomuto_desc(key, val, n, n); }
