// hash.cu -- lib/generators + lib/data_structures of the reference on the GPU:
//   K1 batched projection (Euclidean floor((v.x+t)/w), cosine sign) with certified boundaries,
//   K2 phi / g combination and bucket index, K3 table build (stable radix sort + offsets),
//   hypercube vertex map with the order-dependent f(h) draws reproduced on the host.
#include <algorithm>
#include <climits>
#include <cub/cub.cuh>

#include "tables.cuh"

// ------------------------------------------------------------------------------------------------
// K1 + K2: projections of every row, 2 rows per thread, one table (k <= KC hashes) at a time.
//
// Arithmetic: FP64 FMA chain in index order with the a-priori bound
//     |computed - exact| <= (D+2) 2^-53 ||x|| ||r||
// A projection whose decision (sign / floor) lies inside the bound is recomputed with a
// compensated double-double dot product (counted in counters[CRX_CNT_HASH_DD]); the reference's
// own x87 accumulation (cust_vector.hpp:107-121) is less accurate than that fallback.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load4(const float* p, double o[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void load4(const double* p, double o[4]) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// int(floor(long double)) on x86-64 is an x87 FISTP: out-of-range values give the "integer indefinite"
// 0x80000000 for either sign (euclidean_h_gen.hpp:81 relies on it for |h| >= 2^31); a CUDA cast saturates.
__device__ __forceinline__ int f2i_x87(double f) {
    return (f >= -2147483648.0 && f < 2147483648.0) ? (int)f : INT_MIN;
}

template <typename T>
__device__ int cosine_bit_dd(const T* x, const double* r, int D) {
    double hi, lo;
    dot2(r, x, D, hi, lo);
    return (hi > 0.0 || (hi == 0.0 && lo >= 0.0)) ? 1 : 0;
}
template <typename T>
__device__ int euclid_h_dd(const T* x, const double* v, int D, double t, double w) {
    double hi, lo, s, e;
    dot2(v, x, D, hi, lo);
    two_sum(hi, t, s, e);
    e = __dadd_rn(e, lo);
    double f = floor(__ddiv_rn(s, w));
    double rem = __dadd_rn(__fma_rn(-w, f, s), e);  // (s + e) - w f
    if (rem < 0.0) f -= 1.0;
    else if (rem >= w) f += 1.0;
    return f2i_x87(f);
}

#define PHI_M 2147483647  // int(pow(2,32)-5) as GCC folds it (euclidean_phi_gen.hpp:70)

template <typename T, int KC>
__global__ void __launch_bounds__(128)
hash_rows_kernel(const T* __restrict__ x, int ld, const double* __restrict__ sqn, int64_t N, int D, int metric, int k,
                 int L, const double* __restrict__ proj, int ldp, const double* __restrict__ pnorm,
                 const float* __restrict__ tt, const int32_t* __restrict__ rr, float w, int nbuckets,
                 int32_t* __restrict__ hvals, int32_t* __restrict__ bucket, unsigned long long* counters,
                 const int32_t* __restrict__ rowlist /* nullable: only these rows */, const int* __restrict__ nlist) {
    extern __shared__ double sproj[];  // [k][ldp]
    int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    int64_t i1 = i0 + 1;
    bool v0 = i0 < N, v1 = i1 < N;
    if (rowlist) {  // the rows the fp32 filter could not decide
        const int64_t cnt = *nlist;
        v0 = i0 < cnt; v1 = i1 < cnt;
        i0 = v0 ? rowlist[i0] : 0;
        i1 = v1 ? rowlist[i1] : 0;
    }
    const T* row0 = x + (v0 ? i0 : 0) * ld;
    const T* row1 = x + (v1 ? i1 : 0) * ld;
    double nx0 = v0 ? sqrt(sqn[i0]) : 0.0, nx1 = v1 ? sqrt(sqn[i1]) : 0.0;
    const double cbound = (double)(D + 2) * 1.1102230246251565e-16 * 1.001;
    const double dw = (double)w;

    for (int l = 0; l < L; l++) {
        __syncthreads();
        for (int e = threadIdx.x; e < k * ldp; e += blockDim.x) sproj[e] = proj[(size_t)l * k * ldp + e];
        __syncthreads();
        double a0[KC], a1[KC];
#pragma unroll
        for (int h = 0; h < KC; h++) { a0[h] = 0.0; a1[h] = 0.0; }
        for (int i = 0; i < ld; i += 4) {
            double p0[4], p1[4];
            load4(row0 + i, p0);
            load4(row1 + i, p1);
#pragma unroll
            for (int u = 0; u < 4; u++) {
#pragma unroll
                for (int h = 0; h < KC; h++) {
                    if (h < k) {
                        double r = sproj[h * ldp + i + u];
                        a0[h] = __fma_rn(p0[u], r, a0[h]);
                        a1[h] = __fma_rn(p1[u], r, a1[h]);
                    }
                }
            }
        }
        // decisions
        int g0 = 0, g1 = 0;
        unsigned int phi0 = 0, phi1 = 0;
#pragma unroll
        for (int h = 0; h < KC; h++) {
            if (h < k) {
                double pn = pnorm[l * k + h];
                double E0 = cbound * nx0 * pn, E1 = cbound * nx1 * pn;
                int r0, r1;
                if (metric == CRX_COSINE) {
                    if (fabs(a0[h]) <= E0) { r0 = cosine_bit_dd(row0, sproj + h * ldp, D); if (v0) atomicAdd(&counters[CRX_CNT_HASH_DD], 1ull); }
                    else r0 = a0[h] >= 0.0 ? 1 : 0;
                    if (fabs(a1[h]) <= E1) { r1 = cosine_bit_dd(row1, sproj + h * ldp, D); if (v1) atomicAdd(&counters[CRX_CNT_HASH_DD], 1ull); }
                    else r1 = a1[h] >= 0.0 ? 1 : 0;
                    g0 = (g0 << 1) + r0;  // cosine_g_gen.hpp:62-72: function 0 is the MSB
                    g1 = (g1 << 1) + r1;
                } else {
                    double t = (double)tt[l * k + h];
                    {
                        double s = a0[h] + t, y = s / dw, f = floor(y), fr = y - f;
                        double Ey = (E0 + fabs(s) * 2.3e-16) / dw + fabs(y) * 2.3e-16;
                        if (fmin(fr, 1.0 - fr) <= Ey) { r0 = euclid_h_dd(row0, sproj + h * ldp, D, t, dw); if (v0) atomicAdd(&counters[CRX_CNT_HASH_DD], 1ull); }
                        else r0 = f2i_x87(f);
                    }
                    {
                        double s = a1[h] + t, y = s / dw, f = floor(y), fr = y - f;
                        double Ey = (E1 + fabs(s) * 2.3e-16) / dw + fabs(y) * 2.3e-16;
                        if (fmin(fr, 1.0 - fr) <= Ey) { r1 = euclid_h_dd(row1, sproj + h * ldp, D, t, dw); if (v1) atomicAdd(&counters[CRX_CNT_HASH_DD], 1ull); }
                        else r1 = f2i_x87(f);
                    }
                    if (hvals) {
                        if (v0) hvals[((size_t)l * N + i0) * k + h] = r0;
                        if (v1) hvals[((size_t)l * N + i1) * k + h] = r1;
                    }
                    if (rr) {  // euclidean_phi_gen.hpp:85-96
                        int ri = rr[l * k + h];
                        long long t0 = (long long)(int)((unsigned int)r0 * (unsigned int)ri);
                        long long t1 = (long long)(int)((unsigned int)r1 * (unsigned int)ri);
                        phi0 += (unsigned int)(int)((t0 % PHI_M + PHI_M) % PHI_M);
                        phi1 += (unsigned int)(int)((t1 % PHI_M + PHI_M) % PHI_M);
                    }
                }
            }
        }
        if (bucket) {
            if (metric == CRX_COSINE) {
                if (v0) bucket[(size_t)l * N + i0] = g0 % nbuckets;
                if (v1) bucket[(size_t)l * N + i1] = g1 % nbuckets;
            } else if (rr) {
                unsigned int M = (unsigned int)PHI_M;
                unsigned int f0 = (phi0 % M + M) % M, f1 = (phi1 % M + M) % M;
                if (v0) bucket[(size_t)l * N + i0] = (int)((unsigned long long)f0 % (unsigned long long)nbuckets);
                if (v1) bucket[(size_t)l * N + i1] = (int)((unsigned long long)f1 % (unsigned long long)nbuckets);
            }
        }
    }
}

// ---- fp32 generation of K1/K2 -----------------------------------------------------------------------------------
// All H = L*k projections of a row in ONE pass over its fp32 copy: a thread owns two rows, the projection vectors sit in
// shared memory as floats and are read four coordinates at a time (one broadcast LDS.128 feeds 8 FFMA), which is what
// lifts the kernel off the shared-memory bandwidth the FP64 version is bound by.  Every decision (sign, floor) is taken
// only when the fp32 value clears its boundary by more than the rounding bound
//   |fp32 dot - exact| <= (3 D + 8) 2^-24 |x| |r|      (FMA chain + the float conversions of x and r);
// rows with an undecided projection are listed and redone by the exact kernel above.
template <int KH>
__global__ void __launch_bounds__(128)
hash_rows32_kernel(const float* __restrict__ x, int ld, const double* __restrict__ sqn, int64_t N, int D, int metric, int k,
                   int L, const double* __restrict__ proj, int ldp, const double* __restrict__ pnorm,
                   const float* __restrict__ tt, const int32_t* __restrict__ rr, float w, int nbuckets,
                   int32_t* __restrict__ hvals, int32_t* __restrict__ bucket, int32_t* __restrict__ rowlist, int* __restrict__ nlist) {
    extern __shared__ float sp32[];          // [H][ldp] projections, then the row tile
    const int H = L * k;
    float* tile = sp32 + H * ldp;             // 2 x [256 rows][20]: 16 coordinates of the block's 256 rows; the 80-byte row pitch keeps
                                              // the 128-bit reads of a quarter warp (8 rows) on 8 different bank groups
    for (int e = threadIdx.x; e < H * ldp; e += blockDim.x) sp32[e] = (float)proj[e];
    // thread t owns rows base + t and base + 128 + t; the global reads are cooperative and coalesced (64-byte pieces).
    // A block walks groups of 256 rows (grid = a few blocks per SM): the projections are staged once per block, not per group.
    const int t = threadIdx.x;
    for (int64_t base = (int64_t)blockIdx.x * 256; base < N; base += (int64_t)gridDim.x * 256) {
    int64_t i0 = base + t, i1 = base + 128 + t;
    bool v0 = i0 < N, v1 = i1 < N;
    float a0[KH], a1[KH];
#pragma unroll
    for (int h = 0; h < KH; h++) { a0[h] = 0.f; a1[h] = 0.f; }
    // two tiles: the 16-byte asynchronous copies of chunk c+1 are in flight while chunk c is multiplied (the kernel was bound by
    // the exposed latency of its global loads at 16 warps per SM, not by HBM or the FMA pipe)
    auto fetch = [&](int c0, float* dst_tile) {
#pragma unroll
        for (int it = 0; it < 8; it++) {
            int idx = it * 128 + t;
            int row = idx >> 2, piece = (idx & 3) * 4;
            const bool in = base + row < N && c0 + piece < ld;
            const float* src = in ? x + (base + row) * ld + c0 + piece : x;
            unsigned dst = (unsigned)__cvta_generic_to_shared(dst_tile + row * 20 + piece);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");   // src-size 0: zero fill
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(0, tile);
    int buf = 0;
    for (int c0 = 0; c0 < ld; c0 += 16, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();   // chunk c0 landed for every thread; the other tile is no longer read (and, first time, the projections are staged)
        const float* cur = tile + buf * (256 * 20);
        float p0[16], p1[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float4 u0 = *reinterpret_cast<const float4*>(cur + t * 20 + 4 * i), u1 = *reinterpret_cast<const float4*>(cur + (128 + t) * 20 + 4 * i);
            p0[4 * i] = u0.x; p0[4 * i + 1] = u0.y; p0[4 * i + 2] = u0.z; p0[4 * i + 3] = u0.w;
            p1[4 * i] = u1.x; p1[4 * i + 1] = u1.y; p1[4 * i + 2] = u1.z; p1[4 * i + 3] = u1.w;
        }
        if (c0 + 16 < ld) fetch(c0 + 16, tile + (buf ^ 1) * (256 * 20));
        if (c0 + 16 <= ld) {
#pragma unroll
            for (int h = 0; h < KH; h++) {
                if (h < H) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {   // one broadcast LDS.128 feeds 8 FFMA
                        float4 r = *reinterpret_cast<const float4*>(sp32 + h * ldp + c0 + 4 * j);
                        a0[h] = fmaf(p0[4 * j], r.x, a0[h]); a1[h] = fmaf(p1[4 * j], r.x, a1[h]);
                        a0[h] = fmaf(p0[4 * j + 1], r.y, a0[h]); a1[h] = fmaf(p1[4 * j + 1], r.y, a1[h]);
                        a0[h] = fmaf(p0[4 * j + 2], r.z, a0[h]); a1[h] = fmaf(p1[4 * j + 2], r.z, a1[h]);
                        a0[h] = fmaf(p0[4 * j + 3], r.w, a0[h]); a1[h] = fmaf(p1[4 * j + 3], r.w, a1[h]);
                    }
                }
            }
        } else {   // last, partial chunk (ld is a multiple of 4)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (c0 + 4 * j < ld) {
#pragma unroll
                    for (int h = 0; h < KH; h++) {
                        if (h < H) {
                            float4 r = *reinterpret_cast<const float4*>(sp32 + h * ldp + c0 + 4 * j);
                            a0[h] = fmaf(p0[4 * j], r.x, a0[h]); a1[h] = fmaf(p1[4 * j], r.x, a1[h]);
                            a0[h] = fmaf(p0[4 * j + 1], r.y, a0[h]); a1[h] = fmaf(p1[4 * j + 1], r.y, a1[h]);
                            a0[h] = fmaf(p0[4 * j + 2], r.z, a0[h]); a1[h] = fmaf(p1[4 * j + 2], r.z, a1[h]);
                            a0[h] = fmaf(p0[4 * j + 3], r.w, a0[h]); a1[h] = fmaf(p1[4 * j + 3], r.w, a1[h]);
                        }
                    }
                }
            }
        }
    }
    const double cb = (double)(3 * D + 8) * 5.9604644775390625e-8 * 1.001;
    const double nx0 = v0 ? sqrt(sqn[i0]) : 0.0, nx1 = v1 ? sqrt(sqn[i1]) : 0.0;
    // y = (v.x + t) / w is formed as (v.x + t) * (1 / w): two roundings instead of one, both inside the bound Ey that a
    // decision has to clear (a double division costs ~30 instructions and there are 4 per row and projection otherwise)
    const double inv_w = 1.0 / (double)w;
    bool unsure0 = false, unsure1 = false;
    for (int l = 0; l < L; l++) {
        int g0 = 0, g1 = 0;
        unsigned int phi0 = 0, phi1 = 0;
#pragma unroll
        for (int h = 0; h < KH; h++) {
            if (h >= l * k && h < (l + 1) * k) {   // (compile-time h, run-time table bounds: registers stay registers)
                double pn = pnorm[h];
                double E0 = cb * nx0 * pn, E1 = cb * nx1 * pn;
                int r0, r1;
                if (metric == CRX_COSINE) {
                    unsure0 |= fabs((double)a0[h]) <= E0;
                    unsure1 |= fabs((double)a1[h]) <= E1;
                    r0 = a0[h] >= 0.f ? 1 : 0;
                    r1 = a1[h] >= 0.f ? 1 : 0;
                    g0 = (g0 << 1) + r0;
                    g1 = (g1 << 1) + r1;
                } else {
                    double t = (double)tt[h];
                    {
                        double sv = (double)a0[h] + t, y = sv * inv_w, f = floor(y), fr = y - f;
                        double Ey = (E0 + fabs(sv) * 2.3e-16) * (inv_w * 1.000001) + fabs(y) * 4.6e-16;
                        unsure0 |= !(fmin(fr, 1.0 - fr) > Ey);
                        r0 = f2i_x87(f);
                    }
                    {
                        double sv = (double)a1[h] + t, y = sv * inv_w, f = floor(y), fr = y - f;
                        double Ey = (E1 + fabs(sv) * 2.3e-16) * (inv_w * 1.000001) + fabs(y) * 4.6e-16;
                        unsure1 |= !(fmin(fr, 1.0 - fr) > Ey);
                        r1 = f2i_x87(f);
                    }
                    if (hvals) {
                        if (v0) hvals[((size_t)l * N + i0) * k + (h - l * k)] = r0;
                        if (v1) hvals[((size_t)l * N + i1) * k + (h - l * k)] = r1;
                    }
                    if (rr) {  // euclidean_phi_gen.hpp:85-96
                        int ri = rr[h];
                        long long t0 = (long long)(int)((unsigned int)r0 * (unsigned int)ri);
                        long long t1 = (long long)(int)((unsigned int)r1 * (unsigned int)ri);
                        phi0 += (unsigned int)(int)((t0 % PHI_M + PHI_M) % PHI_M);
                        phi1 += (unsigned int)(int)((t1 % PHI_M + PHI_M) % PHI_M);
                    }
                }
            }
        }
        if (bucket) {
            if (metric == CRX_COSINE) {
                if (v0) bucket[(size_t)l * N + i0] = g0 % nbuckets;
                if (v1) bucket[(size_t)l * N + i1] = g1 % nbuckets;
            } else if (rr) {
                unsigned int M = (unsigned int)PHI_M;
                unsigned int f0 = (phi0 % M + M) % M, f1 = (phi1 % M + M) % M;
                if (v0) bucket[(size_t)l * N + i0] = (int)((unsigned long long)f0 % (unsigned long long)nbuckets);
                if (v1) bucket[(size_t)l * N + i1] = (int)((unsigned long long)f1 % (unsigned long long)nbuckets);
            }
        }
    }
    if (v0 && unsure0) rowlist[atomicAdd(nlist, 1)] = (int32_t)i0;
    if (v1 && unsure1) rowlist[atomicAdd(nlist, 1)] = (int32_t)i1;
    __syncthreads();   // the tiles are free for the next group
    }
}

template <typename T>
static int launch_hash(crx_ctx* c, const T* x, const crx_points* pts, int metric, int k, int L, const double* d_proj,
                       int ldp, const double* d_pnorm, const float* d_t, const int32_t* d_r, float w, int nbuckets,
                       int32_t* hvals, int32_t* bucket, const int32_t* rowlist = nullptr, const int* nlist = nullptr,
                       int64_t list_cap = 0) {
    int64_t N = pts->n;
    int grid = (int)(((rowlist ? list_cap : N) + 255) / 256);
    if (grid == 0) return CRX_OK;
    size_t smem = (size_t)k * ldp * sizeof(double);
    CRX_KERNEL(c, "hash_rows");
#define LAUNCH_H(KC)                                                                                              \
    do {                                                                                                          \
        CRX_CUDA(cudaFuncSetAttribute(hash_rows_kernel<T, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        hash_rows_kernel<T, KC><<<grid, 128, smem, c->stream>>>(x, pts->ld, pts->sqn, N, pts->d, metric, k, L, d_proj, ldp, \
                                                                 d_pnorm, d_t, d_r, w, nbuckets, hvals, bucket, c->counters, rowlist, nlist); \
    } while (0)
    if (k <= 4) LAUNCH_H(4);
    else if (k <= 8) LAUNCH_H(8);
    else LAUNCH_H(16);
#undef LAUNCH_H
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

int crx_hash_rows(crx_ctx* c, const crx_points* pts, int metric, int k, int L, const double* d_proj, int ldp,
                  const double* d_pnorm, const float* d_t, const int32_t* d_r, float w, int nbuckets, int32_t* hvals,
                  int32_t* bucket) {
    CRX_REQUIRE(k >= 1 && k <= 16, "k (hash functions per table / cube dimension) must be in [1,16]");
    CRX_REQUIRE(ldp == pts->ld, "projection stride");
    const int H = L * k;
    static const bool f32_off = getenv("CRX_HASH_F64") != nullptr && getenv("CRX_HASH_F64")[0] == '1';
    const int32_t* rowlist = nullptr;
    const int* nlist = nullptr;
    DevBuf<int32_t> list;
    DevBuf<int> count;
    int64_t cap = 0;
    if (!f32_off && H <= 32 && pts->n >= 1024) {
        // fp32 pass over every row, then the exact kernel over the rows it could not decide
        int64_t N = pts->n;
        CRX_TRY(list.alloc(c, N)); CRX_TRY(count.alloc(c, 1));
        CRX_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int), c->stream));
        int grid = (int)std::min<int64_t>((N + 255) / 256, (int64_t)c->sm_count * 4);
        size_t smem = ((size_t)H * ldp + 2 * 256 * 20) * sizeof(float);
        {
            CRX_KERNEL(c, "hash_rows32");
#define LAUNCH_F(KH)                                                                                                       \
    do {                                                                                                                   \
        CRX_CUDA(cudaFuncSetAttribute(hash_rows32_kernel<KH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        hash_rows32_kernel<KH><<<grid, 128, smem, c->stream>>>(pts->x32, pts->ld, pts->sqn, N, pts->d, metric, k, L, d_proj, ldp, d_pnorm, d_t, \
                                                              d_r, w, nbuckets, hvals, bucket, list.p, count.p);           \
    } while (0)
            if (H <= 8) LAUNCH_F(8);
            else if (H <= 16) LAUNCH_F(16);
            else if (H <= 24) LAUNCH_F(24);
            else LAUNCH_F(32);
#undef LAUNCH_F
            CRX_CUDA(cudaGetLastError());
        }
        int h_count = 0;
        CRX_CUDA(cudaMemcpyAsync(&h_count, count.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        if (h_count == 0) return CRX_OK;
        rowlist = list.p; nlist = count.p; cap = h_count;
    }
    int st;
    if (pts->x64)
        st = launch_hash<double>(c, pts->x64, pts, metric, k, L, d_proj, ldp, d_pnorm, d_t, d_r, w, nbuckets, hvals, bucket, rowlist, nlist, cap);
    else
        st = launch_hash<float>(c, pts->x32, pts, metric, k, L, d_proj, ldp, d_pnorm, d_t, d_r, w, nbuckets, hvals, bucket, rowlist, nlist, cap);
    if (st == CRX_OK && rowlist) CRX_CUDA(cudaStreamSynchronize(c->stream));   // the list is freed on return
    return st;
}

// ------------------------------------------------------------------------------------------------
// K3: segments (stable sort by key + offsets)
// ------------------------------------------------------------------------------------------------
__global__ void iota_kernel(int32_t* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}
// off[b] = lower_bound(sorted, b) for b in [0, nkeys]
__global__ void offsets_kernel(const int32_t* __restrict__ sorted, int64_t n, int nkeys, int32_t* __restrict__ off) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nkeys) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t m = (lo + hi) >> 1;
        if (sorted[m] < b) lo = m + 1; else hi = m;
    }
    off[b] = (int32_t)lo;
}

static int sort_pairs(crx_ctx* c, const int32_t* kin, int32_t* kout, const int32_t* vin, int32_t* vout, int64_t n,
                      int end_bit) {
    size_t bytes = 0;
    CRX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, c->stream));
    DevBuf<char> tmp;
    CRX_TRY(tmp.alloc(c, bytes));
    CRX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, c->stream));
    return CRX_OK;
}

int crx_build_segments(crx_ctx* c, const int32_t* keys, int64_t n, int nkeys, Segments* out) {
    out->n = n; out->nkeys = nkeys; out->owner = c;
    CRX_TRY(crx_alloc(c, &out->perm, n));
    CRX_TRY(crx_alloc(c, &out->sorted, n));
    CRX_TRY(crx_alloc(c, &out->off, (size_t)nkeys + 1));
    DevBuf<int32_t> iota;
    CRX_TRY(iota.alloc(c, n));
    { CRX_KERNEL(c, "iota"); iota_kernel<<<crx_grid(n, 256), 256, 0, c->stream>>>(iota.p, n); }
    int bits = 1;
    while ((1ll << bits) < nkeys && bits < 31) bits++;
    CRX_TRY(sort_pairs(c, keys, out->sorted, iota.p, out->perm, n, bits));
    { CRX_KERNEL(c, "bucket_offsets"); offsets_kernel<<<crx_grid(nkeys + 1, 256), 256, 0, c->stream>>>(out->sorted, n, nkeys, out->off); }
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

// ------------------------------------------------------------------------------------------------
// euclidean filtered groups: dense rank of the k-tuple of h values (cust_hashtable.hpp:81-97)
// ------------------------------------------------------------------------------------------------
__global__ void gather_h_kernel(const int32_t* __restrict__ hv, const int32_t* __restrict__ perm, int64_t n, int k, int j,
                                int32_t* __restrict__ keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = hv[(size_t)perm[i] * k + j];
}
__global__ void tuple_flag_kernel(const int32_t* __restrict__ hv, const int32_t* __restrict__ perm, int64_t n, int k,
                                  int32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = 0;
    if (i > 0) {
        const int32_t* a = hv + (size_t)perm[i] * k;
        const int32_t* b = hv + (size_t)perm[i - 1] * k;
        for (int j = 0; j < k; j++) f |= (a[j] != b[j]);
    }
    flag[i] = f;
}
__global__ void scatter_rank_kernel(const int32_t* __restrict__ rank, const int32_t* __restrict__ perm, int64_t n,
                                    int32_t* __restrict__ gid) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gid[perm[i]] = rank[i];
}

// minimum / maximum of the j-th h value over all rows, j = 0..k-1 (hv is [rows][k]; mn / mx preset to INT_MAX / INT_MIN)
__global__ void cube_minmax_kernel(const int32_t* __restrict__ hv, int64_t total, int k, int* __restrict__ mn, int* __restrict__ mx) {
    __shared__ int smn[16], smx[16];
    if (threadIdx.x < 16) { smn[threadIdx.x] = INT_MAX; smx[threadIdx.x] = INT_MIN; }
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        int f = (int)(p % k), h = hv[p];
        if (h < smn[f]) atomicMin(&smn[f], h);
        if (h > smx[f]) atomicMax(&smx[f], h);
    }
    __syncthreads();
    if (threadIdx.x < k) { atomicMin(&mn[threadIdx.x], smn[threadIdx.x]); atomicMax(&mx[threadIdx.x], smx[threadIdx.x]); }
}
// lexicographic key of the k-tuple in as few bits as its value ranges allow: h_0 most significant
__global__ void pack_tuple_kernel(const int32_t* __restrict__ hv, int64_t n, int k, const int* __restrict__ mn,
                                  const int* __restrict__ shift, unsigned long long* __restrict__ key, int32_t* __restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long v = 0;
    for (int j = 0; j < k; j++) v |= (unsigned long long)((long long)hv[i * k + j] - (long long)mn[j]) << shift[j];
    key[i] = v;
    idx[i] = (int32_t)i;
}

static int build_tuple_groups(crx_ctx* c, const int32_t* hv /* [N][k] */, int64_t n, int k, int32_t* gid, Segments* seg,
                              int* ngroups) {
    DevBuf<int32_t> pa, pb, ka, kb;
    CRX_TRY(pa.alloc(c, n)); CRX_TRY(pb.alloc(c, n)); CRX_TRY(ka.alloc(c, n)); CRX_TRY(kb.alloc(c, n));
    int g = crx_grid(n, 256);
    int32_t* pin = pa.p; int32_t* pout = pb.p;
    // value ranges of the k components: when the whole tuple fits 64 bits, ONE stable sort of the packed key replaces the
    // k LSD passes over 32-bit keys
    bool packed = false;
    if (k <= 16) {
        DevBuf<int> dmn, dmx, dshift;
        CRX_TRY(dmn.alloc(c, 16)); CRX_TRY(dmx.alloc(c, 16)); CRX_TRY(dshift.alloc(c, 16));
        std::vector<int> hmn(16, INT_MAX), hmx(16, INT_MIN), shift(16, 0);
        CRX_CUDA(cudaMemcpyAsync(dmn.p, hmn.data(), 16 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CRX_CUDA(cudaMemcpyAsync(dmx.p, hmx.data(), 16 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        { CRX_KERNEL(c, "cube_minmax"); cube_minmax_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(hv, n * k, k, dmn.p, dmx.p); }
        CRX_CUDA(cudaMemcpyAsync(hmn.data(), dmn.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaMemcpyAsync(hmx.data(), dmx.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        int bits = 0;
        for (int j = k - 1; j >= 0; j--) {
            unsigned long long range = (unsigned long long)((long long)hmx[j] - (long long)hmn[j]);
            int b = 1;
            while (b < 33 && (range >> b) != 0) b++;
            shift[j] = bits;
            bits += b;
        }
        if (bits <= 64) {
            packed = true;
            DevBuf<unsigned long long> k0, k1;
            CRX_TRY(k0.alloc(c, n)); CRX_TRY(k1.alloc(c, n));
            CRX_CUDA(cudaMemcpyAsync(dshift.p, shift.data(), 16 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            { CRX_KERNEL(c, "pack_tuple"); pack_tuple_kernel<<<g, 256, 0, c->stream>>>(hv, n, k, dmn.p, dshift.p, k0.p, pin); }
            size_t sb = 0;
            CRX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sb, k0.p, k1.p, pin, pout, (int)n, 0, bits, c->stream));
            DevBuf<char> stmp;
            CRX_TRY(stmp.alloc(c, sb));
            CRX_CUDA(cub::DeviceRadixSort::SortPairs(stmp.p, sb, k0.p, k1.p, pin, pout, (int)n, 0, bits, c->stream));
            CRX_CUDA(cudaStreamSynchronize(c->stream));   // the temporaries go out of scope
            std::swap(pin, pout);
        }
    }
    if (!packed) {
        { CRX_KERNEL(c, "iota"); iota_kernel<<<g, 256, 0, c->stream>>>(pa.p, n); }
        for (int j = k - 1; j >= 0; j--) {  // LSD passes, each stable => lexicographic order, ties by row
            { CRX_KERNEL(c, "gather_h"); gather_h_kernel<<<g, 256, 0, c->stream>>>(hv, pin, n, k, j, ka.p); }
            CRX_TRY(sort_pairs(c, ka.p, kb.p, pin, pout, n, 32));
            std::swap(pin, pout);
        }
    }
    { CRX_KERNEL(c, "tuple_flag"); tuple_flag_kernel<<<g, 256, 0, c->stream>>>(hv, pin, n, k, ka.p); }
    size_t bytes = 0;
    CRX_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, ka.p, kb.p, (int)n, c->stream));
    DevBuf<char> tmp;
    CRX_TRY(tmp.alloc(c, bytes));
    CRX_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, bytes, ka.p, kb.p, (int)n, c->stream));
    { CRX_KERNEL(c, "scatter_rank"); scatter_rank_kernel<<<g, 256, 0, c->stream>>>(kb.p, pin, n, gid); }
    int32_t last = 0;
    CRX_CUDA(cudaMemcpyAsync(&last, kb.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    *ngroups = last + 1;
    seg->n = n; seg->nkeys = *ngroups; seg->owner = c;
    CRX_TRY(crx_alloc(c, &seg->perm, n));
    CRX_TRY(crx_alloc(c, &seg->sorted, n));
    CRX_TRY(crx_alloc(c, &seg->off, (size_t)*ngroups + 1));
    CRX_CUDA(cudaMemcpyAsync(seg->perm, pin, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    CRX_CUDA(cudaMemcpyAsync(seg->sorted, kb.p, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    { CRX_KERNEL(c, "bucket_offsets"); offsets_kernel<<<crx_grid(*ngroups + 1, 256), 256, 0, c->stream>>>(seg->sorted, n, *ngroups, seg->off); }
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

// ------------------------------------------------------------------------------------------------
// host parameter generation: the reference's libstdc++ <random> calls in the reference's order
// 
// ------------------------------------------------------------------------------------------------
static void draw_euclid_h(std::default_random_engine& e, int D, float w, float* v, float* t) {
    std::normal_distribution<float> nd(0, 1);  // euclidean_h_gen.hpp:60-64
    for (int i = 0; i < D; i++) v[i] = nd(e);
    std::uniform_real_distribution<float> ud(0, w);  // euclidean_h_gen.hpp:68-69
    *t = ud(e);
}
static void draw_cosine_h(std::default_random_engine& e, int D, double* r) {
    std::normal_distribution<double> nd(0, 1);  // cosine_h_gen.hpp:54-59
    for (int i = 0; i < D; i++) r[i] = nd(e);
}

static int upload_proj(crx_ctx* c, int H, int D, int ld, const double* cos_r, const float* euc_v, double** d_proj,
                       double** d_pnorm) {
    std::vector<double> proj((size_t)H * ld, 0.0), pn(H);
    for (int h = 0; h < H; h++) {
        double s = 0;
        for (int i = 0; i < D; i++) {
            double v = cos_r ? cos_r[(size_t)h * D + i] : (double)euc_v[(size_t)h * D + i];
            proj[(size_t)h * ld + i] = v;
            s += v * v;
        }
        pn[h] = std::sqrt(s) * 1.0000001;
    }
    CRX_TRY(crx_alloc(c, d_proj, proj.size()));
    CRX_TRY(crx_alloc(c, d_pnorm, pn.size()));
    CRX_CUDA(cudaMemcpyAsync(*d_proj, proj.data(), proj.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CRX_CUDA(cudaMemcpyAsync(*d_pnorm, pn.data(), pn.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}

// hamming ball enumeration: utils.cpp:22-50 restated as lexicographic combinations of bit positions
static void hamming_ball(int num, int dist, int min_bit, int bits, std::vector<int>& out) {
    if (dist < 1 || dist > bits - min_bit) return;
    std::vector<int> pos(dist);
    for (int i = 0; i < dist; i++) pos[i] = min_bit + i;
    for (;;) {
        int v = num;
        for (int i = 0; i < dist; i++) v ^= (1 << pos[i]);
        out.push_back(v);
        int i = dist - 1;
        while (i >= 0 && pos[i] == bits - dist + i) i--;
        if (i < 0) break;
        pos[i]++;
        for (int j = i + 1; j < dist; j++) pos[j] = pos[j - 1] + 1;
    }
}

// vertex visit order of get_hypercube_combined_buckets (lsh_cube.hpp:140-177), home first
void crx_cube_probe_sequence(int home, int probes, int k, std::vector<int>& seq) {
    seq.clear();
    seq.push_back(home);
    std::vector<int> neigh;
    size_t ni = 0;
    if (probes > 1) hamming_ball(home, 1, 0, k, neigh);
    int left = probes, dist = 1;
    while (left > 0) {
        if (ni < neigh.size()) { seq.push_back(neigh[ni++]); left--; }
        else {
            dist++;
            neigh.clear();
            hamming_ball(home, dist, 0, k, neigh);
            ni = 0;
            if (neigh.empty()) break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// hypercube f-map kernels
// ------------------------------------------------------------------------------------------------
__global__ void cube_keys_kernel(const int32_t* __restrict__ hv, int64_t total, int k, unsigned long long* __restrict__ keys,
                                 uint32_t* __restrict__ pos) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    int f = (int)(p % k);
    keys[p] = ((unsigned long long)f << 32) | (unsigned long long)((uint32_t)hv[p] ^ 0x80000000u);
    pos[p] = (uint32_t)p;
}
__global__ void cube_heads_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ pos, int64_t total,
                                  unsigned long long* __restrict__ hkeys, uint32_t* __restrict__ hpos, int cap,
                                  int* __restrict__ count) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    if (p == 0 || keys[p] != keys[p - 1]) {
        int s = atomicAdd(count, 1);
        if (s < cap) { hkeys[s] = keys[p]; hpos[s] = pos[p]; }
    }
}
// Dense alternative to the sort when the h values of every f span a small range (w not tiny): per-f minimum / maximum
// (cube_minmax_kernel above), then table[f][h - min_f] = smallest position p = row * k + f at which (f, h) occurs.  The table is read before the
// atomic, so once the early rows have claimed their entries the pass is read-only.
__global__ void cube_first_kernel(const int32_t* __restrict__ hv, int64_t total, int k, const int* __restrict__ mn,
                                  const int64_t* __restrict__ toff, uint32_t* __restrict__ table) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    int f = (int)(p % k);
    int64_t idx = toff[f] + ((int64_t)hv[p] - (int64_t)mn[f]);
    if ((uint32_t)p < table[idx]) atomicMin(&table[idx], (uint32_t)p);
}
// vertex = concatenation of f_i(h_i) bits, f_0 = MSB (hypercube_gen.hpp:63-73); maps: sorted h per f
__global__ void cube_vertex_kernel(const int32_t* __restrict__ hv, int64_t N, int k, const int32_t* __restrict__ map_h,
                                   const int32_t* __restrict__ map_bit, const int32_t* __restrict__ map_off,
                                   int32_t* __restrict__ vertex) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int g = 0;
    for (int f = 0; f < k; f++) {
        int h = hv[i * k + f];
        int lo = map_off[f], hi = map_off[f + 1];
        while (lo < hi) {
            int m = (lo + hi) >> 1;
            if (map_h[m] < h) lo = m + 1; else hi = m;
        }
        g = (g << 1) + map_bit[lo];
    }
    vertex[i] = g;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
// bucket (or h-tuple group) of one stored row in every table and the extent [begin, end) of that bucket's member list
struct BucketMeta { const int32_t* ids[16]; const int32_t* off[16]; int L; };
__global__ void bucket_meta_kernel(BucketMeta m, int32_t* __restrict__ ext) {
    int l = threadIdx.x;
    if (l >= m.L) return;
    int g = *m.ids[l];
    ext[2 * l] = m.off[l][g];
    ext[2 * l + 1] = m.off[l][g + 1];
}

extern "C" {

int crx_get_num_hamming_dist_from(int num, int dist, int min_bit, int bits, int32_t* out, int cap) {
    std::vector<int> r;
    hamming_ball(num, dist, min_bit, bits, r);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = r[i];
    return (int)r.size();
}

int crx_lsh_destroy(crx_lsh* t) {
    if (!t) return CRX_OK;
    cudaSetDevice(t->ctx->device);
    crx_ctx* c = t->ctx;
    crx_free(c, t->d_proj); crx_free(c, t->d_pnorm); crx_free(c, t->d_t); crx_free(c, t->d_r);
    crx_free(c, t->hvals); crx_free(c, t->bucket);
    if (t->gid != t->bucket) crx_free(c, t->gid);
    for (size_t l = 0; l < t->by_bucket.size(); l++) {
        if (l < t->by_group.size() && t->by_group[l].perm != t->by_bucket[l].perm) t->by_group[l].free_all();
        t->by_bucket[l].free_all();
    }
    delete t;
    return CRX_OK;
}

int crx_create_LSH_hashtables(crx_ctx* c, const crx_points* pts, int metric, int k, int L, int lsh_bucket_div,
                              double euclidean_h_w, uint64_t seed, crx_lsh** out) {
    CRX_REQUIRE(pts, "NULL argument");
    CRX_NARROW(pts);
    CRX_REQUIRE(c && pts && out, "NULL argument");
    CRX_REQUIRE(metric == CRX_EUCLIDEAN || metric == CRX_COSINE, "metric");
    CRX_REQUIRE(k >= 1 && k <= 16 && L >= 1 && L <= 16, "k in [1,16], L in [1,16]");
    CRX_CUDA(cudaSetDevice(c->device));
    int64_t N = pts->n;
    int D = pts->d;
    crx_lsh* t = new crx_lsh();
    t->ctx = c; t->pts = pts; t->metric = metric; t->k = k; t->L = L; t->D = D; t->N = N; t->w = (float)euclidean_h_w;
    t->ldp = pts->ld;
    if (metric == CRX_EUCLIDEAN) {
        CRX_REQUIRE(lsh_bucket_div > 0 && (int64_t)((size_t)N / (size_t)lsh_bucket_div) > 0,
                    "euclidean LSH needs N / lsh_bucket_div >= 1 buckets (lsh_cube.hpp:60)");
        t->nbuckets = (int)((size_t)N / (size_t)lsh_bucket_div);
    } else t->nbuckets = (int)std::pow(2, k);  // lsh_cube.hpp:65
    int H = L * k;
    // lsh_cube.hpp:49-51 + App. B consumption order
    std::default_random_engine e;
    e.seed((unsigned long)seed);
    if (metric == CRX_EUCLIDEAN) {
        t->euc_v.resize((size_t)H * D); t->euc_t.resize(H); t->euc_r.resize(H);
        for (int l = 0; l < L; l++) {
            std::uniform_int_distribution<int> ui(0, 100);  // euclidean_phi_gen.hpp:64
            for (int j = 0; j < k; j++) {
                int h = l * k + j;
                draw_euclid_h(e, D, t->w, &t->euc_v[(size_t)h * D], &t->euc_t[h]);
                t->euc_r[h] = ui(e);
            }
        }
    } else {
        t->cos_r.resize((size_t)H * D);
        for (int h = 0; h < H; h++) draw_cosine_h(e, D, &t->cos_r[(size_t)h * D]);
    }
    int st = upload_proj(c, H, D, t->ldp, metric == CRX_COSINE ? t->cos_r.data() : nullptr,
                         metric == CRX_EUCLIDEAN ? t->euc_v.data() : nullptr, &t->d_proj, &t->d_pnorm);
    if (st != CRX_OK) { crx_lsh_destroy(t); return st; }
    if (metric == CRX_EUCLIDEAN) {
        CRX_TRY(crx_alloc(c, &t->d_t, H));
        CRX_TRY(crx_alloc(c, &t->d_r, H));
        CRX_CUDA(cudaMemcpyAsync(t->d_t, t->euc_t.data(), H * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        CRX_CUDA(cudaMemcpyAsync(t->d_r, t->euc_r.data(), H * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CRX_TRY(crx_alloc(c, &t->hvals, (size_t)L * N * k));
    }
    CRX_TRY(crx_alloc(c, &t->bucket, (size_t)L * N));
    st = crx_hash_rows(c, pts, metric, k, L, t->d_proj, t->ldp, t->d_pnorm, t->d_t, t->d_r, t->w, t->nbuckets, t->hvals, t->bucket);
    if (st != CRX_OK) { crx_lsh_destroy(t); return st; }
    t->by_bucket.resize(L); t->by_group.resize(L); t->ngroups.resize(L);
    if (metric == CRX_EUCLIDEAN) CRX_TRY(crx_alloc(c, &t->gid, (size_t)L * N));
    else t->gid = t->bucket;
    for (int l = 0; l < L; l++) {
        st = crx_build_segments(c, t->bucket + (size_t)l * N, N, t->nbuckets, &t->by_bucket[l]);
        if (st != CRX_OK) { crx_lsh_destroy(t); return st; }
        if (metric == CRX_EUCLIDEAN) {
            st = build_tuple_groups(c, t->hvals + (size_t)l * N * k, N, k, t->gid + (size_t)l * N, &t->by_group[l], &t->ngroups[l]);
            if (st != CRX_OK) { crx_lsh_destroy(t); return st; }
        } else {
            t->by_group[l] = t->by_bucket[l];
            t->ngroups[l] = t->nbuckets;
        }
    }
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    *out = t;
    return CRX_OK;
}

int crx_lsh_bucket_ids(const crx_lsh* t, int32_t* out, int mem) {
    CRX_REQUIRE(t && out, "NULL argument");
    size_t bytes = (size_t)t->L * t->N * sizeof(int32_t);
    CRX_CUDA(cudaMemcpyAsync(out, t->bucket, bytes, mem == CRX_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, t->ctx->stream));
    if (mem == CRX_HOST) CRX_CUDA(cudaStreamSynchronize(t->ctx->stream));
    return CRX_OK;
}

int crx_lsh_detailed_hashes(const crx_lsh* t, int32_t* out, int mem) {
    CRX_REQUIRE(t && out, "NULL argument");
    CRX_REQUIRE(t->hvals, "detailed hashes exist for euclidean tables only (hasDetailedHash, cosine_g_gen.hpp:76)");
    size_t bytes = (size_t)t->L * t->N * t->k * sizeof(int32_t);
    CRX_CUDA(cudaMemcpyAsync(out, t->hvals, bytes, mem == CRX_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, t->ctx->stream));
    if (mem == CRX_HOST) CRX_CUDA(cudaStreamSynchronize(t->ctx->stream));
    return CRX_OK;
}

int crx_lsh_hash_vector(const crx_lsh* t, const double* x, int32_t* bucket_ids, int32_t* detailed) {
    CRX_REQUIRE(t && x && bucket_ids, "NULL argument");
    crx_ctx* c = t->ctx;
    CRX_CUDA(cudaSetDevice(c->device));
    crx_points* q = nullptr;
    CRX_TRY(crx_points_create(c, x, CRX_F64, 1, t->D, CRX_HOST, &q));
    DevBuf<int32_t> b, h;
    int st = b.alloc(c, t->L);
    if (st == CRX_OK && t->metric == CRX_EUCLIDEAN) st = h.alloc(c, (size_t)t->L * t->k);
    if (st == CRX_OK)
        st = crx_hash_rows(c, q, t->metric, t->k, t->L, t->d_proj, t->ldp, t->d_pnorm, t->d_t, t->d_r, t->w, t->nbuckets,
                           t->metric == CRX_EUCLIDEAN ? h.p : nullptr, b.p);
    if (st == CRX_OK) {
        cudaMemcpyAsync(bucket_ids, b.p, t->L * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
        if (detailed && t->metric == CRX_EUCLIDEAN)
            cudaMemcpyAsync(detailed, h.p, (size_t)t->L * t->k * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) st = CRX_ERR_CUDA;
    }
    crx_points_destroy(q);
    return st;
}

int crx_lsh_hash_points(const crx_lsh* t, const crx_points* q, int32_t* bucket_ids, int32_t* detailed, int mem) {
    CRX_REQUIRE(t && q && bucket_ids, "NULL argument");
    CRX_REQUIRE(q->d == t->D, "dimension mismatch");
    crx_ctx* c = t->ctx;
    CRX_CUDA(cudaSetDevice(c->device));
    const bool det = detailed && t->metric == CRX_EUCLIDEAN;
    IoBuf<int32_t> b, h;
    CRX_TRY(b.bind(c, bucket_ids, (size_t)t->L * q->n, mem, false));
    if (det) CRX_TRY(h.bind(c, detailed, (size_t)t->L * q->n * t->k, mem, false));
    CRX_TRY(crx_hash_rows(c, q, t->metric, t->k, t->L, t->d_proj, t->ldp, t->d_pnorm, t->d_t, t->d_r, t->w, t->nbuckets, det ? h.dev : nullptr, b.dev));
    if (det) CRX_TRY(h.flush());
    return b.flush();
}

int crx_lsh_params(const crx_lsh* t, double* cos_r, float* euc_v, float* euc_t, int32_t* euc_r) {
    CRX_REQUIRE(t, "NULL argument");
    if (cos_r && !t->cos_r.empty()) memcpy(cos_r, t->cos_r.data(), t->cos_r.size() * sizeof(double));
    if (euc_v && !t->euc_v.empty()) memcpy(euc_v, t->euc_v.data(), t->euc_v.size() * sizeof(float));
    if (euc_t && !t->euc_t.empty()) memcpy(euc_t, t->euc_t.data(), t->euc_t.size() * sizeof(float));
    if (euc_r && !t->euc_r.empty()) memcpy(euc_r, t->euc_r.data(), t->euc_r.size() * sizeof(int32_t));
    return CRX_OK;
}

int crx_get_LSH_combined_buckets(const crx_lsh* t, int64_t q, int filtered, int32_t* out, int64_t cap, int64_t* count) {
    CRX_REQUIRE(t && count, "NULL argument");
    CRX_REQUIRE(q >= 0 && q < t->N, "query_row out of range");
    crx_ctx* c = t->ctx;
    CRX_CUDA(cudaSetDevice(c->device));
    // one kernel reads the query's bucket (or h-tuple group) of every table and that bucket's extent; one sync; then the L
    // member lists come back in one more sync
    BucketMeta m;
    m.L = t->L;
    for (int l = 0; l < t->L; l++) {
        m.ids[l] = (filtered ? t->gid : t->bucket) + (size_t)l * t->N + q;
        m.off[l] = (filtered ? t->by_group[l] : t->by_bucket[l]).off;
    }
    int32_t* d_ext = nullptr;
    CRX_TRY(crx_alloc(c, &d_ext, 32));
    bucket_meta_kernel<<<1, 32, 0, c->stream>>>(m, d_ext);
    int32_t ext[32];
    cudaError_t e = cudaMemcpyAsync(ext, d_ext, 2 * t->L * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    crx_free(c, d_ext);
    CRX_CUDA(e);
    size_t total = 0;
    for (int l = 0; l < t->L; l++) total += (size_t)(ext[2 * l + 1] - ext[2 * l]);
    std::vector<int32_t> all(total);
    size_t base = 0;
    for (int l = 0; l < t->L; l++) {
        const Segments& s = filtered ? t->by_group[l] : t->by_bucket[l];
        size_t len = (size_t)(ext[2 * l + 1] - ext[2 * l]);
        if (len) CRX_CUDA(cudaMemcpyAsync(all.data() + base, s.perm + ext[2 * l], len * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        base += len;
    }
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    std::sort(all.begin(), all.end());  // std::set<CustVector*> order = row order (lsh_cube.hpp:96,104)
    all.erase(std::unique(all.begin(), all.end()), all.end());
    *count = (int64_t)all.size();
    if (out) for (size_t i = 0; i < all.size() && (int64_t)i < cap; i++) out[i] = all[i];
    return CRX_OK;
}

// ---- hypercube ----
int crx_cube_destroy(crx_cube* cu) {
    if (!cu) return CRX_OK;
    cudaSetDevice(cu->ctx->device);
    crx_free(cu->ctx, cu->vertex);
    cu->by_vertex.free_all();
    delete cu;
    return CRX_OK;
}

int crx_create_hypercube(crx_ctx* c, const crx_points* pts, int metric, int k, double euclidean_h_w, uint64_t seed,
                         crx_cube** out) {
    CRX_REQUIRE(pts, "NULL argument");
    CRX_NARROW(pts);
    CRX_REQUIRE(c && pts && out, "NULL argument");
    CRX_REQUIRE(metric == CRX_EUCLIDEAN || metric == CRX_COSINE, "metric");
    CRX_REQUIRE(k >= 1 && k <= 16, "cube dimension k must be in [1,16]");
    CRX_CUDA(cudaSetDevice(c->device));
    int64_t N = pts->n;
    int D = pts->d;
    CRX_REQUIRE((uint64_t)N * (uint64_t)k < (1ull << 32), "N*k must be below 2^32");
    crx_cube* cu = new crx_cube();
    cu->ctx = c; cu->pts = pts; cu->metric = metric; cu->k = k; cu->D = D; cu->N = N; cu->w = (float)euclidean_h_w;
    cu->engine.seed((unsigned long)seed);  // lsh_cube.hpp:112-114
    if (metric == CRX_EUCLIDEAN) {
        cu->euc_v.resize((size_t)k * D); cu->euc_t.resize(k);
        for (int j = 0; j < k; j++) draw_euclid_h(cu->engine, D, cu->w, &cu->euc_v[(size_t)j * D], &cu->euc_t[j]);  // euclidean_f_gen.hpp:54-57
    } else {
        cu->cos_r.resize((size_t)k * D);
        for (int j = 0; j < k; j++) draw_cosine_h(cu->engine, D, &cu->cos_r[(size_t)j * D]);
    }
    double* d_proj = nullptr; double* d_pnorm = nullptr; float* d_t = nullptr;
    int st = upload_proj(c, k, D, pts->ld, metric == CRX_COSINE ? cu->cos_r.data() : nullptr,
                         metric == CRX_EUCLIDEAN ? cu->euc_v.data() : nullptr, &d_proj, &d_pnorm);
    if (st != CRX_OK) { crx_cube_destroy(cu); return st; }
    CRX_TRY(crx_alloc(c, &cu->vertex, N));
    int nvert = 1 << k;
    if (metric == CRX_COSINE) {
        st = crx_hash_rows(c, pts, metric, k, 1, d_proj, pts->ld, d_pnorm, nullptr, nullptr, 0.f, nvert, nullptr, cu->vertex);
    } else {
        CRX_TRY(crx_alloc(c, &d_t, k));
        CRX_CUDA(cudaMemcpyAsync(d_t, cu->euc_t.data(), k * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        DevBuf<int32_t> hv;
        CRX_TRY(hv.alloc(c, (size_t)N * k));
        st = crx_hash_rows(c, pts, metric, k, 1, d_proj, pts->ld, d_pnorm, d_t, nullptr, cu->w, nvert, hv.p, nullptr);
        if (st == CRX_OK) {
            // first-occurrence order of the distinct (f, h) pairs in (row-major, f-minor) order:
            // that is the order in which EuclideanFGen draws its 1-or-2 (euclidean_f_gen.hpp:65-79)
            int64_t total = N * k;
            std::vector<unsigned long long> hk;
            std::vector<uint32_t> hp;
            int nheads = 0;
            // ---- dense tables when the h ranges are small
            bool dense = false;
            {
                DevBuf<int> dmn, dmx;
                CRX_TRY(dmn.alloc(c, 16)); CRX_TRY(dmx.alloc(c, 16));
                std::vector<int> hmn(16, INT_MAX), hmx(16, INT_MIN);
                CRX_CUDA(cudaMemcpyAsync(dmn.p, hmn.data(), 16 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
                CRX_CUDA(cudaMemcpyAsync(dmx.p, hmx.data(), 16 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
                { CRX_KERNEL(c, "cube_minmax"); cube_minmax_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(hv.p, total, k, dmn.p, dmx.p); }
                CRX_CUDA(cudaMemcpyAsync(hmn.data(), dmn.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                CRX_CUDA(cudaMemcpyAsync(hmx.data(), dmx.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                CRX_CUDA(cudaStreamSynchronize(c->stream));
                std::vector<int64_t> toff(k + 1, 0);
                for (int f = 0; f < k; f++) toff[f + 1] = toff[f] + ((int64_t)hmx[f] - (int64_t)hmn[f] + 1);
                if (toff[k] <= (1ll << 22)) {
                    dense = true;
                    DevBuf<int64_t> dtoff;
                    DevBuf<uint32_t> table;
                    CRX_TRY(dtoff.alloc(c, k + 1)); CRX_TRY(table.alloc(c, toff[k]));
                    CRX_CUDA(cudaMemcpyAsync(dtoff.p, toff.data(), (k + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
                    CRX_CUDA(cudaMemsetAsync(table.p, 0xff, toff[k] * sizeof(uint32_t), c->stream));
                    { CRX_KERNEL(c, "cube_first"); cube_first_kernel<<<crx_grid(total, 256), 256, 0, c->stream>>>(hv.p, total, k, dmn.p, dtoff.p, table.p); }
                    std::vector<uint32_t> ht(toff[k]);
                    CRX_CUDA(cudaMemcpyAsync(ht.data(), table.p, toff[k] * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
                    CRX_CUDA(cudaStreamSynchronize(c->stream));
                    for (int f = 0; f < k; f++)
                        for (int64_t j = toff[f]; j < toff[f + 1]; j++)
                            if (ht[j] != 0xffffffffu) {
                                int32_t h = (int32_t)((int64_t)hmn[f] + (j - toff[f]));
                                hk.push_back(((unsigned long long)f << 32) | (unsigned long long)((uint32_t)h ^ 0x80000000u));
                                hp.push_back(ht[j]);
                            }
                    nheads = (int)hk.size();
                }
            }
            if (!dense) {
            DevBuf<unsigned long long> keys, keys2, hkeys;
            DevBuf<uint32_t> pos, pos2, hpos;
            DevBuf<int> cnt;
            CRX_TRY(keys.alloc(c, total)); CRX_TRY(keys2.alloc(c, total));
            CRX_TRY(pos.alloc(c, total)); CRX_TRY(pos2.alloc(c, total));
            CRX_TRY(cnt.alloc(c, 1));
            { CRX_KERNEL(c, "cube_keys"); cube_keys_kernel<<<crx_grid(total, 256), 256, 0, c->stream>>>(hv.p, total, k, keys.p, pos.p); }
            size_t bytes = 0;
            CRX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys2.p, pos.p, pos2.p, (int)total, 0, 37, c->stream));
            DevBuf<char> tmp;
            CRX_TRY(tmp.alloc(c, bytes));
            CRX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, keys2.p, pos.p, pos2.p, (int)total, 0, 37, c->stream));
            int cap = 1 << 20;
            for (int attempt = 0; attempt < 2; attempt++) {
                CRX_TRY(hkeys.alloc(c, cap)); CRX_TRY(hpos.alloc(c, cap));
                CRX_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int), c->stream));
                { CRX_KERNEL(c, "cube_heads"); cube_heads_kernel<<<crx_grid(total, 256), 256, 0, c->stream>>>(keys2.p, pos2.p, total, hkeys.p, hpos.p, cap, cnt.p); }
                CRX_CUDA(cudaMemcpyAsync(&nheads, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                CRX_CUDA(cudaStreamSynchronize(c->stream));
                if (nheads <= cap) break;
                cap = nheads;
            }
            hk.resize(nheads); hp.resize(nheads);
            CRX_CUDA(cudaMemcpyAsync(hk.data(), hkeys.p, nheads * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
            CRX_CUDA(cudaMemcpyAsync(hp.data(), hpos.p, nheads * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            CRX_CUDA(cudaStreamSynchronize(c->stream));
            }  // sort path
            std::vector<int> order(nheads);
            for (int i = 0; i < nheads; i++) order[i] = i;
            std::sort(order.begin(), order.end(), [&](int a, int b) { return hp[a] < hp[b]; });
            cu->fmap.assign(k, std::vector<std::pair<int32_t, int32_t>>());
            for (int oi = 0; oi < nheads; oi++) {
                unsigned long long key = hk[order[oi]];
                int f = (int)(key >> 32);
                int32_t h = (int32_t)((uint32_t)(key & 0xffffffffull) ^ 0x80000000u);
                std::uniform_int_distribution<int> u12(1, 2);  // fresh object per draw (euclidean_f_gen.hpp:72)
                int bit = mod_ii(h, u12(cu->engine));
                cu->fmap[f].emplace_back(h, bit);
            }
            std::vector<int32_t> mh, mb, mo(k + 1, 0);
            for (int f = 0; f < k; f++) {
                std::sort(cu->fmap[f].begin(), cu->fmap[f].end());
                for (auto& pr : cu->fmap[f]) { mh.push_back(pr.first); mb.push_back(pr.second); }
                mo[f + 1] = (int32_t)mh.size();
            }
            DevBuf<int32_t> dmh, dmb, dmo;
            CRX_TRY(dmh.alloc(c, mh.size())); CRX_TRY(dmb.alloc(c, mb.size())); CRX_TRY(dmo.alloc(c, mo.size()));
            CRX_CUDA(cudaMemcpyAsync(dmh.p, mh.data(), mh.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            CRX_CUDA(cudaMemcpyAsync(dmb.p, mb.data(), mb.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            CRX_CUDA(cudaMemcpyAsync(dmo.p, mo.data(), mo.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            { CRX_KERNEL(c, "cube_vertex"); cube_vertex_kernel<<<crx_grid(N, 256), 256, 0, c->stream>>>(hv.p, N, k, dmh.p, dmb.p, dmo.p, cu->vertex); }
            CRX_CUDA(cudaGetLastError());
            CRX_CUDA(cudaStreamSynchronize(c->stream));
        }
    }
    crx_free(c, d_proj); crx_free(c, d_pnorm); crx_free(c, d_t);
    if (st != CRX_OK) { crx_cube_destroy(cu); return st; }
    st = crx_build_segments(c, cu->vertex, N, nvert, &cu->by_vertex);
    if (st != CRX_OK) { crx_cube_destroy(cu); return st; }
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    *out = cu;
    return CRX_OK;
}

int crx_cube_vertex_ids(const crx_cube* cu, int32_t* out, int mem) {
    CRX_REQUIRE(cu && out, "NULL argument");
    CRX_CUDA(cudaMemcpyAsync(out, cu->vertex, cu->N * sizeof(int32_t), mem == CRX_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, cu->ctx->stream));
    if (mem == CRX_HOST) CRX_CUDA(cudaStreamSynchronize(cu->ctx->stream));
    return CRX_OK;
}

int crx_get_hypercube_combined_buckets(const crx_cube* cu, int64_t q, int probes, int32_t* out, int64_t cap, int64_t* count) {
    CRX_REQUIRE(cu && count, "NULL argument");
    CRX_REQUIRE(q >= 0 && q < cu->N, "query_row out of range");
    crx_ctx* c = cu->ctx;
    CRX_CUDA(cudaSetDevice(c->device));
    int32_t home;
    CRX_CUDA(cudaMemcpyAsync(&home, cu->vertex + q, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<int> seq;
    crx_cube_probe_sequence(home, probes, cu->k, seq);
    std::vector<int32_t> off((size_t)(1 << cu->k) + 1);
    CRX_CUDA(cudaMemcpyAsync(off.data(), cu->by_vertex.off, off.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    int64_t n = 0;
    for (int v : seq) {
        int64_t len = off[v + 1] - off[v];
        if (out && n < cap && len > 0) {
            int64_t take = std::min<int64_t>(len, cap - n);
            CRX_CUDA(cudaMemcpyAsync(out + n, cu->by_vertex.perm + off[v], take * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        }
        n += len;
    }
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    *count = n;
    return CRX_OK;
}

} // extern "C"
