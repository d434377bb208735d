// cluster.cu -- lib/clustering_phases of the reference on the GPU:
//   K5 Lloyd assignment (tiled FP64 pair scan + fused warp-shuffle argmin), cluster sums / k-means
//   update, K6 k-means++ rounds, K7 LSH / hypercube range-search assignment, K8 PAM medoid update,
//   K11 silhouette.
#include <algorithm>
#include <climits>
#include <sstream>

#include <cub/cub.cuh>
#include <random>

#include "pair_tile.cuh"
#include "rowwalk.cuh"
#include "tables.cuh"
#include "tc_scan.cuh"

void crx_cube_probe_sequence(int home, int probes, int k, std::vector<int>& seq);  // hash.cu

// ------------------------------------------------------------------------------------------------
// centroid staging: [K][D] doubles -> padded [K][ld] + exact sums of squares
// ------------------------------------------------------------------------------------------------
__global__ void pad_centroids_kernel(const double* __restrict__ C, int K, int D, int ld, double* __restrict__ out,
                                     double* __restrict__ csqn) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    double acc = 0.0;
    for (int i = 0; i < ld; i++) {
        double v = i < D ? C[(size_t)c * D + i] : 0.0;
        out[(size_t)c * ld + i] = v;
        if (i < D) acc = __dadd_rn(acc, __dmul_rn(v, v));
    }
    csqn[c] = acc;
}

struct Centroids {
    DevBuf<double> pad, sqn;
    IoBuf<double> in;
    int K = 0, ld = 0;
    int stage(crx_ctx* c, const double* C, int cmem, int K_, int D, int ld_) {
        K = K_; ld = ld_;
        CRX_TRY(in.bind(c, C, (size_t)K * D, cmem, true));
        CRX_TRY(pad.alloc(c, (size_t)K * ld));
        CRX_TRY(sqn.alloc(c, K));
        CRX_KERNEL(c, "pad_centroids");
        pad_centroids_kernel<<<crx_grid(K, 128), 128, 0, c->stream>>>(in.dev, K, D, ld, pad.p, sqn.p);
        CRX_CUDA(cudaGetLastError());
        return CRX_OK;
    }
};

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ x, int ld, int D, const int32_t* __restrict__ rows, int K,
                                   double* __restrict__ out /* [K][D] */) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K * D) return;
    int c = e / D, i = e - c * D;
    out[e] = (double)x[(size_t)rows[c] * ld + i];
}

// rows of `pts` as a [K][D] double matrix on the device
static int gather_rows(crx_ctx* c, const crx_points* p, const int32_t* d_rows, int K, double* d_out) {
    CRX_KERNEL(c, "gather_rows");
    int g = crx_grid((int64_t)K * p->d, 256);
    if (p->x64) gather_rows_kernel<double><<<g, 256, 0, c->stream>>>(p->x64, p->ld, p->d, d_rows, K, d_out);
    else gather_rows_kernel<float><<<g, 256, 0, c->stream>>>(p->x32, p->ld, p->d, d_rows, K, d_out);
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

// ------------------------------------------------------------------------------------------------
// K5: Lloyd assignment.  One CTA = 64 points, all K centroids streamed through shared memory in
// tiles of 64; each warp owns 8 points, each lane 2 centroids per tile; the argmin over centroids
// is lane-local across tiles and finishes with one warp-shuffle reduction per point.
//   Euclidean: FORM_DIFF_EXACT => the sums are the reference's own values, so labels and distances
//              are bit-exact including ties (lowest index wins, assignment.hpp:67).
//   cosine:    FORM_DOT (plain FP64 FMA chain) for the scan; it differs from the reference's x87 value by up to
//              (D + 4) ulp, so the scan also keeps the second-best value: a row whose margin is inside that bound (equal or
//              scalar-multiple centroids, rating-like ties) is decided again with the reference's own arithmetic over all
//              centroids in index order, strict '<' (assignment.hpp:63-70).  The winner's distance is always the x87 one.
// ------------------------------------------------------------------------------------------------
template <typename T, int METRIC>
__global__ void __launch_bounds__(pt::NT)
lloyd_scan_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, const int32_t* __restrict__ rowmap,
                  int64_t nrows, const double* __restrict__ cent, const double* __restrict__ csqn, int K,
                  int32_t* __restrict__ labels, double* __restrict__ dists) {
    extern __shared__ double sm[];
    double* As = sm;
    double* Bs = sm + pt::BM * ld;
    double* cn = Bs + pt::BN * ld;  // [BN] sqrt of centroid norms (cosine)
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t r0 = (int64_t)blockIdx.x * pt::BM;
    pt::load_a_tile<T>(As, x, ld, rowmap, r0, nrows);

    double best_v[pt::RW], second_v[pt::RW];
    int best_i[pt::RW];
    double xn[pt::RW];
    unsigned nan0 = 0;
#pragma unroll
    for (int r = 0; r < pt::RW; r++) {
        best_v[r] = INFINITY; best_i[r] = INT_MAX; second_v[r] = INFINITY;
        int64_t pos = r0 + warp * pt::RW + r;
        xn[r] = 0.0;
        if (METRIC == CRX_COSINE && pos < nrows) xn[r] = __dsqrt_rn(sqn[rowmap ? rowmap[pos] : pos]);
    }
    for (int c0 = 0; c0 < K; c0 += pt::BN) {
        __syncthreads();
        pt::load_b_tile<double>(Bs, cent, ld, nullptr, c0, K);
        if (METRIC == CRX_COSINE && threadIdx.x < pt::BN) {
            int c = c0 + threadIdx.x;
            cn[threadIdx.x] = c < K ? __dsqrt_rn(csqn[c]) : 1.0;
        }
        __syncthreads();
        double acc[pt::RW][2];
        pt::tile_mac<METRIC == CRX_EUCLIDEAN ? pt::FORM_DIFF_EXACT : pt::FORM_DOT>(As, Bs, ld, warp, lane, acc);
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            int col = c0 + 2 * lane + cc;
            if (col < K) {
#pragma unroll
                for (int r = 0; r < pt::RW; r++) {
                    if (METRIC == CRX_EUCLIDEAN) {
                        if (pt::euclid_better(acc[r][cc], col, best_v[r], best_i[r])) { best_v[r] = acc[r][cc]; best_i[r] = col; }
                    } else {
                        double dist = __dsub_rn(1.0, __ddiv_rn(acc[r][cc], __dmul_rn(xn[r], cn[2 * lane + cc])));
                        if (col == 0 && dist != dist) nan0 |= 1u << r;  // `min == -1` takes centroid 0 even when NaN
                        if (pt::plain_better(dist, col, best_v[r], best_i[r])) { second_v[r] = fmin(second_v[r], best_v[r]); best_v[r] = dist; best_i[r] = col; }
                        else if (dist == dist) second_v[r] = fmin(second_v[r], dist);
                    }
                }
            }
        }
    }
    // warp reduction per row
#pragma unroll
    for (int r = 0; r < pt::RW; r++) {
        double v = best_v[r];
        int i = best_i[r];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, v, off);
            int oi = __shfl_xor_sync(0xffffffffu, i, off);
            bool take = METRIC == CRX_EUCLIDEAN ? pt::euclid_better(ov, oi, v, i) : pt::plain_better(ov, oi, v, i);
            if (METRIC == CRX_COSINE) {
                double os = __shfl_xor_sync(0xffffffffu, second_v[r], off);
                second_v[r] = fmin(fmin(second_v[r], os), take ? v : ov);   // the loser's best is a runner-up
            }
            if (take) { v = ov; i = oi; }
        }
        best_v[r] = v; best_i[r] = i;
    }
    unsigned nan0_all = __shfl_sync(0xffffffffu, nan0, 0);  // column 0 lives in lane 0
    // lane r finishes row r
    if (lane < pt::RW) {
        int r = lane;
        double v = 0.0, sec = INFINITY; int i = 0;
#pragma unroll
        for (int rr = 0; rr < pt::RW; rr++) if (rr == r) { v = best_v[rr]; i = best_i[rr]; sec = second_v[rr]; }
        int64_t pos = r0 + warp * pt::RW + r;
        if (pos < nrows) {
            int64_t row = rowmap ? (int64_t)rowmap[pos] : pos;
            double dist;
            if (METRIC == CRX_EUCLIDEAN) dist = __dsqrt_rn(v);
            else {
                if ((nan0_all >> r) & 1u) { i = 0; dist = v = nan(""); }
                else {
                    const double* a = As + (warp * pt::RW + r) * ld;
                    if (sec - v <= (double)(D + 4) * 2.3e-16) {
                        // near tie under the plain arithmetic: the reference's own scan (assignment.hpp:63-70)
                        i = 0;
                        dist = __dsub_rn(1.0, cos_sim_exact(a, cent, D, sqn[row], csqn[0]));
                        for (int cc = 1; cc < K; cc++) {
                            double dc = __dsub_rn(1.0, cos_sim_exact(a, cent + (size_t)cc * ld, D, sqn[row], csqn[cc]));
                            if (dc < dist) { dist = dc; i = cc; }
                        }
                    } else {
                        dist = __dsub_rn(1.0, cos_sim_exact(a, cent + (size_t)i * ld, D, sqn[row], csqn[i]));
                    }
                }
            }
            labels[row] = i;
            dists[row] = dist;
        }
    }
}

__global__ void self_assign_kernel(const int32_t* __restrict__ crow, int K, int32_t* labels, double* dists) {
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int c = 0; c < K; c++)  // sequential: a later centroid aliasing the same row wins (assignment.hpp:77-78)
            if (crow[c] >= 0) { labels[crow[c]] = c; if (dists) dists[crow[c]] = 0.0; }
}

__global__ void compact_unassigned_kernel(const int32_t* __restrict__ labels, int64_t n, int32_t* __restrict__ rows, int* count,
                                          int64_t row_begin = 0) {
    int64_t i = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && labels[i] == -1) rows[atomicAdd(count, 1)] = (int32_t)i;
}

// D > 128 (e.g. the 203-dimensional tweet vectors clustered in front of the recommendation, main.cpp:78-110): the tile
// engine keeps whole rows of both tiles in shared memory and does not fit.  Plain exact form instead: a warp owns 32
// rows (one per lane), the centroids pass by one at a time through shared memory, every lane walks its row in index
// order (rowwalk.cuh) and keeps the running `min == -1 || d < min` of assignment.hpp:62-71.
constexpr int CRX_MAXD = 512;   // crx_points_create accepts d <= CRX_MAXD; everything but the clustering core stays at d <= 128
template <typename T, int METRIC>
__global__ void __launch_bounds__(128)
lloyd_scan_wide_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, const int32_t* __restrict__ rowmap,
                       int64_t nrows, const double* __restrict__ cent, const double* __restrict__ csqn, int K,
                       int32_t* __restrict__ labels, double* __restrict__ dists) {
    __shared__ rw::WarpTile tiles[4];
    __shared__ double vec[CRX_MAXD];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t pos = ((int64_t)blockIdx.x * 4 + warp) * 32 + lane;
    bool valid = pos < nrows;
    int64_t row = valid ? (rowmap ? (int64_t)rowmap[pos] : pos) : -1;
    double nrow = valid ? sqn[row] : 1.0;
    double best = 0.0;
    int arg = 0;
    for (int cc = 0; cc < K; cc++) {
        __syncthreads();
        for (int k = threadIdx.x; k < ld; k += blockDim.x) vec[k] = cent[(size_t)cc * ld + k];
        __syncthreads();
        double d = rw::dist_rows<T, METRIC>(x, ld, D, row, vec, nrow, csqn[cc], tiles[warp]);
        if (cc == 0 || d < best) { best = d; arg = cc; }   // a NaN at centroid 0 stays (the reference's -1 sentinel)
    }
    if (valid) { labels[row] = arg; dists[row] = best; }
}

static int lloyd_scan(crx_ctx* c, const crx_points* p, const int32_t* d_rowmap, int64_t nrows, const Centroids& cen,
                      int metric, int32_t* d_labels, double* d_dists) {
    if (nrows == 0) return CRX_OK;
    int ld = p->ld;
    if (p->d > 128) {
        CRX_KERNEL(c, "lloyd_scan_wide");
        int g = (int)((nrows + 127) / 128);
#define LAUNCH_W(T, M, xptr) lloyd_scan_wide_kernel<T, M><<<g, 128, 0, c->stream>>>(xptr, ld, p->d, p->sqn, d_rowmap, nrows, cen.pad.p, cen.sqn.p, cen.K, d_labels, d_dists)
        if (p->x64) { if (metric == CRX_EUCLIDEAN) LAUNCH_W(double, CRX_EUCLIDEAN, p->x64); else LAUNCH_W(double, CRX_COSINE, p->x64); }
        else { if (metric == CRX_EUCLIDEAN) LAUNCH_W(float, CRX_EUCLIDEAN, p->x32); else LAUNCH_W(float, CRX_COSINE, p->x32); }
#undef LAUNCH_W
        CRX_CUDA(cudaGetLastError());
        return CRX_OK;
    }
    size_t smem = pt::smem_bytes(ld) + pt::BN * sizeof(double);
    int grid = (int)((nrows + pt::BM - 1) / pt::BM);
    CRX_KERNEL(c, "lloyd_scan");
#define LAUNCH_L(T, M, xptr)                                                                                       \
    do {                                                                                                           \
        CRX_CUDA(cudaFuncSetAttribute(lloyd_scan_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        lloyd_scan_kernel<T, M><<<grid, pt::NT, smem, c->stream>>>(xptr, ld, p->d, p->sqn, d_rowmap, nrows, cen.pad.p,   \
                                                                   cen.sqn.p, cen.K, d_labels, d_dists);          \
    } while (0)
    if (p->x64) { if (metric == CRX_EUCLIDEAN) LAUNCH_L(double, CRX_EUCLIDEAN, p->x64); else LAUNCH_L(double, CRX_COSINE, p->x64); }
    else { if (metric == CRX_EUCLIDEAN) LAUNCH_L(float, CRX_EUCLIDEAN, p->x32); else LAUNCH_L(float, CRX_COSINE, p->x32); }
#undef LAUNCH_L
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

// ------------------------------------------------------------------------------------------------
// K5 on the tensor cores (Euclidean, K >= 32): tc_scan's argmin filter gives, per point, the best and the
// second-best value of  0.5 ||c||^2 - x.c  from split-fp16 products.  A point whose margin exceeds twice
// the filter's error bound keeps the filter's label (its distance is then computed exactly, as the
// reference does); the others -- near-ties and exact ties -- are re-scanned by the exact FP64 kernel, so
// labels and distances stay bit-identical to the reference.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxabs_kernel(const T* __restrict__ x, size_t n, unsigned int* __restrict__ out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf((float)x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));  // non-negative floats order like their bit patterns
}

__global__ void half_norm_kernel(const double* __restrict__ csqn, int K, double scale, float* __restrict__ hn, unsigned int* __restrict__ cmax_bits) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    hn[c] = (float)(0.5 * csqn[c] * scale);
    atomicMax(cmax_bits, __float_as_uint((float)sqrt(csqn[c]) * 1.000001f));
}

// One warp = 32 consecutive points.  The point rows are staged 32 x 16 through shared memory (coalesced
// 64-byte reads), each lane then walks ITS row in index order against the row of ITS winning centroid
// (128 contiguous bytes per 16 coordinates, L2 resident): the reference's own sequence of operations.
template <typename T>
__global__ void __launch_bounds__(128)
lloyd_refine_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, int64_t row_begin, int64_t N /* end row */,
                    const double* __restrict__ cent,
                    const float* __restrict__ best, const float* __restrict__ second, const int32_t* __restrict__ bidx, int K,
                    double scale, const unsigned int* __restrict__ cmax_bits, int32_t* __restrict__ labels, double* __restrict__ dists,
                    int32_t* __restrict__ amb_rows, int* __restrict__ amb_count) {
    __shared__ rw::WarpTile tiles[4];    // 32 rows x 16 coordinates of x
    __shared__ rw::WarpTile ctiles[4];   // the same 16 coordinates of each row's winning centroid
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t row0 = row_begin + ((int64_t)blockIdx.x * 4 + warp) * 32;
    if (row0 >= N) return;
    int64_t i = row0 + lane;
    bool valid = i < N;
    double cmax = (double)__uint_as_float(*cmax_bits);
    double xn = valid ? sqrt(sqn[i]) : 0.0;
    // |filter value - true value| <= E (scaled units): split + dropped lo*lo (3 * 2^-22), fp32 accumulation over
    // 21 K-slices (21 * 2^-23, truncation assumed), fp32 rounding of the half norm and of the subtraction
    double E = scale * (6e-6 * xn * cmax + 2.5e-7 * (0.5 * cmax * cmax + xn * cmax));
    int b = valid ? bidx[i] : 0;
    double margin = valid ? (double)second[i] - (double)best[i] : 0.0;
    bool sure = valid && b >= 0 && b < K && margin > 2.0 * E;
    const int bsel = sure ? b : 0;
    double acc = 0.0;
    if constexpr (sizeof(T) == 4) {
        // software pipeline: the global loads of chunk c + 1 (point pieces from HBM, centroid pieces from L2) are in flight while
        // chunk c is walked out of shared memory
        float4 xv[4];
        double2 cv[8];
        int brow[8];
#pragma unroll
        for (int it = 0; it < 8; it++) brow[it] = __shfl_sync(0xffffffffu, bsel, it * 4 + (lane >> 3));
        auto fetch = [&](int c0) {
#pragma unroll
            for (int it = 0; it < 4; it++) {   // 16-byte pieces: 8 rows x 64 bytes per load instruction
                const int64_t src = row0 + it * 8 + (lane >> 2);
                const int piece = (lane & 3) * 4;
                xv[it] = (src < N && c0 + piece < ld) ? *reinterpret_cast<const float4*>(x + src * ld + c0 + piece) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < 8; it++) {   // 8 lanes read the 128 contiguous bytes of one row's centroid (ld is a multiple of 4, padding = 0)
                const int piece = (lane & 7) * 2;
                cv[it] = (c0 + piece < ld) ? *reinterpret_cast<const double2*>(cent + (size_t)brow[it] * ld + c0 + piece) : make_double2(0.0, 0.0);
            }
        };
        fetch(0);
        for (int c0 = 0; c0 < D; c0 += 16) {
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const int r = it * 8 + (lane >> 2), piece = (lane & 3) * 4;
                tiles[warp][r][piece] = (double)xv[it].x; tiles[warp][r][piece + 1] = (double)xv[it].y;
                tiles[warp][r][piece + 2] = (double)xv[it].z; tiles[warp][r][piece + 3] = (double)xv[it].w;
            }
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int r = it * 4 + (lane >> 3), piece = (lane & 7) * 2;
                ctiles[warp][r][piece] = cv[it].x;
                ctiles[warp][r][piece + 1] = cv[it].y;
            }
            __syncwarp();
            if (c0 + 16 < D) fetch(c0 + 16);
            const int lim = min(16, D - c0);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k < lim) {
                    double t = __dsub_rn(tiles[warp][lane][k], ctiles[warp][lane][k]);
                    acc = __dadd_rn(acc, __dmul_rn(t, t));
                }
            }
            __syncwarp();
        }
    } else {
    for (int c0 = 0; c0 < D; c0 += 16) {
#pragma unroll 4
        for (int it = 0; it < 16; it++) {
            int r = it * 2 + (lane >> 4);
            int64_t src = row0 + r;
            int col = c0 + (lane & 15);
            tiles[warp][r][lane & 15] = (src < N && col < D) ? (double)x[src * ld + col] : 0.0;
        }
#pragma unroll
        for (int it = 0; it < 8; it++) {  // ld is a multiple of 4 and the padding holds zeros
            int r = it * 4 + (lane >> 3);
            int piece = (lane & 7) * 2;
            int br = __shfl_sync(0xffffffffu, bsel, r);
            double2 v = (c0 + piece < ld) ? *reinterpret_cast<const double2*>(cent + (size_t)br * ld + c0 + piece) : make_double2(0.0, 0.0);
            ctiles[warp][r][piece] = v.x;
            ctiles[warp][r][piece + 1] = v.y;
        }
        __syncwarp();
        int lim = min(16, D - c0);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k < lim) {
                double t = __dsub_rn(tiles[warp][lane][k], ctiles[warp][lane][k]);
                acc = __dadd_rn(acc, __dmul_rn(t, t));
            }
        }
        __syncwarp();
    }
    }
    if (sure) {
        labels[i] = b;
        dists[i] = __dsqrt_rn(acc);
    } else if (valid) {
        amb_rows[atomicAdd(amb_count, 1)] = (int32_t)i;
    }
}

// labels only (the caller passed no distance buffer): the same certification as lloyd_refine_kernel without the
// exact distance of the winner -- 20 bytes per point instead of the whole row
__global__ void lloyd_label_kernel(const double* __restrict__ sqn, int64_t row_begin, int64_t row_end, const float* __restrict__ best,
                                   const float* __restrict__ second, const int32_t* __restrict__ bidx, int K, double scale,
                                   const unsigned int* __restrict__ cmax_bits, int32_t* __restrict__ labels,
                                   int32_t* __restrict__ amb_rows, int* __restrict__ amb_count) {
    int64_t i = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_end) return;
    double cmax = (double)__uint_as_float(*cmax_bits);
    double xn = sqrt(sqn[i]);
    double E = scale * (6e-6 * xn * cmax + 2.5e-7 * (0.5 * cmax * cmax + xn * cmax));
    int b = bidx[i];
    double margin = (double)second[i] - (double)best[i];
    const bool sure = b >= 0 && b < K && margin > 2.0 * E;
    if (sure) labels[i] = b;
    else amb_rows[atomicAdd(amb_count, 1)] = (int32_t)i;
}

static int points_maxabs(crx_ctx* c, const crx_points* p, double* out) {
    if (p->maxabs < 0) {
        DevBuf<unsigned int> m;
        CRX_TRY(m.alloc(c, 1));
        CRX_CUDA(cudaMemsetAsync(m.p, 0, sizeof(unsigned int), c->stream));
        size_t n = (size_t)p->n * p->ld;
        int g = (int)std::min<size_t>((size_t)c->sm_count * 16, (n + 255) / 256);
        { CRX_KERNEL(c, "maxabs"); maxabs_kernel<float><<<g, 256, 0, c->stream>>>(p->x32, n, m.p); }
        unsigned int bits = 0;
        CRX_CUDA(cudaMemcpyAsync(&bits, m.p, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        float f;
        memcpy(&f, &bits, 4);
        p->maxabs = (double)f;
    }
    *out = p->maxabs;
    return CRX_OK;
}

static int scale_for(double maxabs) {  // 2^s * maxabs in [4096, 8192): far from fp16 overflow, lo parts stay normal
    if (!(maxabs > 0.0) || !std::isfinite(maxabs)) return 0;
    int e;
    frexp(maxabs, &e);  // maxabs = m * 2^e, m in [0.5, 1)
    return 13 - e;
}

static bool tc_disabled() {
    static const bool off = getenv("CRX_NO_TC") != nullptr && getenv("CRX_NO_TC")[0] == '1';
    return off;
}

// rows [r0, r1) only (r1 < 0: all rows); labels / dists are indexed by absolute row
static int lloyd_scan_tc(crx_ctx* c, const crx_points* p, const Centroids& cen, int32_t* d_labels, double* d_dists, int64_t r0 = 0,
                         int64_t r1 = -1) {
    int64_t N = p->n;
    if (r1 < 0) r1 = N;
    if (r1 <= r0) return CRX_OK;
    int K = cen.K, D = p->d, ld = p->ld;
    double mx = 0;
    CRX_TRY(points_maxabs(c, p, &mx));
    if (!p->tc_l2) {
        TcOperand* op = new TcOperand();
        int sx = scale_for(mx);
        int st = crx_tc_prepare(c, p, 1, (double)sx, op);
        if (st != CRX_OK) { delete op; return st; }
        p->tc_l2 = op;
        p->tc_l2_scale = sx;
    }
    // centroids: scale from their own magnitude
    DevBuf<unsigned int> cm, cmaxn;
    CRX_TRY(cm.alloc(c, 1)); CRX_TRY(cmaxn.alloc(c, 1));
    CRX_CUDA(cudaMemsetAsync(cm.p, 0, sizeof(unsigned int), c->stream));
    CRX_CUDA(cudaMemsetAsync(cmaxn.p, 0, sizeof(unsigned int), c->stream));
    { CRX_KERNEL(c, "maxabs"); maxabs_kernel<double><<<std::min(64, crx_grid((int64_t)K * ld, 256)), 256, 0, c->stream>>>(cen.pad.p, (size_t)K * ld, cm.p); }
    unsigned int cbits = 0;
    CRX_CUDA(cudaMemcpyAsync(&cbits, cm.p, sizeof(cbits), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    float cf;
    memcpy(&cf, &cbits, 4);
    int sc = scale_for((double)cf), sx = p->tc_l2_scale;
    double scale = ldexp(1.0, sx + sc);
    TcOperand opC;
    int st = crx_tc_prepare_matrix(c, cen.pad.p, K, D, ld, (double)sc, &opC);
    if (st != CRX_OK) return st;
    DevBuf<float> hn, best, second;
    DevBuf<int32_t> bidx, amb;
    DevBuf<int> amb_count;
    CRX_TRY(hn.alloc(c, K)); CRX_TRY(best.alloc(c, N)); CRX_TRY(second.alloc(c, N)); CRX_TRY(bidx.alloc(c, N));
    CRX_TRY(amb.alloc(c, N)); CRX_TRY(amb_count.alloc(c, 1));
    CRX_CUDA(cudaMemsetAsync(amb_count.p, 0, sizeof(int), c->stream));
    { CRX_KERNEL(c, "half_norm"); half_norm_kernel<<<crx_grid(K, 128), 128, 0, c->stream>>>(cen.sqn.p, K, scale, hn.p, cmaxn.p); }
    st = crx_tc_argmin(c, *p->tc_l2, r0, r1 - r0, opC, hn.p, best.p + r0, second.p + r0, bidx.p + r0);
    if (st == CRX_OK && !d_dists) {
        CRX_KERNEL(c, "lloyd_label");
        lloyd_label_kernel<<<crx_grid(r1 - r0, 256), 256, 0, c->stream>>>(p->sqn, r0, r1, best.p, second.p, bidx.p, K, scale, cmaxn.p, d_labels, amb.p, amb_count.p);
        CRX_CUDA(cudaGetLastError());
    } else if (st == CRX_OK) {
        CRX_KERNEL(c, "lloyd_refine");
        int g = (int)((r1 - r0 + 127) / 128);
        if (p->x64) lloyd_refine_kernel<double><<<g, 128, 0, c->stream>>>(p->x64, ld, D, p->sqn, r0, r1, cen.pad.p, best.p, second.p, bidx.p, K, scale, cmaxn.p, d_labels, d_dists, amb.p, amb_count.p);
        else lloyd_refine_kernel<float><<<g, 128, 0, c->stream>>>(p->x32, ld, D, p->sqn, r0, r1, cen.pad.p, best.p, second.p, bidx.p, K, scale, cmaxn.p, d_labels, d_dists, amb.p, amb_count.p);
        CRX_CUDA(cudaGetLastError());
    }
    int h_amb = 0;
    CRX_CUDA(cudaMemcpyAsync(&h_amb, amb_count.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    opC.free_all();
    if (st != CRX_OK) return st;
    if (h_amb > 0) {
        unsigned long long add = (unsigned long long)h_amb;
        // exact FP64 scan of the ambiguous rows (ties included: lowest index wins there)
        DevBuf<double> scratch;
        if (!d_dists) CRX_TRY(scratch.alloc(c, N));
        CRX_TRY(lloyd_scan(c, p, amb.p, h_amb, cen, CRX_EUCLIDEAN, d_labels, d_dists ? d_dists : scratch.p));
        unsigned long long cur = 0;
        CRX_CUDA(cudaMemcpyAsync(&cur, c->counters + CRX_CNT_LLOYD_EXACT, sizeof(cur), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        cur += add;
        CRX_CUDA(cudaMemcpyAsync(c->counters + CRX_CNT_LLOYD_EXACT, &cur, sizeof(cur), cudaMemcpyHostToDevice, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return CRX_OK;
}

// ------------------------------------------------------------------------------------------------
// cluster sums (the data-parallel half of k_means, update.hpp:50-58): rows are grouped by label
// with the stable segment sort, every (cluster, chunk of CH rows) is summed by one CTA whose thread
// j owns coordinate j and walks the rows IN INPUT ORDER; chunk partials are then added in order.
// A cluster that fits one chunk therefore gets the reference's own sequential sum, bit for bit: always up to 16384
// members, and at any size as long as cutting less still leaves two chunks per SM.
// ------------------------------------------------------------------------------------------------
constexpr int SUM_CH = 16384;       // a cluster up to this size is ALWAYS one sequential sum (bit-exact)
constexpr int SUM_CH_SHORT = 1024;  // chunk used when some cluster is too long for that
constexpr int SUM_CH_MAX = 1 << 15; // largest sequential chain (~130 ns per row: a longer one would stall the whole step)

template <typename T>
__global__ void __launch_bounds__(CRX_MAXD)
chunk_sums_kernel(const T* __restrict__ x, int ld, int D, const int32_t* __restrict__ perm, const int32_t* __restrict__ off,
                  const int32_t* __restrict__ chunk_cluster, const int32_t* __restrict__ chunk_index, int whole_max, int ch_len,
                  double* __restrict__ partial /* [nchunks][D] */) {
    int ch = blockIdx.x;
    int cl = chunk_cluster[ch];
    int len = (off[cl + 1] - off[cl]) <= whole_max ? whole_max : ch_len;   // whole cluster in one chain, or short chunks
    int begin = off[cl] + chunk_index[ch] * len;
    int end = min(off[cl + 1], begin + len);
    int j = threadIdx.x;
    if (j >= D) return;
    double acc = 0.0;
    int r = begin;
    for (; r + 16 <= end; r += 16) {  // sixteen independent row reads in flight, added in row order
        T v[16];
#pragma unroll
        for (int u = 0; u < 16; u++) v[u] = x[(size_t)perm[r + u] * ld + j];
#pragma unroll
        for (int u = 0; u < 16; u++) acc = __dadd_rn(acc, (double)v[u]);
    }
    for (; r < end; r++) acc = __dadd_rn(acc, (double)x[(size_t)perm[r] * ld + j]);
    partial[(size_t)ch * D + j] = acc;
}

__global__ void combine_sums_kernel(const double* __restrict__ partial, const int32_t* __restrict__ chunk_first, int K, int D,
                                    const int32_t* __restrict__ off, int whole_max, int ch_len, double* __restrict__ sums,
                                    long long* __restrict__ counts) {
    int cl = blockIdx.x, j = threadIdx.x;
    int n = off[cl + 1] - off[cl];
    if (j == 0) counts[cl] = n;
    if (j >= D) return;
    int len = n <= whole_max ? whole_max : ch_len;
    int nch = (n + len - 1) / len;
    double acc = 0.0;
    for (int ch = 0; ch < nch; ch++) {
        double v = partial[(size_t)(chunk_first[cl] + ch) * D + j];
        acc = ch == 0 ? v : __dadd_rn(acc, v);
    }
    sums[(size_t)cl * D + j] = acc;
}

// means, convergence (update.hpp:57-80)
__global__ void kmeans_finish_kernel(const double* __restrict__ sums, const long long* __restrict__ counts,
                                     const double* __restrict__ oldc, int K, int D, int metric, double min_dist,
                                     double* __restrict__ newc, int* __restrict__ moved) {
    int cl = blockIdx.x * blockDim.x + threadIdx.x;
    if (cl >= K) return;
    double div = (double)counts[cl];
    double na = 0.0, nb = 0.0, acc = 0.0;
    X87 ip = {0.0, 0.0};
    for (int i = 0; i < D; i++) {
        double v = sums[(size_t)cl * D + i];
        if (div != 0.0) v = __ddiv_rn(v, div);  // divDimensionsByD skips count 0 (cust_vector.hpp:189)
        newc[(size_t)cl * D + i] = v;
        double o = oldc[(size_t)cl * D + i];
        if (metric == CRX_EUCLIDEAN) {
            double t = __dsub_rn(v, o);
            acc = __dadd_rn(acc, __dmul_rn(t, t));
        } else {
            na = __dadd_rn(na, __dmul_rn(v, v));
            nb = __dadd_rn(nb, __dmul_rn(o, o));
            x87_add(ip, __dmul_rn(v, o));
        }
    }
    double dist = metric == CRX_EUCLIDEAN ? __dsqrt_rn(acc) : __dsub_rn(1.0, cos_sim_x87(ip, na, nb));
    if (dist > min_dist) atomicExch(moved, 1);
}

__global__ void select_centroids_kernel(const double* __restrict__ newc, const double* __restrict__ oldc, size_t n,
                                        const int* __restrict__ moved, double* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = *moved ? newc[i] : oldc[i];
}

// ------------------------------------------------------------------------------------------------
// K6: k-means++ (initialization.hpp:72-156)
// ------------------------------------------------------------------------------------------------
template <typename T, int METRIC>
__global__ void __launch_bounds__(256)
kpp_update_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, int64_t N,
                  const double* __restrict__ cvec /* [ld] coordinates, [ld] = exact sum of squares */, int first,
                  double* __restrict__ mind, unsigned long long* __restrict__ maxbits,
                  const int32_t* __restrict__ rowmap /* nullable: only these rows */, const int* __restrict__ nrows_dev,
                  int32_t* __restrict__ nearest /* nullable: which chosen centroid gives mind */, int cur) {
    __shared__ rw::WarpTile tiles[8];
    __shared__ double vec[CRX_MAXD];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < ld; k += blockDim.x) vec[k] = cvec[k];
    __syncthreads();
    const int64_t total = rowmap ? (int64_t)*nrows_dev : N;
    for (int64_t base = ((int64_t)blockIdx.x * 8 + warp) * 32; base < total; base += (int64_t)gridDim.x * 256) {
        int64_t idx = base + lane;
        bool valid = idx < total;
        int64_t row = valid ? (rowmap ? (int64_t)rowmap[idx] : idx) : -1;
        double d = rw::dist_rows<T, METRIC>(x, ld, D, row, vec, valid ? sqn[row] : 1.0, cvec[ld], tiles[warp]);
        double m = 0.0;
        if (valid) {
            m = mind[row];
            if (first || d < m) { m = d; mind[row] = d; if (nearest) nearest[row] = cur; }  // running form of the `min == -1 || d < min` scan
        }
        // max over values > 0 (max_for_normalizing starts at 0 and uses '>', initialization.hpp:116-117)
        double mx = (valid && m > 0.0) ? m : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if (lane == 0 && mx > 0.0) atomicMax(maxbits, (unsigned long long)__double_as_longlong(mx));
    }
}

// Euclidean rounds after the first: by the triangle inequality a point whose nearest chosen centroid c_j lies at
// d(c_new, c_j) >= 2 mind cannot be closer to the new centroid than mind, so its row need not even be read: the pass
// below touches 12 bytes per point (mind, nearest) instead of the row, lists the rest for the exact update and feeds
// the unchanged minima into the normaliser.  The 1e-9 slack dwarfs the rounding of the three distances involved.
__global__ void kpp_cdist_kernel(const double* __restrict__ cmat, int ld, int D, int ncent, const double* __restrict__ cvec,
                                 double* __restrict__ cd) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncent) return;
    cd[j] = euclid_exact(cmat + (size_t)j * ld, cvec, D);
}
__global__ void __launch_bounds__(256)
kpp_prune_kernel(const double* __restrict__ mind, const int32_t* __restrict__ nearest, const double* __restrict__ cd, int64_t N,
                 int32_t* __restrict__ flagged, int* __restrict__ nflag, unsigned long long* __restrict__ maxbits) {
    int lane = threadIdx.x & 31;
    double wmax = 0.0;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; base < N; base += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = base + lane;
        bool flag = false;
        double m = 0.0;
        if (i < N) {
            m = mind[i];
            flag = !(cd[nearest[i]] >= 2.0 * m * (1.0 + 1e-9));   // NaN is listed too
            if (!flag && m > wmax) wmax = m;
        }
        unsigned fm = __ballot_sync(0xffffffffu, flag);
        if (fm) {
            int start = 0;
            if (lane == 0) start = atomicAdd(nflag, __popc(fm));
            start = __shfl_sync(0xffffffffu, start, 0);
            if (flag) flagged[start + __popc(fm & ((1u << lane) - 1u))] = (int32_t)i;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
    if (lane == 0 && wmax > 0.0) atomicMax(maxbits, (unsigned long long)__double_as_longlong(wmax));
}

// Rounds after the first, fp32 data: a point's minimum only changes when the new centroid is closer, which is rare.
// One streaming pass computes the distance in fp32 (a warp per row, one 16-byte load per lane, 8 rows in flight;
// relative error <= (D + 4) 2^-24 for the difference form, absolute for the cosine form) and lists only the rows that
// are not CERTAINLY farther than their current minimum; kpp_update_kernel then evaluates those exactly.  Rows that
// are not listed contribute their unchanged minimum to the normaliser here.
template <int METRIC>
__global__ void __launch_bounds__(256)
kpp_filter_kernel(const float* __restrict__ x, int ld, int D, const double* __restrict__ sqn, int64_t N,
                  const double* __restrict__ cvec, const double* __restrict__ mind, int32_t* __restrict__ flagged,
                  int* __restrict__ nflag, unsigned long long* __restrict__ maxbits,
                  const int32_t* __restrict__ rowlist /* nullable: only these rows */, const int* __restrict__ nlist) {
    constexpr int R = 8;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * lane < ld) cv = make_float4((float)cvec[4 * lane], (float)cvec[4 * lane + 1], (float)cvec[4 * lane + 2], (float)cvec[4 * lane + 3]);
    const double cn = cvec[ld];
    const float slack = (float)(D + 4) * 1.2e-7f;
    double wmax = 0.0;
    const int64_t total = rowlist ? (int64_t)*nlist : N;
    for (int64_t base = ((int64_t)blockIdx.x * 8 + warp) * R; base < total; base += (int64_t)gridDim.x * 8 * R) {
        float acc[R];
        // the decision inputs of this lane's row travel together with the row data (one latency, not two)
        const int64_t pos = base + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        const bool have = (lane & 3) == 0 && pos < total;
        const int64_t row = have ? (rowlist ? (int64_t)rowlist[pos] : pos) : 0;
        double m = 0.0, nrow = 1.0;
        if (have) { m = mind[row]; if (METRIC == CRX_COSINE) nrow = sqn[row]; }
#pragma unroll
        for (int r = 0; r < R; r++) {
            int64_t rpos = base + r;
            int64_t row = rpos < total ? (rowlist ? (int64_t)rowlist[rpos] : rpos) : -1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0 && 4 * lane < ld) v = __ldg(reinterpret_cast<const float4*>(x + row * ld + 4 * lane));
            if (METRIC == CRX_EUCLIDEAN) {
                float a = v.x - cv.x, b = v.y - cv.y, c2 = v.z - cv.z, d2 = v.w - cv.w;
                acc[r] = fmaf(a, a, fmaf(b, b, fmaf(c2, c2, d2 * d2)));
            } else {
                acc[r] = fmaf(v.x, cv.x, fmaf(v.y, cv.y, fmaf(v.z, cv.z, v.w * cv.w)));
            }
        }
        // 8 rows x 32 partial sums -> lane L ends with the total of row 4 b4 + 2 b3 + b2 (bits of L): halving exchanges
        // (4 + 2 + 1 shuffles), then two plain butterfly steps over bits 1 and 0
        const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
        float k4[4], k2[2], mine;
#pragma unroll
        for (int r = 0; r < 4; r++) k4[r] = (h4 ? acc[r + 4] : acc[r]) + __shfl_xor_sync(0xffffffffu, h4 ? acc[r] : acc[r + 4], 16);
#pragma unroll
        for (int r = 0; r < 2; r++) k2[r] = (h3 ? k4[r + 2] : k4[r]) + __shfl_xor_sync(0xffffffffu, h3 ? k4[r] : k4[r + 2], 8);
        mine = (h2 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? k2[0] : k2[1], 4);
        mine += __shfl_xor_sync(0xffffffffu, mine, 2);
        mine += __shfl_xor_sync(0xffffffffu, mine, 1);
        bool flag = false;
        if (have) {
            double lower;  // a certain lower bound of the exact distance
            if (METRIC == CRX_EUCLIDEAN) lower = (double)(sqrtf(fmaxf(mine, 0.f)) * (1.f - slack));
            else lower = 1.0 - ((double)mine / (sqrt(nrow) * sqrt(cn)) + (double)slack + 1e-7);
            flag = !(lower > m);  // NaN (zero vector under cosine) is listed too
        }
        unsigned fm = __ballot_sync(0xffffffffu, flag);
        if (fm) {
            int start = 0;
            if (lane == 0) start = atomicAdd(nflag, __popc(fm));
            start = __shfl_sync(0xffffffffu, start, 0);
            if (flag) flagged[start + __popc(fm & ((1u << lane) - 1u))] = (int32_t)row;
        }
        if (have && !flag && m > wmax) wmax = m;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
    if (lane == 0 && wmax > 0.0) atomicMax(maxbits, (unsigned long long)__double_as_longlong(wmax));
}

// coordinates of one stored row widened to double, followed by its exact sum of squares
template <typename T>
__global__ void stage_row_kernel(const T* __restrict__ x, int ld, const double* __restrict__ sqn, int64_t row, double* __restrict__ out) {
    for (int k = threadIdx.x; k < ld; k += blockDim.x) out[k] = (double)x[(size_t)row * ld + k];
    if (threadIdx.x == 0) out[ld] = sqn[row];
}

// p_v = (min_d / max)^2 (initialization.hpp:122-127, before the running sum)
__global__ void kpp_prob_kernel(const double* __restrict__ mind, int64_t N, const unsigned long long* __restrict__ maxbits,
                                double* __restrict__ p) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double mx = __longlong_as_double((long long)*maxbits);
    double t = __ddiv_rn(mind[i], mx);
    p[i] = __dmul_rn(t, t);
}

// ---- the tail of a k-means++ round on the device (no host round trip per round) ----
// st[]: [KPP_RNG] state of std::default_random_engine (minstd_rand0) as a double, [KPP_TOT + r] own shard total written into
// slot r = this rank, [KPP_ALL + r] the totals of all shards (all-gathered; world == 1: slot KPP_ALL is written directly)
constexpr int KPP_RNG = 0, KPP_TOT = 1, KPP_MAXW = 64, KPP_ALL = KPP_TOT + KPP_MAXW;
__global__ void kpp_total_kernel(const double* __restrict__ P, int64_t N, double* __restrict__ out) { *out = N > 0 ? P[N - 1] : 0.0; }

// initialization.hpp:131-149 for sharded rows: the running sum of the shard totals is the prefix sum at every shard boundary;
// x ~ U(0, total) from the reference's engine (libstdc++: generate_canonical<double, 53> takes two draws of minstd_rand0,
// sum = (x1 - 1) + (x2 - 1) R with R = 2147483646, canonical = sum / R^2, then canonical * (b - a) + a); the shard that holds
// x searches its own prefix sums (the reference's binary search) and stages the picked row: coordinates, exact sum of
// squares, global row -- the other ranks stage zeros, so an all-reduce(sum) of `share` is the broadcast.
template <typename T>
__global__ void kpp_draw_pick_kernel(const T* __restrict__ x, int ld, const double* __restrict__ sqn, const double* __restrict__ P, int64_t N,
                                     int64_t row_offset, double* __restrict__ st, int world, int me, const int64_t* __restrict__ lens,
                                     double* __restrict__ share, unsigned long long* counters) {
    __shared__ int64_t s_row;
    __shared__ int s_owner;
    if (threadIdx.x == 0) {
        const double* tot = world > 1 ? st + KPP_ALL : st + KPP_TOT + me;
        double cum_prev = 0.0, total = 0.0;
        for (int r = 0; r < world; r++) total = r == 0 ? tot[0] : __dadd_rn(total, tot[r]);
        unsigned long long s = (unsigned long long)st[KPP_RNG];
        s = s * 16807ull % 2147483647ull;
        const double x1 = (double)(s - 1ull);
        s = s * 16807ull % 2147483647ull;
        const double x2 = (double)(s - 1ull);
        st[KPP_RNG] = (double)s;
        const double R = 2147483646.0;
        const double sum = __dadd_rn(x1, __dmul_rn(x2, R));
        double canon = __ddiv_rn(sum, __dmul_rn(R, R));
        if (canon >= 1.0) canon = 0.99999999999999988897769753748434595763683319091796875;   // nextafter(1, 0)
        const double xr = __dadd_rn(__dmul_rn(canon, total), 0.0);
        int pick = world - 1;
        double cum = 0.0, before = 0.0;
        bool found = false;
        for (int r = 0; r < world; r++) {
            cum = r == 0 ? tot[0] : __dadd_rn(cum, tot[r]);
            if (!found && lens[r] > 0 && xr <= cum) { pick = r; before = cum_prev; found = true; }
            cum_prev = cum;
        }
        if (!found && world > 1) {   // the last shard: its prefix starts at the running sum in front of it
            double c2 = 0.0;
            for (int r = 0; r < world - 1; r++) c2 = r == 0 ? tot[0] : __dadd_rn(c2, tot[r]);
            before = c2;
        }
        s_owner = pick;
        int64_t ch = 0;
        if (pick == me && N > 0) {
            const double xl = pick > 0 ? __dsub_rn(xr, before) : xr;
            int64_t lo = 0, hi = N - 1;
            if (xl > P[lo]) {
                while (hi - lo > 1) {
                    int64_t m = lo + (hi - lo) / 2;
                    if (xl <= P[m]) hi = m; else lo = m;
                }
                ch = hi;
            }
            const double tol = 1e-12 * P[N - 1];
            if (fabs(xl - P[ch]) <= tol || (ch > 0 && fabs(xl - P[ch - 1]) <= tol)) atomicAdd(&counters[CRX_CNT_KPP_NEAR], 1ull);
        }
        s_row = ch;
    }
    __syncthreads();
    const bool mine = s_owner == me;
    const int64_t row = s_row;
    for (int k = threadIdx.x; k < ld; k += blockDim.x) share[k] = mine ? (double)x[(size_t)row * ld + k] : 0.0;
    if (threadIdx.x == 0) { share[ld] = mine ? sqn[row] : 0.0; share[ld + 1] = mine ? (double)(row_offset + row) : 0.0; }
}

// after the exchange: `share` holds the new centroid on every rank
__global__ void kpp_share_kernel(const double* __restrict__ share, int ld, double* __restrict__ cvec, double* __restrict__ vec_out, int D,
                                 int64_t* __restrict__ out_row) {
    for (int k = threadIdx.x; k <= ld; k += blockDim.x) cvec[k] = share[k];
    if (vec_out) for (int k = threadIdx.x; k < D; k += blockDim.x) vec_out[k] = share[k];
    if (threadIdx.x == 0) *out_row = (int64_t)share[ld + 1];
}

// ------------------------------------------------------------------------------------------------
// K7: range-search assignment (assignment.hpp:156-217).
//
// The reference sweeps the centroids in order and doubles the radius after EVERY centroid, so the
// global step j = sweep*K + c tests the annulus [r0 2^(j-1), r0 2^j) ([0, r0) for j = 0).  The
// annuli partition [0, inf), an assigned vector always has dist < min_radius at every later step,
// hence it is never reconsidered: a vector ends up at the FIRST step j at which some centroid
// c = j mod K has it in its bucket list with d(c, v) inside annulus j.  That is an order-free
// statement: every (centroid, bucket member) pair is evaluated independently (exact FP64
// distance), fires at most at one step a(d) and only if a(d) == c (mod K); atomicMin over the
// firing steps gives the assignment.  The reference stops after the first sweep that assigned
// nothing; the per-sweep counts reproduce that cut-off afterwards.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int annulus_index(double d, double r0) {
    if (!(r0 > 0.0) || !(d >= 0.0)) return -1;  // radius stays <= 0 for ever / NaN or negative distance never fires
    if (d < r0) return 0;
    int j = ilogb(d / r0) + 1;
    if (j < 1) j = 1;
    while (j > 1 && d < ldexp(r0, j - 1)) j--;
    while (!(d < ldexp(r0, j))) { j++; if (j > 4096) return -1; }
    return j;
}

template <typename T, int METRIC>
__global__ void __launch_bounds__(256)
range_fire_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, int K,
                  const double* __restrict__ cvec,            // [K][ld] centroid coordinates (stored rows gathered, or any vectors)
                  const double* __restrict__ csqn,            // [K] their sums of squares
                  const int32_t* __restrict__ cbucket,        // [L][K] bucket of every centroid in every table (LSH) or NULL
                  double r0, int nseg,                        // segments per centroid
                  const int32_t* __restrict__ seg_begin,      // [K*nseg] start position in seg_perm space
                  const int32_t* __restrict__ seg_end,        // [K*nseg]
                  const int32_t* const* __restrict__ seg_perm,  // [nseg] position -> row (per table) or one shared perm
                  const int32_t* __restrict__ bucket, int64_t N,  // [L][N] bucket ids for the duplicate filter (LSH) or NULL
                  int chunk, int cs0, int* __restrict__ key) {
    __shared__ rw::WarpTile tiles[8];
    __shared__ double vec[128];
    int cs = cs0 + blockIdx.x;  // centroid * nseg + seg (cs0: first one of this rank's share of the centroids)
    int c = cs / nseg, sgi = cs - c * nseg;
    int begin = seg_begin[cs] + blockIdx.y * chunk;
    int end = min(seg_end[cs], begin + chunk);
    if (begin >= end) return;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < ld; k += blockDim.x) vec[k] = cvec[(size_t)c * ld + k];
    __syncthreads();
    const int32_t* perm = seg_perm[bucket ? sgi : 0];
    double ncr = csqn[c];
    for (int base = begin + warp * 32; base < end; base += 256) {
        int pos = base + lane;
        int64_t row = pos < end ? (int64_t)perm[pos] : -1;
        if (row >= 0 && bucket) {  // union of the L buckets: count a vector in the first table that holds it
            for (int l = 0; l < sgi; l++)
                if (bucket[(size_t)l * N + row] == cbucket[(size_t)l * K + c]) { row = -1; break; }
        }
        double d = rw::dist_rows<T, METRIC>(x, ld, D, row, vec, row >= 0 ? sqn[row] : 1.0, ncr, tiles[warp]);
        if (row >= 0) {
            int a = annulus_index(d, r0);
            if (a >= 0 && a % K == c) atomicMin(&key[row], a);
        }
    }
}

// Shared-id mode: the reference caches distances under the key "<centroid id>to<vector id>"
// (assignment.hpp:183-194) and k_means names every new centre "k_means_center" (update.hpp:48), so from the second
// iteration of {range assignment, k_means} all centroids share one cache entry per vector: the distance every centroid
// uses for vector v is d0(v) = d(c0(v), v), c0(v) = the lowest centroid whose bucket list holds v (the first to evaluate it;
// an assigned vector is never evaluated again before that, see above).  v is then assigned at step a(d0) to centroid
// a(d0) mod K -- if that centroid's list holds v -- with distance d0.
__global__ void __launch_bounds__(256)
range_first_kernel(int nseg, const int32_t* __restrict__ seg_begin, const int32_t* __restrict__ seg_end,
                   const int32_t* const* __restrict__ seg_perm, int per_table, int chunk, int* __restrict__ first) {
    int cs = blockIdx.x;
    int c = cs / nseg, sgi = cs - c * nseg;
    int begin = seg_begin[cs] + blockIdx.y * chunk;
    int end = min(seg_end[cs], begin + chunk);
    const int32_t* perm = seg_perm[per_table ? sgi : 0];
    for (int pos = begin + threadIdx.x; pos < end; pos += blockDim.x) atomicMin(&first[perm[pos]], c);
}

template <typename T, int METRIC>
__global__ void range_shared_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, int64_t N, int K, int L,
                                    const double* __restrict__ cvec, const double* __restrict__ csqn, const int32_t* __restrict__ bucket,
                                    const int32_t* __restrict__ cbucket, double r0, const int* __restrict__ first,
                                    int* __restrict__ key, double* __restrict__ d0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c0 = first[i];
    if (c0 == INT_MAX) return;
    double d = metric_dist_exact(METRIC, cvec + (size_t)c0 * ld, x + (size_t)i * ld, D, csqn[c0], sqn[i]);
    d0[i] = d;
    int a = annulus_index(d, r0);
    if (a < 0) return;
    int c = a % K;
    bool member = false;
    for (int l = 0; l < L; l++) member |= bucket[(size_t)l * N + i] == cbucket[(size_t)l * K + c];
    if (member) key[i] = a;
}

__global__ void range_hist_kernel(const int* __restrict__ key, int64_t N, int K, int* __restrict__ hist, int nh) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int a = key[i];
    if (a != INT_MAX) { int s = a / K; if (s < nh) atomicAdd(&hist[s], 1); }
}

template <typename T, int METRIC>
__global__ void __launch_bounds__(256)
range_finalize_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, int64_t N, int K,
                      const double* __restrict__ cvec, const double* __restrict__ csqn, const int* __restrict__ key, int sweep_limit,
                      const double* __restrict__ shared_d0,   // shared-id mode: the one distance every centroid used for row i
                      int32_t* __restrict__ labels, double* __restrict__ dists) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int a = key[i];
    if (a != INT_MAX && a / K < sweep_limit) {
        int c = a % K;
        const T* xi = x + (size_t)i * ld;
        labels[i] = c;
        dists[i] = shared_d0 ? shared_d0[i] : metric_dist_exact(METRIC, cvec + (size_t)c * ld, xi, D, csqn[c], sqn[i]);
    } else {
        labels[i] = -1;
        dists[i] = 0.0;
    }
}

// min over centroid pairs (utils.hpp:161-178): one thread per (a, b > a) pair.  Distances are >= 0 (or NaN /
// -0-ish for cosine, which `d < min` never selects unless first), so the minimum of the non-negative ones is an
// atomicMin on their bit patterns; the host applies the `-1` sentinel rule.
template <typename T>
__global__ void min_pair_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn,
                                int K, int metric, unsigned long long* __restrict__ minbits,
                                int* __restrict__ any_negative_or_nan) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long npairs = (long long)K * (K - 1) / 2;
    if (t >= npairs) return;
    // unrank t -> (a, b), a < b, row-major over the strict upper triangle
    int a = (int)((2.0 * K - 1.0 - sqrt((2.0 * K - 1.0) * (2.0 * K - 1.0) - 8.0 * (double)t)) / 2.0);
    while ((long long)a * (2 * K - a - 1) / 2 > t) a--;
    while ((long long)(a + 1) * (2 * K - a - 2) / 2 <= t) a++;
    int b = (int)(t - (long long)a * (2 * K - a - 1) / 2) + a + 1;
    double d = metric_dist_exact(metric, x + (size_t)a * ld, x + (size_t)b * ld, D, sqn[a], sqn[b]);
    if (d >= 0.0) atomicMin(minbits, (unsigned long long)__double_as_longlong(d));
    else atomicExch(any_negative_or_nan, 1);
}

// the reference's scan itself (one thread): used when some pair distance is negative or NaN, where the
// `min == -1 || d < min` rule is order dependent
template <typename T>
__global__ void min_pair_seq_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn,
                                    int K, int metric, double* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double m = -1;
    for (int a = 0; a < K; a++)
        for (int b = a + 1; b < K; b++) {
            double d = metric_dist_exact(metric, x + (size_t)a * ld, x + (size_t)b * ld, D, sqn[a], sqn[b]);
            if (m == -1 || d < m) m = d;
        }
    *out = m;
}

// ------------------------------------------------------------------------------------------------
// K8: PAM medoid update (update.hpp:90-142).  Row sums of exact pairwise distances: one warp per
// candidate medoid, lanes over the cluster members; then per cluster the candidates within the
// summation-order tolerance of the minimum are re-summed sequentially in member order (the
// reference's order) and the first minimum wins.
// ------------------------------------------------------------------------------------------------
template <typename T, int METRIC>
__global__ void __launch_bounds__(256)
pam_rowsum_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, const int32_t* __restrict__ perm,
                  const int32_t* __restrict__ sorted_label, const int32_t* __restrict__ off, int64_t pos_begin, int64_t pos_end,
                  double* __restrict__ rowsum, double* __restrict__ rowerr) {
    __shared__ rw::WarpTile tiles[8];
    __shared__ double vecs[8][128];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t pos = pos_begin + (int64_t)blockIdx.x * 8 + warp;
    if (pos >= pos_end) return;
    int cl = sorted_label[pos];
    int begin = off[cl], end = off[cl + 1];
    int64_t mrow = perm[pos];
    rw::stage_vector<T>(x, ld, mrow, vecs[warp]);
    double nm = sqn[mrow];
    double acc = 0.0;
    for (int base = begin; base < end; base += 32) {
        int p = base + lane;
        int64_t row = p < end ? (int64_t)perm[p] : -1;
        double d = rw::dist_rows<T, METRIC>(x, ld, D, row, vecs[warp], row >= 0 ? sqn[row] : 1.0, nm, tiles[warp]);
        if (row >= 0) acc += d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    // exact distances, summed in a different order than the reference: off by at most ~n ulps
    if (lane == 0) { rowsum[pos] = acc; rowerr[pos] = fabs(acc) * (8.0 * (double)(end - begin + D) * 1.1102230246251565e-16); }
}

// ---- picking the medoid from row sums with error bars (exact FP64 distances summed in parallel order, or the
// tcgen05 row-sum scan for Euclidean): only the rows whose interval reaches below the smallest upper end are re-summed
// exactly, sequentially in member order (the reference's own floating-point sum), and the first minimum wins
// (update.hpp:110-127).
__global__ void __launch_bounds__(256)
pam_bounds_kernel(const int32_t* __restrict__ off, const double* __restrict__ rowsum, const double* __restrict__ rowerr, int D,
                  int32_t* __restrict__ cand /* [N] positions */, int* __restrict__ ncand_total, int32_t* __restrict__ winner /* [K] */) {
    __shared__ double red[256];
    __shared__ int count, base;
    int cl = blockIdx.x;
    int begin = off[cl], end = off[cl + 1];
    int n = end - begin;
    if (n == 0) { if (threadIdx.x == 0) winner[cl] = -1; return; }
    double m = INFINITY;
    for (int p = begin + threadIdx.x; p < end; p += blockDim.x) m = fmin(m, rowsum[p] + rowerr[p]);
    red[threadIdx.x] = m;
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    // the sequential double sums of the reference differ from the real sums by at most n ulps each
    double U = red[0] * (1.0 + 8.0 * (double)(n + D) * 1.1102230246251565e-16);
    int mine = 0, first = INT_MAX;
    for (int p = begin + threadIdx.x; p < end; p += blockDim.x)
        if (rowsum[p] - rowerr[p] <= U) { mine++; first = min(first, p); }
    if (mine) atomicAdd(&count, mine);
    __syncthreads();
    if (count == 0) {  // every sum is NaN (cosine with a zero vector): `min == -1 || s < min` keeps member 0
        if (threadIdx.x == 0) winner[cl] = begin;
        return;
    }
    if (count == 1) {  // its real sum is below every other row's by more than the rounding of the reference's sums
        if (mine) winner[cl] = first;
        return;
    }
    if (threadIdx.x == 0) { base = atomicAdd(ncand_total, count); winner[cl] = -2; count = 0; }
    __syncthreads();
    for (int p = begin + threadIdx.x; p < end; p += blockDim.x)
        if (rowsum[p] - rowerr[p] <= U) cand[base + atomicAdd(&count, 1)] = p;
}

// one CTA per candidate: the 8 warps compute 256 exact distances at a time into shared memory, thread 0 adds them in
// member order (update.hpp:118's own sequential sum) while the warps are already working on the next 256
template <typename T, int METRIC>
__global__ void __launch_bounds__(256)
pam_exact_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, const int32_t* __restrict__ perm,
                 const int32_t* __restrict__ sorted_label, const int32_t* __restrict__ off, const int32_t* __restrict__ cand,
                 int ncand, double* __restrict__ exact /* [ncand] */, unsigned long long* counters) {
    __shared__ rw::WarpTile tiles[8];
    __shared__ double vecs[8][128];
    __shared__ double dist[2][256];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int i = blockIdx.x;
    int pos = cand[i];
    int cl = sorted_label[pos];
    int begin = off[cl], end = off[cl + 1];
    int64_t mrow = perm[pos];
    rw::stage_vector<T>(x, ld, mrow, vecs[warp]);
    double nm = sqn[mrow], s = 0.0;
    int buf = 0;
    for (int b = begin; b < end + 256; b += 256, buf ^= 1) {
        if (b < end) {
            int p = b + warp * 32 + lane;
            int64_t row = p < end ? (int64_t)perm[p] : -1;
            double d = rw::dist_rows<T, METRIC>(x, ld, D, row, vecs[warp], row >= 0 ? sqn[row] : 1.0, nm, tiles[warp]);
            dist[buf][threadIdx.x] = d;
        }
        if (threadIdx.x == 0 && b > begin) {  // previous batch
            int cntv = min(256, end - (b - 256));
            const double* dv = dist[buf ^ 1];
            for (int j = 0; j < cntv; j++) s = __dadd_rn(s, dv[j]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { exact[i] = s; atomicAdd(&counters[CRX_CNT_PAM_EXACT], 1ull); }
}

__global__ void pam_final_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ sorted_label,
                                 const int32_t* __restrict__ off, const int32_t* __restrict__ cand, int ncand,
                                 const double* __restrict__ exact, const int32_t* __restrict__ winner,
                                 const int32_t* __restrict__ crow, int K, int32_t* __restrict__ new_crow, int* __restrict__ swapped) {
    int cl = blockIdx.x * blockDim.x + threadIdx.x;
    if (cl >= K) return;
    int w = winner[cl];
    if (w == -1) { new_crow[cl] = crow[cl]; return; }
    if (w == -2) {
        double best = 0.0;
        int bp = INT_MAX;
        for (int i = 0; i < ncand; i++) {  // few candidates overall; each cluster's are contiguous but unordered
            int p = cand[i];
            if (sorted_label[p] != cl) continue;
            double v = exact[i];
            if (bp == INT_MAX || v < best || (v == best && p < bp)) { best = v; bp = p; }
        }
        w = bp;
    }
    int row = perm[w];
    new_crow[cl] = row;
    if (row != crow[cl]) atomicExch(swapped, 1);
}

// ------------------------------------------------------------------------------------------------
// K11: silhouette (silhouette.hpp:32-144)
// ------------------------------------------------------------------------------------------------
__global__ void near_centroid_kernel(const double* __restrict__ cent, const double* __restrict__ csqn, int K, int ld, int D,
                                     int metric, int32_t* __restrict__ near) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= K) return;
    double mn = 0.0;
    int arg = 0;
    bool first = true;
    for (int b = 0; b < K; b++) {
        if (b == a) continue;
        double d = metric_dist_exact(metric, cent + (size_t)a * ld, cent + (size_t)b * ld, D, csqn[a], csqn[b]);
        if (first || d < mn) { mn = d; arg = b; first = false; }
    }
    near[a] = arg;
}

template <typename T, int METRIC>
__global__ void __launch_bounds__(256)
silhouette_point_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, const int32_t* __restrict__ perm,
                        const int32_t* __restrict__ sorted_label, const int32_t* __restrict__ off,
                        const int32_t* __restrict__ near, int64_t pos_begin, int64_t pos_end, double* __restrict__ s_out) {
    __shared__ rw::WarpTile tiles[8];
    __shared__ double vecs[8][128];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t pos = pos_begin + (int64_t)blockIdx.x * 8 + warp;
    if (pos >= pos_end) return;
    int cl = sorted_label[pos];
    int64_t mrow = perm[pos];
    rw::stage_vector<T>(x, ld, mrow, vecs[warp]);
    double nm = sqn[mrow];
    double sums[2];
    for (int which = 0; which < 2; which++) {
        int tc = which == 0 ? cl : near[cl];
        int begin = off[tc], end = off[tc + 1];
        double acc = 0.0;
        for (int base = begin; base < end; base += 32) {
            int p = base + lane;
            int64_t row = p < end ? (int64_t)perm[p] : -1;
            double d = rw::dist_rows<T, METRIC>(x, ld, D, row, vecs[warp], row >= 0 ? sqn[row] : 1.0, nm, tiles[warp]);
            if (row >= 0) acc += d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        sums[which] = acc;
    }
    if (lane == 0) {
        int n_own = off[cl + 1] - off[cl];
        int n_nb = off[near[cl] + 1] - off[near[cl]];
        double a = sums[0], b = sums[1];
        if (n_own != 1) a = a / (double)(n_own - 1);
        b = b / (double)n_nb;
        double mx = a;
        if (b > a) mx = b;
        s_out[pos] = (b - a) / mx;
    }
}

__global__ void silhouette_reduce_kernel(const double* __restrict__ s, const int32_t* __restrict__ off, int K,
                                         double* __restrict__ sils, double* __restrict__ raw) {
    int cl = blockIdx.x * blockDim.x + threadIdx.x;
    if (cl >= K) return;
    double acc = 0.0;
    for (int p = off[cl]; p < off[cl + 1]; p++) acc = acc + s[p];  // member order = input order
    raw[cl] = acc;
    sils[cl] = acc / (double)(off[cl + 1] - off[cl]);
}

__global__ void merge_remaining_kernel(const int32_t* __restrict__ tl, const double* __restrict__ td, int64_t n, int32_t* __restrict__ labels,
                                       double* __restrict__ dists) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && labels[i] == -1) { labels[i] = tl[i]; dists[i] = td[i]; }
}

__global__ void range_iota_kernel(int32_t* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

__global__ void fill_int_kernel(int* p, int64_t n, int v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void gather_int_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int n, int32_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

// ------------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------------
template <typename F32, typename F64>
static int by_type(const crx_points* p, F32 f32, F64 f64) { return p->x64 ? f64() : f32(); }

// ---- exchange steps of the sharded entry points (include/crx.h, crx_comm) ----
static inline int comm_world(const crx_comm* cm) { return (cm && cm->world > 1) ? cm->world : 1; }
static inline int comm_rank(const crx_comm* cm) { return (cm && cm->world > 1) ? cm->rank : 0; }
static int comm_done(int rc, const char* what) {
    if (rc == 0) return CRX_OK;
    crx_set_error("collective %s failed in the caller's callback (%d)", what, rc);
    return CRX_ERR_COMM;
}
static int comm_allreduce(crx_ctx* c, const crx_comm* cm, void* buf, int64_t n, int dtype, int op, int mem) {
    if (comm_world(cm) == 1 || n == 0) return CRX_OK;
    CRX_REQUIRE(cm->allreduce, "crx_comm.allreduce is NULL");
    if (mem == CRX_DEVICE && !cm->stream_ordered) CRX_CUDA(cudaStreamSynchronize(c->stream));
    return comm_done(cm->allreduce(cm->user, buf, n, dtype, op, mem), "allreduce");
}
static int comm_allgather(crx_ctx* c, const crx_comm* cm, const void* send, void* recv, int64_t n, int dtype, int mem) {
    CRX_REQUIRE(cm && cm->allgather, "crx_comm.allgather is NULL");
    if (mem == CRX_DEVICE && !cm->stream_ordered) CRX_CUDA(cudaStreamSynchronize(c->stream));
    return comm_done(cm->allgather(cm->user, send, recv, n, dtype, mem), "allgather");
}
// cp: the K centroids as an FP64 point set of their own (stored rows gathered, or the caller's vectors); h_crow[c] = the stored
// row centroid c aliases (it is assigned to its own cluster at the end, assignment.hpp:126-128) or -1; d_cbucket: [L][K]
// bucket of every centroid in every table (LSH) or NULL (hypercube); shared_ids: SURVEY App. A-2, see range_first_kernel
static int range_assign_common(crx_ctx* c, const crx_points* p, const int32_t* h_crow, const crx_points* cp, const int32_t* d_cbucket,
                               bool shared_ids, int K, int metric, int nseg,
                               const std::vector<int32_t>& h_begin, const std::vector<int32_t>& h_end,
                               const std::vector<const int32_t*>& h_perm, const int32_t* d_bucket, int32_t* labels,
                               double* dists, int mem, int32_t* before, const crx_comm* comm) {
    const int world = comm_world(comm), me = comm_rank(comm);
    int64_t N = p->n;
    int D = p->d, ld = p->ld;
    CRX_REQUIRE(cp && cp->x64 && cp->n == K && cp->d == D && cp->ld == ld, "centroid set");
    CRX_REQUIRE(!shared_ids || (d_bucket && d_cbucket), "the shared-id mode needs LSH tables");
    IoBuf<int32_t> lab, bef;
    IoBuf<double> dis;
    CRX_TRY(lab.bind(c, labels, N, mem, false));
    CRX_TRY(dis.bind(c, dists, N, mem, false));
    CRX_TRY(bef.bind(c, before, N, mem, false));
    DevBuf<int32_t> d_crow, d_begin, d_end;
    DevBuf<const int32_t*> d_perm;
    DevBuf<int> key, hist;
    const int NH = 4096;
    CRX_TRY(d_crow.alloc(c, K)); CRX_TRY(d_begin.alloc(c, h_begin.size())); CRX_TRY(d_end.alloc(c, h_end.size()));
    CRX_TRY(d_perm.alloc(c, h_perm.size())); CRX_TRY(key.alloc(c, N)); CRX_TRY(hist.alloc(c, NH));
    CRX_CUDA(cudaMemcpyAsync(d_crow.p, h_crow, K * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CRX_CUDA(cudaMemcpyAsync(d_begin.p, h_begin.data(), h_begin.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CRX_CUDA(cudaMemcpyAsync(d_end.p, h_end.data(), h_end.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CRX_CUDA(cudaMemcpyAsync(d_perm.p, h_perm.data(), h_perm.size() * sizeof(const int32_t*), cudaMemcpyHostToDevice, c->stream));
    // r0 = min centroid-centroid distance / 2 (assignment.hpp:161)
    DevBuf<unsigned long long> minbits;
    DevBuf<int> oddflag;
    CRX_TRY(minbits.alloc(c, 1)); CRX_TRY(oddflag.alloc(c, 1));
    CRX_CUDA(cudaMemsetAsync(minbits.p, 0xff, sizeof(unsigned long long), c->stream));
    CRX_CUDA(cudaMemsetAsync(oddflag.p, 0, sizeof(int), c->stream));
    long long npairs = (long long)K * (K - 1) / 2;
    if (npairs > 0) {
        CRX_KERNEL(c, "min_pair");
        int g = crx_grid(npairs, 128);
        min_pair_kernel<double><<<g, 128, 0, c->stream>>>(cp->x64, ld, D, cp->sqn, K, metric, minbits.p, oddflag.p);
    }
    unsigned long long h_minbits = ~0ull;
    int h_odd = 0;
    CRX_CUDA(cudaMemcpyAsync(&h_minbits, minbits.p, sizeof(h_minbits), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaMemcpyAsync(&h_odd, oddflag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    double mn = -1;  // find_min_vector_distance's sentinel: K == 1 leaves -1
    if (h_minbits != ~0ull) memcpy(&mn, &h_minbits, sizeof(double));
    if (h_odd) {
        DevBuf<double> seq;
        CRX_TRY(seq.alloc(c, 1));
        CRX_KERNEL(c, "min_pair_seq");
        min_pair_seq_kernel<double><<<1, 32, 0, c->stream>>>(cp->x64, ld, D, cp->sqn, K, metric, seq.p);
        CRX_CUDA(cudaMemcpyAsync(&mn, seq.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
    }
    double r0 = mn / 2;
    CRX_CUDA(cudaMemsetAsync(hist.p, 0, NH * sizeof(int), c->stream));
    { CRX_KERNEL(c, "fill_key"); fill_int_kernel<<<crx_grid(N, 256), 256, 0, c->stream>>>(key.p, N, INT_MAX); }
    // Centroid c only fires for distances in [r0 2^(c-1+sK), r0 2^(c+sK)): with the radius doubling after every centroid
    //  the annuli of all but the first few centroids lie beyond the diameter of the data, so their
    // buckets need not be probed at all -- nothing in them can fire.  dmax bounds every distance from above.
    int c_limit = K;
    {
        double dmax = 2.0 + 1e-9;   // cosine distance
        if (metric == CRX_EUCLIDEAN) {
            double mx = 0;
            CRX_TRY(points_maxabs(c, p, &mx));
            dmax = 2.0 * sqrt((double)D) * mx * 1.00001;
        }
        if (std::isfinite(dmax)) {
            int a_max = annulus_index(dmax, r0);           // largest step any pair can fire at (-1: the radius never grows)
            if (a_max < 0) c_limit = (r0 > 0.0) ? K : 0;   // -1 with r0 > 0 means "beyond the step cap": probe everything
            else if (a_max < K) c_limit = a_max + 1;
        }
    }
    DevBuf<double> d0;
    const int chunk = 8192;
    if (shared_ids) {
        // every rank walks every bucket list (the lists are short next to the Lloyd pass that follows); c0(v) needs all K
        // centroids, whatever their annuli
        DevBuf<int> first;
        CRX_TRY(first.alloc(c, N)); CRX_TRY(d0.alloc(c, N));
        { CRX_KERNEL(c, "fill_key"); fill_int_kernel<<<crx_grid(N, 256), 256, 0, c->stream>>>(first.p, N, INT_MAX); }
        int maxlen = 0;
        for (size_t i = 0; i < (size_t)K * nseg; i++) maxlen = std::max(maxlen, h_end[i] - h_begin[i]);
        if (maxlen > 0) {
            CRX_KERNEL(c, "range_first");
            dim3 grid((unsigned)(K * nseg), (unsigned)((maxlen + chunk - 1) / chunk));
            range_first_kernel<<<grid, 256, 0, c->stream>>>(nseg, d_begin.p, d_end.p, d_perm.p, 1, chunk, first.p);
        }
        {
            CRX_KERNEL(c, "range_shared");
#define LAUNCH_S(T, M, xptr) range_shared_kernel<T, M><<<crx_grid(N, 256), 256, 0, c->stream>>>(xptr, ld, D, p->sqn, N, K, nseg, cp->x64, cp->sqn, d_bucket, d_cbucket, r0, first.p, key.p, d0.p)
            if (p->x64) { if (metric == CRX_EUCLIDEAN) LAUNCH_S(double, CRX_EUCLIDEAN, p->x64); else LAUNCH_S(double, CRX_COSINE, p->x64); }
            else { if (metric == CRX_EUCLIDEAN) LAUNCH_S(float, CRX_EUCLIDEAN, p->x32); else LAUNCH_S(float, CRX_COSINE, p->x32); }
#undef LAUNCH_S
        }
        CRX_CUDA(cudaGetLastError());
    } else {
    // this rank probes the buckets of its share of those centroids; the firing steps are combined with a min
    const int c_lo = (int)((int64_t)c_limit * me / world), c_hi = (int)((int64_t)c_limit * (me + 1) / world);
    int maxlen = 0;
    for (size_t i = (size_t)c_lo * nseg; i < (size_t)c_hi * nseg; i++) maxlen = std::max(maxlen, h_end[i] - h_begin[i]);
    dim3 grid((unsigned)std::max(1, (c_hi - c_lo) * nseg), (unsigned)std::max(1, (maxlen + chunk - 1) / chunk));
    if (maxlen > 0) {
        CRX_KERNEL(c, "range_fire");
#define LAUNCH_R(T, M, xptr) range_fire_kernel<T, M><<<grid, 256, 0, c->stream>>>(xptr, ld, D, p->sqn, K, cp->x64, cp->sqn, d_cbucket, r0, nseg, d_begin.p, d_end.p, d_perm.p, d_bucket, N, chunk, c_lo * nseg, key.p)
        if (p->x64) { if (metric == CRX_EUCLIDEAN) LAUNCH_R(double, CRX_EUCLIDEAN, p->x64); else LAUNCH_R(double, CRX_COSINE, p->x64); }
        else { if (metric == CRX_EUCLIDEAN) LAUNCH_R(float, CRX_EUCLIDEAN, p->x32); else LAUNCH_R(float, CRX_COSINE, p->x32); }
#undef LAUNCH_R
    }
    }
    CRX_TRY(comm_allreduce(c, comm, key.p, N, CRX_I32, CRX_MIN, CRX_DEVICE));
    { CRX_KERNEL(c, "range_hist"); range_hist_kernel<<<crx_grid(N, 256), 256, 0, c->stream>>>(key.p, N, K, hist.p, NH); }
    std::vector<int> h_hist(NH);
    CRX_CUDA(cudaMemcpyAsync(h_hist.data(), hist.p, NH * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    int limit = 0;  // first sweep that assigns nothing ends the do-while (assignment.hpp:216)
    while (limit < NH && h_hist[limit] > 0) limit++;
    {
        CRX_KERNEL(c, "range_finalize");
#define LAUNCH_F(T, M, xptr) range_finalize_kernel<T, M><<<crx_grid(N, 256), 256, 0, c->stream>>>(xptr, ld, D, p->sqn, N, K, cp->x64, cp->sqn, key.p, limit, shared_ids ? d0.p : nullptr, lab.dev, dis.dev)
        if (p->x64) { if (metric == CRX_EUCLIDEAN) LAUNCH_F(double, CRX_EUCLIDEAN, p->x64); else LAUNCH_F(double, CRX_COSINE, p->x64); }
        else { if (metric == CRX_EUCLIDEAN) LAUNCH_F(float, CRX_EUCLIDEAN, p->x32); else LAUNCH_F(float, CRX_COSINE, p->x32); }
#undef LAUNCH_F
    }
    CRX_CUDA(cudaGetLastError());
    if (before) {
        CRX_CUDA(cudaMemcpyAsync(bef.dev, lab.dev, N * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
        CRX_TRY(bef.flush());
    }
    // lloyds_for_remaining over the unassigned rows (assignment.hpp:121-124), then the self-assignment
    DevBuf<int32_t> rows;
    DevBuf<int> cnt;
    DevBuf<double> cmat;
    CRX_TRY(rows.alloc(c, N)); CRX_TRY(cnt.alloc(c, 1)); CRX_TRY(cmat.alloc(c, (size_t)K * D));
    CRX_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int), c->stream));
    { CRX_KERNEL(c, "compact_unassigned"); compact_unassigned_kernel<<<crx_grid(N, 256), 256, 0, c->stream>>>(lab.dev, N, rows.p, cnt.p); }
    int h_cnt = 0;
    CRX_CUDA(cudaMemcpyAsync(&h_cnt, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    {   // the centroid coordinates as a compact [K][D] matrix for the Lloyd pass
        DevBuf<int32_t> iota;
        CRX_TRY(iota.alloc(c, K));
        { CRX_KERNEL(c, "iota"); range_iota_kernel<<<crx_grid(K, 128), 128, 0, c->stream>>>(iota.p, K); }
        CRX_TRY(gather_rows(c, cp, iota.p, K, cmat.p));
    }
    Centroids cen;
    CRX_TRY(cen.stage(c, cmat.p, CRX_DEVICE, K, D, ld));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    const bool tensor_rest = metric == CRX_EUCLIDEAN && K >= 32 && N >= 1024 && (int64_t)h_cnt * 4 >= N && !tc_disabled();
    if (tensor_rest || world > 1) {
        // results for this rank's share of the rows go into scratch (zeros elsewhere), are summed over the ranks, and
        // the unassigned rows take theirs.  Tensor path (most rows are left -- the radius doubles per centroid,
        // SURVEY App. A-7): filter + exact refine over all rows of the share; otherwise the exact scan of the
        // unassigned rows of the share.
        const int64_t lo = N * me / world, hi = N * (me + 1) / world;
        DevBuf<int32_t> tl;
        DevBuf<double> td;
        CRX_TRY(tl.alloc(c, N)); CRX_TRY(td.alloc(c, N));
        if (world > 1) {
            CRX_CUDA(cudaMemsetAsync(tl.p, 0, N * sizeof(int32_t), c->stream));
            CRX_CUDA(cudaMemsetAsync(td.p, 0, N * sizeof(double), c->stream));
        }
        if (tensor_rest) {
            CRX_TRY(lloyd_scan_tc(c, p, cen, tl.p, td.p, lo, hi));
        } else {
            CRX_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int), c->stream));
            if (hi > lo) { CRX_KERNEL(c, "compact_unassigned"); compact_unassigned_kernel<<<crx_grid(hi - lo, 256), 256, 0, c->stream>>>(lab.dev, hi, rows.p, cnt.p, lo); }
            int mine = 0;
            CRX_CUDA(cudaMemcpyAsync(&mine, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CRX_CUDA(cudaStreamSynchronize(c->stream));
            CRX_TRY(lloyd_scan(c, p, rows.p, mine, cen, metric, tl.p, td.p));
        }
        CRX_TRY(comm_allreduce(c, comm, tl.p, N, CRX_I32, CRX_SUM, CRX_DEVICE));
        CRX_TRY(comm_allreduce(c, comm, td.p, N, CRX_F64, CRX_SUM, CRX_DEVICE));
        CRX_KERNEL(c, "merge_remaining");
        merge_remaining_kernel<<<crx_grid(N, 256), 256, 0, c->stream>>>(tl.p, td.p, N, lab.dev, dis.dev);
    } else {
        CRX_TRY(lloyd_scan(c, p, rows.p, h_cnt, cen, metric, lab.dev, dis.dev));
    }
    { CRX_KERNEL(c, "self_assign"); self_assign_kernel<<<1, 32, 0, c->stream>>>(d_crow.p, K, lab.dev, dis.dev); }
    CRX_CUDA(cudaGetLastError());
    CRX_TRY(lab.flush());
    CRX_TRY(dis.flush());
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}

extern "C" {

int crx_rand_selection(crx_ctx* c, const crx_points* p, int K, uint64_t seed, int32_t* out) {
    CRX_REQUIRE(c && p && out, "NULL argument");
    CRX_REQUIRE(K >= 1 && K <= p->n, "cluster_num must be in [1, N]");
    // initialization.hpp:40-68: host-only integer logic (ids are unique row indices)
    std::default_random_engine e;
    e.seed((unsigned long)seed);
    std::uniform_int_distribution<int> ui(0, (int)p->n - 1);
    out[0] = ui(e);
    for (int i = 1; i < K; i++) {
        int r;
        bool clash;
        do {
            r = ui(e);
            clash = false;
            for (int j = 0; j < i; j++) if (out[j] == r) { clash = true; break; }
        } while (clash);
        out[i] = r;
    }
    return CRX_OK;
}

int crx_k_means_pp_sharded(crx_ctx* c, const crx_points* p, int64_t row_offset, int64_t n_global, int K, int metric, uint64_t seed,
                           const crx_comm* comm, int64_t* out, double* out_vectors) {
    CRX_REQUIRE(c && p && out, "NULL argument");
    CRX_REQUIRE(K >= 1, "cluster_num");
    CRX_REQUIRE(metric == CRX_EUCLIDEAN || metric == CRX_COSINE, "metric");
    CRX_REQUIRE(n_global >= 1 && n_global < (1ll << 31) && row_offset >= 0 && row_offset + p->n <= n_global, "row range");
    CRX_CUDA(cudaSetDevice(c->device));
    const int world = comm_world(comm), me = comm_rank(comm);
    int64_t N = p->n;
    int ld = p->ld, D = p->d;
    // who owns which rows
    std::vector<int64_t> offs(world, 0), lens(world, N);
    if (world > 1) {
        int64_t mine[2] = {row_offset, N};
        std::vector<int64_t> all((size_t)2 * world);
        CRX_TRY(comm_allgather(c, comm, mine, all.data(), 2, CRX_I64, CRX_HOST));
        for (int r = 0; r < world; r++) { offs[r] = all[2 * r]; lens[r] = all[2 * r + 1]; }
    } else offs[0] = row_offset;
    auto owner_of = [&](int64_t g) { for (int r = 0; r < world; r++) if (g >= offs[r] && g < offs[r] + lens[r]) return r; return -1; };
    std::default_random_engine e;
    e.seed((unsigned long)seed);
    std::uniform_int_distribution<int> ui(0, (int)n_global - 1);
    out[0] = ui(e);
    DevBuf<double> mind, prob, P, cvec;
    DevBuf<unsigned long long> mx;
    DevBuf<char> tmp;
    CRX_TRY(mind.alloc(c, N)); CRX_TRY(prob.alloc(c, N)); CRX_TRY(P.alloc(c, N)); CRX_TRY(mx.alloc(c, 1));
    CRX_TRY(cvec.alloc(c, ld + 1));
    DevBuf<int32_t> flagged, nearest;
    DevBuf<int> nflag;
    DevBuf<double> cmat, cd;   // coordinates of the centroids chosen so far, and their distances to the newest one
    CRX_TRY(flagged.alloc(c, N)); CRX_TRY(nflag.alloc(c, 1));
    static const bool prune_off = getenv("CRX_KPP_NOPRUNE") != nullptr && getenv("CRX_KPP_NOPRUNE")[0] == '1';
    const bool prune = metric == CRX_EUCLIDEAN && N >= 4096 && !prune_off;
    DevBuf<int32_t> flagged2;
    DevBuf<int> nflag2;
    if (prune) {
        CRX_TRY(nearest.alloc(c, N)); CRX_TRY(cmat.alloc(c, (size_t)K * ld)); CRX_TRY(cd.alloc(c, K));
        CRX_TRY(flagged2.alloc(c, N)); CRX_TRY(nflag2.alloc(c, 1));
    }
    size_t bytes = 0;
    if (N > 0) CRX_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, prob.p, P.p, (int)N, c->stream));
    CRX_TRY(tmp.alloc(c, bytes));
    int gridu = (int)((N + 255) / 256);
    // device-side state of the rounds: RNG state, shard totals (own slot + gathered), lengths of the shards
    DevBuf<double> st, share, vecs_dev;
    DevBuf<int64_t> rows_dev, lens_dev;
    CRX_TRY(st.alloc(c, KPP_ALL + world)); CRX_TRY(share.alloc(c, ld + 2)); CRX_TRY(rows_dev.alloc(c, K)); CRX_TRY(lens_dev.alloc(c, world));
    if (out_vectors) CRX_TRY(vecs_dev.alloc(c, (size_t)K * D));
    {
        // the engine's state after the first draw (minstd_rand0: one integer), handed to the device
        std::stringstream ss;
        ss << e;
        unsigned long long state = 0;
        ss >> state;
        std::vector<double> h_st(KPP_ALL + world, 0.0);
        h_st[KPP_RNG] = (double)state;
        CRX_CUDA(cudaMemcpyAsync(st.p, h_st.data(), h_st.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CRX_CUDA(cudaMemcpyAsync(lens_dev.p, lens.data(), world * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));   // h_st / lens are stack / heap temporaries
    }
    {
        // the first centroid: its owner stages the row, the others contribute zeros
        int own = owner_of(out[0]);
        CRX_REQUIRE(own >= 0, "the shards do not cover the chosen row");
        CRX_CUDA(cudaMemsetAsync(share.p, 0, (ld + 2) * sizeof(double), c->stream));
        if (own == me) {
            if (p->x64) stage_row_kernel<double><<<1, 128, 0, c->stream>>>(p->x64, ld, p->sqn, out[0] - row_offset, share.p);
            else stage_row_kernel<float><<<1, 128, 0, c->stream>>>(p->x32, ld, p->sqn, out[0] - row_offset, share.p);
            const double g = (double)out[0];
            CRX_CUDA(cudaMemcpyAsync(share.p + ld + 1, &g, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            CRX_CUDA(cudaStreamSynchronize(c->stream));
        }
        if (world > 1) CRX_TRY(comm_allreduce(c, comm, share.p, ld + 2, CRX_F64, CRX_SUM, CRX_DEVICE));
        { CRX_KERNEL(c, "kpp_share"); kpp_share_kernel<<<1, 128, 0, c->stream>>>(share.p, ld, cvec.p, vecs_dev.p, D, rows_dev.p); }
    }
    for (int i = 1; i < K; i++) {
        // cvec holds centroid i-1 (shared at the end of the previous round)
        CRX_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(unsigned long long), c->stream));
        const bool filter = i > 1 && p->x64 == nullptr && N >= 4096 && ld <= 128;   // fp32 data only (its fp32 copy is exact); one 16-byte piece per lane covers the row
        if (prune) CRX_CUDA(cudaMemcpyAsync(cmat.p + (size_t)(i - 1) * ld, cvec.p, ld * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        if (prune && i > 1) {
            CRX_CUDA(cudaMemsetAsync(nflag.p, 0, sizeof(int), c->stream));
            { CRX_KERNEL(c, "kpp_cdist"); kpp_cdist_kernel<<<crx_grid(i - 1, 64), 64, 0, c->stream>>>(cmat.p, ld, D, i - 1, cvec.p, cd.p); }
            { CRX_KERNEL(c, "kpp_prune"); kpp_prune_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(mind.p, nearest.p, cd.p, N, flagged.p, nflag.p, mx.p); }
            const int32_t* list = flagged.p;
            const int* nl = nflag.p;
            if (p->x64 == nullptr && ld <= 128) {   // fp32 data: the survivors' rows are read once in fp32; what that cannot rule out is exact
                CRX_CUDA(cudaMemsetAsync(nflag2.p, 0, sizeof(int), c->stream));
                CRX_KERNEL(c, "kpp_filter");
                kpp_filter_kernel<CRX_EUCLIDEAN><<<c->sm_count * 8, 256, 0, c->stream>>>(p->x32, ld, D, p->sqn, N, cvec.p, mind.p, flagged2.p, nflag2.p, mx.p, flagged.p, nflag.p);
                list = flagged2.p; nl = nflag2.p;
            }
            CRX_KERNEL(c, "kpp_update");
            int ge = c->sm_count * 4;
            if (p->x64) kpp_update_kernel<double, CRX_EUCLIDEAN><<<ge, 256, 0, c->stream>>>(p->x64, ld, D, p->sqn, N, cvec.p, 0, mind.p, mx.p, list, nl, nearest.p, i - 1);
            else kpp_update_kernel<float, CRX_EUCLIDEAN><<<ge, 256, 0, c->stream>>>(p->x32, ld, D, p->sqn, N, cvec.p, 0, mind.p, mx.p, list, nl, nearest.p, i - 1);
        } else if (N > 0 && filter) {
            CRX_CUDA(cudaMemsetAsync(nflag.p, 0, sizeof(int), c->stream));
            {
                CRX_KERNEL(c, "kpp_filter");
                int gf = (int)std::min<int64_t>((N + 63) / 64, (int64_t)c->sm_count * 8);
                if (metric == CRX_EUCLIDEAN) kpp_filter_kernel<CRX_EUCLIDEAN><<<gf, 256, 0, c->stream>>>(p->x32, ld, D, p->sqn, N, cvec.p, mind.p, flagged.p, nflag.p, mx.p, nullptr, nullptr);
                else kpp_filter_kernel<CRX_COSINE><<<gf, 256, 0, c->stream>>>(p->x32, ld, D, p->sqn, N, cvec.p, mind.p, flagged.p, nflag.p, mx.p, nullptr, nullptr);
            }
            CRX_KERNEL(c, "kpp_update");
            int ge = c->sm_count * 2;  // the list is short; its length stays on the device
            if (metric == CRX_EUCLIDEAN) kpp_update_kernel<float, CRX_EUCLIDEAN><<<ge, 256, 0, c->stream>>>(p->x32, ld, D, p->sqn, N, cvec.p, 0, mind.p, mx.p, flagged.p, nflag.p, nullptr, 0);
            else kpp_update_kernel<float, CRX_COSINE><<<ge, 256, 0, c->stream>>>(p->x32, ld, D, p->sqn, N, cvec.p, 0, mind.p, mx.p, flagged.p, nflag.p, nullptr, 0);
        } else if (N > 0) {
            CRX_KERNEL(c, "kpp_update");
#define LAUNCH_K(T, M, xptr) kpp_update_kernel<T, M><<<gridu, 256, 0, c->stream>>>(xptr, ld, D, p->sqn, N, cvec.p, i == 1, mind.p, mx.p, nullptr, nullptr, prune ? nearest.p : nullptr, i - 1)
            if (p->x64) { if (metric == CRX_EUCLIDEAN) LAUNCH_K(double, CRX_EUCLIDEAN, p->x64); else LAUNCH_K(double, CRX_COSINE, p->x64); }
            else { if (metric == CRX_EUCLIDEAN) LAUNCH_K(float, CRX_EUCLIDEAN, p->x32); else LAUNCH_K(float, CRX_COSINE, p->x32); }
#undef LAUNCH_K
        }
        // max_for_normalizing over all shards (initialization.hpp:116-117): positive doubles order like their bit patterns
        if (world > 1) CRX_TRY(comm_allreduce(c, comm, mx.p, 1, CRX_F64, CRX_MAX, CRX_DEVICE));
        if (N > 0) {
            { CRX_KERNEL(c, "kpp_prob"); kpp_prob_kernel<<<crx_grid(N, 256), 256, 0, c->stream>>>(mind.p, N, mx.p, prob.p); }
            CRX_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, bytes, prob.p, P.p, (int)N, c->stream));
        }
        // shard totals -> draw -> owner's binary search -> the picked row and its coordinates, all on the device: the round
        // has no host synchronisation (the draws replicate std::default_random_engine + uniform_real_distribution<double>)
        { CRX_KERNEL(c, "kpp_total"); kpp_total_kernel<<<1, 1, 0, c->stream>>>(P.p, N, st.p + KPP_TOT + me); }
        if (world > 1) CRX_TRY(comm_allgather(c, comm, st.p + KPP_TOT + me, st.p + KPP_ALL, 1, CRX_F64, CRX_DEVICE));
        {
            CRX_KERNEL(c, "kpp_pick");
            if (p->x64) kpp_draw_pick_kernel<double><<<1, 128, 0, c->stream>>>(p->x64, ld, p->sqn, P.p, N, row_offset, st.p, world, me, lens_dev.p, share.p, c->counters);
            else kpp_draw_pick_kernel<float><<<1, 128, 0, c->stream>>>(p->x32, ld, p->sqn, P.p, N, row_offset, st.p, world, me, lens_dev.p, share.p, c->counters);
        }
        // the owner's buffer holds the vector, everybody else's zeros: the sum IS the broadcast, with no root to know on the host
        if (world > 1) CRX_TRY(comm_allreduce(c, comm, share.p, ld + 2, CRX_F64, CRX_SUM, CRX_DEVICE));
        { CRX_KERNEL(c, "kpp_share"); kpp_share_kernel<<<1, 128, 0, c->stream>>>(share.p, ld, cvec.p, vecs_dev.p ? vecs_dev.p + (size_t)i * D : nullptr, D, rows_dev.p + i); }
    }
    CRX_CUDA(cudaGetLastError());
    CRX_CUDA(cudaMemcpyAsync(out, rows_dev.p, (size_t)K * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    if (out_vectors) CRX_CUDA(cudaMemcpyAsync(out_vectors, vecs_dev.p, (size_t)K * D * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

int crx_k_means_pp(crx_ctx* c, const crx_points* p, int K, int metric, uint64_t seed, int32_t* out) {
    CRX_REQUIRE(c && p && out, "NULL argument");
    CRX_REQUIRE(K >= 1, "cluster_num");
    std::vector<int64_t> rows(K);
    CRX_TRY(crx_k_means_pp_sharded(c, p, 0, p->n, K, metric, seed, nullptr, rows.data(), nullptr));
    for (int i = 0; i < K; i++) out[i] = (int32_t)rows[i];
    return CRX_OK;
}

__global__ void counts_to_double_kernel(long long* __restrict__ cnt, int K, double* __restrict__ d, int back) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    if (back) cnt[k] = (long long)d[k]; else d[k] = (double)cnt[k];
}

int crx_k_means_sharded(crx_ctx* c, const crx_points* p, const int32_t* labels, int lmem, const double* oldc, int K, int metric,
                        double min_dist, const crx_comm* comm, double* newc, int cmem, int* cont) {
    CRX_REQUIRE(c && p && labels && oldc && newc && cont, "NULL argument");
    DevBuf<double> sums, o, n;
    DevBuf<long long> cnt;
    size_t kd = (size_t)K * p->d;
    CRX_TRY(sums.alloc(c, kd + K)); CRX_TRY(cnt.alloc(c, K)); CRX_TRY(o.alloc(c, kd)); CRX_TRY(n.alloc(c, kd));
    CRX_CUDA(cudaMemcpyAsync(o.p, oldc, kd * sizeof(double), cmem == CRX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
    CRX_TRY(crx_cluster_sums(c, p, labels, lmem, K, sums.p, (int64_t*)cnt.p, CRX_DEVICE));
    if (comm_world(comm) > 1) {
        // ONE all-reduce per iteration: the K member counts ride behind the K*D sums as doubles (exact below 2^53)
        { CRX_KERNEL(c, "pack_counts"); counts_to_double_kernel<<<crx_grid(K, 256), 256, 0, c->stream>>>(cnt.p, K, sums.p + kd, 0); }
        CRX_TRY(comm_allreduce(c, comm, sums.p, (int64_t)(kd + K), CRX_F64, CRX_SUM, CRX_DEVICE));
        { CRX_KERNEL(c, "pack_counts"); counts_to_double_kernel<<<crx_grid(K, 256), 256, 0, c->stream>>>(cnt.p, K, sums.p + kd, 1); }
    }
    CRX_TRY(crx_k_means_finish(c, sums.p, (const int64_t*)cnt.p, o.p, K, p->d, metric, min_dist, n.p, CRX_DEVICE, cont));
    CRX_CUDA(cudaMemcpyAsync(newc, n.p, kd * sizeof(double), cmem == CRX_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}

int crx_lloyds_assignment(crx_ctx* c, const crx_points* p, const double* centroids, int cmem, int K, const int32_t* crow,
                          int metric, int32_t* labels, double* dists, int mem) {
    CRX_REQUIRE(c && p && centroids && labels, "NULL argument");
    CRX_REQUIRE(K >= 1, "K");
    CRX_REQUIRE(metric == CRX_EUCLIDEAN || metric == CRX_COSINE, "metric");
    CRX_CUDA(cudaSetDevice(c->device));
    Centroids cen;
    CRX_TRY(cen.stage(c, centroids, cmem, K, p->d, p->ld));
    IoBuf<int32_t> lab;
    IoBuf<double> dis;
    CRX_TRY(lab.bind(c, labels, p->n, mem, false));
    CRX_TRY(dis.bind(c, dists, p->n, mem, false));   // dists == NULL: labels only
    if (metric == CRX_EUCLIDEAN && K >= 32 && p->n >= 1024 && p->d <= 128 && !tc_disabled()) CRX_TRY(lloyd_scan_tc(c, p, cen, lab.dev, dis.dev));
    else {
        DevBuf<double> scratch;
        if (!dis.dev) CRX_TRY(scratch.alloc(c, p->n));
        CRX_TRY(lloyd_scan(c, p, nullptr, p->n, cen, metric, lab.dev, dis.dev ? dis.dev : scratch.p));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (crow) {
        DevBuf<int32_t> d_crow;
        CRX_TRY(d_crow.alloc(c, K));
        for (int i = 0; i < K; i++) CRX_REQUIRE(crow[i] >= -1 && crow[i] < p->n, "centroid_rows out of range");
        CRX_CUDA(cudaMemcpyAsync(d_crow.p, crow, K * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CRX_KERNEL(c, "self_assign");
        self_assign_kernel<<<1, 32, 0, c->stream>>>(d_crow.p, K, lab.dev, dis.dev);
        CRX_CUDA(cudaStreamSynchronize(c->stream));
    }
    CRX_CUDA(cudaGetLastError());
    CRX_TRY(lab.flush());
    CRX_TRY(dis.flush());
    if (mem == CRX_HOST || cmem == CRX_HOST) CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}

int crx_lloyds_for_remaining(crx_ctx* c, const crx_points* p, const double* centroids, int cmem, int K, int metric,
                             int32_t* labels, double* dists, int mem) {
    CRX_REQUIRE(c && p && centroids && labels && dists, "NULL argument");
    CRX_CUDA(cudaSetDevice(c->device));
    Centroids cen;
    CRX_TRY(cen.stage(c, centroids, cmem, K, p->d, p->ld));
    IoBuf<int32_t> lab;
    IoBuf<double> dis;
    CRX_TRY(lab.bind(c, labels, p->n, mem, true));
    CRX_TRY(dis.bind(c, dists, p->n, mem, true));
    DevBuf<int32_t> rows;
    DevBuf<int> cnt;
    CRX_TRY(rows.alloc(c, p->n)); CRX_TRY(cnt.alloc(c, 1));
    CRX_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int), c->stream));
    { CRX_KERNEL(c, "compact_unassigned"); compact_unassigned_kernel<<<crx_grid(p->n, 256), 256, 0, c->stream>>>(lab.dev, p->n, rows.p, cnt.p); }
    int h_cnt = 0;
    CRX_CUDA(cudaMemcpyAsync(&h_cnt, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    CRX_TRY(lloyd_scan(c, p, rows.p, h_cnt, cen, metric, lab.dev, dis.dev));
    CRX_TRY(lab.flush());
    CRX_TRY(dis.flush());
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}

// the K stored rows `d_crow` as an FP64 point set of their own
static int centroid_set_from_rows(crx_ctx* c, const crx_points* p, const int32_t* d_crow, int K, crx_points** out) {
    DevBuf<double> cm;
    CRX_TRY(cm.alloc(c, (size_t)K * p->d));
    CRX_TRY(gather_rows(c, p, d_crow, K, cm.p));
    return crx_points_create(c, cm.p, CRX_F64, K, p->d, CRX_DEVICE, out);
}
struct PointsGuard {
    crx_points* p = nullptr;
    ~PointsGuard() { if (p) crx_points_destroy(p); }
};

// d_cb: [L][K] bucket of every centroid in every table (get_LSH_combined_buckets -> getBucketFor, unfiltered)
static int lsh_range_with_buckets(crx_ctx* c, const crx_points* p, const crx_lsh* t, const int32_t* h_crow, const crx_points* cp,
                                  const int32_t* d_cb, bool shared_ids, int K, int metric, const crx_comm* comm, int32_t* labels,
                                  double* dists, int mem, int32_t* before) {
    int L = t->L;
    std::vector<int32_t> cb((size_t)K * L);
    CRX_CUDA(cudaMemcpyAsync(cb.data(), d_cb, cb.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    std::vector<std::vector<int32_t>> offs(L);
    for (int l = 0; l < L; l++) {
        offs[l].resize((size_t)t->nbuckets + 1);
        CRX_CUDA(cudaMemcpyAsync(offs[l].data(), t->by_bucket[l].off, offs[l].size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<int32_t> hb((size_t)K * L), he((size_t)K * L);
    std::vector<const int32_t*> perms(L);
    for (int l = 0; l < L; l++) perms[l] = t->by_bucket[l].perm;
    for (int cc = 0; cc < K; cc++)
        for (int l = 0; l < L; l++) {
            int b = cb[(size_t)l * K + cc];
            hb[(size_t)cc * L + l] = offs[l][b];
            he[(size_t)cc * L + l] = offs[l][b + 1];
        }
    return range_assign_common(c, p, h_crow, cp, d_cb, shared_ids, K, metric, L, hb, he, perms, t->bucket, labels, dists, mem, before, comm);
}

int crx_lsh_range_assignment_sharded(crx_ctx* c, const crx_points* p, const crx_lsh* t, const int32_t* crow, int K, int metric,
                                     const crx_comm* comm, int32_t* labels, double* dists, int mem, int32_t* before) {
    CRX_REQUIRE(c && p && t && crow && labels && dists, "NULL argument");
    CRX_REQUIRE(t->pts == p, "the tables were built over a different point set");
    CRX_REQUIRE(K >= 1, "K");
    for (int i = 0; i < K; i++) CRX_REQUIRE(crow[i] >= 0 && crow[i] < p->n, "centroid_rows must be stored rows");
    CRX_CUDA(cudaSetDevice(c->device));
    int L = t->L;
    int64_t N = p->n;
    DevBuf<int32_t> d_crow, d_cb;
    CRX_TRY(d_crow.alloc(c, K)); CRX_TRY(d_cb.alloc(c, (size_t)K * L));
    CRX_CUDA(cudaMemcpyAsync(d_crow.p, crow, K * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    for (int l = 0; l < L; l++) {
        CRX_KERNEL(c, "gather_int");
        gather_int_kernel<<<crx_grid(K, 128), 128, 0, c->stream>>>(t->bucket + (size_t)l * N, d_crow.p, K, d_cb.p + (size_t)l * K);
    }
    PointsGuard cp;
    CRX_TRY(centroid_set_from_rows(c, p, d_crow.p, K, &cp.p));
    return lsh_range_with_buckets(c, p, t, crow, cp.p, d_cb.p, false, K, metric, comm, labels, dists, mem, before);
}

int crx_lsh_range_assignment(crx_ctx* c, const crx_points* p, const crx_lsh* t, const int32_t* crow, int K, int metric,
                             int32_t* labels, double* dists, int mem, int32_t* before) {
    return crx_lsh_range_assignment_sharded(c, p, t, crow, K, metric, nullptr, labels, dists, mem, before);
}

int crx_lsh_range_assignment_vectors(crx_ctx* c, const crx_points* p, const crx_lsh* t, const double* centroids, const int32_t* crow,
                                     int K, int metric, int shared_ids, int32_t* labels, double* dists, int mem, int32_t* before) {
    CRX_REQUIRE(c && p && t && centroids && labels && dists, "NULL argument");
    CRX_REQUIRE(t->pts == p, "the tables were built over a different point set");
    CRX_REQUIRE(K >= 1, "K");
    std::vector<int32_t> rows((size_t)K, -1);
    if (crow) for (int i = 0; i < K; i++) { CRX_REQUIRE(crow[i] >= -1 && crow[i] < p->n, "centroid_rows out of range"); rows[i] = crow[i]; }
    CRX_CUDA(cudaSetDevice(c->device));
    PointsGuard cp;
    CRX_TRY(crx_points_create(c, centroids, CRX_F64, K, p->d, CRX_HOST, &cp.p));
    // CustHashtable::getHash of every centroid vector (cust_hashtable.hpp:123), all tables in one pass
    DevBuf<int32_t> d_cb;
    CRX_TRY(d_cb.alloc(c, (size_t)K * t->L));
    CRX_TRY(crx_hash_rows(c, cp.p, t->metric, t->k, t->L, t->d_proj, t->ldp, t->d_pnorm, t->d_t, t->d_r, t->w, t->nbuckets, nullptr, d_cb.p));
    return lsh_range_with_buckets(c, p, t, rows.data(), cp.p, d_cb.p, shared_ids != 0, K, metric, nullptr, labels, dists, mem, before);
}

int crx_cube_range_assignment_sharded(crx_ctx* c, const crx_points* p, const crx_cube* cu, const int32_t* crow, int K, int metric,
                                      int probes, const crx_comm* comm, int32_t* labels, double* dists, int mem, int32_t* before) {
    CRX_REQUIRE(c && p && cu && crow && labels && dists, "NULL argument");
    CRX_REQUIRE(cu->pts == p, "the hypercube was built over a different point set");
    CRX_REQUIRE(K >= 1, "K");
    for (int i = 0; i < K; i++) CRX_REQUIRE(crow[i] >= 0 && crow[i] < p->n, "centroid_rows must be stored rows");
    CRX_CUDA(cudaSetDevice(c->device));
    DevBuf<int32_t> d_crow, d_home;
    CRX_TRY(d_crow.alloc(c, K)); CRX_TRY(d_home.alloc(c, K));
    CRX_CUDA(cudaMemcpyAsync(d_crow.p, crow, K * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    { CRX_KERNEL(c, "gather_int"); gather_int_kernel<<<crx_grid(K, 128), 128, 0, c->stream>>>(cu->vertex, d_crow.p, K, d_home.p); }
    std::vector<int32_t> home(K), off((size_t)(1 << cu->k) + 1);
    CRX_CUDA(cudaMemcpyAsync(home.data(), d_home.p, K * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaMemcpyAsync(off.data(), cu->by_vertex.off, off.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    // probe sequences (lsh_cube.hpp:140-177); every centroid gets the same number of segment slots
    std::vector<std::vector<int>> seqs(K);
    size_t nseg = 1;
    for (int cc = 0; cc < K; cc++) { crx_cube_probe_sequence(home[cc], probes, cu->k, seqs[cc]); nseg = std::max(nseg, seqs[cc].size()); }
    std::vector<int32_t> hb((size_t)K * nseg, 0), he((size_t)K * nseg, 0);
    for (int cc = 0; cc < K; cc++)
        for (size_t s = 0; s < seqs[cc].size(); s++) { hb[cc * nseg + s] = off[seqs[cc][s]]; he[cc * nseg + s] = off[seqs[cc][s] + 1]; }
    std::vector<const int32_t*> perms(1, cu->by_vertex.perm);
    PointsGuard cp;
    CRX_TRY(centroid_set_from_rows(c, p, d_crow.p, K, &cp.p));
    return range_assign_common(c, p, crow, cp.p, nullptr, false, K, metric, (int)nseg, hb, he, perms, nullptr, labels, dists, mem, before, comm);
}

int crx_cube_range_assignment(crx_ctx* c, const crx_points* p, const crx_cube* cu, const int32_t* crow, int K, int metric,
                              int probes, int32_t* labels, double* dists, int mem, int32_t* before) {
    return crx_cube_range_assignment_sharded(c, p, cu, crow, K, metric, probes, nullptr, labels, dists, mem, before);
}

int crx_cluster_sums(crx_ctx* c, const crx_points* p, const int32_t* labels, int lmem, int K, double* sums, int64_t* counts, int mem) {
    CRX_REQUIRE(c && p && labels && sums && counts, "NULL argument");
    CRX_REQUIRE(K >= 1, "K");
    CRX_CUDA(cudaSetDevice(c->device));
    int64_t N = p->n;
    int D = p->d;
    IoBuf<int32_t> lab;
    IoBuf<double> out;
    IoBuf<long long> cnt;
    CRX_TRY(lab.bind(c, labels, N, lmem, true));
    CRX_TRY(out.bind(c, sums, (size_t)K * D, mem, false));
    CRX_TRY(cnt.bind(c, (long long*)counts, K, mem, false));
    Segments seg;
    SegmentsGuard seg_guard(seg);   // freed on every return below
    int st = crx_build_segments(c, lab.dev, N, K, &seg);
    if (st != CRX_OK) return st;
    std::vector<int32_t> off(K + 1);
    CRX_CUDA(cudaMemcpyAsync(off.data(), seg.off, (K + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    // A cluster of up to `whole_max` members is summed as ONE sequential chain = the reference's own sum, bit for bit:
    // 16384 always (a chain of that length costs ~2 ms at most), 32768 when there are at least two clusters per SM to
    // keep the GPU busy.  Longer clusters are cut into 1024-row chunks (deterministic, ~1e-15 relative).
    int maxc = 0, nonempty = 0;
    for (int cl = 0; cl < K; cl++) { int n = off[cl + 1] - off[cl]; maxc = std::max(maxc, n); nonempty += n > 0; }
    int whole_max = SUM_CH;
    if (nonempty >= 2 * c->sm_count) while (whole_max < maxc && whole_max < SUM_CH_MAX) whole_max *= 2;
    const int ch_len = SUM_CH_SHORT;
    std::vector<int32_t> ch_cluster, ch_index, ch_first(K + 1, 0);
    for (int cl = 0; cl < K; cl++) {
        int n = off[cl + 1] - off[cl];
        int len = n <= whole_max ? whole_max : ch_len;
        int nch = (n + len - 1) / len;
        ch_first[cl] = (int32_t)ch_cluster.size();
        for (int i = 0; i < nch; i++) { ch_cluster.push_back(cl); ch_index.push_back(i); }
    }
    ch_first[K] = (int32_t)ch_cluster.size();
    int nchunks = (int)ch_cluster.size();
    const int sums_block = std::max(128, ((D + 31) / 32) * 32);   // thread j owns coordinate j
    DevBuf<int32_t> d_cc, d_ci, d_cf;
    DevBuf<double> partial;
    CRX_TRY(d_cc.alloc(c, nchunks)); CRX_TRY(d_ci.alloc(c, nchunks)); CRX_TRY(d_cf.alloc(c, K + 1));
    CRX_TRY(partial.alloc(c, (size_t)nchunks * D));
    if (nchunks) {
        CRX_CUDA(cudaMemcpyAsync(d_cc.p, ch_cluster.data(), nchunks * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CRX_CUDA(cudaMemcpyAsync(d_ci.p, ch_index.data(), nchunks * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    }
    CRX_CUDA(cudaMemcpyAsync(d_cf.p, ch_first.data(), (K + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    if (nchunks) {
        CRX_KERNEL(c, "chunk_sums");
        if (p->x64) chunk_sums_kernel<double><<<nchunks, sums_block, 0, c->stream>>>(p->x64, p->ld, D, seg.perm, seg.off, d_cc.p, d_ci.p, whole_max, ch_len, partial.p);
        else chunk_sums_kernel<float><<<nchunks, sums_block, 0, c->stream>>>(p->x32, p->ld, D, seg.perm, seg.off, d_cc.p, d_ci.p, whole_max, ch_len, partial.p);
    }
    { CRX_KERNEL(c, "combine_sums"); combine_sums_kernel<<<K, sums_block, 0, c->stream>>>(partial.p, d_cf.p, K, D, seg.off, whole_max, ch_len, out.dev, cnt.dev); }
    CRX_CUDA(cudaGetLastError());
    st = out.flush();
    if (st == CRX_OK) st = cnt.flush();
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return st;
}

int crx_k_means_finish(crx_ctx* c, const double* sums, const int64_t* counts, const double* oldc, int K, int D, int metric,
                       double min_dist, double* newc, int mem, int* cont) {
    CRX_REQUIRE(c && sums && counts && oldc && newc && cont, "NULL argument");
    CRX_CUDA(cudaSetDevice(c->device));
    IoBuf<double> s, o, n;
    IoBuf<long long> cn;
    CRX_TRY(s.bind(c, sums, (size_t)K * D, mem, true));
    CRX_TRY(o.bind(c, oldc, (size_t)K * D, mem, true));
    CRX_TRY(cn.bind(c, (const long long*)counts, K, mem, true));
    CRX_TRY(n.bind(c, newc, (size_t)K * D, mem, false));
    DevBuf<double> cand;
    DevBuf<int> moved;
    CRX_TRY(cand.alloc(c, (size_t)K * D)); CRX_TRY(moved.alloc(c, 1));
    CRX_CUDA(cudaMemsetAsync(moved.p, 0, sizeof(int), c->stream));
    { CRX_KERNEL(c, "kmeans_finish"); kmeans_finish_kernel<<<crx_grid(K, 64), 64, 0, c->stream>>>(s.dev, cn.dev, o.dev, K, D, metric, min_dist, cand.p, moved.p); }
    { CRX_KERNEL(c, "select_centroids"); select_centroids_kernel<<<crx_grid((int64_t)K * D, 256), 256, 0, c->stream>>>(cand.p, o.dev, (size_t)K * D, moved.p, n.dev); }
    CRX_CUDA(cudaGetLastError());
    int h = 0;
    CRX_CUDA(cudaMemcpyAsync(&h, moved.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CRX_TRY(n.flush());
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    *cont = h;
    return CRX_OK;
}

int crx_k_means(crx_ctx* c, const crx_points* p, const int32_t* labels, int lmem, const double* oldc, int K, int metric,
                double min_dist, double* newc, int cmem, int* cont) {
    return crx_k_means_sharded(c, p, labels, lmem, oldc, K, metric, min_dist, nullptr, newc, cmem, cont);
}

// pam_lloyds: row sums with error bars for this rank's share of the candidate rows (tensor engine: operand rows
// permuted into cluster order, one CTA per 128-row tile of a cluster against the cluster's own columns; otherwise exact
// FP64 distances, one warp per candidate), all-reduce, then the pick (identical on every rank)
int crx_pam_lloyds_sharded(crx_ctx* c, const crx_points* p, const int32_t* labels, int lmem, const int32_t* crow, int K, int metric,
                           const crx_comm* comm, int32_t* new_crow, int* swapped) {
    CRX_REQUIRE(c && p && labels && crow && new_crow && swapped, "NULL argument");
    CRX_REQUIRE(metric == CRX_EUCLIDEAN || metric == CRX_COSINE, "metric");
    CRX_NARROW(p);
    CRX_CUDA(cudaSetDevice(c->device));
    const int world = comm_world(comm), me = comm_rank(comm);
    int64_t N = p->n;
    IoBuf<int32_t> lab;
    CRX_TRY(lab.bind(c, labels, N, lmem, true));
    struct SegGuard { Segments s; ~SegGuard() { s.free_all(); } } sg;
    Segments& seg = sg.s;
    CRX_TRY(crx_build_segments(c, lab.dev, N, K, &seg));
    DevBuf<double> rowsum, rowerr, exact;
    DevBuf<int32_t> d_crow, d_new, cand, winner;
    DevBuf<int> d_sw, ncand;
    CRX_TRY(rowsum.alloc(c, N)); CRX_TRY(rowerr.alloc(c, N)); CRX_TRY(d_crow.alloc(c, K)); CRX_TRY(d_new.alloc(c, K));
    CRX_TRY(d_sw.alloc(c, 1)); CRX_TRY(cand.alloc(c, N)); CRX_TRY(winner.alloc(c, K)); CRX_TRY(ncand.alloc(c, 1));
    CRX_CUDA(cudaMemcpyAsync(d_crow.p, crow, K * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CRX_CUDA(cudaMemsetAsync(d_sw.p, 0, sizeof(int), c->stream));
    CRX_CUDA(cudaMemsetAsync(ncand.p, 0, sizeof(int), c->stream));
    if (world > 1) {
        CRX_CUDA(cudaMemsetAsync(rowsum.p, 0, N * sizeof(double), c->stream));
        CRX_CUDA(cudaMemsetAsync(rowerr.p, 0, N * sizeof(double), c->stream));
    }
    double mx = 0;
    if (metric == CRX_EUCLIDEAN && N >= 4096 && p->d <= 128 && !tc_disabled()) CRX_TRY(points_maxabs(c, p, &mx));
    if (mx > 0.0 && std::isfinite(mx)) {
        std::vector<int32_t> off(K + 1);
        CRX_CUDA(cudaMemcpyAsync(off.data(), seg.off, (K + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<int4> jobs, mine;
        for (int cl = 0; cl < K; cl++)
            for (int r = off[cl]; r < off[cl + 1]; r += 128) jobs.push_back(make_int4(r, off[cl + 1], off[cl], off[cl + 1]));
        // longest clusters first: the tail of the grid is then made of short jobs; ranks take them round-robin
        std::stable_sort(jobs.begin(), jobs.end(), [](const int4& a, const int4& b) { return a.w - a.z > b.w - b.z; });
        for (size_t j = (size_t)me; j < jobs.size(); j += world) mine.push_back(jobs[j]);
        DevBuf<int4> d_jobs;
        DevBuf<float> norm_s, errw_s;
        CRX_TRY(d_jobs.alloc(c, mine.size() + 1)); CRX_TRY(norm_s.alloc(c, N)); CRX_TRY(errw_s.alloc(c, N));
        if (!mine.empty()) CRX_CUDA(cudaMemcpyAsync(d_jobs.p, mine.data(), mine.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
        TcOperand op;
        int st = crx_tc_prepare(c, p, 1, (double)scale_for(mx), &op, seg.perm, norm_s.p, errw_s.p);
        // every scaled squared norm is at most D (max|x| 2^scale)^2
        const double nn_max = (double)p->d * ldexp(mx, scale_for(mx)) * ldexp(mx, scale_for(mx)) * 1.0001;
        if (st == CRX_OK) st = crx_tc_rowsum(c, op, d_jobs.p, (int)mine.size(), norm_s.p, errw_s.p, (float)(crx_tc_errw(nn_max, p->d) * 1.0001), rowsum.p, rowerr.p);
        cudaError_t e = cudaStreamSynchronize(c->stream);  // the job list and the operand are done with
        op.free_all();
        if (st != CRX_OK) return st;
        CRX_CUDA(e);
    } else {
        int64_t lo = N * me / world, hi = N * (me + 1) / world;
        int g = (int)((hi - lo + 7) / 8);
        if (g > 0) {
            CRX_KERNEL(c, "pam_rowsum");
#define LAUNCH_P(T, M, xptr) pam_rowsum_kernel<T, M><<<g, 256, 0, c->stream>>>(xptr, p->ld, p->d, p->sqn, seg.perm, seg.sorted, seg.off, lo, hi, rowsum.p, rowerr.p)
            if (p->x64) { if (metric == CRX_EUCLIDEAN) LAUNCH_P(double, CRX_EUCLIDEAN, p->x64); else LAUNCH_P(double, CRX_COSINE, p->x64); }
            else { if (metric == CRX_EUCLIDEAN) LAUNCH_P(float, CRX_EUCLIDEAN, p->x32); else LAUNCH_P(float, CRX_COSINE, p->x32); }
#undef LAUNCH_P
        }
    }
    CRX_TRY(comm_allreduce(c, comm, rowsum.p, N, CRX_F64, CRX_SUM, CRX_DEVICE));
    CRX_TRY(comm_allreduce(c, comm, rowerr.p, N, CRX_F64, CRX_SUM, CRX_DEVICE));
    {
        CRX_KERNEL(c, "pam_bounds");
        pam_bounds_kernel<<<K, 256, 0, c->stream>>>(seg.off, rowsum.p, rowerr.p, p->d, cand.p, ncand.p, winner.p);
    }
    int h = 0;
    CRX_CUDA(cudaMemcpyAsync(&h, ncand.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    CRX_TRY(exact.alloc(c, std::max(h, 1)));
    if (h > 0) {
        CRX_KERNEL(c, "pam_exact");
#define LAUNCH_E(T, M, xptr) pam_exact_kernel<T, M><<<h, 256, 0, c->stream>>>(xptr, p->ld, p->d, p->sqn, seg.perm, seg.sorted, seg.off, cand.p, h, exact.p, c->counters)
        if (p->x64) { if (metric == CRX_EUCLIDEAN) LAUNCH_E(double, CRX_EUCLIDEAN, p->x64); else LAUNCH_E(double, CRX_COSINE, p->x64); }
        else { if (metric == CRX_EUCLIDEAN) LAUNCH_E(float, CRX_EUCLIDEAN, p->x32); else LAUNCH_E(float, CRX_COSINE, p->x32); }
#undef LAUNCH_E
    }
    {
        CRX_KERNEL(c, "pam_final");
        pam_final_kernel<<<crx_grid(K, 128), 128, 0, c->stream>>>(seg.perm, seg.sorted, seg.off, cand.p, h, exact.p, winner.p, d_crow.p, K, d_new.p, d_sw.p);
    }
    CRX_CUDA(cudaGetLastError());
    int hs = 0;
    CRX_CUDA(cudaMemcpyAsync(new_crow, d_new.p, K * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaMemcpyAsync(&hs, d_sw.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    *swapped = hs;
    return CRX_OK;
}

int crx_pam_lloyds(crx_ctx* c, const crx_points* p, const int32_t* labels, int lmem, const int32_t* crow, int K, int metric,
                   int32_t* new_crow, int* swapped) {
    return crx_pam_lloyds_sharded(c, p, labels, lmem, crow, K, metric, nullptr, new_crow, swapped);
}

int crx_silhouette_cluster_sharded(crx_ctx* c, const crx_points* p, const int32_t* labels, int lmem, const double* centroids,
                                   int cmem, int K, int metric, const crx_comm* comm, double* sils) {
    CRX_REQUIRE(c && p && labels && centroids && sils, "NULL argument");
    CRX_NARROW(p);
    CRX_CUDA(cudaSetDevice(c->device));
    const int world = comm_world(comm), me = comm_rank(comm);
    int64_t N = p->n;
    IoBuf<int32_t> lab;
    CRX_TRY(lab.bind(c, labels, N, lmem, true));
    Centroids cen;
    CRX_TRY(cen.stage(c, centroids, cmem, K, p->d, p->ld));
    struct SegGuard { Segments s; ~SegGuard() { s.free_all(); } } sg;
    Segments& seg = sg.s;
    CRX_TRY(crx_build_segments(c, lab.dev, N, K, &seg));
    DevBuf<int32_t> near;
    DevBuf<double> s, d_sils, raw;
    CRX_TRY(near.alloc(c, K)); CRX_TRY(s.alloc(c, N)); CRX_TRY(d_sils.alloc(c, K)); CRX_TRY(raw.alloc(c, K));
    if (world > 1) CRX_CUDA(cudaMemsetAsync(s.p, 0, N * sizeof(double), c->stream));
    { CRX_KERNEL(c, "near_centroid"); near_centroid_kernel<<<crx_grid(K, 64), 64, 0, c->stream>>>(cen.pad.p, cen.sqn.p, K, p->ld, p->d, metric, near.p); }
    // this rank's share of the points (positions in cluster order); the others contribute zeros to the all-reduce
    int64_t lo = N * me / world, hi = N * (me + 1) / world;
    int g = (int)((hi - lo + 7) / 8);
    if (g > 0) {
        CRX_KERNEL(c, "silhouette_point");
#define LAUNCH_S(T, M, xptr) silhouette_point_kernel<T, M><<<g, 256, 0, c->stream>>>(xptr, p->ld, p->d, p->sqn, seg.perm, seg.sorted, seg.off, near.p, lo, hi, s.p)
        if (p->x64) { if (metric == CRX_EUCLIDEAN) LAUNCH_S(double, CRX_EUCLIDEAN, p->x64); else LAUNCH_S(double, CRX_COSINE, p->x64); }
        else { if (metric == CRX_EUCLIDEAN) LAUNCH_S(float, CRX_EUCLIDEAN, p->x32); else LAUNCH_S(float, CRX_COSINE, p->x32); }
#undef LAUNCH_S
    }
    CRX_TRY(comm_allreduce(c, comm, s.p, N, CRX_F64, CRX_SUM, CRX_DEVICE));
    { CRX_KERNEL(c, "silhouette_reduce"); silhouette_reduce_kernel<<<crx_grid(K, 64), 64, 0, c->stream>>>(s.p, seg.off, K, d_sils.p, raw.p); }
    CRX_CUDA(cudaGetLastError());
    std::vector<double> h_raw(K);
    CRX_CUDA(cudaMemcpyAsync(sils, d_sils.p, K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaMemcpyAsync(h_raw.data(), raw.p, K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    double total = 0;  // sils[K] accumulates the per-cluster sums in cluster order (silhouette.hpp:78)
    for (int cl = 0; cl < K; cl++) total = total + h_raw[cl];
    sils[K] = total / (double)N;
    return CRX_OK;
}

int crx_silhouette_cluster(crx_ctx* c, const crx_points* p, const int32_t* labels, int lmem, const double* centroids,
                           int cmem, int K, int metric, double* sils) {
    return crx_silhouette_cluster_sharded(c, p, labels, lmem, centroids, cmem, K, metric, nullptr, sils);
}

} // extern "C"
