// rowwalk.cuh -- "each lane owns one row and walks it in index order" against one vector held in
// shared memory.  Global reads are staged 32 rows x 16 columns at a time through shared memory, so
// that they are coalesced in 64-byte pieces while every lane still performs the reference's scalar
// loop over ITS row (cust_vector.hpp:107-174) -- which is what bit-exact distances require.
#pragma once
#include "common.cuh"

namespace rw {

constexpr int TW = 17;  // padded tile width (doubles)
typedef double WarpTile[32][TW];

// result of a walk: Euclidean => acc = sum (x_k - v_k)^2 exactly as the reference;
// cosine => (h, l) = the reference's x87 inner product x . v (64-bit mantissa, sequentially rounded)
struct Walk {
    double a, b;
};

// myrow < 0 => lane idle (still participates in the staging shuffles).
template <typename T, int METRIC>
__device__ __forceinline__ Walk walk_rows(const T* __restrict__ x, int ld, int d, int64_t myrow, const double* __restrict__ vec,
                                          WarpTile& tile) {
    int lane = threadIdx.x & 31;
    double s = 0.0;
    X87 acc = {0.0, 0.0};
    const int myrow32 = (int)myrow;   // rows are < 2^31 (crx_points_create)
    for (int c0 = 0; c0 < d; c0 += 16) {
        // 16-byte pieces: one load instruction fetches 64 contiguous bytes of 8 rows (fp32) / 128 bytes of 4 rows (fp64);
        // ld is a multiple of 4 and the padding holds zeros, so a piece that starts inside the row is readable
        if constexpr (sizeof(T) == 4) {
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const int r = it * 8 + (lane >> 2), piece = (lane & 3) * 4;
                const int src = __shfl_sync(0xffffffffu, myrow32, r);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src >= 0 && c0 + piece < ld) v = *reinterpret_cast<const float4*>(x + (size_t)src * ld + c0 + piece);
                tile[r][piece] = (double)v.x; tile[r][piece + 1] = (double)v.y; tile[r][piece + 2] = (double)v.z; tile[r][piece + 3] = (double)v.w;
            }
        } else {
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int r = it * 4 + (lane >> 3), piece = (lane & 7) * 2;
                const int src = __shfl_sync(0xffffffffu, myrow32, r);
                double2 v = make_double2(0.0, 0.0);
                if (src >= 0 && c0 + piece < ld) v = *reinterpret_cast<const double2*>(x + (size_t)src * ld + c0 + piece);
                tile[r][piece] = v.x; tile[r][piece + 1] = v.y;
            }
        }
        __syncwarp();
        int lim = min(16, d - c0);
        if (METRIC == CRX_EUCLIDEAN) {
            for (int k = 0; k < lim; k++) {
                double t = __dsub_rn(tile[lane][k], vec[c0 + k]);
                s = __dadd_rn(s, __dmul_rn(t, t));
            }
        } else {
            for (int k = 0; k < lim; k++) x87_add(acc, __dmul_rn(tile[lane][k], vec[c0 + k]));
        }
        __syncwarp();
    }
    Walk w;
    if (METRIC == CRX_EUCLIDEAN) { w.a = s; w.b = 0.0; }
    else { w.a = acc.h; w.b = acc.l; }
    return w;
}

// distance of the lane's row to `vec` in the reference's metric; nrow / nvec = exact sums of squares
template <typename T, int METRIC>
__device__ __forceinline__ double dist_rows(const T* __restrict__ x, int ld, int d, int64_t myrow, const double* __restrict__ vec,
                                            double nrow, double nvec, WarpTile& tile) {
    Walk w = walk_rows<T, METRIC>(x, ld, d, myrow, vec, tile);
    if (METRIC == CRX_EUCLIDEAN) return __dsqrt_rn(w.a);
    X87 ip = {w.a, w.b};
    return __dsub_rn(1.0, cos_sim_x87(ip, nrow, nvec));
}

// stage one row of x into shared memory as doubles (all 32 lanes of the calling warp)
template <typename T>
__device__ __forceinline__ void stage_vector(const T* __restrict__ x, int ld, int64_t row, double* vec) {
    int lane = threadIdx.x & 31;
    for (int k = lane; k < ld; k += 32) vec[k] = (double)x[row * ld + k];
    __syncwarp();
}

}  // namespace rw
