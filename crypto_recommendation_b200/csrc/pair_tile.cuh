// pair_tile.cuh -- the FP64 "all pairs of a row tile x a column tile" engine shared by the Lloyd
// assignment scan, the top-P similarity scan and the PAM row sums.
//
// CTA = 256 threads = 8 warps.  A tile: BM = 64 rows, resident in shared memory for the CTA's
// lifetime, row-major [row][k].  B tile: BN = 64 rows ("columns" of the pair matrix), k-major
// [k][col] so that the 32 lanes of a warp read 64 consecutive doubles.  Warp w owns rows
// 8w..8w+7 of the A tile for ALL columns, lane owns columns 2*lane and 2*lane+1: every lane
// accumulates an 8x2 micro-tile over k in INDEX ORDER, so each pair's accumulator sees exactly the
// sequence of operations of the reference's scalar loops (cust_vector.hpp:107-136).  Because a warp
// owns whole rows, the row reductions (argmin, top-P, sum) need warp shuffles only.
//
//   FORM_DOT        acc = fma(a_k, b_k, acc)                   dot product (cosine, projections)
//   FORM_DIFF_EXACT t = a_k - b_k; acc = acc + t*t  (no FMA)   bit-exact Euclidean accumulation
#pragma once
#include "common.cuh"

namespace pt {

constexpr int BM = 64;
constexpr int BN = 64;
constexpr int NT = 256;
constexpr int RW = 8;  // rows per warp
constexpr int FORM_DOT = 0;
constexpr int FORM_DIFF_EXACT = 1;

__device__ __forceinline__ void ld4(const float* p, double o[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void ld4(const double* p, double o[4]) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// shared memory needed for one A tile + one B tile of leading dimension ld (multiple of 4)
__host__ __device__ inline size_t smem_bytes(int ld) { return (size_t)(BM + BN) * ld * sizeof(double); }

// A tile: rows r0 .. r0+BM-1 of the (optionally permuted) row sequence; positions >= rend give zeros.
// rowmap (nullable): position -> row of x.
template <typename T>
__device__ __forceinline__ void load_a_tile(double* As, const T* __restrict__ x, int ld, const int32_t* __restrict__ rowmap,
                                            int64_t r0, int64_t rend) {
    int q = ld >> 2;
    for (int e = threadIdx.x; e < BM * q; e += NT) {
        int r = e / q, c4 = e - r * q;
        int64_t pos = r0 + r;
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (pos < rend) {
            int64_t row = rowmap ? (int64_t)rowmap[pos] : pos;
            ld4(x + row * ld + c4 * 4, v);
        }
        double* dst = As + r * ld + c4 * 4;
        dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; dst[3] = v[3];
    }
}

// B tile: columns c0 .. c0+BN-1 of the (optionally permuted) column sequence, stored k-major.
template <typename T>
__device__ __forceinline__ void load_b_tile(double* Bs, const T* __restrict__ x, int ld, const int32_t* __restrict__ colmap,
                                            int64_t c0, int64_t cend) {
    int q = ld >> 2;
    for (int e = threadIdx.x; e < BN * q; e += NT) {
        int c4 = e / BN, col = e - c4 * BN;
        int64_t pos = c0 + col;
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (pos < cend) {
            int64_t row = colmap ? (int64_t)colmap[pos] : pos;
            ld4(x + row * ld + c4 * 4, v);
        }
        int k = c4 * 4;
        Bs[(k + 0) * BN + col] = v[0];
        Bs[(k + 1) * BN + col] = v[1];
        Bs[(k + 2) * BN + col] = v[2];
        Bs[(k + 3) * BN + col] = v[3];
    }
}

// acc[r][cc] for rows 8*warp + r, columns 2*lane + cc
template <int FORM>
__device__ __forceinline__ void tile_mac(const double* __restrict__ As, const double* __restrict__ Bs, int ld, int warp,
                                         int lane, double acc[RW][2]) {
#pragma unroll
    for (int r = 0; r < RW; r++) { acc[r][0] = 0.0; acc[r][1] = 0.0; }
    const double* arow = As + (warp * RW) * ld;
    for (int k = 0; k < ld; k += 2) {
        double2 b0 = *reinterpret_cast<const double2*>(Bs + k * BN + 2 * lane);
        double2 b1 = *reinterpret_cast<const double2*>(Bs + (k + 1) * BN + 2 * lane);
#pragma unroll
        for (int r = 0; r < RW; r++) {
            double2 a = *reinterpret_cast<const double2*>(arow + r * ld + k);
            if (FORM == FORM_DOT) {
                acc[r][0] = __fma_rn(a.x, b0.x, acc[r][0]);
                acc[r][1] = __fma_rn(a.x, b0.y, acc[r][1]);
                acc[r][0] = __fma_rn(a.y, b1.x, acc[r][0]);
                acc[r][1] = __fma_rn(a.y, b1.y, acc[r][1]);
            } else {
                double t00 = __dsub_rn(a.x, b0.x), t01 = __dsub_rn(a.x, b0.y);
                acc[r][0] = __dadd_rn(acc[r][0], __dmul_rn(t00, t00));
                acc[r][1] = __dadd_rn(acc[r][1], __dmul_rn(t01, t01));
                double t10 = __dsub_rn(a.y, b1.x), t11 = __dsub_rn(a.y, b1.y);
                acc[r][0] = __dadd_rn(acc[r][0], __dmul_rn(t10, t10));
                acc[r][1] = __dadd_rn(acc[r][1], __dmul_rn(t11, t11));
            }
        }
    }
}

// "is (va, ia) a better nearest-centroid candidate than (vb, ib)" for squared Euclidean sums, with
// the reference's semantics: it compares sqrt(acc) with strict '<' scanning centroids in index
// order, i.e. minimum of (sqrt(acc), index) lexicographically.  sqrt is only evaluated when the two
// sums are within a few ulps of each other.
__device__ __forceinline__ bool euclid_better(double va, int ia, double vb, int ib) {
    if (va == vb) return ia < ib;
    double hi = fmax(va, vb), lo = fmin(va, vb);
    if (lo >= hi * (1.0 - 1e-15)) {
        double sa = __dsqrt_rn(va), sb = __dsqrt_rn(vb);
        if (sa == sb) return ia < ib;
        return sa < sb;
    }
    return va < vb;
}
// generic (value, index) lexicographic minimum; NaN never wins
__device__ __forceinline__ bool plain_better(double va, int ia, double vb, int ib) {
    if (va != va) return (vb != vb) ? ia < ib : false;
    if (vb != vb) return true;
    if (va < vb) return true;
    if (va == vb) return ia < ib;
    return false;
}

}  // namespace pt
