// recommend_pass2.cuh -- second, targeted pass of the batched top-P (included by recommend.cu).
//
// get_P_closest (crypto_rec.hpp:213-232) sorts ALL candidates of a user -- in row order (lsh_cube.hpp:96-104) -- with the
// Lomuto quicksort of crypto_rec.hpp:235-277 and keeps the first P.  A partition keeps the ">= pivot" elements in their order,
// so while that side still holds P elements only it is consumed: reading the similarities from the last row backwards, every
// weak running maximum (r_j, v_j) is a pivot that merely drops what is below it, and after pivot j the sequence is
//     S_j = { rows < r_j with similarity >= v_j }      (in row order).
// The first P outputs of the literal sort of ANY S_j that is still in front of the first pivot with fewer than P elements on
// its ">=" side equal the reference's.  With H = the P best (ties at the P-th place included) and e* = its last row, the last
// such state is R = { rows < r' with similarity >= t' }, t' = the best similarity behind e*, r' = the first row behind e* that
// reaches it (when e* itself is a minimum of H, H alone decides: DESIGN.md section 3).
//
// rec_finalize resolves a query when R lies inside its 64-entry list.  The others (R larger than the list: the chance is
// ~P/64 per tied query; or a plateau of equal similarities larger than the list) come here:
//   collect   every candidate whose filter score reaches a per-query threshold theta (tcgen05 threshold scan over the
//             compacted queries, tc_scan.cu MODE_COLLECT; exact FP64 SIMT scan for tables the tensor path does not take),
//   evaluate  the reference's own similarity (x87 accumulation) of every collected candidate,
//   resolve   one warp per query: P-th best T_P, H, e*, (t', r') from the collected tail, R, and the literal quicksort on R
//             (warp-parallel form of the Lomuto partition, below); a query whose collected set does not yet decide it is
//             queued again with a lower threshold / an "everything behind e*" column bound.
// Rounds repeat until the queue is empty: every query ends with the reference's list.
#pragma once

struct P2Queue {
    int32_t* q;        // query row relative to q_begin
    double* theta;     // collect candidates with exact similarity >= theta + 2 eps  (filter score >= theta)
    int32_t* colx;     // ... and every candidate in a row > colx
    int32_t* tries;
    unsigned int* count;
};

__device__ __forceinline__ void p2_emit(const P2Queue& w, int q, double theta, int colx, int tries) {
    unsigned int i = atomicAdd(w.count, 1u);
    w.q[i] = q; w.theta[i] = theta; w.colx[i] = colx; w.tries[i] = tries;
}

constexpr int P2_MAXP = 64;   // neighbours per query the batched path can return

// The `keep` (<= 64) largest values seen so far, sorted descending over the warp: slot s lives in lane s (lo) for s < 32 and
// in lane s - 32 (hi) beyond.  thr = the keep-th largest once `keep` values are in (else -inf).
struct WarpTop {
    double lo = -INFINITY, hi = -INFINITY, thr = -INFINITY;
    int filled = 0;
    __device__ __forceinline__ double slot(int s) const {   // every lane receives the value of slot s
        const double a = __shfl_sync(0xffffffffu, lo, s & 31), b = __shfl_sync(0xffffffffu, hi, s & 31);
        return s < 32 ? a : b;
    }
    // every lane offers k (taken when `in`); all 32 lanes call together
    __device__ __forceinline__ void offer(double k, bool in, int keep) {
        const int lane = threadIdx.x & 31;
        unsigned want = __ballot_sync(0xffffffffu, in && k == k && (k > thr || filled < keep));
        while (want) {
            const int src = __ffs(want) - 1;
            want &= want - 1;
            const double x = __shfl_sync(0xffffffffu, k, src);
            if (x > thr || filled < keep) {
                const int pos = __popc(__ballot_sync(0xffffffffu, lo >= x)) + __popc(__ballot_sync(0xffffffffu, hi >= x));
                const double up_lo = __shfl_up_sync(0xffffffffu, lo, 1), up_hi = __shfl_up_sync(0xffffffffu, hi, 1);
                const double carry = __shfl_sync(0xffffffffu, lo, 31);   // old slot 31 moves to slot 32 when the shift crosses it
                if (pos < 32) {
                    lo = lane < pos ? lo : (lane == pos ? x : up_lo);
                    hi = lane == 0 ? carry : up_hi;
                } else {
                    const int ph = pos - 32;
                    hi = lane < ph ? hi : (lane == ph ? x : up_hi);
                }
                if (filled < keep) filled++;
                thr = filled < keep ? -INFINITY : slot(keep - 1);
            }
        }
    }
};

struct P2Blocks {
    const float* blockmax;   // [nq][2][TC_NBLK] filter units, or NULL
    const uint32_t* qcode;   // packed table codes of the query / base rows and the field masks (tail scan of rec_finalize), or NULL
    const uint32_t* ccode;
    uint32_t low, high;
    int64_t nb;              // rows of the table
    int nblk;
    int tile_cols;           // columns per tile of the scan that produced blockmax
    int bt[TC_NBLK + 1];
    double unscale;          // filter units -> similarity
};

// threshold that certainly reaches the best candidate behind row `estar`: the largest filter score of the whole column blocks
// behind it (an upper bound that a candidate normally attains), never above T_P.  No whole block behind e*: colx = e*.
__device__ __forceinline__ double p2_tail_theta(const P2Blocks& b, int64_t qrel, int estar, double TP, double eps, double theta_old, int& colx) {
    colx = 0x7fffffff;
    double cap = fmin(TP - 2.02 * eps, theta_old);
    if (b.blockmax != nullptr) {
        const int tile = estar / b.tile_cols;
        float bm = -INFINITY;
        bool any = false;
        for (int j = 1; j < b.nblk; j++)
            if (b.bt[j] > tile) {
                any = true;
                bm = fmaxf(bm, fmaxf(b.blockmax[(qrel * 2 + 0) * TC_NBLK + j], b.blockmax[(qrel * 2 + 1) * TC_NBLK + j]));
            }
        if (any && bm > -INFINITY) {
            double th = (double)bm * b.unscale - 2.02 * eps;
            if (th < theta_old - 0.5 * eps || theta_old == INFINITY) return fmin(th, cap);   // new information
        }
    }
    colx = estar;
    return cap;
}

// ------------------------------------------------------------------------------------------------
// Warp-parallel literal Lomuto quicksort (crypto_rec.hpp:235-277), first `need` positions final, arrays of any length in
// shared or global memory.  One partition of [lo, hi], pivot = key[hi], cnt = #{>= pivot in front}, p = lo + cnt:
//   * the ">=" elements land in [lo, p) in their order, the pivot at p;
//   * the "<" elements: with q1 = the first of them and rho = #{>= behind q1} + 1, the swaps of the literal loop rotate the
//     "<" block once per ">=" element met behind q1 and once more for the pivot.  In closed form: every "<" element at a
//     position >= q1 + rho STAYS where it is, and the holes left there by ">=" elements (and the pivot's slot hi) receive the
//     "<" elements of [q1, q1 + rho): hole number r (r-th ">=" element behind q1; the pivot slot is number rho) receives
//     the content of position q1 + r - 1, which, when that position holds the r''-th ">=" element itself, is the content of
//     hole r''.  (tests/test_qs_model.py checks this form against the literal loop.)
//   * ranges that start at or beyond `need` are never consumed: when p + 1 >= need only the ">=" side is produced;
//   * a pivot that is a minimum of its range stays in place together with all equal elements behind the last larger one.
// `need` <= 126.  posge / hk / hv: per-warp scratch of 128 entries each.
// ------------------------------------------------------------------------------------------------
// A warp that walks a list of 10^5..10^6 entries alone is bound by the latency of its loads (the lists of one round are GBs:
// every step misses L2).  Once per 128-entry step the lanes ask for the lines eight steps ahead: 32 x 128 B of keys (four
// steps' worth) and 32 x 128 B of rows (eight steps' worth) on their way to L2.
__device__ __forceinline__ void p2_prefetch_ahead(const double* key, const int* val, int base0, int end, int lane) {
    const int ek = base0 + 1024 + lane * 16, ev = base0 + 1024 + lane * 32;
    if (ek < end) asm volatile("prefetch.global.L2 [%0];" ::"l"(key + ek));
    if (val && ev < end) asm volatile("prefetch.global.L2 [%0];" ::"l"(val + ev));
}

// per-warp scratch of warp_qs_topn_big (shared memory): posge/hk/hv [128] for the literal steps, sk/sv [128] = the head of the
// range as it was before a partition step, gp [128] = where the first 128 ">=" elements of the step came from
struct QsScratch { int* posge; double* hk; int* hv; double* sk; int* sv; int* gp; };

__device__ __noinline__ void warp_qs_topn_big(double* key, int* val, int n, int need, const QsScratch S) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int* posge = S.posge; double* hk = S.hk; int* hv = S.hv;
    int st_lo[128], st_hi[128];
    int sp = 1;
    st_lo[0] = 0; st_hi[0] = n - 1;
    while (sp > 0) {
        sp--;
        int lo = st_lo[sp], hi = st_hi[sp];
        while (lo < hi && lo < need) {
            const double pivot = key[hi];
            const int pv = val[hi];
            const int m = hi - lo;
            // ONE pass per partition step: the ">=" elements are counted AND moved to the front as they come (stable, in place: a
            // destination is never behind its source, and a step of 128 entries is in registers before anything is written).
            // Nothing moves when every element is ">=" (the two cases that end or shorten the range without a partition).  A
            // step that turns out to be literal needs the "<" elements the compaction has overwritten: they lie in the head of
            // the range (fewer than `need` <= 126 elements moved), which is saved first.
            for (int j = lane; j < min(m, 128); j += 32) { S.sk[j] = key[lo + j]; S.sv[j] = val[lo + j]; }
            __syncwarp();
            int cnt = 0, lastgt = -1, q1 = 0x7fffffff;
            bool alleq = true;
            for (int base0 = lo; base0 < hi; base0 += 128) {   // four groups of 32 per step: enough loads in flight on long lists
                double k4[4];
                int v4[4];
                bool g4[4];
                p2_prefetch_ahead(key, val, base0, hi, lane);
#pragma unroll
                for (int u = 0; u < 4; u++) { const int e = base0 + u * 32 + lane; k4[u] = e < hi ? key[e] : 0.0; g4[u] = e < hi && k4[u] >= pivot; }
#pragma unroll
                for (int u = 0; u < 4; u++) { const int e = base0 + u * 32 + lane; v4[u] = g4[u] ? val[e] : 0; }
                __syncwarp();
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int base = base0 + u * 32;
                    if (base >= hi) break;
                    const int e = base + lane;
                    const bool in = e < hi;
                    const double k = k4[u];
                    const unsigned bg = __ballot_sync(0xffffffffu, g4[u]);
                    const unsigned bgt = __ballot_sync(0xffffffffu, in && k > pivot);
                    const unsigned blt = __ballot_sync(0xffffffffu, in && !(k >= pivot));
                    alleq = alleq && __all_sync(0xffffffffu, !in || k == pivot);
                    if (g4[u]) {
                        const int r = cnt + __popc(bg & lt);
                        if (r < 128) S.gp[r] = e;
                        if (lo + r != e) { key[lo + r] = k; val[lo + r] = v4[u]; }
                    }
                    cnt += __popc(bg);
                    if (bgt) lastgt = base + 31 - __clz(bgt);
                    if (blt && q1 == 0x7fffffff) q1 = base + __ffs(blt) - 1;
                }
                __syncwarp();
            }
            if (alleq) break;
            if (cnt == m) { hi = lastgt; continue; }   // the pivot and its equals behind the last larger element stay
            const int p = lo + cnt;
            const bool literal = p + 1 < need && p < hi;
            int rho = 0;
            if (literal) {
                // the ">=" elements behind q1 (the first q1 - lo elements of the range are ">=" and did not move), by rank
                const int lead = q1 - lo;
                rho = cnt - lead + 1;
                for (int r = 1 + lane; r < rho; r += 32) posge[r] = S.gp[lead + r - 1];
                if (lane == 0) posge[rho] = hi;
                __syncwarp();
                if (lane == 0) {   // the elements the literal algorithm leaves at those places (read from the saved head)
                    int pp = 1;
                    for (int r = 1; r <= rho; r++) {
                        const int src = q1 + r - 1;
                        const double ks = S.sk[src - lo];
                        if (!(ks >= pivot)) { hk[r] = ks; hv[r] = S.sv[src - lo]; }
                        else {
                            while (posge[pp] < src) pp++;
                            hk[r] = hk[pp]; hv[r] = hv[pp];
                        }
                    }
                }
                __syncwarp();
            }
            if (lane == 0) { key[p] = pivot; val[p] = pv; }
            if (literal) {
                for (int r = 1 + lane; r <= rho; r += 32)
                    if (posge[r] > p) { key[posge[r]] = hk[r]; val[posge[r]] = hv[r]; }
                __syncwarp();
                if (sp < 128) { st_lo[sp] = p + 1; st_hi[sp] = hi; sp++; }
            }
            __syncwarp();
            hi = p - 1;
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// crypto_rec.hpp:281-345 for one query and its (<= 32) neighbours in s_idx / s_sim: weighted rating prediction of the unknown
// coins (neighbours in the given order, the reference's sequential double operations) and the literal Lomuto top-N over them.
// All 32 lanes of the warp; lane owns coins 4*lane .. 4*lane+3; every neighbour row is read as one 16-byte piece per lane.
// ------------------------------------------------------------------------------------------------
template <typename TB>
__device__ __forceinline__ void predict_and_recommend(const TB* __restrict__ xb, int ldb, const double* __restrict__ mean_b,
                                                      const uint8_t* __restrict__ unk_row, double mq, int D, const int* s_idx,
                                                      const double* s_sim, int keep, int Nrec, double* s_pred_q, int* s_coin_q,
                                                      int32_t* __restrict__ recs_row) {
    const int lane = threadIdx.x & 31;
    double main_sum[4] = {0.0, 0.0, 0.0, 0.0}, abs_sum = 0.0;
    for (int i0 = 0; i0 < keep; i0 += 4) {
        double nv[4][4], nm[4], ns[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int i = i0 + u;
            nv[u][0] = nv[u][1] = nv[u][2] = nv[u][3] = 0.0;
            ns[u] = 0.0; nm[u] = 0.0;
            if (i < keep) {
                int nb = s_idx[i];
                ns[u] = s_sim[i];
                nm[u] = mean_b[nb];
                if (4 * lane < ldb) pt::ld4(xb + (size_t)nb * ldb + 4 * lane, nv[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (i0 + u < keep) {
                abs_sum = __dadd_rn(abs_sum, fabs(ns[u]));
#pragma unroll
                for (int t = 0; t < 4; t++) main_sum[t] = __dadd_rn(main_sum[t], __dmul_rn(ns[u], __dsub_rn(nv[u][t], nm[u])));
            }
        }
    }
    unsigned um[4];
    bool unk[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int j = 4 * lane + t;
        unk[t] = j < D && unk_row[j] != 0;
        um[t] = __ballot_sync(0xffffffffu, unk[t]);
    }
    unsigned lt = (1u << lane) - 1u;
    int before = __popc(um[0] & lt) + __popc(um[1] & lt) + __popc(um[2] & lt) + __popc(um[3] & lt);
    int nu = __popc(um[0]) + __popc(um[1]) + __popc(um[2]) + __popc(um[3]);
#pragma unroll
    for (int t = 0; t < 4; t++) {
        if (unk[t]) {
            s_pred_q[before] = __dadd_rn(__ddiv_rn(main_sum[t], abs_sum), mq);
            s_coin_q[before] = 4 * lane + t;
            before++;
        }
    }
    __syncwarp();
    warp_lomuto_topn(s_pred_q, s_coin_q, nu, Nrec);  // crypto_rec.hpp:320
    for (int j = lane; j < Nrec; j += 32) recs_row[j] = j < nu ? s_coin_q[j] : 0;  // resize(N) pads with coin 0
}

// ------------------------------------------------------------------------------------------------
// round plumbing
// ------------------------------------------------------------------------------------------------
constexpr int P2_KIND_SHIFT = 16;   // tries = attempts | (status the query carried out of rec_finalize << 16)

// theta = -inf asks for EVERY candidate of the query ("no candidate behind e*: R = every candidate"): such a row takes no part
// in the threshold scan (its threshold becomes +inf) -- its list is the table mask itself, written by p2_fill_all_kernel
__global__ void p2_prepare_kernel(P2Queue w, unsigned int n, int64_t q_begin, double scale, float* __restrict__ theta_f,
                                  int32_t* __restrict__ qrow_abs, int32_t* __restrict__ allf, int32_t* __restrict__ colx_eff) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool all = w.theta[i] == -INFINITY;
    double th = w.theta[i] * scale;
    float f = (float)th;
    if (isfinite(f)) f = f - fabsf(f) * 1.2e-7f - 1e-30f;   // never above the double threshold
    theta_f[i] = all ? INFINITY : f;
    colx_eff[i] = all ? 0x7fffffff : w.colx[i];
    allf[i] = all;
    qrow_abs[i] = (int32_t)(q_begin + w.q[i]);
}

__global__ void p2_all_count_kernel(const int32_t* __restrict__ allf, unsigned int n, const int32_t* __restrict__ wq, const int32_t* __restrict__ ncand,
                                    int32_t* __restrict__ count, int32_t* __restrict__ ovf) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && allf[i]) { count[i] = ncand[wq[i]]; ovf[i] = 0; }
}

// every row that shares a bucket with the query, ascending: one block per "all" query, block-wide compaction of the table mask
__global__ void __launch_bounds__(256)
p2_fill_all_kernel(const int32_t* __restrict__ allf, const int32_t* __restrict__ ovf, const int32_t* __restrict__ qrow_abs, const uint32_t* __restrict__ qcode,
                   const uint32_t* __restrict__ ccode, int64_t nb, uint32_t low, uint32_t high, const int64_t* __restrict__ off, int32_t* __restrict__ cols) {
    const unsigned int i = blockIdx.x;
    if (!allf[i] || ovf[i]) return;
    __shared__ int wsum[8];
    __shared__ int running;
    const uint32_t cq = qcode[qrow_abs[i]];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    int32_t* out = cols + off[i];
    for (int64_t c0 = 0; c0 < nb; c0 += 256) {
        const int64_t c = c0 + threadIdx.x;
        bool pass = false;
        if (c < nb) { const uint32_t x = cq ^ ccode[c]; pass = ((x - low) & ~x & high) != 0u; }
        const unsigned bm = __ballot_sync(0xffffffffu, pass);
        if (lane == 0) wsum[warp] = __popc(bm);
        __syncthreads();
        int before = running;
        for (int x = 0; x < warp; x++) before += wsum[x];
        if (pass) out[before + __popc(bm & ((1u << lane) - 1u))] = (int32_t)c;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int x = 0; x < 8; x++) t += wsum[x]; running += t; }
        __syncthreads();
    }
}

__global__ void p2_sizes_kernel(const int32_t* __restrict__ count, const int32_t* __restrict__ ovf, unsigned int n, int64_t* __restrict__ seg) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    seg[i] = (i < n && !(ovf && ovf[i])) ? (((int64_t)count[i] + 31) / 32) * 32 : 0;
}

// pass masks of the threshold scan -> ascending column lists.  cmask[(tile * wpt + word)][rows_pad]: one block takes 32
// consecutive rows; 128 mask words per row are staged at a time (every warp reads whole 128-byte lines of the matrix), then a
// warp walks 4 of the rows: lane = one word of 32, prefix sum of the popcounts, the set bits written in order.
__global__ void __launch_bounds__(256)
p2_expand_kernel(const uint32_t* __restrict__ cmask, int64_t rows_pad, int ntiles, int tile_cols, const int32_t* __restrict__ allf,
                 const int32_t* __restrict__ count, const int32_t* __restrict__ ovf, const int64_t* __restrict__ off, unsigned int n,
                 int32_t* __restrict__ cols) {
    __shared__ uint32_t sm[128][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    const int wpt = tile_cols / 32;
    const int64_t W = (int64_t)ntiles * wpt;
    int run[4] = {0, 0, 0, 0};
    bool live[4];
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int64_t row = row0 + warp * 4 + rr;
        live[rr] = row < (int64_t)n && !ovf[row] && !allf[row] && count[row] > 0;
    }
    const bool block_live = __syncthreads_or(live[0] || live[1] || live[2] || live[3]);
    if (block_live) {
        for (int64_t w0 = 0; w0 < W; w0 += 128) {
#pragma unroll 4
            for (int j = 0; j < 16; j++) {
                const int64_t w = w0 + warp * 16 + j;
                sm[warp * 16 + j][lane] = w < W ? cmask[w * rows_pad + row0 + lane] : 0u;
            }
            __syncthreads();
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                if (!live[rr]) continue;
                const int r = warp * 4 + rr;
                int32_t* out = cols + off[row0 + r];
#pragma unroll
                for (int sstep = 0; sstep < 4; sstep++) {
                    uint32_t bits = sm[sstep * 32 + lane][r];
                    const unsigned any = __ballot_sync(0xffffffffu, bits != 0u);
                    if (!any) continue;
                    int pc = __popc(bits), before = pc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, before, o); if (lane >= o) before += v; }
                    const int tot = __shfl_sync(0xffffffffu, before, 31);
                    before -= pc;
                    const int64_t w = w0 + sstep * 32 + lane;
                    const int colbase = (int)(w / wpt) * tile_cols + (int)(w % wpt) * 32;
                    int pos = run[rr] + before;
                    while (bits != 0u) {
                        const int j = __ffs(bits) - 1;
                        bits &= bits - 1u;
                        out[pos++] = colbase + j;
                    }
                    run[rr] += tot;
                }
            }
            __syncthreads();
        }
    }
    // segments are padded to multiples of 32 entries
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int64_t row = row0 + warp * 4 + rr;
        if (row < (int64_t)n && !ovf[row]) {
            const int total = count[row];
            const int j = total + lane;
            if (j < (total + 31) / 32 * 32) cols[off[row] + j] = -1;
        }
    }
    (void)lt;
}

// plain FP64 filter for tables the tensor path does not take: one block per queued query, columns in order.
// FILL = false: count[i] = number of passing candidates;  FILL = true: their rows, ascending, at cols[off[i] ...]
template <typename TQ, typename TB, bool FILL>
__global__ void __launch_bounds__(256)
p2_collect_simt_kernel(const TQ* __restrict__ xq, int ldq, const double* __restrict__ sqn_q, const TB* __restrict__ xb, int ldb,
                       const double* __restrict__ sqn_b, int D, int64_t nb, int64_t q_begin, P2Queue w, int L,
                       const int32_t* __restrict__ qgid, int64_t qstride, const int32_t* __restrict__ cgid, int64_t cstride,
                       int32_t* __restrict__ count, const int32_t* __restrict__ ovf, const int64_t* __restrict__ off,
                       int32_t* __restrict__ cols) {
    if (FILL && ovf[blockIdx.x]) return;   // waits for the next round (its segment is empty)
    __shared__ double qv[128];
    __shared__ int qg[MAXL];
    __shared__ int wsum[8];
    __shared__ int running;
    const unsigned int i = blockIdx.x;
    const int64_t qrow = q_begin + w.q[i];
    const double theta = w.theta[i];
    const int colx = w.colx[i];
    for (int k = threadIdx.x; k < 128; k += 256) qv[k] = k < D ? (double)xq[qrow * ldq + k] : 0.0;
    if (threadIdx.x < L) qg[threadIdx.x] = qgid[(size_t)threadIdx.x * qstride + qrow];
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const double qinv = 1.0 / sqrt(sqn_q[qrow]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t* out = FILL ? cols + off[i] : nullptr;
    for (int64_t c0 = 0; c0 < nb; c0 += 256) {
        const int64_t c = c0 + threadIdx.x;
        bool pass = false;
        if (c < nb) {
            bool match = false;
            for (int l = 0; l < L; l++) { int g = cgid[(size_t)l * cstride + c]; match |= (g == qg[l] && g >= 0); }
            if (match) {
                double acc = 0.0;
                const TB* row = xb + (size_t)c * ldb;
                for (int k = 0; k < D; k++) acc = fma((double)row[k], qv[k], acc);
                double s = acc * qinv / sqrt(sqn_b[c]);
                pass = s >= theta || c > colx;
            }
        }
        const unsigned bm = __ballot_sync(0xffffffffu, pass);
        if (lane == 0) wsum[warp] = __popc(bm);
        __syncthreads();
        int before = running;
        for (int x = 0; x < warp; x++) before += wsum[x];
        if (FILL && pass) out[before + __popc(bm & ((1u << lane) - 1u))] = (int32_t)c;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int x = 0; x < 8; x++) t += wsum[x]; running += t; }
        __syncthreads();
    }
    if (!FILL && threadIdx.x == 0) count[i] = running;
    if (FILL) for (int j = running + threadIdx.x; j < (running + 31) / 32 * 32; j += 256) out[j] = -1;
}

// uval[row] = c when every coordinate of the row equals c (a user with ONE known coin: the unknown coins hold the user's
// mean, crypto_rec.hpp:118-125, so the whole vector is that rating), NaN otherwise.  One warp per row.
template <typename T>
__global__ void __launch_bounds__(256) uniform_rows_kernel(const T* __restrict__ x, int ld, int D, int64_t n, double* __restrict__ uval) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const T first = x[row * ld];
    bool same = true;
    for (int k = lane; k < D; k += 32) same = same && x[row * ld + k] == first;
    same = __all_sync(0xffffffffu, same);
    if (lane == 0) uval[row] = same ? (double)first : CRX_FROM_BITS(0x7ff8000000000000LL);
}

// the reference's own similarity of every collected candidate (crypto_rec.hpp:220, cust_vector.hpp:160-174): a warp takes runs
// of 32 consecutive entries (segments are padded to multiples of 32, so each run belongs to one query), each lane walks its row.
// Two rows whose coordinates are all equal (c_a and c_b) need no walk: every product is p = fl(c_a c_b), and the extended
// accumulator holds k p exactly after k additions (53 + 7 bits), so the inner product is D p = h + l with h = fl(D p),
// l = fma(D, p, -h) -- what the walk would leave in (h, l), checked against long double arithmetic (tests/test_x87_cpu.py).
// Single-coin users form one large clique of mutually tied neighbours; this keeps their second pass off the FP64 pipe.
constexpr int P2_RUN = 64;   // 32-entry blocks a warp takes at a time (one binary search per run)
template <typename TQ, typename TB>
__global__ void __launch_bounds__(256)
p2_exact_kernel(const TQ* __restrict__ xq, int ldq, const double* __restrict__ sqn_q, const TB* __restrict__ xb, int ldb,
                const double* __restrict__ sqn_b, int D, int64_t q_begin, const int32_t* __restrict__ wq, const int64_t* __restrict__ off,
                unsigned int n, int64_t nblocks, const int32_t* __restrict__ cols, double* __restrict__ xs,
                const double* __restrict__ uval_q, const double* __restrict__ uval_b) {
    __shared__ rw::WarpTile tiles[8];
    __shared__ double qvec[8][128];
    __shared__ int q_ent[8][64], q_col[8][64];   // pairs that need the walk, gathered until a full warp of them is waiting
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t nruns = (nblocks + P2_RUN - 1) / P2_RUN;
    for (int64_t run = (int64_t)blockIdx.x * 8 + warp; run < nruns; run += (int64_t)gridDim.x * 8) {
        const int64_t b0 = run * P2_RUN, b1 = min(nblocks, b0 + P2_RUN);
        const int64_t run_base = b0 * 32;
        // item of the first block = the last i with off[i] <= base  (off[n] = total > base)
        unsigned int lo = 0, hi = n;
        while (lo < hi) { unsigned int m = (lo + hi + 1) >> 1; if (off[m] <= run_base) lo = m; else hi = m - 1; }
        int64_t next_off = off[lo + 1];
        bool staged = false;
        int64_t qrow = q_begin + wq[lo];
        double uq = uval_q[qrow], nq = sqn_q[qrow];
        int qn = 0;
        // the literal walk of (up to) 32 waiting pairs of the current query: queue slots [k0, k0 + cnt)
        auto walk_some = [&](int k0, int cnt) {
            if (!staged) { rw::stage_vector<TQ>(xq, ldq, qrow, qvec[warp]); staged = true; }
            const bool on = lane < cnt;
            const int col = on ? q_col[warp][k0 + lane] : -1;
            const int ent = on ? q_ent[warp][k0 + lane] : 0;
            rw::Walk wk = rw::walk_rows<TB, CRX_COSINE>(xb, ldb, D, on ? (int64_t)col : -1, qvec[warp], tiles[warp]);
            if (on) { X87 ip = {wk.a, wk.b}; xs[run_base + ent] = cos_sim_x87(ip, sqn_b[col], nq); }
            __syncwarp();
        };
        for (int64_t b = b0; b < b1; b++) {
            const int64_t base = b * 32;
            if (base >= next_off) {
                if (qn > 0) { walk_some(0, qn); qn = 0; }
                do { lo++; next_off = off[lo + 1]; } while (base >= next_off);
                qrow = q_begin + wq[lo];
                uq = uval_q[qrow]; nq = sqn_q[qrow];
                staged = false;
            }
            const int mine = cols[base + lane];
            bool walk = mine >= 0;
            if (mine >= 0 && uq == uq) {
                const double ub = uval_b[mine];
                if (ub == ub) {
                    const double pr = __dmul_rn(uq, ub);
                    X87 ip;
                    ip.h = __dmul_rn((double)D, pr);
                    ip.l = __fma_rn((double)D, pr, -ip.h);
                    xs[base + lane] = cos_sim_x87(ip, sqn_b[mine], nq);
                    walk = false;
                }
            }
            if (mine < 0) xs[base + lane] = -INFINITY;
            const unsigned bm = __ballot_sync(0xffffffffu, walk);
            if (bm) {
                if (walk) { const int pos = qn + __popc(bm & lt); q_ent[warp][pos] = (int)(base - run_base) + lane; q_col[warp][pos] = mine; }
                qn += __popc(bm);
                __syncwarp();
                if (qn >= 32) { walk_some(qn - 32, 32); qn -= 32; }
            }
        }
        if (qn > 0) walk_some(0, qn);
        __syncwarp();
    }
}

struct P2Resolve {
    const uint8_t* unk_q; const double* mean_q; const double* mean_b;
    int ldb, D, P, Nrec;
    int64_t q_begin;
    const int32_t* ncand;
    P2Queue cur, next;
    unsigned int n;
    const int64_t* off; const int32_t* count; const int32_t* ovf;
    int32_t* cols; double* xs;
    double eps;
    const double* eps_q;   // per absolute query row (centred operands), or NULL: eps for every query
    P2Blocks blocks;
    int32_t* recs; int32_t* nbr_rows; double* nbr_sims; int32_t* qstatus;
    unsigned long long* counters;
    unsigned long long* dbg;   // nullable, per kind (tie-order at [0], plateau at [4]): sum of |R|, sum of the list lengths, resolved, of which "H alone decides"
};

template <typename TB>
__global__ void __launch_bounds__(128)
p2_resolve_kernel(const TB* __restrict__ xb, P2Resolve a) {
    constexpr int QW = 4;
    __shared__ int posge[QW][128];
    __shared__ double hk[QW][128];
    __shared__ int hv[QW][128];
    __shared__ double qs_sk[QW][128];
    __shared__ int qs_sv[QW][128];
    __shared__ int qs_gp[QW][128];
    __shared__ int s_idx[QW][P2_MAXP];
    __shared__ double s_sim[QW][P2_MAXP];
    __shared__ double s_pred[QW][128];
    __shared__ int s_coin[QW][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned int i = blockIdx.x * QW + warp;
    if (i >= a.n) return;
    const int qrel = a.cur.q[i];
    const double theta = a.cur.theta[i];
    const int colx = a.cur.colx[i];
    const int tries = a.cur.tries[i];
    const double eps = a.eps_q ? a.eps_q[a.q_begin + qrel] : a.eps;
    if (a.ovf && a.ovf[i]) {   // the chunk pool ran dry in this round: again (the next round has fewer lists to hold)
        if (lane == 0) p2_emit(a.next, qrel, theta, colx, tries + 1);
        return;
    }
    const int cnt = a.count[i];
    double* key = a.xs + a.off[i];
    int* val = a.cols + a.off[i];
    const int nc = a.ncand[qrel];
    const int keep = min(a.P, nc);
    const bool complete = cnt >= nc;
    const double kappa = theta + 1.05 * eps;   // every candidate with an exact similarity above kappa is in the list
    const double lower = ldexp(64.0 * eps, 2 * min(tries & 0xffff, 8));
    if (cnt < keep) {
        if (lane == 0) p2_emit(a.next, qrel, (tries & 0xffff) > 8 ? -INFINITY : theta - lower, colx, tries + 1);
        return;
    }
    // ---- T_P = the keep-th best similarity (WarpTop: the best `keep` sorted over the warp).  Long lists (a plateau of tied
    // candidates) are read four 32-entry groups at a time so that enough loads are in flight
    WarpTop top;
    for (int base = 0; base < cnt; base += 128) {
        double k4[4];
        p2_prefetch_ahead(key, nullptr, base, cnt, lane);
#pragma unroll
        for (int u = 0; u < 4; u++) { const int e = base + u * 32 + lane; k4[u] = e < cnt ? key[e] : -INFINITY; }
#pragma unroll
        for (int u = 0; u < 4; u++) top.offer(k4[u], base + u * 32 + lane < cnt, keep);
    }
    const double TP = top.slot(keep - 1);
    if (!(TP > kappa) && !complete) {   // the P best are not all known yet
        if (lane == 0) p2_emit(a.next, qrel, fmin(theta - lower, TP - 2.1 * eps), colx, tries + 1);
        return;
    }
    // ---- H = { >= T_P }, e* = its last row: read from the back (on a plateau it is found at once)
    int idx_e = -1;
    for (int base = (cnt - 1) & ~31; base >= 0 && idx_e < 0; base -= 32) {
        const int e = base + lane;
        const unsigned b = __ballot_sync(0xffffffffu, e < cnt && key[e] >= TP);
        if (b) idx_e = base + 31 - __clz(b);
    }
    const int estar = val[idx_e];
    const double s_estar = key[idx_e];
    // ---- the tail behind e*: best known similarity and the first row that reaches it (not needed when e* is a minimum of H)
    const bool tail_all = complete || estar >= colx;
    double tk = -INFINITY;
    int te = 0x7fffffff;
    if (s_estar != TP) {
        for (int base = (idx_e + 1) & ~31; base < cnt; base += 32) {
            const int e = base + lane;
            if (e > idx_e && e < cnt) {
                const double k = key[e];
                if ((tail_all || k > kappa) && k > tk) { tk = k; te = e; }   // ascending e per lane: the first one is kept
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ok = __shfl_xor_sync(0xffffffffu, tk, o);
            const int oe = __shfl_xor_sync(0xffffffffu, te, o);
            if (ok > tk || (ok == tk && oe < te)) { tk = ok; te = oe; }
        }
    }
    const bool found = te != 0x7fffffff;
    double tp;
    int rlim;   // R = entries e < rlim with key >= tp
    if (s_estar == TP) {            // e* is a minimum of H: H alone decides (no candidate below T_P is ever consumed)
        tp = TP; rlim = idx_e + 1;
    } else if (!(tail_all || found)) {
        if (lane == 0) {
            int cx;
            double th = p2_tail_theta(a.blocks, qrel, estar, TP, eps, theta, cx);
            p2_emit(a.next, qrel, th, cx, tries + 1);
        }
        return;
    } else if (!found) {            // no candidate behind e*: R = every candidate
        if (!complete) { if (lane == 0) p2_emit(a.next, qrel, -INFINITY, 0x7fffffff, tries + 1); return; }
        tp = -INFINITY; rlim = cnt;
    } else {
        tp = tk; rlim = te;
        if (!(tp > kappa) && !complete) {   // t' is known, the rows in front that reach it are not all listed
            if (lane == 0) p2_emit(a.next, qrel, tp - 2.1 * eps, 0x7fffffff, tries + 1);
            return;
        }
    }
    // ---- R to the front, in row order (four groups of 32 entries per step; rows are read for the members only)
    int m = 0;
    if (tp == -INFINITY && rlim == cnt) m = cnt;   // R = every candidate: the list is R already
    else
    for (int base = 0; base < rlim; base += 128) {
        double k4[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int e = base + u * 32 + lane; k4[u] = e < rlim ? key[e] : -INFINITY; }
        int v4[4];
        bool g4[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = base + u * 32 + lane;
            g4[u] = e < rlim && (k4[u] >= tp || tp == -INFINITY);
            v4[u] = g4[u] ? val[e] : 0;
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = base + u * 32 + lane;
            const unsigned bg = __ballot_sync(0xffffffffu, g4[u]);
            if (g4[u]) { const int dst = m + __popc(bg & lt); if (dst != e) { key[dst] = k4[u]; val[dst] = v4[u]; } }
            m += __popc(bg);
        }
        __syncwarp();
    }
    if (a.dbg && lane == 0) {
        const int kd = (tries >> P2_KIND_SHIFT) == CRX_Q_PLATEAU ? 4 : 0;   // [0..3] tie-order queries, [4..7] plateau queries
        atomicAdd(&a.dbg[kd + 0], (unsigned long long)m); atomicAdd(&a.dbg[kd + 1], (unsigned long long)cnt); atomicAdd(&a.dbg[kd + 2], 1ull);
        if (s_estar == TP) atomicAdd(&a.dbg[kd + 3], 1ull);
    }
    warp_qs_topn_big(key, val, m, keep, QsScratch{posge[warp], hk[warp], hv[warp], qs_sk[warp], qs_sv[warp], qs_gp[warp]});
    for (int j = lane; j < keep; j += 32) { s_idx[warp][j] = val[j]; s_sim[warp][j] = key[j]; }
    __syncwarp();
    if (lane == 0) {
        atomicAdd(&a.counters[CRX_CNT_TOPP_PASS2], 1ull);
        if (a.qstatus) a.qstatus[qrel] = CRX_Q_EXACT;
    }
    if (a.nbr_rows) for (int j = lane; j < a.P; j += 32) a.nbr_rows[(size_t)qrel * a.P + j] = j < keep ? s_idx[warp][j] : -1;
    if (a.nbr_sims) for (int j = lane; j < a.P; j += 32) a.nbr_sims[(size_t)qrel * a.P + j] = j < keep ? s_sim[warp][j] : 0.0;
    if (a.recs) {
        const int64_t qrow = a.q_begin + qrel;
        predict_and_recommend<TB>(xb, a.ldb, a.mean_b, a.unk_q + qrow * a.D, a.mean_q[qrow], a.D, s_idx[warp], s_sim[warp], keep, a.Nrec,
                                  s_pred[warp], s_coin[warp], a.recs + (size_t)qrel * a.Nrec);
    }
}

// ------------------------------------------------------------------------------------------------
// Constant query rows (single-coin users) against the constant rows of the table: the one LARGE class of mutually tied
// neighbours this domain produces (5% of the users, all parallel to (1, ..., 1)).  For such a query q = c_q (1, ..., 1), c_q > 0:
//   * the similarity to a constant row c_b (1, ..., 1), c_b > 0, has the closed form of p2_exact_kernel: 1 within a few ulp;
//   * the similarity to ANY other row x is (sum x_k) / (sqrt(D) |x|) = alpha(x) up to rounding, the same for every such query:
//     rows with alpha(x) < 1 - 2^-39 (computed once, plain FP64, error < 2^-44) lie below 1 - 2^-40 for all of them.
// So when the P-th best similarity T_P over M = { constant rows with c > 0 } u { the (rare) rows with alpha >= 1 - 2^-39 } is
// >= 1 - 2^-40, the P best and every candidate tied with them are in M; and when the last row e* among them carries T_P itself
// ("H alone decides", recommend_pass2.cuh header), the reference's list is the literal sort of R = { members of M in front of
// e* (e* included) with similarity >= T_P }.  One warp per queued query walks M (row order) instead of sending 50k-entry lists
// through the collection scan; a query that does not meet the two conditions stays queued for the general rounds.
// ------------------------------------------------------------------------------------------------
struct P2Uniform {
    const int32_t* mrow;     // [nm] rows of M, ascending
    const double* mval;      // [nm] constant value of the row, NaN for a near-constant row (exact walk)
    const double* mroot;     // [nm] sqrt of the row's sum of squares
    const uint32_t* mcode;   // [nm] packed table code of the row
    int nm;
    const uint32_t* qcode;   // packed codes of the query rows
    uint32_t low, high;
    const double* uval_q; const double* sqn_q;
};

template <typename TQ, typename TB>
__device__ __forceinline__ double p2u_sim(const P2Uniform& u, int j, double uq, double rq, int D, const TQ* __restrict__ xq_row,
                                          const TB* __restrict__ xb, int ldb) {
    const double ub = u.mval[j];
    if (ub == ub) {
        const double pr = __dmul_rn(uq, ub);
        X87 ip;
        ip.h = __dmul_rn((double)D, pr);
        ip.l = __fma_rn((double)D, pr, -ip.h);
        return cos_sim_x87_roots(ip, u.mroot[j], rq);
    }
    return cos_sim_x87_roots(dot_x87(xb + (size_t)u.mrow[j] * ldb, xq_row, D), u.mroot[j], rq);   // near-constant row: rare
}

// phase A: T_P, e*, and the answer when R is one run of equal similarities (need[i] = 0); otherwise need[i] = |R| for phase B,
// or -1 when the query stays with the general rounds
template <typename TQ, typename TB>
__global__ void __launch_bounds__(128)
p2u_select_kernel(const TQ* __restrict__ xq, int ldq, const TB* __restrict__ xb, P2Resolve a, P2Uniform u, int32_t* __restrict__ need,
                  int32_t* __restrict__ rlim_out, double* __restrict__ tp_out) {
    constexpr int QW = 4;
    __shared__ int s_idx[QW][P2_MAXP];
    __shared__ double s_sim[QW][P2_MAXP];
    __shared__ double s_pred[QW][128];
    __shared__ int s_coin[QW][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int i = blockIdx.x * QW + warp;
    if (i >= a.n) return;
    const int qrel = a.cur.q[i];
    const int64_t qrow = a.q_begin + qrel;
    if (lane == 0) need[i] = -1;
    const double uq = u.uval_q[qrow];
    const int nc = a.ncand[qrel];
    if (!(uq > 0.0) || nc < a.P) return;
    const int keep = a.P;
    const double rq = __dsqrt_rn(u.sqn_q[qrow]);
    const uint32_t cq = u.qcode[qrow];
    const TQ* xq_row = xq + (size_t)qrow * ldq;
    const double FLOOR = 1.0 - 9.094947017729282e-13;   // 1 - 2^-40
    // ---- T_P over the candidates in M
    WarpTop top;
    for (int base = 0; base < u.nm; base += 32) {
        const int j = base + lane;
        double k = -INFINITY;
        bool cand = false;
        if (j < u.nm) {
            const uint32_t x = cq ^ u.mcode[j];
            cand = ((x - u.low) & ~x & u.high) != 0u;
            if (cand) k = p2u_sim<TQ, TB>(u, j, uq, rq, a.D, xq_row, xb, a.ldb);
        }
        top.offer(k, cand, keep);
    }
    if (top.filled < keep) return;
    const double TP = top.slot(keep - 1);
    const double best = top.slot(0);
    if (!(TP >= FLOOR)) return;
    // ---- e* = the last candidate of M with similarity >= T_P (read from the back), which must carry T_P itself
    int je = -1;
    double se = 0.0;
    for (int base = (u.nm - 1) & ~31; base >= 0 && je < 0; base -= 32) {
        const int j = base + lane;
        double k = -INFINITY;
        if (j < u.nm) {
            const uint32_t x = cq ^ u.mcode[j];
            if (((x - u.low) & ~x & u.high) != 0u) k = p2u_sim<TQ, TB>(u, j, uq, rq, a.D, xq_row, xb, a.ldb);
        }
        const unsigned b = __ballot_sync(0xffffffffu, k >= TP);
        if (b) { const int src = 31 - __clz(b); je = base + src; se = __shfl_sync(0xffffffffu, k, src); }
    }
    if (je < 0 || se != TP) return;
    if (best != TP) {
        // similarities above T_P among the P best: R is not one run of equal keys -- its literal sort runs in phase B
        int m = 0;
        for (int base = 0; base <= je; base += 32) {
            const int j = base + lane;
            double k = -INFINITY;
            if (j <= je) {
                const uint32_t x = cq ^ u.mcode[j];
                if (((x - u.low) & ~x & u.high) != 0u) k = p2u_sim<TQ, TB>(u, j, uq, rq, a.D, xq_row, xb, a.ldb);
            }
            m += __popc(__ballot_sync(0xffffffffu, k >= TP));
        }
        if (lane == 0) { need[i] = m; rlim_out[i] = je + 1; tp_out[i] = TP; }
        return;
    }
    // ---- every member of R equals T_P: the literal sort leaves it untouched, the answer is its first P rows
    int got = 0;
    for (int base = 0; base <= je && got < keep; base += 32) {
        const int j = base + lane;
        double k = -INFINITY;
        if (j <= je) {
            const uint32_t x = cq ^ u.mcode[j];
            if (((x - u.low) & ~x & u.high) != 0u) k = p2u_sim<TQ, TB>(u, j, uq, rq, a.D, xq_row, xb, a.ldb);
        }
        const bool hit = k >= TP;
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (hit) { const int pos = got + __popc(b & ((1u << lane) - 1u)); if (pos < keep) { s_idx[warp][pos] = u.mrow[j]; s_sim[warp][pos] = k; } }
        got += __popc(b);
    }
    __syncwarp();
    if (lane == 0) {
        need[i] = 0;
        atomicAdd(&a.counters[CRX_CNT_TOPP_PASS2], 1ull);
        if (a.qstatus) a.qstatus[qrel] = CRX_Q_EXACT;
    }
    if (a.nbr_rows) for (int j = lane; j < a.P; j += 32) a.nbr_rows[(size_t)qrel * a.P + j] = s_idx[warp][j];
    if (a.nbr_sims) for (int j = lane; j < a.P; j += 32) a.nbr_sims[(size_t)qrel * a.P + j] = s_sim[warp][j];
    if (a.recs)
        predict_and_recommend<TB>(xb, a.ldb, a.mean_b, a.unk_q + qrow * a.D, a.mean_q[qrow], a.D, s_idx[warp], s_sim[warp], keep, a.Nrec,
                                  s_pred[warp], s_coin[warp], a.recs + (size_t)qrel * a.Nrec);
}

// phase B: R written out (row order), literal quicksort, answer.  off[i] .. off[i + 1]: the query's slice of the scratch
template <typename TQ, typename TB>
__global__ void __launch_bounds__(128)
p2u_sort_kernel(const TQ* __restrict__ xq, int ldq, const TB* __restrict__ xb, P2Resolve a, P2Uniform u, const int32_t* __restrict__ need,
                const int32_t* __restrict__ rlim_in, const double* __restrict__ tp_in, const int64_t* __restrict__ off,
                double* __restrict__ keys, int32_t* __restrict__ vals) {
    constexpr int QW = 4;
    __shared__ int posge[QW][128];
    __shared__ double hk[QW][128];
    __shared__ int hv[QW][128];
    __shared__ double qs_sk[QW][128];
    __shared__ int qs_sv[QW][128];
    __shared__ int qs_gp[QW][128];
    __shared__ int s_idx[QW][P2_MAXP];
    __shared__ double s_sim[QW][P2_MAXP];
    __shared__ double s_pred[QW][128];
    __shared__ int s_coin[QW][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int i = blockIdx.x * QW + warp;
    if (i >= a.n || need[i] <= 0) return;
    const int qrel = a.cur.q[i];
    const int64_t qrow = a.q_begin + qrel;
    const double uq = u.uval_q[qrow], rq = __dsqrt_rn(u.sqn_q[qrow]), TP = tp_in[i];
    const uint32_t cq = u.qcode[qrow];
    const TQ* xq_row = xq + (size_t)qrow * ldq;
    double* key = keys + off[i];
    int* val = vals + off[i];
    const int rlim = rlim_in[i], keep = a.P;
    int m = 0;
    for (int base = 0; base < rlim; base += 32) {
        const int j = base + lane;
        double k = -INFINITY;
        if (j < rlim) {
            const uint32_t x = cq ^ u.mcode[j];
            if (((x - u.low) & ~x & u.high) != 0u) k = p2u_sim<TQ, TB>(u, j, uq, rq, a.D, xq_row, xb, a.ldb);
        }
        const bool hit = k >= TP;
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (hit) { const int pos = m + __popc(b & ((1u << lane) - 1u)); key[pos] = k; val[pos] = u.mrow[j]; }
        m += __popc(b);
    }
    __syncwarp();
    warp_qs_topn_big(key, val, m, keep, QsScratch{posge[warp], hk[warp], hv[warp], qs_sk[warp], qs_sv[warp], qs_gp[warp]});
    for (int j = lane; j < keep; j += 32) { s_idx[warp][j] = val[j]; s_sim[warp][j] = key[j]; }
    __syncwarp();
    if (lane == 0) {
        atomicAdd(&a.counters[CRX_CNT_TOPP_PASS2], 1ull);
        if (a.qstatus) a.qstatus[qrel] = CRX_Q_EXACT;
    }
    if (a.nbr_rows) for (int j = lane; j < a.P; j += 32) a.nbr_rows[(size_t)qrel * a.P + j] = s_idx[warp][j];
    if (a.nbr_sims) for (int j = lane; j < a.P; j += 32) a.nbr_sims[(size_t)qrel * a.P + j] = s_sim[warp][j];
    if (a.recs)
        predict_and_recommend<TB>(xb, a.ldb, a.mean_b, a.unk_q + qrow * a.D, a.mean_q[qrow], a.D, s_idx[warp], s_sim[warp], keep, a.Nrec,
                                  s_pred[warp], s_coin[warp], a.recs + (size_t)qrel * a.Nrec);
}

// alpha(x) = (sum x_k) / (sqrt(D) |x|) of the rows that are not constant; flag[row] = 1 for the members of M
template <typename T>
__global__ void __launch_bounds__(256) p2u_flag_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, const double* __restrict__ uval,
                                                       int64_t n, uint8_t* __restrict__ flag, int* __restrict__ refuse) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double uv = uval[row];
    if (uv == uv) {
        if (lane == 0) { flag[row] = uv > 0.0; if (uv == 0.0) *refuse = 1; }   // an all-zero row has no similarity: the general rounds decide
        return;
    }
    double s = 0.0;
    for (int k = lane; k < D; k += 32) s += (double)x[row * ld + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double alpha = s / (sqrt((double)D) * sqrt(sqn[row]));
    if (lane == 0) flag[row] = alpha >= 1.0 - 1.8189894035458565e-12;   // 1 - 2^-39 (NaN: not a member)
}

__global__ void p2u_gather_kernel(const int32_t* __restrict__ mrow, int nm, const double* __restrict__ uval, const double* __restrict__ sqn,
                                  const uint32_t* __restrict__ ccode, double* __restrict__ mval, double* __restrict__ mroot, uint32_t* __restrict__ mcode) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nm) return;
    const int r = mrow[j];
    mval[j] = uval[r];
    mroot[j] = __dsqrt_rn(sqn[r]);
    mcode[j] = ccode[r];
}

// queue entries that phase A / B answered leave the queue
__global__ void p2u_compact_kernel(P2Queue cur, unsigned int n, const int32_t* __restrict__ need, P2Queue out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || need[i] >= 0) return;
    p2_emit(out, cur.q[i], cur.theta[i], cur.colx[i], cur.tries[i]);
}

// queries still queued after the last round keep the provisional answer of rec_finalize and are counted
__global__ void p2_leftover_kernel(P2Queue w, unsigned int n, unsigned long long* counters) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int kind = w.tries[i] >> P2_KIND_SHIFT;
    atomicAdd(&counters[kind == CRX_Q_PLATEAU ? CRX_CNT_TOPP_RESCAN : CRX_CNT_TOPP_TIES], 1ull);
}
