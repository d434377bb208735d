// recommend.cu -- crypto_rec.hpp:214-345 on the GPU:
//   K9  masked cosine-similarity scan with a register-resident top-32 list per query (filter),
//   K10 exact re-evaluation of the survivors, descending order, weighted rating prediction and the
//       literal Lomuto top-N over the unknown coins (refine + predict),
//   cluster-neighbour recommendation (get_top_N_recom without similarities).
#include <algorithm>
#include <climits>
#include <ctime>

#include <cub/cub.cuh>

#include "pair_tile.cuh"
#include "rowwalk.cuh"
#include "tables.cuh"
#include "tc_scan.cuh"

constexpr int LIST = 32;   // per-query candidate list length (one entry per lane)
constexpr int MAXL = 16;

struct ScanArgs {
    int L, pass;                      // pass = table index, or -1 for the dense any-table scan
    int64_t nq;                       // queries in this launch (positions in qperm)
    int64_t nb;                       // candidate positions
    const int32_t* qperm;             // position -> query row (absolute row of the query point set)
    int64_t q_begin;                  // first query row of the batch (lists / ncand are relative to it)
    const int32_t* qgid[MAXL];        // [row] group id of the query rows per table
    const int32_t* cgid[MAXL];        // [row] group id of the base rows per table
    const int32_t* cperm;             // position -> base row (NULL = identity)
    const int32_t* csorted;           // group ids of the pass's table in cperm order (NULL for dense)
    double* list_s;                   // [nq_total][LIST]
    int32_t* list_i;                  // [nq_total][LIST]
    int32_t* ncand;                   // [nq_total]
    int load_state;                   // lists already hold the results of earlier passes
};

__device__ __forceinline__ void warp_argmin(double v, double& mn, int& ml) {
    int lane = threadIdx.x & 31;
    double bv = v;
    int bl = lane;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        int ol = __shfl_xor_sync(0xffffffffu, bl, off);
        if (ov < bv || (ov == bv && ol < bl)) { bv = ov; bl = ol; }
    }
    mn = bv; ml = bl;
}

template <typename TQ, typename TB>
__global__ void __launch_bounds__(pt::NT)
topp_scan_kernel(const TQ* __restrict__ xq, int ldq, const double* __restrict__ sqn_q, const TB* __restrict__ xb, int ldb,
                 const double* __restrict__ sqn_b, ScanArgs a) {
    extern __shared__ double sm[];
    int ld = ldb;
    double* As = sm;
    double* Bs = sm + pt::BM * ld;
    double* cinv = Bs + pt::BN * ld;                       // [BN]
    int32_t* crow = reinterpret_cast<int32_t*>(cinv + pt::BN);  // [BN]
    int32_t* cg = crow + pt::BN;                           // [L][BN]
    int32_t* qg = cg + MAXL * pt::BN;                      // [L][BM]
    __shared__ int64_t range[2];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t p0 = (int64_t)blockIdx.x * pt::BM;
    pt::load_a_tile<TQ>(As, xq, ld, a.qperm, p0, a.nq);
    for (int e = threadIdx.x; e < a.L * pt::BM; e += pt::NT) {
        int l = e / pt::BM, r = e - l * pt::BM;
        int64_t pos = p0 + r;
        qg[l * pt::BM + r] = pos < a.nq ? a.qgid[l][a.qperm[pos]] : -2;
    }
    if (threadIdx.x == 0) {
        int64_t c0 = 0, c1 = a.nb;
        if (a.pass >= 0) {
            int64_t plast = min(p0 + pt::BM, a.nq) - 1;
            int gfirst = a.qgid[a.pass][a.qperm[p0]], glast = a.qgid[a.pass][a.qperm[plast]];
            if (glast < 0) { c0 = 0; c1 = 0; }
            else {
                if (gfirst < 0) gfirst = 0;
                int64_t lo = 0, hi = a.nb;
                while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a.csorted[m] < gfirst) lo = m + 1; else hi = m; }
                c0 = lo;
                hi = a.nb;
                while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a.csorted[m] <= glast) lo = m + 1; else hi = m; }
                c1 = lo;
            }
        }
        range[0] = c0; range[1] = c1;
    }
    // per-row state (this lane's slot of each of the warp's 8 lists)
    double ls[pt::RW], thr[pt::RW], qinv[pt::RW];
    int li[pt::RW], minlane[pt::RW], cnt[pt::RW];
    int64_t qrel[pt::RW];
#pragma unroll
    for (int r = 0; r < pt::RW; r++) {
        int64_t pos = p0 + warp * pt::RW + r;
        cnt[r] = 0;
        if (pos < a.nq) {
            int64_t qrow = a.qperm[pos];
            qrel[r] = qrow - a.q_begin;
            qinv[r] = 1.0 / sqrt(sqn_q[qrow]);
            if (a.load_state) { ls[r] = a.list_s[qrel[r] * LIST + lane]; li[r] = a.list_i[qrel[r] * LIST + lane]; }
            else { ls[r] = -INFINITY; li[r] = -1; }
        } else { qrel[r] = -1; qinv[r] = 0.0; ls[r] = -INFINITY; li[r] = -1; }
        warp_argmin(ls[r], thr[r], minlane[r]);
    }
    __syncthreads();
    int64_t c0 = range[0], c1 = range[1];
    for (int64_t cb = c0; cb < c1; cb += pt::BN) {
        __syncthreads();
        pt::load_b_tile<TB>(Bs, xb, ld, a.cperm, cb, c1);
        if (threadIdx.x < pt::BN) {
            int64_t pos = cb + threadIdx.x;
            int row = -1;
            double inv = 0.0;
            if (pos < c1) { row = a.cperm ? a.cperm[pos] : (int)pos; inv = 1.0 / sqrt(sqn_b[row]); }
            crow[threadIdx.x] = row;
            cinv[threadIdx.x] = inv;
        }
        for (int e = threadIdx.x; e < a.L * pt::BN; e += pt::NT) {
            int l = e / pt::BN, col = e - l * pt::BN;
            int64_t pos = cb + col;
            int g = -3;
            if (pos < c1) { int row = a.cperm ? a.cperm[pos] : (int)pos; g = a.cgid[l][row]; }
            cg[l * pt::BN + col] = g;
        }
        __syncthreads();
        double acc[pt::RW][2];
        pt::tile_mac<pt::FORM_DOT>(As, Bs, ld, warp, lane, acc);
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            int col = 2 * lane + cc;
            int brow = crow[col];
            double binv = cinv[col];
#pragma unroll
            for (int r = 0; r < pt::RW; r++) {
                int rr = warp * pt::RW + r;
                bool match = false;
                if (brow >= 0 && qrel[r] >= 0) {
                    if (a.pass < 0) {
                        for (int l = 0; l < a.L; l++) match |= (qg[l * pt::BM + rr] == cg[l * pt::BN + col]);
                    } else {
                        match = qg[a.pass * pt::BM + rr] == cg[a.pass * pt::BN + col];
                        for (int l = 0; l < a.pass; l++) match &= (qg[l * pt::BM + rr] != cg[l * pt::BN + col]);
                    }
                }
                double s = acc[r][cc] * qinv[r] * binv;
                if (match) cnt[r]++;
                bool want = match && s > thr[r];
                unsigned m = __ballot_sync(0xffffffffu, want);
                while (m) {
                    int src = __ffs(m) - 1;
                    m &= m - 1;
                    double sv = __shfl_sync(0xffffffffu, s, src);
                    int iv = __shfl_sync(0xffffffffu, brow, src);
                    if (sv > thr[r]) {
                        if (lane == minlane[r]) { ls[r] = sv; li[r] = iv; }
                        warp_argmin(ls[r], thr[r], minlane[r]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < pt::RW; r++) {
        int total = cnt[r];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
        if (qrel[r] >= 0) {
            a.list_s[qrel[r] * LIST + lane] = ls[r];
            a.list_i[qrel[r] * LIST + lane] = li[r];
            if (lane == 0 && total) atomicAdd(&a.ncand[qrel[r]], total);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-cooperative form of the literal Lomuto quicksort (crypto_rec.hpp:235-277) when only the first `need`
// positions of the result are consumed (resize(N), crypto_rec.hpp:322).  key/val live in shared memory, n <= 128.
//
// One partition of [lo, hi] with pivot key[hi] leaves (a) the elements >= pivot in their ORIGINAL order in
// [lo, p), (b) the pivot at p = lo + #{>= pivot}, (c) the elements < pivot, permuted, in (p, hi].  Ranges that
// start at or beyond `need` never influence the consumed prefix, so (c) only matters when p + 1 < need; then the
// step is done literally by one lane.  Otherwise (a) is a stable compaction done by the whole warp with ballots.
// A range whose keys all equal the pivot is left untouched by the literal algorithm and is skipped.
// ------------------------------------------------------------------------------------------------
template <int NT = 4>   // n <= 32 * NT
__device__ __forceinline__ void warp_lomuto_topn(double* key, int* val, int n, int need) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int st_lo[128], st_hi[128];  // at most one pending right part per consumed position (need <= 128)
    int sp = 1;
    st_lo[0] = 0; st_hi[0] = n - 1;
    while (sp > 0) {
        sp--;
        int lo = st_lo[sp], hi = st_hi[sp];
        while (lo < hi && lo < need) {
            const double pivot = key[hi];
            const int m = hi - lo;  // elements in front of the pivot
            double kk[NT];
            int vv[NT];
            unsigned ge[NT];
            bool alleq = true;
            int cnt = 0;
#pragma unroll
            for (int t = 0; t < NT; t++) {
                int e = t * 32 + lane;
                bool in = e < m;
                kk[t] = in ? key[lo + e] : 0.0;
                vv[t] = in ? val[lo + e] : 0;
                bool g = in && (kk[t] >= pivot);
                ge[t] = __ballot_sync(0xffffffffu, g);
                alleq = alleq && __all_sync(0xffffffffu, !in || (kk[t] == pivot));
                cnt += __popc(ge[t]);
            }
            if (alleq) break;
            const int p = lo + cnt;
            if (p + 1 < need && p < hi) {
                // the right part reaches into the consumed prefix: literal partition by one lane
                __syncwarp();
                if (lane == 0) {
                    int i = lo - 1;
                    for (int j = lo; j < hi; j++) {
                        if (key[j] >= pivot) {
                            i++;
                            double tk = key[i]; key[i] = key[j]; key[j] = tk;
                            int tv = val[i]; val[i] = val[j]; val[j] = tv;
                        }
                    }
                    double tk = key[i + 1]; key[i + 1] = key[hi]; key[hi] = tk;
                    int tv = val[i + 1]; val[i + 1] = val[hi]; val[hi] = tv;
                }
                __syncwarp();
                if (sp < 128) { st_lo[sp] = p + 1; st_hi[sp] = hi; sp++; }  // right part afterwards
                hi = p - 1;                                                  // left part first
                continue;
            }
            if (cnt < m) {  // something moves: stable compaction of the ">= pivot" elements, pivot behind them
                const int pv = val[hi];
                __syncwarp();
                int base = 0;
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    if ((ge[t] >> lane) & 1u) {
                        int dst = lo + base + __popc(ge[t] & lt);
                        key[dst] = kk[t];
                        val[dst] = vv[t];
                    }
                    base += __popc(ge[t]);
                }
                if (lane == 0) { key[p] = pivot; val[p] = pv; }
                __syncwarp();
            }
            hi = p - 1;  // [p + 1, hi] starts at or beyond `need` (or is empty): never consumed
        }
    }
    __syncwarp();
}

#include "recommend_pass2.cuh"

// ------------------------------------------------------------------------------------------------
// K10: refine + predict + top-N.  One warp per query.  LISTN = 32 (FP64 scan, double scores) or
// 64 (tensor-core filter, float scores scaled by `approx_scale`, filter error `approx_eps`).
// ------------------------------------------------------------------------------------------------
template <typename TQ, typename TB, int LISTN, typename TS>
__global__ void __launch_bounds__(128)
rec_finalize_kernel(const TQ* __restrict__ xq, int ldq, const double* __restrict__ sqn_q, const uint8_t* __restrict__ unk_q,
                    const double* __restrict__ mean_q, const TB* __restrict__ xb, int ldb, const double* __restrict__ sqn_b,
                    const double* __restrict__ mean_b, int D, int64_t q_begin, int64_t nq, int P, int Nrec,
                    const TS* __restrict__ list_s, const int32_t* __restrict__ list_i, const int32_t* __restrict__ ncand,
                    double approx_scale, double approx_eps_all, int32_t* __restrict__ recs, int32_t* __restrict__ nbr_rows,
                    double* __restrict__ nbr_sims, unsigned long long* counters, int tie_order, int32_t* __restrict__ qstatus,
                    int pass2, P2Queue w2, P2Blocks blk, const double* __restrict__ eps_q /* per query row, or NULL */) {
    constexpr int EPL = LISTN / 32;  // list entries per lane
    constexpr int QW = 4;            // queries (warps) per block
    __shared__ int a_idx[QW][LISTN];
    __shared__ double a_sim[QW][LISTN];
    __shared__ int s_idx[QW][P2_MAXP];      // the P best in their final order (P <= 64 on the tensor path, <= 32 otherwise)
    __shared__ double s_sim[QW][P2_MAXP];
    __shared__ double s_pred[QW][128];
    __shared__ int s_coin[QW][128];
    __shared__ rw::WarpTile tiles[QW];
    __shared__ double qvec[QW][128];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t qrel = (int64_t)blockIdx.x * QW + warp;
    if (qrel >= nq) return;
    int64_t qrow = q_begin + qrel;
    const double approx_eps = eps_q ? eps_q[qrow] : approx_eps_all;   // the filter's error bound for this query (centred operands: per row)
    // every 32-slot group of the list is an independent "best 32 of its share of the columns" list
    // (the FP64 scan has one, the tensor-core filter one per column half): a candidate outside the lists
    // scores at most max over FULL groups of (smallest approximate score of the group)
    double floor_s = -INFINITY;
    int nvalid = 0;
    rw::stage_vector<TQ>(xq, ldq, qrow, qvec[warp]);
    const double nq_ = sqn_q[qrow];
    int idx_e[EPL];
    double val_e[EPL];   // approximate score, -inf for an empty slot
#pragma unroll
    for (int e = 0; e < EPL; e++) {
        int slot = e * 32 + lane;
        idx_e[e] = list_i[qrel * LISTN + slot];
        double ap = idx_e[e] >= 0 ? (double)list_s[qrel * LISTN + slot] : INFINITY;
        val_e[e] = idx_e[e] >= 0 ? ap : -INFINITY;
        int gcount = __popc(__ballot_sync(0xffffffffu, idx_e[e] >= 0));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ap = fmin(ap, __shfl_xor_sync(0xffffffffu, ap, off));
        nvalid += gcount;
        if (gcount == 32) floor_s = fmax(floor_s, ap);
    }
    // Listed candidates whose approximate score is more than twice the filter error below the P-th best approximate
    // score cannot be among the exact P best: they are dropped before the exact (expensive) evaluation.  When at most 32
    // survive, one row walk serves them all.
    bool one_walk = false;
    int walked = 0;              // one-walk path: slots of a_idx / a_sim in use
    double cut = -INFINITY;      // filter units: listed candidates below it are not evaluated (one-walk path)
    if (EPL == 2 && nvalid > P) {
        int g0 = 0, g1 = 0;
        for (int t = 0; t < 32; t++) {
            double o0 = __shfl_sync(0xffffffffu, val_e[0], t), o1 = __shfl_sync(0xffffffffu, val_e[EPL - 1], t);
            g0 += (o0 > val_e[0]) + (o1 > val_e[0]);
            g1 += (o0 > val_e[EPL - 1]) + (o1 > val_e[EPL - 1]);
        }
        // P-th largest = the smallest value that has fewer than P strictly larger ones
        double T = fmin(g0 < P ? val_e[0] : INFINITY, g1 < P ? val_e[EPL - 1] : INFINITY);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) T = fmin(T, __shfl_xor_sync(0xffffffffu, T, off));
        const double margin = 2.02 * approx_eps / approx_scale;   // approx_eps is in similarity units, the list in filter units
        cut = T - margin;
        bool k0 = idx_e[0] >= 0 && val_e[0] >= cut, k1 = idx_e[EPL - 1] >= 0 && val_e[EPL - 1] >= cut;
        unsigned b0 = __ballot_sync(0xffffffffu, k0), b1 = __ballot_sync(0xffffffffu, k1);
        int count = __popc(b0) + __popc(b1);
        if (count <= 32) {
            walked = count;
            one_walk = true;
            unsigned lt = (1u << lane) - 1u;
            if (k0) { a_idx[warp][__popc(b0 & lt)] = idx_e[0]; s_pred[warp][__popc(b0 & lt)] = val_e[0]; }
            if (k1) { a_idx[warp][__popc(b0) + __popc(b1 & lt)] = idx_e[EPL - 1]; s_pred[warp][__popc(b0) + __popc(b1 & lt)] = val_e[EPL - 1]; }
            __syncwarp();
            int mine = lane < count ? a_idx[warp][lane] : -1;
            const double mine_ap = lane < count ? s_pred[warp][lane] : 0.0;
            __syncwarp();
            rw::Walk w = rw::walk_rows<TB, CRX_COSINE>(xb, ldb, D, mine >= 0 ? (int64_t)mine : -1, qvec[warp], tiles[warp]);
            double mysim = -INFINITY;
            if (mine >= 0) { X87 ip = {w.a, w.b}; mysim = cos_sim_x87(ip, sqn_b[mine], nq_); }
            // the error bound everything above rests on, checked on every evaluated candidate: a violation is counted as an
            // uncertified query (the tests require the counter to stay 0)
            if (mine >= 0 && fabs(mine_ap * approx_scale - mysim) > approx_eps) atomicAdd(&counters[CRX_CNT_TOPP_RESCAN], 1ull);
            a_idx[warp][lane] = mine; a_sim[warp][lane] = mysim;
            a_idx[warp][32 + lane] = -1; a_sim[warp][32 + lane] = -INFINITY;
        }
    }
    if (!one_walk) {
#pragma unroll
        for (int e = 0; e < EPL; e++) {
            int slot = e * 32 + lane;
            int idx = idx_e[e];
            // the reference's own similarity of every listed candidate (crypto_rec.hpp:220, cust_vector.hpp:160-174): each
            // lane walks ITS candidate in index order with the x87 accumulation, rows staged through shared memory
            double mysim = -INFINITY;
            if (__ballot_sync(0xffffffffu, idx >= 0)) {
                rw::Walk w = rw::walk_rows<TB, CRX_COSINE>(xb, ldb, D, idx >= 0 ? (int64_t)idx : -1, qvec[warp], tiles[warp]);
                if (idx >= 0) { X87 ip = {w.a, w.b}; mysim = cos_sim_x87(ip, sqn_b[idx], nq_); }
                if (idx >= 0 && fabs(val_e[e] * approx_scale - mysim) > approx_eps) atomicAdd(&counters[CRX_CNT_TOPP_RESCAN], 1ull);
            }
            a_idx[warp][slot] = idx;
            a_sim[warp][slot] = mysim;
        }
    }
    __syncwarp();
    int keep = min(P, nvalid);
    bool tied[EPL];
#pragma unroll
    for (int e = 0; e < EPL; e++) tied[e] = false;
    // descending similarity, ties by ascending row
#pragma unroll
    for (int e = 0; e < EPL; e++) {
        int slot = e * 32 + lane;
        int idx = a_idx[warp][slot];
        double sim = a_sim[warp][slot];
        if (idx >= 0) {
            int rank = 0, same = 0;
            for (int t = 0; t < LISTN; t++) {
                int oi = a_idx[warp][t];
                double os = a_sim[warp][t];
                if (oi >= 0 && (os > sim || (os == sim && oi < idx))) rank++;
                same += oi >= 0 && os == sim;
            }
            if (rank < P2_MAXP) { s_idx[warp][rank] = idx; s_sim[warp][rank] = sim; }
            tied[e] = same > 1;
        }
    }
    __syncwarp();
    int nc = ncand[qrel];
    // certification: every candidate outside the list has an approximate similarity <= the smallest
    // approximate one in it; the exact P-th best must clear that by more than the filter's error
    int status = CRX_Q_EXACT;
    if (keep > 0) {
        const double floor_ = floor_s * approx_scale;
        const double unlisted = floor_s > -INFINITY ? floor_ + approx_eps * fmax(1.0, fabs(floor_)) : -INFINITY;
        const double pth = s_sim[warp][keep - 1];
        if (!(pth > unlisted)) {
            status = CRX_Q_PLATEAU;
            if (lane == 0) {
                // second pass: every candidate that can reach the P-th best listed similarity
                if (pass2) p2_emit(w2, (int)qrel, pth - 2.1 * approx_eps, 0x7fffffff, CRX_Q_PLATEAU << P2_KIND_SHIFT);
                else atomicAdd(&counters[CRX_CNT_TOPP_RESCAN], 1ull);
            }
        } else {
            bool t = false;
#pragma unroll
            for (int e = 0; e < EPL; e++) t |= tied[e] && a_sim[warp][e * 32 + lane] >= pth;
            const bool any_tie = __any_sync(0xffffffffu, t);
            if (any_tie && !tie_order) status = CRX_Q_TIE_ORDER;
            if (any_tie && lane == 0) {
                atomicAdd(&counters[CRX_CNT_TOPP_TIED], 1ull);
                if (!tie_order) {
                    if (pass2) p2_emit(w2, (int)qrel, pth - 2.1 * approx_eps, 0x7fffffff, CRX_Q_TIE_ORDER << P2_KIND_SHIFT);
                    else atomicAdd(&counters[CRX_CNT_TOPP_TIES], 1ull);
                }
            }
            if (any_tie && tie_order) {
                // ---- equal similarities among the P best: reproduce the order of the reference's quicksort ----
                // (crypto_rec.hpp:235-277 over the candidates in std::set order = row order, lsh_cube.hpp:96-104.)  While
                // the ">= pivot" side of a partition still holds P elements only that side is consumed, and it keeps
                // the row order; so the consumed prefix equals the literal sort of
                //     R = { candidates before row r' with similarity >= t' },  in row order,
                // t' = the best similarity stored behind the last member e* of the P best (ties included), r' = the
                // first row behind e* that reaches t' (DESIGN.md section 3).  R is known exactly when every candidate
                // above t' has been evaluated: unlisted ones are below `unlisted`, unevaluated listed ones below `cut2`.
                int estar = -1;
#pragma unroll
                for (int e = 0; e < EPL; e++)
                    if (a_idx[warp][e * 32 + lane] >= 0 && a_sim[warp][e * 32 + lane] >= pth) estar = max(estar, a_idx[warp][e * 32 + lane]);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) estar = max(estar, __shfl_xor_sync(0xffffffffu, estar, off));
                double s_estar = -INFINITY;
#pragma unroll
                for (int e = 0; e < EPL; e++) if (a_idx[warp][e * 32 + lane] == estar) s_estar = a_sim[warp][e * 32 + lane];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s_estar = fmax(s_estar, __shfl_xor_sync(0xffffffffu, s_estar, off));
                int before = 0;
#pragma unroll
                for (int e = 0; e < EPL; e++)
                    before += __popc(__ballot_sync(0xffffffffu, a_idx[warp][e * 32 + lane] >= 0 && a_idx[warp][e * 32 + lane] < estar &&
                                                                    a_sim[warp][e * 32 + lane] >= s_estar));
                double known_above = unlisted, tp = -INFINITY;
                int rlim = 0x7fffffff;
                bool resolved = true;
                if (before >= keep) {
                    // e* itself is still a pivot with P elements on its ">=" side (a plateau reaching beyond the P-th place):
                    // R = the members of the P best (ties included) in front of e*
                    tp = s_estar; rlim = estar;
                } else {
                    if (one_walk) {
                        // evaluate the listed candidates down to the best one stored behind e* (by approximate score)
                        double tm = -INFINITY;
#pragma unroll
                        for (int e = 0; e < EPL; e++) if (idx_e[e] > estar) tm = fmax(tm, val_e[e]);
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) tm = fmax(tm, __shfl_xor_sync(0xffffffffu, tm, off));
                        if (tm > -INFINITY || floor_s == -INFINITY) {
                            const double cut2 = tm > -INFINITY ? tm - 2.02 * approx_eps / approx_scale : -INFINITY;
                            if (cut2 > -INFINITY) known_above = fmax(unlisted, cut2 * approx_scale + 1.01 * approx_eps);
#pragma unroll
                            for (int e = 0; e < EPL; e++) {
                                bool add = idx_e[e] >= 0 && val_e[e] >= cut2 && !(val_e[e] >= cut);
                                unsigned bm = __ballot_sync(0xffffffffu, add);
                                if (bm) {
                                    rw::Walk w = rw::walk_rows<TB, CRX_COSINE>(xb, ldb, D, add ? (int64_t)idx_e[e] : -1, qvec[warp], tiles[warp]);
                                    if (add) {
                                        X87 ip = {w.a, w.b};
                                        int dst = walked + __popc(bm & ((1u << lane) - 1u));
                                        a_idx[warp][dst] = idx_e[e];
                                        a_sim[warp][dst] = cos_sim_x87(ip, sqn_b[idx_e[e]], nq_);
                                    }
                                    walked += __popc(bm);
                                }
                            }
                            __syncwarp();
                        } else {
                            resolved = false;   // nothing listed behind e*
                        }
                    }
                    if (resolved) {
#pragma unroll
                        for (int e = 0; e < EPL; e++) {
                            int oi = a_idx[warp][e * 32 + lane];
                            double os = a_sim[warp][e * 32 + lane];
                            if (oi > estar && os > known_above) tp = fmax(tp, os);
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) tp = fmax(tp, __shfl_xor_sync(0xffffffffu, tp, off));
#pragma unroll
                        for (int e = 0; e < EPL; e++) {
                            int oi = a_idx[warp][e * 32 + lane];
                            if (oi > estar && a_sim[warp][e * 32 + lane] == tp) rlim = min(rlim, oi);
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) rlim = min(rlim, __shfl_xor_sync(0xffffffffu, rlim, off));
                        // nothing behind e*: R = everything, provided everything has been evaluated
                        resolved = tp > -INFINITY || known_above == -INFINITY;
                    }
                }
                int m_idx[EPL];
                double m_sim[EPL];
#pragma unroll
                for (int e = 0; e < EPL; e++) {
                    m_idx[e] = a_idx[warp][e * 32 + lane];
                    m_sim[e] = a_sim[warp][e * 32 + lane];
                    if (m_idx[e] >= 0 && !(m_sim[e] > known_above)) m_idx[e] = -1;   // outside the exactly known set
                }
                if (resolved) {
                    double* rk = s_pred[warp];
                    int* rv = s_coin[warp];
                    __syncwarp();
                    // R in row order: position = number of members with a smaller row
                    unsigned mem[EPL];
                    int pos[EPL];
                    int m = 0;
#pragma unroll
                    for (int e = 0; e < EPL; e++) {
                        mem[e] = __ballot_sync(0xffffffffu, m_idx[e] >= 0 && m_idx[e] < rlim && m_sim[e] >= tp);
                        pos[e] = 0;
                        m += __popc(mem[e]);
                    }
#pragma unroll
                    for (int f = 0; f < EPL; f++)
                        for (unsigned left = mem[f]; left; left &= left - 1) {
                            int oi = a_idx[warp][f * 32 + __ffs(left) - 1];
#pragma unroll
                            for (int e = 0; e < EPL; e++) pos[e] += oi < m_idx[e];
                        }
#pragma unroll
                    for (int e = 0; e < EPL; e++)
                        if ((mem[e] >> lane) & 1u) { rk[pos[e]] = m_sim[e]; rv[pos[e]] = m_idx[e]; }
                    __syncwarp();
                    warp_lomuto_topn<EPL>(rk, rv, m, keep);
                    __syncwarp();
                    for (int j = lane; j < keep; j += 32) { s_idx[warp][j] = rv[j]; s_sim[warp][j] = rk[j]; }
                    __syncwarp();
                } else {
                    status = CRX_Q_TIE_ORDER;   // kept: descending similarity, ties by row (second pass: the tail behind e*)
                    if (pass2) {
                        int cx;
                        double th = p2_tail_theta(blk, qrel, estar, pth, approx_eps, INFINITY, cx);
                        if (cx != 0x7fffffff && blk.ccode != nullptr) {
                            // no whole column block lies behind e*: the tail is at most a tile or two, so its best similarity
                            // t' is looked up directly (plain FP64, certain to 1e-13) and the query needs ONE collection round
                            // instead of two ("everything behind e*", then everything that reaches t')
                            const uint32_t cq = blk.qcode[qrow];
                            double tm = -INFINITY;
                            for (int64_t r = (int64_t)estar + 1 + lane; r < blk.nb; r += 32) {
                                const uint32_t x = cq ^ blk.ccode[r];
                                if (((x - blk.low) & ~x & blk.high) == 0u) continue;
                                const TB* row = xb + (size_t)r * ldb;
                                double acc = 0.0;
                                for (int k = 0; k < D; k++) acc = fma((double)row[k], qvec[warp][k], acc);
                                const double sr = acc / (sqrt(sqn_b[r]) * sqrt(nq_));
                                if (sr > tm) tm = sr;
                            }
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) tm = fmax(tm, __shfl_xor_sync(0xffffffffu, tm, off));
                            cx = 0x7fffffff;
                            th = tm > -INFINITY ? fmin(tm - 1e-12 - 2.1 * approx_eps, pth - 2.02 * approx_eps) : -INFINITY;   // nothing behind e*: every candidate
                        }
                        if (lane == 0) p2_emit(w2, (int)qrel, th, cx, CRX_Q_TIE_ORDER << P2_KIND_SHIFT);
                    } else if (lane == 0) atomicAdd(&counters[CRX_CNT_TOPP_TIES], 1ull);
                }
            }
        }
    }
    if (qstatus && lane == 0) qstatus[qrel] = status;
    if (nbr_rows) for (int j = lane; j < P; j += 32) nbr_rows[qrel * P + j] = j < keep ? s_idx[warp][j] : -1;
    if (nbr_sims) for (int j = lane; j < P; j += 32) nbr_sims[qrel * P + j] = j < keep ? s_sim[warp][j] : 0.0;
    if (!recs) return;
    if (nc == 0) {  // `if (!neighbors.empty())` (main.cpp:161)
        for (int j = lane; j < Nrec; j += 32) recs[qrel * Nrec + j] = -1;
        return;
    }
    // crypto_rec.hpp:281-345 over the neighbours in their final order
    predict_and_recommend<TB>(xb, ldb, mean_b, unk_q + qrow * D, mean_q[qrow], D, s_idx[warp], s_sim[warp], keep, Nrec, s_pred[warp],
                              s_coin[warp], recs + qrel * Nrec);
}

// ------------------------------------------------------------------------------------------------
// packed table codes and exact candidate counts for the tensor-core path.
// code(row) = concatenation of the L k-bit bucket ids.  |union_l bucket_l(q)| by inclusion-exclusion over
// the 2^L - 1 non-empty table subsets S: sum (-1)^(|S|+1) #{rows agreeing with q on every table of S};
// the agreement counts are histograms over the sub-codes (sizes 2^(k|S|), (1+2^k)^L - 1 bins in total).
// ------------------------------------------------------------------------------------------------
__global__ void pack_codes_kernel(const int32_t* __restrict__ bucket, int64_t stride, int64_t n, int k, int L, uint32_t* __restrict__ code) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = 0;
    for (int l = 0; l < L; l++) c |= (uint32_t)bucket[(size_t)l * stride + i] << (l * k);
    code[i] = c;
}
__device__ __forceinline__ uint32_t subcode(uint32_t code, int S, int k, int L) {
    uint32_t key = 0;
    int pos = 0;
    uint32_t fm = (1u << k) - 1u;
    for (int l = 0; l < L; l++)
        if ((S >> l) & 1) { key |= ((code >> (l * k)) & fm) << pos; pos += k; }
    return key;
}
__global__ void subset_hist_kernel(const uint32_t* __restrict__ code, int64_t n, int k, int L, const int64_t* __restrict__ hoff,
                                   int* __restrict__ hist) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = code[i];
    for (int S = 1; S < (1 << L); S++) atomicAdd(&hist[hoff[S] + subcode(c, S, k, L)], 1);
}
__global__ void sum_counts_kernel(const int32_t* __restrict__ v, int64_t n, unsigned long long* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long x = i < n ? (unsigned long long)v[i] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(out, x);
}
__global__ void subset_count_kernel(const uint32_t* __restrict__ qcode, int64_t q_begin, int64_t nq, int k, int L,
                                    const int64_t* __restrict__ hoff, const int* __restrict__ hist, int32_t* __restrict__ ncand) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    uint32_t c = qcode[q_begin + i];
    long long total = 0;
    for (int S = 1; S < (1 << L); S++) {
        int h = hist[hoff[S] + subcode(c, S, k, L)];
        total += (__popc(S) & 1) ? h : -h;
    }
    ncand[i] = (int32_t)total;
}

// ------------------------------------------------------------------------------------------------
// cluster-neighbour recommendation (crypto_rec.hpp:328-345): similarities to ALL co-members in
// input order, no sort, no cut.  One warp per query; members are consumed 32 at a time (each lane
// computes one exact similarity), then every lane owning a coin accumulates the 32 in order.
// ------------------------------------------------------------------------------------------------
template <typename TQ, typename TB>
__global__ void __launch_bounds__(256)
rec_cluster_kernel(const TQ* __restrict__ xq, int ldq, const double* __restrict__ sqn_q, const uint8_t* __restrict__ unk_q,
                   const double* __restrict__ mean_q, const int32_t* __restrict__ qlabels, const TB* __restrict__ xb, int ldb,
                   const double* __restrict__ sqn_b, const double* __restrict__ mean_b, const int32_t* __restrict__ perm,
                   const int32_t* __restrict__ off, int K, int D, int64_t nq, int Nrec, int32_t* __restrict__ recs) {
    __shared__ int s_idx[8][32];
    __shared__ double s_sim[8][32];
    __shared__ double s_pred[8][128];
    __shared__ int s_coin[8][128];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t qrow = (int64_t)blockIdx.x * 8 + warp;
    if (qrow >= nq) return;
    int cl = qlabels[qrow];
    const bool known = cl >= 0 && cl < K;   // a label outside [0, K) has no cluster: as an empty one
    int begin = known ? off[cl] : 0, end = known ? off[cl + 1] : 0;
    if (begin == end) {
        for (int j = lane; j < Nrec; j += 32) recs[qrow * Nrec + j] = -1;
        return;
    }
    const TQ* q = xq + qrow * ldq;
    double nqv = sqn_q[qrow];
    double main_sum[4] = {0, 0, 0, 0}, abs_sum = 0.0;  // coins lane, lane+32, lane+64, lane+96
    for (int base = begin; base < end; base += 32) {
        int p = base + lane;
        int nb = p < end ? perm[p] : -1;
        s_idx[warp][lane] = nb;
        s_sim[warp][lane] = nb >= 0 ? cos_sim_exact(xb + (size_t)nb * ldb, q, D, sqn_b[nb], nqv) : 0.0;
        __syncwarp();
        int lim = min(32, end - base);
        for (int i = 0; i < lim; i++) {
            double s = s_sim[warp][i];
            int r = s_idx[warp][i];
            abs_sum = __dadd_rn(abs_sum, fabs(s));
            double mb = mean_b[r];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int j = lane + 32 * t;
                if (j < D) main_sum[t] = __dadd_rn(main_sum[t], __dmul_rn(s, __dsub_rn((double)xb[(size_t)r * ldb + j], mb)));
            }
        }
        __syncwarp();
    }
    double mq = mean_q[qrow];
    int nu = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int j = lane + 32 * t;
        bool u = j < D && unk_q[qrow * D + j] != 0;
        unsigned um = __ballot_sync(0xffffffffu, u);
        if (u) {
            int slot = nu + __popc(um & ((1u << lane) - 1));
            s_pred[warp][slot] = __dadd_rn(__ddiv_rn(main_sum[t], abs_sum), mq);
            s_coin[warp][slot] = j;
        }
        nu += __popc(um);
    }
    __syncwarp();
    warp_lomuto_topn(s_pred[warp], s_coin[warp], nu, Nrec);
    for (int j = lane; j < Nrec; j += 32) recs[qrow * Nrec + j] = j < nu ? s_coin[warp][j] : 0;
}

// get_predicted_user_sim (crypto_rec.hpp:281-306) + get_top_N_recom (:310-345) for ONE user and an explicit
// neighbour list (rows of `xb`, in the given order).  sims == NULL: similarities to all neighbours are computed
// first (crypto_rec.hpp:331-333).  One warp.
template <typename TQ, typename TB>
__global__ void rec_list_kernel(const TQ* __restrict__ q, int D, const double* __restrict__ sqn_q_p, const uint8_t* __restrict__ unk_q, const double* __restrict__ mean_q_p,
                                const TB* __restrict__ xb, int ldb, const double* __restrict__ sqn_b, const double* __restrict__ mean_b,
                                const int32_t* __restrict__ nbr, const double* __restrict__ sims_in, int n, int Nrec,
                                double* __restrict__ predicted /* [D] or NULL */, int32_t* __restrict__ recs /* [Nrec] or NULL */) {
    __shared__ int s_idx[32];
    __shared__ double s_sim[32];
    __shared__ double s_pred[128];
    __shared__ int s_coin[128];
    int lane = threadIdx.x & 31;
    const double sqn_q = *sqn_q_p, mean_q = *mean_q_p;
    double main_sum[4] = {0, 0, 0, 0}, abs_sum = 0.0;
    for (int base = 0; base < n; base += 32) {
        int p = base + lane;
        int nb = p < n ? nbr[p] : -1;
        s_idx[lane] = nb;
        s_sim[lane] = nb >= 0 ? (sims_in ? sims_in[p] : cos_sim_exact(xb + (size_t)nb * ldb, q, D, sqn_b[nb], sqn_q)) : 0.0;
        __syncwarp();
        int lim = min(32, n - base);
        for (int i = 0; i < lim; i++) {
            double s = s_sim[i];
            int r = s_idx[i];
            abs_sum = __dadd_rn(abs_sum, fabs(s));
            double mb = mean_b[r];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int j = lane + 32 * t;
                if (j < D) main_sum[t] = __dadd_rn(main_sum[t], __dmul_rn(s, __dsub_rn((double)xb[(size_t)r * ldb + j], mb)));
            }
        }
        __syncwarp();
    }
    int nu = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int j = lane + 32 * t;
        bool u = j < D && unk_q[j] != 0;
        double pr = u ? __dadd_rn(__ddiv_rn(main_sum[t], abs_sum), mean_q) : (j < D ? (double)q[j] : 0.0);
        if (predicted && j < D) predicted[j] = pr;
        unsigned um = __ballot_sync(0xffffffffu, u);
        if (u) {
            int slot = nu + __popc(um & ((1u << lane) - 1));
            s_pred[slot] = pr;
            s_coin[slot] = j;
        }
        nu += __popc(um);
    }
    __syncwarp();
    if (recs) {
        warp_lomuto_topn(s_pred, s_coin, nu, Nrec);
        for (int j = lane; j < Nrec; j += 32) recs[j] = j < nu ? s_coin[j] : 0;
    }
}

// The literal Lomuto quicksort is one sequential walk; it runs on one thread, in shared memory when the arrays fit (12
// bytes per element: up to ~18k neighbours), and stops refining ranges that lie beyond the first `need` positions.
__global__ void __launch_bounds__(256)
quicksort_kernel(double* sims, int32_t* ids, int n, int need, int in_smem) {
    extern __shared__ double qs_keys[];
    if (in_smem) {
        int32_t* qs_ids = reinterpret_cast<int32_t*>(qs_keys + n);
        for (int i = threadIdx.x; i < n; i += blockDim.x) { qs_keys[i] = sims[i]; qs_ids[i] = ids[i]; }
        __syncthreads();
        if (threadIdx.x == 0) lomuto_desc(qs_keys, qs_ids, n, need);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) { sims[i] = qs_keys[i]; ids[i] = qs_ids[i]; }
    } else if (threadIdx.x == 0) {
        lomuto_desc(sims, ids, n, need);
    }
}
// similarity of every listed neighbour to one user (crypto_rec.hpp:219-220): neighbors[i]->cosineSimilarity(&user), i.e. the
// neighbour is the left operand of cust_vector.hpp:160-174
template <typename TQ, typename TB>
__global__ void list_sims_kernel(const TQ* __restrict__ q, const double* __restrict__ sqn_q, const TB* __restrict__ xb, int ldb,
                                 const double* __restrict__ sqn_b, const int32_t* __restrict__ nbr, int64_t n, int D, double* __restrict__ sims) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t r = nbr[i];
    sims[i] = cos_sim_exact(xb + (int64_t)r * ldb, q, D, sqn_b[r], *sqn_q);
}

static int launch_quicksort(crx_ctx* c, double* d_sims, int32_t* d_ids, int n, int need) {
    size_t smem = (size_t)n * (sizeof(double) + sizeof(int32_t));
    int in_smem = smem <= 200 * 1024;
    if (in_smem) CRX_CUDA(cudaFuncSetAttribute(quicksort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CRX_KERNEL(c, "quicksort");
    quicksort_kernel<<<1, 256, in_smem ? smem : 0, c->stream>>>(d_sims, d_ids, n, need, in_smem);
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

// one warp: the warp-parallel form of the literal quicksort (recommend_pass2.cuh) on arrays in global memory
__global__ void __launch_bounds__(32) warp_qs_kernel(double* sims, int32_t* ids, int n, int need) {
    __shared__ int posge[128];
    __shared__ double hk[128];
    __shared__ int hv[128];
    __shared__ double sk[128];
    __shared__ int sv[128];
    __shared__ int gp[128];
    warp_qs_topn_big(sims, ids, n, need, QsScratch{posge, hk, hv, sk, sv, gp});
}

// group id of an external query in one Euclidean table: the stored rows of its bucket are walked in insertion order; the
// first one with the query's k-tuple of h values gives its group; `none` (= the number of groups: a key no stored row has, and
// still a valid sort key) when no stored row shares the tuple
__global__ void query_groups_kernel(const int32_t* __restrict__ qh, const int32_t* __restrict__ qb, int64_t nq, int k,
                                    const int32_t* __restrict__ hvals, const int32_t* __restrict__ gid, const int32_t* __restrict__ perm,
                                    const int32_t* __restrict__ off, int nbuckets, int none, int32_t* __restrict__ qgid) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int b = qb[q];
    int g = none;
    if (b >= 0 && b < nbuckets) {
        for (int pos = off[b]; pos < off[b + 1] && g == none; pos++) {
            const int r = perm[pos];
            bool same = true;
            for (int j = 0; j < k; j++) same = same && hvals[(size_t)r * k + j] == qh[(size_t)q * k + j];
            if (same) g = gid[r];
        }
    }
    qgid[q] = g;
}

__global__ void fill_lists_kernel(double* s, int32_t* i, int32_t* nc, int64_t nq) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nq * LIST) { s[t] = -INFINITY; i[t] = -1; }
    if (t < nq) nc[t] = 0;
}

__global__ void iota_off_kernel(int32_t* p, int64_t n, int32_t off) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i + off;
}

__global__ void add_off_kernel(int32_t* p, int64_t n, int32_t off) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += off;
}

// sum over groups of (group size)^2, as double: cost model for the per-table passes
__global__ void sq_sizes_kernel(const int32_t* __restrict__ off, int ngroups, double* __restrict__ out) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (g < ngroups) { double n = (double)(off[g + 1] - off[g]); v = n * n; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(out, v);
}

// ------------------------------------------------------------------------------------------------
// second pass of the batched top-P: rounds of collect -> evaluate -> resolve over the queued queries
// (recommend_pass2.cuh) until every query carries the reference's list
// ------------------------------------------------------------------------------------------------
// CRX_P2_DEBUG: wall-clock marks (stream synchronised) through one crx_recommend_lsh call; CRX_P2_DEBUG=host: the same marks
// without the extra synchronisations, i.e. where the HOST spends its time (its own waits included)
struct DebugClock {
    bool on, sync; cudaStream_t s; double t0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    DebugClock(cudaStream_t st) : on(getenv("CRX_P2_DEBUG") != nullptr), sync(on && strcmp(getenv("CRX_P2_DEBUG"), "host") != 0), s(st), t0(0) {
        if (on) { if (sync) cudaStreamSynchronize(s); t0 = now(); }
    }
    void mark(const char* what) { if (!on) return; if (sync) cudaStreamSynchronize(s); double t = now(); fprintf(stderr, "[crx clock] %-28s %8.2f ms\n", what, t - t0); t0 = t; }
};

struct P2Host {
    const crx_points* base; const crx_points* queries;
    int64_t q_begin, nq;
    int P, Nrec;
    const int32_t* nc;
    bool use_tc;
    const TcOperand* opQ; const TcOperand* opB;        // tensor path: query / base operands
    const uint32_t* qcode; const uint32_t* ccode; int k, L; bool dense;
    const int32_t* qgid; int64_t qstride; const int32_t* cgid; int64_t cstride;   // SIMT path: group ids
    double eps, unscale;
    const double* eps_q;   // per query row (centred operands) or NULL
    size_t free_bytes;     // device memory that was free when the first pass had been enqueued
    P2Blocks blocks;
    int32_t* recs; int32_t* rows; double* sims; int32_t* status;
};

struct P2QueueBuf {
    DevBuf<int32_t> q, colx, tries;
    DevBuf<double> theta;
    DevBuf<unsigned int> count;
    int alloc(crx_ctx* c, size_t n) {
        CRX_TRY(q.alloc(c, n)); CRX_TRY(colx.alloc(c, n)); CRX_TRY(tries.alloc(c, n)); CRX_TRY(theta.alloc(c, n)); CRX_TRY(count.alloc(c, 1));
        CRX_CUDA(cudaMemsetAsync(count.p, 0, sizeof(unsigned int), c->stream));
        return CRX_OK;
    }
    P2Queue view() const { P2Queue w; w.q = q.p; w.theta = theta.p; w.colx = colx.p; w.tries = tries.p; w.count = count.p; return w; }
};

__global__ void p2_trim_kernel(const int64_t* __restrict__ off, unsigned int n, int64_t limit, int32_t* __restrict__ ovf) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && off[i + 1] > limit) ovf[i] = 1;
}

static int p2_scan_sizes(crx_ctx* c, const int32_t* count, const int32_t* ovf, unsigned int n, DevBuf<int64_t>& seg, DevBuf<int64_t>& off,
                         DevBuf<char>& tmp, size_t& tmp_bytes, int64_t* h_total) {
    { CRX_KERNEL(c, "p2_sizes"); p2_sizes_kernel<<<crx_grid((int64_t)n + 1, 256), 256, 0, c->stream>>>(count, ovf, n, seg.p); }
    size_t need = 0;
    CRX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, seg.p, off.p, (int)(n + 1), c->stream));
    if (need > tmp_bytes) { CRX_TRY(tmp.alloc(c, need)); tmp_bytes = need; }
    CRX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, seg.p, off.p, (int)(n + 1), c->stream));
    CRX_CUDA(cudaMemcpyAsync(h_total, off.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}

static int pass2_run(crx_ctx* c, const P2Host& h, P2QueueBuf* qa, P2QueueBuf* qb) {
    const crx_points* base = h.base;
    const crx_points* queries = h.queries;
    const int64_t N = base->n;
    P2QueueBuf* cur = qa;
    P2QueueBuf* nxt = qb;
    unsigned int n = 0;
    CRX_CUDA(cudaMemcpyAsync(&n, cur->count.p, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    static const bool p2_debug = getenv("CRX_P2_DEBUG") != nullptr;
    DebugClock clk(c->stream);
    // rows whose coordinates are all equal (single-coin users): their mutual similarities have a closed form (p2_exact_kernel)
    DevBuf<double> uval_b, uval_q;
    if (n > 0) {
        CRX_KERNEL(c, "uniform_rows");
        CRX_TRY(uval_b.alloc(c, (size_t)N));
        if (base->x64) uniform_rows_kernel<double><<<crx_grid(N, 8), 256, 0, c->stream>>>(base->x64, base->ld, base->d, N, uval_b.p);
        else uniform_rows_kernel<float><<<crx_grid(N, 8), 256, 0, c->stream>>>(base->x32, base->ld, base->d, N, uval_b.p);
        if (queries != base) {
            CRX_TRY(uval_q.alloc(c, (size_t)queries->n));
            if (queries->x64) uniform_rows_kernel<double><<<crx_grid(queries->n, 8), 256, 0, c->stream>>>(queries->x64, queries->ld, queries->d, queries->n, uval_q.p);
            else uniform_rows_kernel<float><<<crx_grid(queries->n, 8), 256, 0, c->stream>>>(queries->x32, queries->ld, queries->d, queries->n, uval_q.p);
        }
        CRX_CUDA(cudaGetLastError());
    }
    const double* d_uval_q = queries != base ? uval_q.p : uval_b.p;
    // ---- constant query rows against the constant rows of the table (recommend_pass2.cuh, P2Uniform): answered without lists
    static const bool uni_off = getenv("CRX_P2_UNIFORM") != nullptr && getenv("CRX_P2_UNIFORM")[0] == '0';
    if (n > 0 && h.use_tc && !uni_off && N < (int64_t)INT_MAX) {
        DevBuf<uint8_t> flag;
        DevBuf<int> refuse, nsel;
        DevBuf<int32_t> mrow;
        CRX_TRY(flag.alloc(c, (size_t)N)); CRX_TRY(refuse.alloc(c, 1)); CRX_TRY(nsel.alloc(c, 1)); CRX_TRY(mrow.alloc(c, (size_t)N));
        CRX_CUDA(cudaMemsetAsync(refuse.p, 0, sizeof(int), c->stream));
        {
            CRX_KERNEL(c, "p2u_flag");
            if (base->x64) p2u_flag_kernel<double><<<crx_grid(N, 8), 256, 0, c->stream>>>(base->x64, base->ld, base->d, base->sqn, uval_b.p, N, flag.p, refuse.p);
            else p2u_flag_kernel<float><<<crx_grid(N, 8), 256, 0, c->stream>>>(base->x32, base->ld, base->d, base->sqn, uval_b.p, N, flag.p, refuse.p);
        }
        {
            size_t bytes = 0;
            cub::CountingInputIterator<int32_t> rows_it(0);
            CRX_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, rows_it, flag.p, mrow.p, nsel.p, (int)N, c->stream));
            DevBuf<char> tmp;
            CRX_TRY(tmp.alloc(c, bytes));
            CRX_CUDA(cub::DeviceSelect::Flagged(tmp.p, bytes, rows_it, flag.p, mrow.p, nsel.p, (int)N, c->stream));
        }
        int h_info[2] = {0, 0};
        CRX_CUDA(cudaMemcpyAsync(&h_info[0], nsel.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaMemcpyAsync(&h_info[1], refuse.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        const int nm = h_info[0];
        if (nm >= h.P && h_info[1] == 0) {
            DevBuf<double> mval, mroot, tp;
            DevBuf<uint32_t> mcode;
            DevBuf<int32_t> need, rlim;
            CRX_TRY(mval.alloc(c, nm)); CRX_TRY(mroot.alloc(c, nm)); CRX_TRY(mcode.alloc(c, nm));
            CRX_TRY(need.alloc(c, n)); CRX_TRY(rlim.alloc(c, n)); CRX_TRY(tp.alloc(c, n));
            { CRX_KERNEL(c, "p2u_gather"); p2u_gather_kernel<<<crx_grid(nm, 256), 256, 0, c->stream>>>(mrow.p, nm, uval_b.p, base->sqn, h.ccode, mval.p, mroot.p, mcode.p); }
            P2Uniform u;
            u.mrow = mrow.p; u.mval = mval.p; u.mroot = mroot.p; u.mcode = mcode.p; u.nm = nm; u.qcode = h.qcode;
            u.low = 0; u.high = 0;
            for (int l = 0; l < h.L; l++) { u.low |= 1u << (l * h.k); u.high |= 1u << (l * h.k + h.k - 1); }
            u.uval_q = d_uval_q; u.sqn_q = queries->sqn;
            P2Resolve a;
            memset(&a, 0, sizeof(a));
            a.unk_q = queries->unknown; a.mean_q = queries->mean; a.mean_b = base->mean;
            a.ldb = base->ld; a.D = base->d; a.P = h.P; a.Nrec = h.Nrec; a.q_begin = h.q_begin; a.ncand = h.nc;
            a.cur = cur->view(); a.n = n; a.eps = h.eps; a.eps_q = h.eps_q;
            a.recs = h.recs; a.nbr_rows = h.rows; a.nbr_sims = h.sims; a.qstatus = h.status; a.counters = c->counters;
#define LAUNCH_U(KERNEL, TQ, TB, xqp, xbp, ...) KERNEL<TQ, TB><<<crx_grid(n, 4), 128, 0, c->stream>>>(xqp, queries->ld, xbp, a, u, __VA_ARGS__)
#define DISPATCH_U(KERNEL, ...)                                                                                  \
    do {                                                                                                         \
        if (queries->x64 && base->x64) LAUNCH_U(KERNEL, double, double, queries->x64, base->x64, __VA_ARGS__);    \
        else if (queries->x64) LAUNCH_U(KERNEL, double, float, queries->x64, base->x32, __VA_ARGS__);             \
        else if (base->x64) LAUNCH_U(KERNEL, float, double, queries->x32, base->x64, __VA_ARGS__);                \
        else LAUNCH_U(KERNEL, float, float, queries->x32, base->x32, __VA_ARGS__);                                \
    } while (0)
            { CRX_KERNEL(c, "p2u_select"); DISPATCH_U(p2u_select_kernel, need.p, rlim.p, tp.p); }
            CRX_CUDA(cudaGetLastError());
            // phase B for the queries whose R mixes similarities: scratch sized by a scan of |R|
            DevBuf<int64_t> seg, off;
            DevBuf<char> tmp;
            size_t tmp_bytes = 0;
            int64_t total = 0;
            CRX_TRY(seg.alloc(c, (size_t)n + 1)); CRX_TRY(off.alloc(c, (size_t)n + 1));
            CRX_TRY(p2_scan_sizes(c, need.p, nullptr, n, seg, off, tmp, tmp_bytes, &total));
            if (total > 0) {
                DevBuf<double> keys;
                DevBuf<int32_t> vals;
                CRX_TRY(keys.alloc(c, (size_t)total)); CRX_TRY(vals.alloc(c, (size_t)total));
                { CRX_KERNEL(c, "p2u_sort"); DISPATCH_U(p2u_sort_kernel, need.p, rlim.p, tp.p, off.p, keys.p, vals.p); }
                CRX_CUDA(cudaGetLastError());
            }
#undef DISPATCH_U
#undef LAUNCH_U
            CRX_CUDA(cudaMemsetAsync(nxt->count.p, 0, sizeof(unsigned int), c->stream));
            { CRX_KERNEL(c, "p2u_compact"); p2u_compact_kernel<<<crx_grid(n, 256), 256, 0, c->stream>>>(cur->view(), n, need.p, nxt->view()); }
            std::swap(cur, nxt);
            const unsigned int n_before = n;
            CRX_CUDA(cudaMemcpyAsync(&n, cur->count.p, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
            CRX_CUDA(cudaStreamSynchronize(c->stream));
            if (p2_debug) fprintf(stderr, "[crx pass2] constant rows: |M| = %d, %u of %u queued queries answered (%lld keys sorted in phase B)\n", nm, n_before - n, n_before, (long long)total);
        }
    }
    static const int max_rounds = getenv("CRX_P2_ROUNDS") ? atoi(getenv("CRX_P2_ROUNDS")) : 16;
    clk.mark("  constant rows");
    DevBuf<unsigned long long> dbg;
    if (p2_debug) { CRX_TRY(dbg.alloc(c, 8)); }
    for (int round = 0; n > 0 && round < max_rounds; round++) {
        CRX_CUDA(cudaMemsetAsync(nxt->count.p, 0, sizeof(unsigned int), c->stream));
        if (p2_debug) CRX_CUDA(cudaMemsetAsync(dbg.p, 0, 8 * sizeof(unsigned long long), c->stream));
        // (cudaMemGetInfo costs 15-50 ms of host time per call with a populated memory pool -- and the GPU idles meanwhile: the
        // query is made once, while the first pass runs, see recommend_lsh_impl; a round frees what it allocates)
        size_t free_b = h.free_bytes;
        for (auto& b : c->big_free) free_b += b.second;   // blocks the context keeps between calls are available to this round
        // budget for the collected lists: one third for the pass masks of the threshold scan (32 B per row and column tile), the
        // rest for rows + similarities (12 B per entry).  A queue that needs more is processed in slices.
        const int64_t budget = (int64_t)std::max<size_t>(free_b / 2, (size_t)64 << 20);
        unsigned int slice = n;
        if (h.use_tc) {
            const int64_t per_row = crx_tc_collect_mask_words(128, N) / 128 * 4;
            slice = (unsigned int)std::max<int64_t>(128, std::min<int64_t>(n, (budget / 3 / per_row) / 128 * 128));
        }
        int64_t collected = 0;
        for (unsigned int s0 = 0; s0 < n; s0 += slice) {
        const unsigned int ns = std::min(slice, n - s0);
        const bool tc_round = h.use_tc;
        DevBuf<int32_t> count, ovf, qrow_abs, cols, allf, colx_eff;
        DevBuf<uint32_t> cmask;
        DevBuf<float> theta_f;
        DevBuf<int64_t> seg, off;
        DevBuf<char> tmp;
        DevBuf<double> xs;
        size_t tmp_bytes = 0;
        int64_t total = 0;
        CRX_TRY(count.alloc(c, ns)); CRX_TRY(ovf.alloc(c, ns)); CRX_TRY(seg.alloc(c, (size_t)ns + 1)); CRX_TRY(off.alloc(c, (size_t)ns + 1));
        CRX_CUDA(cudaMemsetAsync(ovf.p, 0, (size_t)ns * sizeof(int32_t), c->stream));
        CRX_CUDA(cudaMemsetAsync(count.p, 0, (size_t)ns * sizeof(int32_t), c->stream));
        P2Queue w = cur->view();
        w.q += s0; w.theta += s0; w.colx += s0; w.tries += s0;
        int64_t rows_pad = 0;
        if (tc_round) {
            CRX_TRY(qrow_abs.alloc(c, ns)); CRX_TRY(theta_f.alloc(c, ns)); CRX_TRY(allf.alloc(c, ns)); CRX_TRY(colx_eff.alloc(c, ns));
            { CRX_KERNEL(c, "p2_prepare"); p2_prepare_kernel<<<crx_grid(ns, 256), 256, 0, c->stream>>>(w, ns, h.q_begin, 1.0 / h.unscale, theta_f.p, qrow_abs.p, allf.p, colx_eff.p); }
            TcOperand opA;
            clk.mark("  round buffers");
            CRX_TRY(crx_tc_gather(c, *h.opQ, qrow_abs.p, ns, &opA));
            clk.mark("  gather operand");
            CRX_TRY(cmask.alloc(c, (size_t)crx_tc_collect_mask_words(ns, N)));
            clk.mark("  mask alloc");
            rows_pad = ((int64_t)ns + 127) / 128 * 128;
            CRX_TRY(crx_tc_collect(c, opA, ns, *h.opB, h.qcode, qrow_abs.p, h.ccode, h.k, h.L, h.dense, theta_f.p, colx_eff.p, cmask.p, count.p));
            { CRX_KERNEL(c, "p2_all_count"); p2_all_count_kernel<<<crx_grid(ns, 256), 256, 0, c->stream>>>(allf.p, ns, w.q, h.nc, count.p, ovf.p); }
            clk.mark("  collect");
        } else {
            CRX_KERNEL(c, "p2_collect_simt");
#define LAUNCH_S(TQ, TB, xqp, xbp, FILL)                                                                                          \
    p2_collect_simt_kernel<TQ, TB, FILL><<<ns, 256, 0, c->stream>>>(xqp, queries->ld, queries->sqn, xbp, base->ld, base->sqn, base->d, N, \
        h.q_begin, w, h.L, h.qgid, h.qstride, h.cgid, h.cstride, count.p, ovf.p, off.p, cols.p)
            if (queries->x64 && base->x64) LAUNCH_S(double, double, queries->x64, base->x64, false);
            else if (queries->x64) LAUNCH_S(double, float, queries->x64, base->x32, false);
            else if (base->x64) LAUNCH_S(float, double, queries->x32, base->x64, false);
            else LAUNCH_S(float, float, queries->x32, base->x32, false);
            CRX_CUDA(cudaGetLastError());
        }
        CRX_TRY(p2_scan_sizes(c, count.p, ovf.p, ns, seg, off, tmp, tmp_bytes, &total));
        const int64_t list_budget = budget - (int64_t)cmask.count * 4;
        if (total * 12 > list_budget) {   // too many rows for this round: the queries behind the budget wait for the next one
            { CRX_KERNEL(c, "p2_trim"); p2_trim_kernel<<<crx_grid(ns, 256), 256, 0, c->stream>>>(off.p, ns, list_budget / 12, ovf.p); }
            CRX_TRY(p2_scan_sizes(c, count.p, ovf.p, ns, seg, off, tmp, tmp_bytes, &total));
        }
        collected += total;
        clk.mark("  sizes");
        CRX_TRY(cols.alloc(c, (size_t)total)); CRX_TRY(xs.alloc(c, (size_t)total));
        clk.mark("  alloc lists");
        if (tc_round) {
            {
                CRX_KERNEL(c, "p2_expand");
                const int tn = crx_tc_tile_cols();
                p2_expand_kernel<<<(unsigned int)(rows_pad / 32), 256, 0, c->stream>>>(cmask.p, rows_pad, (int)((N + tn - 1) / tn), tn, allf.p, count.p, ovf.p, off.p, ns, cols.p);
                CRX_CUDA(cudaGetLastError());
            }
            cmask.release();
            {
                uint32_t low = 0, high = 0;
                for (int l = 0; l < h.L; l++) { low |= 1u << (l * h.k); high |= 1u << (l * h.k + h.k - 1); }
                CRX_KERNEL(c, "p2_fill_all");
                p2_fill_all_kernel<<<ns, 256, 0, c->stream>>>(allf.p, ovf.p, qrow_abs.p, h.qcode, h.ccode, N, low, high, off.p, cols.p);
                CRX_CUDA(cudaGetLastError());
            }
        } else {
            CRX_KERNEL(c, "p2_collect_simt");
            // (a trimmed query writes nothing: its segment is empty)
            if (queries->x64 && base->x64) LAUNCH_S(double, double, queries->x64, base->x64, true);
            else if (queries->x64) LAUNCH_S(double, float, queries->x64, base->x32, true);
            else if (base->x64) LAUNCH_S(float, double, queries->x32, base->x64, true);
            else LAUNCH_S(float, float, queries->x32, base->x32, true);
#undef LAUNCH_S
            CRX_CUDA(cudaGetLastError());
        }
        clk.mark("  expand");
        if (total > 0) {
            CRX_KERNEL(c, "p2_exact");
            const int64_t nblocks = total / 32;
            int grid = (int)std::min<int64_t>(((nblocks + P2_RUN - 1) / P2_RUN + 7) / 8, (int64_t)c->sm_count * 16);
#define LAUNCH_E(TQ, TB, xqp, xbp) \
    p2_exact_kernel<TQ, TB><<<grid, 256, 0, c->stream>>>(xqp, queries->ld, queries->sqn, xbp, base->ld, base->sqn, base->d, h.q_begin, w.q, off.p, ns, nblocks, cols.p, xs.p, d_uval_q, uval_b.p)
            if (queries->x64 && base->x64) LAUNCH_E(double, double, queries->x64, base->x64);
            else if (queries->x64) LAUNCH_E(double, float, queries->x64, base->x32);
            else if (base->x64) LAUNCH_E(float, double, queries->x32, base->x64);
            else LAUNCH_E(float, float, queries->x32, base->x32);
#undef LAUNCH_E
            CRX_CUDA(cudaGetLastError());
        }
        clk.mark("  exact");
        {
            CRX_KERNEL(c, "p2_resolve");
            P2Resolve a;
            memset(&a, 0, sizeof(a));
            a.unk_q = queries->unknown; a.mean_q = queries->mean; a.mean_b = base->mean;
            a.ldb = base->ld; a.D = base->d; a.P = h.P; a.Nrec = h.Nrec; a.q_begin = h.q_begin; a.ncand = h.nc;
            a.cur = w; a.next = nxt->view(); a.n = ns;
            a.off = off.p; a.count = count.p; a.ovf = ovf.p; a.cols = cols.p; a.xs = xs.p;
            a.eps = h.eps; a.eps_q = h.eps_q; a.blocks = h.blocks;
            a.recs = h.recs; a.nbr_rows = h.rows; a.nbr_sims = h.sims; a.qstatus = h.status; a.counters = c->counters;
            a.dbg = p2_debug ? dbg.p : nullptr;
            if (base->x64) p2_resolve_kernel<double><<<crx_grid(ns, 4), 128, 0, c->stream>>>(base->x64, a);
            else p2_resolve_kernel<float><<<crx_grid(ns, 4), 128, 0, c->stream>>>(base->x32, a);
            CRX_CUDA(cudaGetLastError());
        }
        clk.mark("  resolve");
        }   // slices
        std::swap(cur, nxt);
        const unsigned int n_in = n;
        CRX_CUDA(cudaMemcpyAsync(&n, cur->count.p, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        if (p2_debug) {
            unsigned long long hd[8];
            CRX_CUDA(cudaMemcpy(hd, dbg.p, sizeof(hd), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[crx pass2] round %d: %u queued (slices of %u), %lld collected (%.1f per query), %u left; tie-order: %llu resolved (%llu by H alone), sum |R| = %llu, "
                            "lists = %llu; plateau: %llu resolved (%llu by H alone), sum |R| = %llu, lists = %llu\n",
                    round, n_in, slice, (long long)collected, (double)collected / n_in, n, hd[2], hd[3], hd[0], hd[1], hd[6], hd[7], hd[4], hd[5]);
        }
    }
    if (n > 0) {
        CRX_KERNEL(c, "p2_leftover");
        p2_leftover_kernel<<<crx_grid(n, 256), 256, 0, c->stream>>>(cur->view(), n, c->counters);
        CRX_CUDA(cudaGetLastError());
    }
    return CRX_OK;
}

extern "C" {

static int recommend_lsh_impl(crx_ctx* c, const crx_lsh* t, const crx_points* queries, int64_t q_begin, int64_t q_end, int P,
                              int Nrec, int32_t* recs, int32_t* nbr_rows, double* nbr_sims, int32_t* ncand, int32_t* status, int mem) {
    CRX_REQUIRE(c && t, "NULL argument");
    const crx_points* base = t->pts;
    bool self = queries == nullptr || queries == base;
    if (!queries) queries = base;
    CRX_REQUIRE(queries->d == base->d, "query / table dimension mismatch");
    CRX_REQUIRE(q_begin >= 0 && q_begin <= q_end && q_end <= queries->n, "query range");
    CRX_REQUIRE(P >= 1 && P <= P2_MAXP, "P must be in [1, 64]");
    CRX_REQUIRE(Nrec >= 0 && Nrec <= 128, "Nrec");
    if (recs) CRX_REQUIRE(base->mean && queries->unknown && queries->mean, "ratings metadata missing (crx_points_set_ratings)");
    CRX_NARROW(queries);
    CRX_CUDA(cudaSetDevice(c->device));
    int64_t nq = q_end - q_begin;
    if (nq == 0) return CRX_OK;
    DebugClock clk(c->stream);
    int L = t->L;
    int64_t N = base->n;

    // group ids of the query rows
    DevBuf<int32_t> qgid_buf;
    const int32_t* qgid_base = t->gid;
    int64_t qstride = N;
    if (!self) {
        CRX_TRY(qgid_buf.alloc(c, (size_t)L * queries->n));
        if (t->metric == CRX_COSINE) {
            CRX_TRY(crx_hash_rows(c, queries, t->metric, t->k, L, t->d_proj, t->ldp, t->d_pnorm, t->d_t, t->d_r, t->w, t->nbuckets, nullptr, qgid_buf.p));
        } else {
            // Euclidean tables filter a bucket by the k-tuple of h values (cust_hashtable.hpp:81-97): an external query takes the
            // group of any stored row of its bucket that has its tuple, or none
            DevBuf<int32_t> qh, qb;
            CRX_TRY(qh.alloc(c, (size_t)L * queries->n * t->k)); CRX_TRY(qb.alloc(c, (size_t)L * queries->n));
            CRX_TRY(crx_hash_rows(c, queries, t->metric, t->k, L, t->d_proj, t->ldp, t->d_pnorm, t->d_t, t->d_r, t->w, t->nbuckets, qh.p, qb.p));
            for (int l = 0; l < L; l++) {
                CRX_KERNEL(c, "query_groups");
                query_groups_kernel<<<crx_grid(queries->n, 128), 128, 0, c->stream>>>(qh.p + (size_t)l * queries->n * t->k, qb.p + (size_t)l * queries->n, queries->n, t->k,
                    t->hvals + (size_t)l * N * t->k, t->gid + (size_t)l * N, t->by_bucket[l].perm, t->by_bucket[l].off, t->nbuckets, t->ngroups[l], qgid_buf.p + (size_t)l * queries->n);
            }
            CRX_CUDA(cudaGetLastError());
            CRX_CUDA(cudaStreamSynchronize(c->stream));   // qh / qb are released on leaving this scope
        }
        qgid_base = qgid_buf.p;
        qstride = queries->n;
    }
    DevBuf<double> list_s;
    DevBuf<int32_t> list_i, nc;
    DevBuf<float> tl_s, blockmax;
    DevBuf<int32_t> tl_i;
    DevBuf<uint32_t> ccode, qcode_buf;     // packed table codes: also read by the second pass
    const uint32_t* qcode = nullptr;
    TcOperand opB, opA;                    // split-fp16 operands: kept until the second pass is through
    bool dense_cols = false;
    double tc_unscale = 1.0;
    DevBuf<double> eps_b, eps_a;           // centred operands: the filter's error bound per base / query row
    const double* eps_rows = nullptr;
    CRX_TRY(nc.alloc(c, nq));
    // second, targeted pass for the queries whose list does not decide the reference's order (recommend_pass2.cuh);
    // CRX_TOPP_EXACT=0 leaves them counted instead (counters [1] and [5])
    static const int pass2 = !(getenv("CRX_TOPP_EXACT") != nullptr && getenv("CRX_TOPP_EXACT")[0] == '0');
    P2QueueBuf p2a, p2b;
    P2Blocks blk;
    memset(&blk, 0, sizeof(blk));
    if (pass2) { CRX_TRY(p2a.alloc(c, (size_t)nq)); CRX_TRY(p2b.alloc(c, (size_t)nq)); }
    // Tensor-core filter (tcgen05, tc_scan.cu): cosine tables whose L k-bit bucket ids pack into 32 bits and
    // whose sub-code histograms stay small.  CRX_NO_TC=1 forces the FP64 scan (tests compare the two paths).
    static const bool tc_off = getenv("CRX_NO_TC") != nullptr && getenv("CRX_NO_TC")[0] == '1';
    // CRX_TIE_ORDER=0: skip the reconstruction of the reference's order among equal similarities (they are then counted)
    static const int tie_order = !(getenv("CRX_TIE_ORDER") != nullptr && getenv("CRX_TIE_ORDER")[0] == '0');
    const bool use_tc = !tc_off && t->metric == CRX_COSINE && t->k * L <= 32 && L <= 8 &&
                        pow(1.0 + (double)(1 << t->k), (double)L) <= (double)(1 << 24);
    CRX_REQUIRE(use_tc || P <= LIST, "P <= 32 for tables the tensor path does not take (Euclidean tables, k*L > 32, CRX_NO_TC)");
    if (use_tc) {
        int k = t->k;
        CRX_TRY(tl_s.alloc(c, (size_t)nq * TC_LIST)); CRX_TRY(tl_i.alloc(c, (size_t)nq * TC_LIST));
        CRX_TRY(ccode.alloc(c, N));
        { CRX_KERNEL(c, "pack_codes"); pack_codes_kernel<<<crx_grid(N, 256), 256, 0, c->stream>>>(t->bucket, N, N, k, L, ccode.p); }
        qcode = ccode.p;
        if (!self) {
            CRX_TRY(qcode_buf.alloc(c, queries->n));
            { CRX_KERNEL(c, "pack_codes"); pack_codes_kernel<<<crx_grid(queries->n, 256), 256, 0, c->stream>>>(qgid_base, qstride, queries->n, k, L, qcode_buf.p); }
            qcode = qcode_buf.p;
        }
        // exact |candidates| per query by inclusion-exclusion over table subsets
        std::vector<int64_t> hoff((size_t)1 << L, 0);
        int64_t bins = 0;
        for (int S = 1; S < (1 << L); S++) { hoff[S] = bins; bins += (int64_t)1 << (k * __builtin_popcount(S)); }
        DevBuf<int64_t> d_hoff;
        DevBuf<int> hist;
        CRX_TRY(d_hoff.alloc(c, hoff.size())); CRX_TRY(hist.alloc(c, bins));
        CRX_CUDA(cudaMemcpyAsync(d_hoff.p, hoff.data(), hoff.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        CRX_CUDA(cudaMemsetAsync(hist.p, 0, bins * sizeof(int), c->stream));
        { CRX_KERNEL(c, "subset_hist"); subset_hist_kernel<<<crx_grid(N, 256), 256, 0, c->stream>>>(ccode.p, N, k, L, d_hoff.p, hist.p); }
        { CRX_KERNEL(c, "subset_count"); subset_count_kernel<<<crx_grid(nq, 256), 256, 0, c->stream>>>(qcode, q_begin, nq, k, L, d_hoff.p, hist.p, nc.p); }
        CRX_CUDA(cudaGetLastError());
        // split-fp16 operands: unit rows times 2^10, so the accumulators hold 2^20 * cosine
        // centred on (1,..,1)/sqrt(D) when three spare operand columns exist (tc_prep_rows_kernel): per-query error bounds
        static const bool center_off = getenv("CRX_TC_CENTER") != nullptr && getenv("CRX_TC_CENTER")[0] == '0';
        const bool centered = !center_off && base->d + 3 <= 128;
        if (centered) { CRX_TRY(eps_b.alloc(c, (size_t)N)); if (!self) CRX_TRY(eps_a.alloc(c, (size_t)queries->n)); }
        int st = crx_tc_prepare(c, base, centered ? 2 : 0, 10.0, &opB, nullptr, nullptr, nullptr, centered ? eps_b.p : nullptr);
        if (st == CRX_OK && !self) st = crx_tc_prepare(c, queries, centered ? 2 : 0, 10.0, &opA, nullptr, nullptr, nullptr, centered ? eps_a.p : nullptr);
        if (centered) eps_rows = self ? eps_b.p : eps_a.p;
        // candidate density decides the hot-loop variant of the filter
        DevBuf<unsigned long long> tot;
        CRX_TRY(tot.alloc(c, 1));
        CRX_CUDA(cudaMemsetAsync(tot.p, 0, sizeof(unsigned long long), c->stream));
        { CRX_KERNEL(c, "sum_counts"); sum_counts_kernel<<<crx_grid(nq, 256), 256, 0, c->stream>>>(nc.p, nq, tot.p); }
        unsigned long long h_tot = 0;
        CRX_CUDA(cudaMemcpyAsync(&h_tot, tot.p, sizeof(h_tot), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        // the unmasked running maximum (1 instruction per score, table mask only in the rare path) pays as long as scores above the
        // row's threshold stay rare: their number grows like 32 ln(N) / f.  Measured at C2-(ii), f = 0.276: 761 ms masked in the
        // hot loop, dense variant below
        dense_cols = (double)h_tot >= 0.2 * (double)nq * (double)N;
        if (pass2) {
            CRX_TRY(blockmax.alloc(c, (size_t)nq * 2 * TC_NBLK));
            blk.blockmax = blockmax.p;
            blk.nblk = crx_tc_blocks(N, blk.bt);
            blk.tile_cols = crx_tc_tile_cols();
            blk.qcode = qcode; blk.ccode = ccode.p; blk.nb = N;
            for (int l = 0; l < L; l++) { blk.low |= 1u << (l * k); blk.high |= 1u << (l * k + k - 1); }
            blk.unscale = ldexp(1.0, -20);
        }
        clk.mark("codes, counts, operands");
        if (st == CRX_OK) st = crx_tc_topp(c, self ? opB : opA, q_begin, nq, opB, qcode, ccode.p, k, L, dense_cols, tl_s.p, tl_i.p, 3, blockmax.p);
        clk.mark("first-pass scan");
        tc_unscale = ldexp(1.0, -20);
        if (st != CRX_OK) return st;
    } else {
        CRX_TRY(list_s.alloc(c, (size_t)nq * LIST)); CRX_TRY(list_i.alloc(c, (size_t)nq * LIST));
        { CRX_KERNEL(c, "fill_lists"); fill_lists_kernel<<<crx_grid(nq * LIST, 256), 256, 0, c->stream>>>(list_s.p, list_i.p, nc.p, nq); }

        // cost model: dense any-table scan (nq * N pairs) vs one pass per table (sum of squared group sizes,
        // scaled to the query share)
        DevBuf<double> cost;
        CRX_TRY(cost.alloc(c, 1));
        CRX_CUDA(cudaMemsetAsync(cost.p, 0, sizeof(double), c->stream));
        for (int l = 0; l < L; l++) {
            CRX_KERNEL(c, "sq_sizes");
            sq_sizes_kernel<<<crx_grid(t->ngroups[l], 256), 256, 0, c->stream>>>(t->by_group[l].off, t->ngroups[l], cost.p);
        }
        double h_cost = 0;
        CRX_CUDA(cudaMemcpyAsync(&h_cost, cost.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        double table_cost = h_cost * ((double)nq / (double)N) + (double)L * 64.0 * (double)nq;
        double dense_cost = (double)nq * (double)N;
        bool dense = dense_cost <= table_cost;

        ScanArgs a;
        memset(&a, 0, sizeof(a));
        a.L = L; a.nq = nq; a.nb = N; a.q_begin = q_begin;
        for (int l = 0; l < L; l++) { a.qgid[l] = qgid_base + (size_t)l * qstride; a.cgid[l] = t->gid + (size_t)l * N; }
        a.list_s = list_s.p; a.list_i = list_i.p; a.ncand = nc.p;
        int ld = base->ld;
        size_t smem = pt::smem_bytes(ld) + pt::BN * sizeof(double) + pt::BN * sizeof(int32_t) + (size_t)MAXL * (pt::BN + pt::BM) * sizeof(int32_t);
        int grid = (int)((nq + pt::BM - 1) / pt::BM);
        DevBuf<int32_t> qperm, qkeys, qsorted;
        CRX_TRY(qperm.alloc(c, nq));
        auto launch_scan = [&](const ScanArgs& args) -> int {
            CRX_KERNEL(c, "topp_scan");
#define LAUNCH_T(TQ, TB, xqp, xbp)                                                                                      \
        do {                                                                                                                \
            CRX_CUDA(cudaFuncSetAttribute(topp_scan_kernel<TQ, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            topp_scan_kernel<TQ, TB><<<grid, pt::NT, smem, c->stream>>>(xqp, queries->ld, queries->sqn, xbp, base->ld, base->sqn, args); \
        } while (0)
            if (queries->x64 && base->x64) LAUNCH_T(double, double, queries->x64, base->x64);
            else if (queries->x64) LAUNCH_T(double, float, queries->x64, base->x32);
            else if (base->x64) LAUNCH_T(float, double, queries->x32, base->x64);
            else LAUNCH_T(float, float, queries->x32, base->x32);
#undef LAUNCH_T
            CRX_CUDA(cudaGetLastError());
            return CRX_OK;
        };
        if (dense) {
            { CRX_KERNEL(c, "iota"); iota_off_kernel<<<crx_grid(nq, 256), 256, 0, c->stream>>>(qperm.p, nq, (int32_t)q_begin); }
            a.pass = -1; a.qperm = qperm.p; a.cperm = nullptr; a.csorted = nullptr; a.load_state = 0;
            CRX_TRY(launch_scan(a));
        } else {
            for (int l = 0; l < L; l++) {
                // queries of the batch ordered by their group in table l
                Segments qs;
                int st = crx_build_segments(c, a.qgid[l] + q_begin, nq, t->ngroups[l] + 1, &qs);   // + the "no group" key of external queries
                if (st != CRX_OK) { qs.free_all(); return st; }
                { CRX_KERNEL(c, "add_off"); add_off_kernel<<<crx_grid(nq, 256), 256, 0, c->stream>>>(qs.perm, nq, (int32_t)q_begin); }
                a.pass = l; a.qperm = qs.perm; a.cperm = t->by_group[l].perm; a.csorted = t->by_group[l].sorted; a.load_state = l > 0;
                st = launch_scan(a);
                CRX_CUDA(cudaStreamSynchronize(c->stream));
                qs.free_all();
                if (st != CRX_OK) return st;
            }
        }
    }
    IoBuf<int32_t> o_recs, o_rows, o_nc, o_status;
    IoBuf<double> o_sims;
    CRX_TRY(o_status.bind(c, status, (size_t)nq, mem, false));

    CRX_TRY(o_recs.bind(c, recs, (size_t)nq * Nrec, mem, false));
    CRX_TRY(o_rows.bind(c, nbr_rows, (size_t)nq * P, mem, false));
    CRX_TRY(o_sims.bind(c, nbr_sims, (size_t)nq * P, mem, false));
    {
        CRX_KERNEL(c, "rec_finalize");
        int g = (int)((nq + 3) / 4);
#define LAUNCH_F(TQ, TB, xqp, xbp)                                                                                         \
    do {                                                                                                                   \
        if (use_tc)                                                                                                        \
            rec_finalize_kernel<TQ, TB, TC_LIST, float><<<g, 128, 0, c->stream>>>(xqp, queries->ld, queries->sqn, queries->unknown, queries->mean, xbp, \
                base->ld, base->sqn, base->mean, base->d, q_begin, nq, P, Nrec, tl_s.p, tl_i.p, nc.p, tc_unscale, 8e-6, o_recs.dev, o_rows.dev, o_sims.dev, c->counters, tie_order, o_status.dev, pass2, p2a.view(), blk, eps_rows); \
        else                                                                                                               \
            rec_finalize_kernel<TQ, TB, LIST, double><<<g, 128, 0, c->stream>>>(xqp, queries->ld, queries->sqn, queries->unknown, queries->mean, xbp, \
                base->ld, base->sqn, base->mean, base->d, q_begin, nq, P, Nrec, list_s.p, list_i.p, nc.p, 1.0, 1e-12, o_recs.dev, o_rows.dev, o_sims.dev, c->counters, tie_order, o_status.dev, pass2, p2a.view(), blk, nullptr); \
    } while (0)
        if (queries->x64 && base->x64) LAUNCH_F(double, double, queries->x64, base->x64);
        else if (queries->x64) LAUNCH_F(double, float, queries->x64, base->x32);
        else if (base->x64) LAUNCH_F(float, double, queries->x32, base->x64);
        else LAUNCH_F(float, float, queries->x32, base->x32);
#undef LAUNCH_F
    }
    CRX_CUDA(cudaGetLastError());
    size_t free_now = 0, total_now = 0;
    if (pass2) CRX_CUDA(cudaMemGetInfo(&free_now, &total_now));   // slow call: made here, behind the work enqueued above
    clk.mark("rec_finalize");
    if (pass2) {
        P2Host h;
        memset(&h, 0, sizeof(h));
        h.free_bytes = free_now;
        h.base = base; h.queries = queries; h.q_begin = q_begin; h.nq = nq; h.P = P; h.Nrec = Nrec; h.nc = nc.p; h.use_tc = use_tc;
        h.opQ = self ? &opB : &opA; h.opB = &opB; h.qcode = qcode; h.ccode = ccode.p; h.k = t->k; h.L = L; h.dense = dense_cols;
        h.qgid = qgid_base; h.qstride = qstride; h.cgid = t->gid; h.cstride = N;
        h.eps = use_tc ? 8e-6 : 1e-12; h.eps_q = eps_rows; h.unscale = tc_unscale; h.blocks = blk;
        h.recs = o_recs.dev; h.rows = o_rows.dev; h.sims = o_sims.dev; h.status = o_status.dev;
        CRX_TRY(pass2_run(c, h, &p2a, &p2b));
        clk.mark("second pass");
    }
    opB.free_all();
    opA.free_all();
    if (ncand) {
        CRX_CUDA(cudaMemcpyAsync(ncand, nc.p, nq * sizeof(int32_t), mem == CRX_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    }
    CRX_TRY(o_recs.flush());
    CRX_TRY(o_rows.flush());
    CRX_TRY(o_sims.flush());
    CRX_TRY(o_status.flush());
    if (mem == CRX_HOST) CRX_CUDA(cudaStreamSynchronize(c->stream));
    clk.mark("results out");
    return CRX_OK;
}

int crx_recommend_lsh(crx_ctx* c, const crx_lsh* t, const crx_points* queries, int64_t q_begin, int64_t q_end, int P,
                      int Nrec, int32_t* recs, int32_t* nbr_rows, double* nbr_sims, int32_t* ncand, int mem) {
    return recommend_lsh_impl(c, t, queries, q_begin, q_end, P, Nrec, recs, nbr_rows, nbr_sims, ncand, nullptr, mem);
}
int crx_recommend_lsh_status(crx_ctx* c, const crx_lsh* t, const crx_points* queries, int64_t q_begin, int64_t q_end, int P,
                             int Nrec, int32_t* recs, int32_t* nbr_rows, double* nbr_sims, int32_t* ncand, int32_t* status, int mem) {
    return recommend_lsh_impl(c, t, queries, q_begin, q_end, P, Nrec, recs, nbr_rows, nbr_sims, ncand, status, mem);
}

int crx_recommend_cluster(crx_ctx* c, const crx_points* users, const int32_t* labels, int lmem, int K,
                          const crx_points* queries, const int32_t* qlabels, int Nrec, int32_t* recs, int mem) {
    CRX_REQUIRE(c && users && labels && recs, "NULL argument");
    bool self = queries == nullptr || queries == users;
    if (!queries) queries = users;
    CRX_REQUIRE(self || qlabels, "qlabels required for external queries");
    CRX_REQUIRE(users->mean && queries->unknown && queries->mean, "ratings metadata missing (crx_points_set_ratings)");
    CRX_REQUIRE(queries->d == users->d && users->d <= 128, "dimension");
    CRX_REQUIRE(Nrec >= 0 && Nrec <= 128, "Nrec");
    CRX_CUDA(cudaSetDevice(c->device));
    int64_t N = users->n, nq = queries->n;
    IoBuf<int32_t> lab, qlab, out;
    CRX_TRY(lab.bind(c, labels, N, lmem, true));
    if (!self) CRX_TRY(qlab.bind(c, qlabels, nq, lmem, true));
    CRX_TRY(out.bind(c, recs, (size_t)nq * Nrec, mem, false));
    Segments seg;
    SegmentsGuard seg_guard(seg);   // freed on every return below
    int st = crx_build_segments(c, lab.dev, N, K, &seg);
    if (st != CRX_OK) return st;
    {
        CRX_KERNEL(c, "rec_cluster");
        int g = (int)((nq + 7) / 8);
        const int32_t* ql = self ? lab.dev : qlab.dev;
#define LAUNCH_C(TQ, TB, xqp, xbp)                                                                                          \
    rec_cluster_kernel<TQ, TB><<<g, 256, 0, c->stream>>>(xqp, queries->ld, queries->sqn, queries->unknown, queries->mean, ql, xbp, \
                                                          users->ld, users->sqn, users->mean, seg.perm, seg.off, K, users->d, nq, Nrec, out.dev)
        if (queries->x64 && users->x64) LAUNCH_C(double, double, queries->x64, users->x64);
        else if (queries->x64) LAUNCH_C(double, float, queries->x64, users->x32);
        else if (users->x64) LAUNCH_C(float, double, queries->x32, users->x64);
        else LAUNCH_C(float, float, queries->x32, users->x32);
#undef LAUNCH_C
    }
    CRX_CUDA(cudaGetLastError());
    st = out.flush();
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return st;
}

int crx_get_P_closest(crx_ctx* c, const crx_points* users, int32_t* neighbor_rows, int64_t n, const crx_points* query_set,
                      int64_t query_row, int P, double* similarities, int64_t* kept) {
    CRX_REQUIRE(c && users && neighbor_rows && query_set && similarities && kept, "NULL argument");
    CRX_REQUIRE(users->d == query_set->d, "dimension mismatch");
    CRX_REQUIRE(query_row >= 0 && query_row < query_set->n && n >= 0 && P >= 0, "argument range");
    *kept = std::min<int64_t>(n, P);
    if (n == 0) return CRX_OK;
    CRX_CUDA(cudaSetDevice(c->device));
    // similarities (crypto_rec.hpp:219-220), then the literal co-sort (:223): both on the device, one upload of the rows, one
    // download of the reordered rows and of the kept similarities, one synchronisation
    IoBuf<int32_t> ib;
    DevBuf<double> sims;
    CRX_TRY(ib.bind(c, neighbor_rows, (size_t)n, CRX_HOST, true));
    CRX_TRY(sims.alloc(c, (size_t)n));
    {
        CRX_KERNEL(c, "list_sims");
        int grid = crx_grid(n, 128);
#define LAUNCH_S(TQ, TB, xqp, xbp) \
    list_sims_kernel<TQ, TB><<<grid, 128, 0, c->stream>>>(xqp + (size_t)query_row * query_set->ld, query_set->sqn + query_row, xbp, users->ld, \
                                                          users->sqn, ib.dev, n, users->d, sims.p)
        if (query_set->x64 && users->x64) LAUNCH_S(double, double, query_set->x64, users->x64);
        else if (query_set->x64) LAUNCH_S(double, float, query_set->x64, users->x32);
        else if (users->x64) LAUNCH_S(float, double, query_set->x32, users->x64);
        else LAUNCH_S(float, float, query_set->x32, users->x32);
#undef LAUNCH_S
        CRX_CUDA(cudaGetLastError());
    }
    // only the first min(n, P) positions are consumed (crypto_rec.hpp:225-228): ranges beyond them are left unrefined
    CRX_TRY(launch_quicksort(c, sims.p, ib.dev, (int)n, (int)*kept));
    if (*kept > 0) CRX_CUDA(cudaMemcpyAsync(similarities, sims.p, (size_t)*kept * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return ib.flush();
}

int crx_get_top_N_recom(crx_ctx* c, const crx_points* users, const int32_t* neighbor_rows, const double* similarities,
                        int64_t n, const crx_points* query_set, int64_t query_row, int N, double* predicted, int32_t* recs) {
    CRX_REQUIRE(c && users && query_set && (neighbor_rows || n == 0), "NULL argument");
    CRX_REQUIRE(users->d == query_set->d && users->d <= 128, "dimension");
    CRX_REQUIRE(users->mean && query_set->unknown && query_set->mean, "ratings metadata missing (crx_points_set_ratings)");
    CRX_REQUIRE(query_row >= 0 && query_row < query_set->n && N >= 0 && N <= 128, "argument range");
    CRX_CUDA(cudaSetDevice(c->device));
    int D = users->d;
    IoBuf<int32_t> nb, out;
    IoBuf<double> sm, pr;
    CRX_TRY(nb.bind(c, neighbor_rows, (size_t)n, CRX_HOST, true));
    CRX_TRY(sm.bind(c, similarities, (size_t)n, CRX_HOST, true));
    CRX_TRY(pr.bind(c, predicted, (size_t)D, CRX_HOST, false));
    CRX_TRY(out.bind(c, recs, (size_t)N, CRX_HOST, false));
    {
        CRX_KERNEL(c, "rec_list");
        const uint8_t* unk = query_set->unknown + (size_t)query_row * D;
#define LAUNCH_L(TQ, TB, xqp, xbp) \
    rec_list_kernel<TQ, TB><<<1, 32, 0, c->stream>>>(xqp + (size_t)query_row * query_set->ld, D, query_set->sqn + query_row, unk, query_set->mean + query_row, xbp, users->ld, users->sqn, \
                                                      users->mean, nb.dev, sm.dev, (int)n, N, pr.dev, out.dev)
        if (query_set->x64 && users->x64) LAUNCH_L(double, double, query_set->x64, users->x64);
        else if (query_set->x64) LAUNCH_L(double, float, query_set->x64, users->x32);
        else if (users->x64) LAUNCH_L(float, double, query_set->x32, users->x64);
        else LAUNCH_L(float, float, query_set->x32, users->x32);
#undef LAUNCH_L
    }
    CRX_CUDA(cudaGetLastError());
    if (predicted) CRX_CUDA(cudaMemcpyAsync(predicted, pr.dev, (size_t)D * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (recs && N > 0) CRX_CUDA(cudaMemcpyAsync(recs, out.dev, (size_t)N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}

int crx_parallel_quickSort_topn(crx_ctx* c, double* sims, int32_t* ids, int n, int need) {
    CRX_REQUIRE(c && sims && ids && n >= 0 && need >= 0 && need <= 126, "argument");
    if (n == 0) return CRX_OK;
    CRX_CUDA(cudaSetDevice(c->device));
    IoBuf<double> s;
    IoBuf<int32_t> d;
    CRX_TRY(s.bind(c, sims, n, CRX_HOST, true));
    CRX_TRY(d.bind(c, ids, n, CRX_HOST, true));
    {
        CRX_KERNEL(c, "warp_qs");
        warp_qs_kernel<<<1, 32, 0, c->stream>>>(s.dev, d.dev, n, need);
        CRX_CUDA(cudaGetLastError());
    }
    CRX_TRY(s.flush());
    return d.flush();
}

int crx_parallel_quickSort(crx_ctx* c, double* sims, int32_t* ids, int n) {
    CRX_REQUIRE(c && sims && ids && n >= 0, "argument");
    if (n == 0) return CRX_OK;
    CRX_CUDA(cudaSetDevice(c->device));
    IoBuf<double> s;
    IoBuf<int32_t> d;
    CRX_TRY(s.bind(c, sims, n, CRX_HOST, true));
    CRX_TRY(d.bind(c, ids, n, CRX_HOST, true));
    CRX_TRY(launch_quicksort(c, s.dev, d.dev, n, n));
    CRX_TRY(s.flush());
    return d.flush();
}

} // extern "C"
