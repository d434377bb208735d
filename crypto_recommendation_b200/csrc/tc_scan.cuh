// tc_scan.cuh -- declarations of the tcgen05 (5th-gen tensor core) pair-scan engine.
//
// The contraction-bound kernels of this path (K9 cosine top-P scan, K5 Lloyd assignment at large K)
// are "row block x ALL columns" GEMMs followed by a per-row reduction.  tc_scan.cu runs them on the
// tensor cores as a FILTER: operands are split into fp16 hi + lo parts (22 significant bits), three
// products hi*hi + hi*lo + lo*hi are accumulated in fp32 in TMEM, and the epilogue keeps, per row,
// either the 64 best masked scores (top-P) or the best and second-best value (argmin).  The exact
// FP64 kernels then re-evaluate only the survivors, so results stay those of the reference.
#pragma once
#include "common.cuh"

// [rows][4*64 or 2*64] fp16: hi blocks first, then lo blocks, 64 columns per block (one 128-byte
// swizzle atom); rows padded to a multiple of 128 with zeros.
struct TcOperand {
    void* data = nullptr;     // __half [rows_pad][nkb*2*64]
    int64_t rows = 0, rows_pad = 0;
    int nkb = 0;              // 64-column blocks per part (1 for D <= 64, 2 for D <= 128)
    int d = 0;                // true number of columns
    double scale_log2 = 0;    // values were multiplied by 2^scale_log2 before the split
    bool centered = false;    // crx_tc_prepare mode 2: rows centred on (1,..,1)/sqrt(D), alpha in three extra columns (d counts them)
    crx_ctx* owner = nullptr;
    TcOperand() {}
    TcOperand(const TcOperand&) = delete;
    TcOperand& operator=(const TcOperand&) = delete;
    ~TcOperand() { free_all(); }   // stream-ordered free: safe after the launches that read it were enqueued
    void free_all() { if (owner && data) crx_free(owner, data); data = nullptr; }
};

// mode 0: rows scaled to unit length times 2^scale_log2 (cosine);  mode 1: all rows times 2^scale_log2;  mode 2: unit rows
// centred on (1,..,1)/sqrt(D) (cosine top-P scans, D <= 125: see tc_prep_rows_kernel), eps_out[row] = the filter's error
// bound (similarity units) for that row as a query
// rowmap (nullable): operand row i holds point rowmap[i];  norm_s / errw_s (nullable, together): scaled squared norms and
// the row-sum error weights
int crx_tc_prepare(crx_ctx* c, const crx_points* p, int mode, double scale_log2, TcOperand* out, const int32_t* rowmap = nullptr,
                   float* norm_s = nullptr, float* errw_s = nullptr, double* eps_out = nullptr);
// same for a [K][ld] double matrix (centroids)
int crx_tc_prepare_matrix(crx_ctx* c, const double* m, int K, int D, int ld, double scale_log2, TcOperand* out);

constexpr int TC_LIST = 64;  // per-row candidate list of the top-P filter
constexpr int TC_NBLK = 16;  // geometric column blocks whose score maxima the top-P filter reports (crx_tc_blocks)

// top-P filter: for query rows [q0, q0+nq) of A against all rows of B, keep the TC_LIST best scores among
// the columns whose packed code shares at least one k-bit field with the query's code.
//   list_s[nq][TC_LIST] (float, scaled by 2^(sa+sb)), list_i[nq][TC_LIST] (column, -1 = empty)
// dense: the caller knows that (nearly) every column is a candidate of every row (mean |cand| / N > 0.9)
// nprod: 3 = split-fp16 products hi.hi + lo.hi + hi.lo (error ~6e-6 |a||b|); 1 = hi.hi only (error ~2^-10 |a||b|): kept for
//        the measurement recorded in DESIGN.md section 8 -- TMEM reads, not the MMAs, bound it, so it is not used
// blockmax (nullable): [nq][2][TC_NBLK] maxima of the raw scores (table mask applied only in the non-dense variant) of each
//        row over the geometric column blocks of crx_tc_blocks, per epilogue half -- an upper bound of every candidate's score
int crx_tc_topp(crx_ctx* c, const TcOperand& A, int64_t q0, int64_t nq, const TcOperand& B, const uint32_t* qcode,
                const uint32_t* ccode, int k, int L, bool dense, float* list_s, int32_t* list_i, int nprod = 3,
                float* blockmax = nullptr);
// first tile of each geometric block into bt[0..nblk], bt[nblk] = number of 256-column tiles; returns nblk <= TC_NBLK
int crx_tc_blocks(int64_t b_rows, int* bt);
// columns per tile of the top-P / collection scans (256; 192 with CRX_TC_ATM=1: A operand in tensor memory)
int crx_tc_tile_cols();

// operand rows d_rows[0..n) of `src` copied into a compact operand (padded with zero rows to a multiple of 128)
int crx_tc_gather(crx_ctx* c, const TcOperand& src, const int32_t* d_rows, int64_t n, TcOperand* out);

// threshold collection (second pass of the top-P): for compact row i of A (= query row d_qrow[i], which indexes qcode) the bit of
// every column j of B that shares a bucket with the query and has score >= d_theta[i] (filter units) or j > d_colx[i] is set in
// cmask[(tile * (tile_cols / 32) + word)][rows_pad] (rows_pad = nrows rounded up to 128; every word of the matrix is written);
// d_count[i] (zeroed by the caller) receives the number of set bits.  crx_tc_collect_mask_words: size of the matrix in words.
int64_t crx_tc_collect_mask_words(int64_t nrows, int64_t b_rows);
int crx_tc_collect(crx_ctx* c, const TcOperand& A, int64_t nrows, const TcOperand& B, const uint32_t* qcode, const int32_t* d_qrow,
                   const uint32_t* ccode, int k, int L, bool dense, const float* d_theta, const int32_t* d_colx, uint32_t* cmask,
                   int32_t* d_count);

// argmin filter: for rows [r0, r0+nr) of A against the K rows of B: best / second-best of
// half_norm[j] - dot(a, b_j) (scaled units) and the best column.
int crx_tc_argmin(crx_ctx* c, const TcOperand& A, int64_t r0, int64_t nr, const TcOperand& B, const float* half_norm,
                  float* best, float* second, int32_t* best_idx);

// PAM row sums: job j = {first operand row of a 128-row tile, end row of its cluster, first column, end column};
// rowsum[row] ~ sum over the cluster of the Euclidean distance (scaled by 2^scale), rowerr[row] bounds its error.
// errw_max: an upper bound of errw_s[] (crx_tc_errw of the largest scaled squared norm)
int crx_tc_rowsum(crx_ctx* c, const TcOperand& A, const int4* d_jobs, int njobs, const float* norm_s, const float* errw_s,
                  float errw_max, double* rowsum, double* rowerr);
// error weight of a row with scaled squared norm nn: the squared distance of two rows is off by at most errw(a) + errw(b).
// 6.3e-6 nn covers the three-product split-fp16 dot with fp32 accumulation (2 x 3e-6 |a||b| <= 3e-6 (na + nb)) and the fp32
// norms; the square-root term covers low parts that fall below the fp16 normal range
__host__ __device__ inline double crx_tc_errw(double nn, int D) { return (6.3e-6 * nn + 5.97e-8 * sqrt((double)D * nn)) * 1.000001; }
