// tc_scan.cu -- tcgen05 / TMEM / TMA pair-scan engine for sm_100a (see tc_scan.cuh).
//
// One CTA = 128 rows of A ("queries" / "points"), resident in shared memory (TMA, 128B swizzle), against
// ALL 256-column tiles of B streamed through a 3-stage ring of 32 KB TMA boxes.  Warp 0 = TMA producer (one
// elected lane), warp 1 = tcgen05.mma issuer (one elected lane; also owns the TMEM allocation), warps 2..9 =
// epilogue: each thread owns ONE row x ONE 128-column half of the 128x256 fp32 accumulator tile
// (tcgen05.ld 32x32b.x32), so the per-row reductions (masked top-32 list per half / best + second best / row
// sums / threshold collection) are thread-private -- no shuffles, no atomics.  The accumulator is
// double-buffered in TMEM (2 x 256 columns = all of TMEM) so that the MMAs of tile t+1 overlap the epilogue of
// tile t.
//
// Split-fp16 arithmetic: every operand row is stored as [hi blocks | lo blocks] with hi = fp16(v),
// lo = fp16(v - hi).  For each B block the issuer runs
//     B.hi_j : D += A.hi_j * B.hi_j ;  D += A.lo_j * B.hi_j
//     B.lo_j : D += A.hi_j * B.lo_j
// i.e. hi*hi + lo*hi + hi*lo (lo*lo <= 2^-22 is dropped) with fp32 accumulation: relative error of the
// dot product <= ~4e-6 of |a||b| -- a FILTER; the exact FP64 kernels decide.
#include <cuda.h>
#include <cuda_fp16.h>

#include "tc_scan.cuh"

namespace {

constexpr int TM = 128;            // rows per CTA
constexpr int TN = 256;            // columns per tile (UMMA N = 256: A 4 KB + B 8 KB of smem reads per 128-cycle MMA)
constexpr int BLK_BYTES = TM * 64 * 2;   // one 128-row x 64-col fp16 block of A = 16 KB
constexpr int BBLK_BYTES = TN * 64 * 2;  // one 256-row x 64-col fp16 block of B = 32 KB
constexpr int NS = 3;              // B ring stages
constexpr uint32_t idesc_for(int n) {   // D = F32, A = B = F16, K-major A and B, N = n, M = 128
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
// "A from tensor memory" variant of the top-P / collection scans (ATM): the 128 query rows are written ONCE into TMEM
// columns [384, 512) (tcgen05.st, thread = row), the two accumulators shrink to 192 columns each, and every MMA reads only
// its B operand from shared memory.  With A and B both in shared memory the operand reads (4 + 8 KB per 128-cycle MMA =
// 96 B/clk) plus the TMA writes (35 B/clk) exceeded the 128 B/clk of the shared-memory pipe and held the tensor pipe at 73%.
constexpr int TN_ATM = 192;
constexpr int NS_ATM = 4;
constexpr int ATM_COL = 2 * TN_ATM;   // first TMEM column of the A operand: 64 words of hi parts, then 64 words of lo parts

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory (row = lane, two fp16 per 32-bit column), B in shared memory
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte swizzle, rows at 128 B pitch, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) /* LBO (ignored) */ | (64ull << 32) /* SBO = 1024 B */ |
           (1ull << 46) /* version */ | (2ull << 61) /* SWIZZLE_128B */;
}
// up to four K=16 MMAs over one 64-column block; descriptors advance by 32 bytes (>> 4 = 2).  `steps` < 4
// skips the all-zero padding columns of the last block (D = 100 needs 7 of the 8 K=16 slices).
__device__ __forceinline__ void mma_block(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t acc_first, int steps, uint32_t idesc) {
    uint64_t ad = smem_desc(a_addr), bd = smem_desc(b_addr);
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < steps) tc_mma(tmem_d, ad + 2 * k, bd + 2 * k, k == 0 ? acc_first : 1u, idesc);
}
// the same with the A block in tensor memory: a K = 16 slice is 8 columns
__device__ __forceinline__ void mma_block_ts(uint32_t tmem_d, uint32_t a_tmem, uint32_t b_addr, uint32_t acc_first, int steps, uint32_t idesc) {
    uint64_t bd = smem_desc(b_addr);
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < steps) tc_mma_ts(tmem_d, a_tmem + 8 * k, bd + 2 * k, k == 0 ? acc_first : 1u, idesc);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

struct TcParams {
    int64_t q0, nq;   // rows of A handled by the grid: [q0, q0 + nq)
    int64_t nb;       // valid rows of B
    int ntiles, nkb;
    int last_steps;   // K=16 slices of the last 64-column block that hold data
    const uint32_t* a_words;   // ATM: the A operand in global memory, [rows_pad][nkb * 64] words (hi part, then lo part)
    int64_t a_rows_pad;
    int lo_first;     // centred operands: the B.lo blocks go first and A.hi_last x B.hi_last is the final product of a tile, so that the
                      // large alpha_q alpha_c term (carried by three extra columns) meets the accumulator in the LAST MMA only
    // top-P
    const uint32_t* qcode;
    const uint32_t* ccode;
    uint32_t low_mask, high_mask;
    float* list_s;
    int32_t* list_i;
    int nprod;             // 3 = hi.hi + lo.hi + hi.lo (split-fp16), 1 = hi.hi only (coarse filter)
    // argmin
    const float* half_norm;
    float* best;
    float* second;
    int32_t* best_idx;
    // row sums of Euclidean distances (PAM): one job per CTA = (first row, end of its cluster, column range of the cluster)
    const int4* jobs;
    const float* norm_s;   // [rows] scaled squared norms of the operand rows
    const float* errw_s;   // [rows] error weight of a row: |d2 error| <= errw[a] + errw[b]
    float errw_max;        // an upper bound of every errw_s[]
    double* rowsum;
    double* rowerr;
    // top-P: maxima of the raw scores over geometric column blocks (block j = tiles [bt[j], bt[j+1])), per row and half
    float* blockmax;       // [nq][2][TC_NBLK] or NULL
    int nblk;
    int bt[TC_NBLK + 1];
    // threshold collection (second pass of the top-P): compact row i of A is query c_qrow[i]; the bit of every masked column with
    // score >= c_theta[i] or column index > c_colx[i] is set in the row's line of the mask matrix
    const float* c_theta;
    const int32_t* c_colx;
    const int32_t* c_qrow;
    uint32_t* cmask;        // [ntiles][TN / 32][c_rows_pad] pass bits of every 32-column chunk (one coalesced store per warp)
    int64_t c_rows_pad;     // rows of the mask matrix (the grid's rows, a multiple of 128)
    int32_t* c_count;       // [rows] number of set bits, added up by the row's epilogue threads (zeroed by the caller)
    int c_split;            // column tiles per CTA: blockIdx.y takes tiles [y c_split, (y + 1) c_split) -- a short queue (the later
                            // rounds) is spread over the SMs by columns; one CTA streaming every tile is bound by its 3-stage ring
};

constexpr int MODE_TOPP = 0, MODE_ARGMIN = 1, MODE_ROWSUM = 2, MODE_COLLECT = 3;
constexpr int HL = TC_LIST / 2;   // entries of one half-list (each epilogue half keeps its own top-HL)

// r[j] for a run-time j: 5-level select tree (registers cannot be indexed dynamically)
__device__ __forceinline__ uint32_t pick32(const uint32_t (&r)[32], int j) {
    uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (j & 1) ? r[2 * i + 1] : r[2 * i];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; i++) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
    for (int i = 0; i < 2; i++) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
    return (j & 16) ? d[1] : d[0];
}

// Replace the smallest entry of the row's (full) list, or append while it is filling, then recompute
// the threshold = smallest kept score.  Rare path (~HL*ln(N/HL) times per row).
__device__ __forceinline__ void list_insert(float* ls, int32_t* li, float s, int32_t idx, int& cnt, float& thr, int& minpos) {
    // ls / li point at this thread's column of the [entry][row] arrays: entry e lives at e * TM
    if (cnt < HL) {
        ls[cnt * TM] = s;
        li[cnt * TM] = idx;
        cnt++;
        if (cnt < HL) return;
    } else {
        ls[minpos * TM] = s;
        li[minpos * TM] = idx;
    }
    float m = INFINITY;
    int mp = 0;
#pragma unroll
    for (int e = 0; e < HL; e++) {
        float v = ls[e * TM];
        if (v < m) { m = v; mp = e; }
    }
    thr = m;
    minpos = mp;
}

constexpr int NEPI = 8;                       // epilogue warps: two per TMEM lane quarter
constexpr int NTHREADS_K = 64 + NEPI * 32;    // + producer warp + MMA warp

// DENSE (top-P only): nearly every column shares a bucket with every row (f = mean|cand|/N close to 1), so the
// hot loop only tracks the running maximum of the raw scores (1 FMNMX per score) and the table mask is
// evaluated in the rare path; otherwise the mask is applied before the maximum (5 instructions per score).
template <int MODE, bool DENSE, bool ATM = false>
__global__ void __launch_bounds__(NTHREADS_K, 1)
tc_scan_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
    static_assert(!ATM || MODE == MODE_TOPP || MODE == MODE_COLLECT, "A from tensor memory: top-P and collection scans only");
    constexpr int TN = ATM ? TN_ATM : ::TN;                 // columns per tile (shadows the file-scope value)
    constexpr int NS = ATM ? NS_ATM : ::NS;
    constexpr int BBLK_BYTES = TN * 64 * 2;
    constexpr int HALF = TN / 2;                            // columns of a tile per epilogue thread
    constexpr int NCH = HALF / 32;                          // 32-column chunks per epilogue thread and tile
    constexpr uint32_t IDESC = idesc_for(TN);
    extern __shared__ __align__(1024) uint8_t smem[];
    const int nblk = 2 * p.nkb;
    // ARGMIN is persistent: the CTA walks row tiles blockIdx.x, blockIdx.x + gridDim.x, ... with TWO A buffers, so the
    // A load, the pipeline fill and the drain of one row tile hide behind the MMAs of its neighbours (K = 1024 columns
    // are only 4 tiles per row tile).  TOPP / ROWSUM keep one row tile per CTA (their lists / long scans fill the CTA).
    constexpr int NA = MODE == MODE_ARGMIN ? 2 : 1;
    constexpr int A_BYTES = ATM ? 0 : NA * 4 * BLK_BYTES;   // ATM: the A operand lives in tensor memory
    uint8_t* sA = smem;                                    // NA x (nblk blocks, <= 64 KB)
    uint8_t* sB = smem + A_BYTES;                          // NS-stage ring of B blocks
    float* ls = reinterpret_cast<float*>(smem + A_BYTES + NS * BBLK_BYTES);   // [2 halves][HL][TM]  (top-P only)
    int32_t* li = reinterpret_cast<int32_t*>(ls + 2 * HL * TM);
    uint32_t* stile = MODE == MODE_TOPP ? reinterpret_cast<uint32_t*>(li + 2 * HL * TM)
                                        : reinterpret_cast<uint32_t*>(smem + A_BYTES + NS * BBLK_BYTES);  // [2][TN] codes / half norms
    constexpr int STW = MODE == MODE_ROWSUM ? 2 : 1;       // staged words per column (row sums: norm + error weight)
    uint64_t* bars = reinterpret_cast<uint64_t*>(stile + 2 * TN * STW);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* full = bars + 4;
    uint64_t* empty = full + NS;
    uint64_t* tfull = empty + NS;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    // hand-over between the column halves: [TM][3] floats (argmin: aliases the staging tile, shared memory is full) or
    // [TM][2] doubles (row sums)
    float* merge = MODE == MODE_ARGMIN ? reinterpret_cast<float*>(stile) : reinterpret_cast<float*>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();  // the 128B-swizzle atoms need a 1024-byte aligned base
        for (int b = 0; b < 2; b++) { mbar_init(&a_full[b], ATM ? NEPI : 1); mbar_init(&a_empty[b], 1); }
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], NEPI); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    int64_t row0 = p.q0 + (int64_t)blockIdx.x * TM;  // first A row of this CTA (of its first row tile: ARGMIN)
    int64_t row_end = p.q0 + p.nq, col0 = 0, col_end = p.nb;
    int ntiles = p.ntiles;
    const int n_row_tiles = MODE == MODE_ARGMIN ? (int)((p.nq + TM - 1) / TM) : (int)blockIdx.x + 1;
    const int rt_step = MODE == MODE_ARGMIN ? (int)gridDim.x : n_row_tiles;   // non-persistent modes: exactly one pass
    int tile0 = 0;   // first column tile of this CTA (collection with column splits)
    if (MODE == MODE_COLLECT) {
        tile0 = (int)blockIdx.y * p.c_split;
        ntiles = max(0, min(p.ntiles, tile0 + p.c_split) - tile0);
        col0 = (int64_t)tile0 * TN;
    }
    if (MODE == MODE_ROWSUM) {
        const int4 job = p.jobs[blockIdx.x];
        row0 = job.x; row_end = job.y; col0 = job.z; col_end = job.w;
        ntiles = (int)((col_end - col0 + TN - 1) / TN);
    }

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            int stage = 0;
            uint32_t phase = 0;
            const int nload = p.nprod == 1 ? p.nkb : nblk;   // single-product filter: high parts only
            int it = 0;
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += rt_step, it++) {
                const int ab = it & (NA - 1);
                const uint32_t use = (uint32_t)(it / NA);
                if (!ATM) {
                    mbar_wait(&a_empty[ab], (use & 1u) ^ 1u);   // the MMAs of the previous row tile in this buffer are done
                    mbar_arrive_expect_tx(&a_full[ab], (uint32_t)(nblk * BLK_BYTES));
                    const int arow = MODE == MODE_ARGMIN ? (int)(p.q0 + (int64_t)rt * TM) : (int)row0;
                    for (int b = 0; b < nblk; b++) tma_load_2d(sA + ab * 4 * BLK_BYTES + b * BLK_BYTES, &tmA, b * 64, arow, &a_full[ab]);
                }
                for (int t = 0; t < ntiles; t++) {
                    for (int sb = 0; sb < nload; sb++) {
                        const int b = p.lo_first ? (sb < p.nkb ? p.nkb + sb : sb - p.nkb) : sb;   // lo blocks first
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full[stage], (uint32_t)BBLK_BYTES);
                        tma_load_2d(sB + stage * BBLK_BYTES, &tmB, b * 64, (int)col0 + t * TN, &full[stage]);
                        if (++stage == NS) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            const uint32_t b_base = smem_u32(sB);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += rt_step, it++) {
            const int ab = it & (NA - 1);
            mbar_wait(&a_full[ab], (uint32_t)(it / NA) & 1u);
            tc_fence_after();
            const uint32_t a_base = smem_u32(sA) + ab * 4 * BLK_BYTES;
            for (int t = 0; t < ntiles; t++) {
                const int g = it * ntiles + t;                 // running tile number: TMEM buffer and its phase
                const int buf = g & 1;
                const uint32_t bphase = (uint32_t)(g >> 1) & 1u;
                mbar_wait(&tempty[buf], bphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * TN);  // TN fp32 columns per buffer
                const int nload = p.nprod == 1 ? p.nkb : nblk;
                uint32_t fresh = 0u;   // 0: the next MMA overwrites the accumulator (first of the tile), 1: it accumulates
                for (int sb = 0; sb < nload; sb++) {
                    const int b = p.lo_first ? (sb < p.nkb ? p.nkb + sb : sb - p.nkb) : sb;
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t bs = b_base + stage * BBLK_BYTES;
                    const int j = b < p.nkb ? b : b - p.nkb;                 // 64-column block index inside a part
                    const int steps = j == p.nkb - 1 ? p.last_steps : 4;
                    if (ATM) {        // A blocks: 32 words (64 fp16) each, hi parts at ATM_COL, lo parts 64 columns behind
                        const uint32_t ta = tmem_base + (uint32_t)(ATM_COL + j * 32);
                        if (b < p.nkb) {   // B.hi_j with A.lo_j and A.hi_j (the high x high product last)
                            if (p.nprod != 1) { mma_block_ts(d, ta + 64u, bs, fresh, steps, IDESC); fresh = 1u; }
                            mma_block_ts(d, ta, bs, fresh, steps, IDESC);
                        } else {
                            mma_block_ts(d, ta, bs, fresh, steps, IDESC);
                        }
                    } else if (b < p.nkb) {  // B.hi_j with A.lo_j and A.hi_j (the high x high product last)
                        if (p.nprod != 1) { mma_block(d, a_base + (p.nkb + j) * BLK_BYTES, bs, fresh, steps, IDESC); fresh = 1u; }
                        mma_block(d, a_base + j * BLK_BYTES, bs, fresh, steps, IDESC);
                    } else {          // B.lo_j with A.hi_j
                        mma_block(d, a_base + j * BLK_BYTES, bs, fresh, steps, IDESC);
                    }
                    fresh = 1u;
                    tc_commit(&empty[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull[buf]);
            }
            tc_commit(&a_empty[ab]);   // arrives when every MMA issued so far has read its operands
            }
        }
    } else {
        // ---------------- epilogue: thread = (row, column half) ----------------
        const int ew = warp - 2;                 // 0..7
        const int etid = threadIdx.x - 64;       // 0..255
        const int quarter = warp & 3;            // TMEM lane quarter this warp may read
        const int half = ew >> 2;                // columns [HALF*half, HALF*half + HALF) of every tile
        const int me = quarter * 32 + lane;      // row inside the tile
        int it = 0;
        for (int rt = blockIdx.x; rt < n_row_tiles; rt += rt_step, it++) {
        const int64_t grow = MODE == MODE_ROWSUM ? row0 + me : (int64_t)rt * TM + me;  // row relative to q0 (absolute for row sums)
        const bool valid = MODE == MODE_ROWSUM ? grow < row_end : grow < p.nq;
        const float na = (MODE == MODE_ROWSUM && valid) ? p.norm_s[grow] : 0.f;
        const float ea = (MODE == MODE_ROWSUM && valid) ? p.errw_s[grow] : 0.f;
        const float thr4 = 4.f * (ea + p.errw_max);   // d2 above this is certainly outside the small regime
        double rs_sum = 0.0, rs_err = 0.0;
        uint32_t cq = 0;
        float thr = -INFINITY, best = INFINITY, second = INFINITY;
        int cnt = 0, minpos = 0, bidx = 0x7fffffff;
        float* myls = ls + half * HL * TM + me;
        int32_t* myli = li + half * HL * TM + me;
        if (MODE == MODE_TOPP) {
            if (valid) cq = p.qcode[p.q0 + grow];
            for (int e = 0; e < HL; e++) { myls[e * TM] = -INFINITY; myli[e * TM] = -1; }
        }
        if (ATM) {
            // this thread's operand row (padded rows hold zeros): the half-0 warp of a lane quarter writes the hi part, the
            // half-1 warp the lo part; a K = 16 slice is 8 consecutive columns of the row's lane
            const int wpp = p.nkb * 32;    // words per part
            const int64_t arow = p.q0 + (int64_t)rt * TM + me;
            const bool arow_ok = arow < p.a_rows_pad;   // a query range that does not start on a tile boundary can reach beyond the operand
            const uint4* src = reinterpret_cast<const uint4*>(p.a_words + (size_t)(arow_ok ? arow : 0) * (size_t)(2 * wpp) + (size_t)half * wpp);
            const uint32_t ta = tmem_base + (uint32_t)(ATM_COL + half * 64) + ((uint32_t)(quarter * 32) << 16);
            for (int blk = 0; blk < p.nkb; blk++) {
                uint32_t w[32];
#pragma unroll
                for (int q4 = 0; q4 < 8; q4++) {
                    const uint4 v = arow_ok ? src[blk * 8 + q4] : make_uint4(0u, 0u, 0u, 0u);
                    w[q4 * 4 + 0] = v.x; w[q4 * 4 + 1] = v.y; w[q4 * 4 + 2] = v.z; w[q4 * 4 + 3] = v.w;
                }
                tmem_st32(ta + blk * 32, w);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[0]);
        }
        // top-P: running maximum of the raw scores of the current geometric column block
        float bm = -INFINITY;
        int bnext = 1;
        // collection: threshold, "everything behind" column, number of bits this thread has set
        float c_th = INFINITY;
        int c_cx = 0x7fffffff, c_total = 0;
        if (MODE == MODE_COLLECT && valid) {
            cq = p.qcode[p.c_qrow[grow]];
            c_th = p.c_theta[grow];
            c_cx = p.c_colx[grow];
        }
        const uint32_t low = p.low_mask, high = p.high_mask;
        // staging value (packed code / half norm) of this thread's column of the NEXT tile, fetched one tile ahead
        auto fetch_col = [&](int t) -> uint32_t {
            int64_t col = col0 + (int64_t)t * TN + etid;
            if (t >= ntiles) return 0u;
            if (MODE == MODE_TOPP || MODE == MODE_COLLECT) return col < p.nb ? p.ccode[col] : 0u;
            if (MODE == MODE_ROWSUM) return __float_as_uint(col < col_end ? p.norm_s[col] : -1e30f);  // outside the cluster: d = 0
            return __float_as_uint(col < p.nb ? p.half_norm[col] : INFINITY);
        };
        auto fetch_col2 = [&](int t) -> float {
            int64_t col = col0 + (int64_t)t * TN + etid;
            return (MODE == MODE_ROWSUM && t < ntiles && col < col_end) ? p.errw_s[col] : 0.f;
        };
        uint32_t next_col = fetch_col(0);
        float next_col2 = fetch_col2(0);
        for (int t = 0; t < ntiles; t++) {
            const int g = it * ntiles + t;             // running tile number (as in the MMA issuer)
            const int buf = g & 1;
            const uint32_t bphase = (uint32_t)(g >> 1) & 1u;
            if (etid < TN) stile[buf * TN * STW + etid] = next_col;
            if (MODE == MODE_ROWSUM) { stile[buf * TN * STW + TN + etid] = __float_as_uint(next_col2); next_col2 = fetch_col2(t + 1); }
            next_col = fetch_col(t + 1);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (MODE == MODE_TOPP && p.blockmax != nullptr && t == p.bt[bnext]) {
                if (valid) p.blockmax[(grow * 2 + half) * TC_NBLK + bnext - 1] = bm;
                bm = -INFINITY;
                bnext++;
            }
            mbar_wait(&tfull[buf], bphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(buf * TN + half * HALF) + ((uint32_t)(quarter * 32) << 16);
            // this thread's HALF columns in two rounds of (up to) two 32-column chunks; the TMEM buffer is handed back as
            // soon as the second round sits in registers
#pragma unroll 1
            for (int rnd = 0; rnd < 2; rnd++) {
            uint32_t r0[32], r1[32];
            const bool two = rnd * 2 + 1 < NCH;    // NCH = 3 (ATM): the second round holds one chunk
            tmem_ld32_nowait(taddr + rnd * 64, r0);
            if (two) tmem_ld32_nowait(taddr + rnd * 64 + 32, r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (rnd == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);
            }
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                if (rnd * 2 + ch >= NCH) continue;
                uint32_t (&r)[32] = ch == 0 ? r0 : r1;
                const uint4* st4 = reinterpret_cast<const uint4*>(stile + buf * TN * STW + half * HALF + rnd * 64 + ch * 32);
                const int cbase = (tile0 + t) * TN + half * HALF + rnd * 64 + ch * 32;
                if (MODE == MODE_TOPP) {
                    // hot loop: running maximum of the (masked) scores, four independent chains
                    float vm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    if (DENSE) {
#pragma unroll
                        for (int j = 0; j < 32; j++) vm[j & 3] = fmaxf(vm[j & 3], __uint_as_float(r[j]));
                    } else {
#pragma unroll
                        for (int g = 0; g < 8; g++) {
                            uint4 cc = st4[g];
                            uint32_t c4[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                uint32_t x = cq ^ c4[u];
                                uint32_t m = (x - low) & ~x & high;  // some k-bit field of x is zero
                                vm[u] = fmaxf(vm[u], m != 0u ? __uint_as_float(r[g * 4 + u]) : -INFINITY);
                            }
                        }
                    }
                    const float mx4 = fmaxf(fmaxf(vm[0], vm[1]), fmaxf(vm[2], vm[3]));
                    bm = fmaxf(bm, mx4);
                    if (mx4 > thr && valid) {
                        // rare path.  vm[u] is the maximum of chain u = entries j with j % 4 == u: only the chains that
                        // beat the threshold are looked at (8 entries each); an entry that is alone in its chain IS the
                        // chain maximum, so its value needs no register select tree
                        uint32_t bits = 0;
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (vm[u] > thr) {
#pragma unroll
                                for (int g = 0; g < 8; g++) {
                                    bool pass = __uint_as_float(r[g * 4 + u]) > thr;
                                    if (!DENSE) {  // the chain maximum was taken over the masked scores
                                        uint32_t x = cq ^ reinterpret_cast<const uint32_t*>(st4)[g * 4 + u];
                                        pass &= ((x - low) & ~x & high) != 0u;
                                    }
                                    bits |= pass ? (1u << (g * 4 + u)) : 0u;
                                }
                            }
                        }
                        const uint32_t all = bits;
                        // lane-local loop over the set bits: all lanes that have a candidate in this chunk walk it TOGETHER
                        while (bits != 0u) {
                            const int j = __ffs(bits) - 1;
                            bits &= bits - 1u;
                            const int u = j & 3;
                            float s;
                            if (__popc(all & (0x11111111u << u)) == 1) s = u == 0 ? vm[0] : (u == 1 ? vm[1] : (u == 2 ? vm[2] : vm[3]));
                            else s = __uint_as_float(pick32(r, j));
                            const int c = cbase + j;
                            bool ok = s > thr && c < p.nb;
                            if (DENSE && ok) {  // the table mask was not part of the hot loop
                                uint32_t x = cq ^ reinterpret_cast<const uint32_t*>(st4)[j];
                                ok = ((x - low) & ~x & high) != 0u;
                            }
                            if (ok) list_insert(myls, myli, s, c, cnt, thr, minpos);
                        }
                    }
                } else if (MODE == MODE_COLLECT) {
                    // pass bits of the 32 scores, no data-dependent branch: rows of a large plateau pass a threshold in nearly
                    // every chunk, and one such row would send its whole warp through a rare path (2 instructions per score
                    // here).  The table mask and the column bound are applied to the set bits only.
                    uint32_t bits = 0;
#pragma unroll
                    for (int j = 0; j < 32; j++) bits |= __uint_as_float(r[j]) >= c_th ? (1u << j) : 0u;
                    if (cbase + 31 > c_cx) bits |= cbase > c_cx ? 0xffffffffu : ~((2u << (c_cx - cbase)) - 1u);   // every column behind c_cx
                    if (cbase + 32 > (int)p.nb) bits &= cbase >= (int)p.nb ? 0u : ((1u << ((int)p.nb - cbase)) - 1u);
                    if (!valid) bits = 0u;
                    uint32_t left = bits;
                    while (left != 0u) {
                        const int j = __ffs(left) - 1;
                        left &= left - 1u;
                        const uint32_t x = cq ^ reinterpret_cast<const uint32_t*>(st4)[j];
                        if (((x - low) & ~x & high) == 0u) bits &= ~(1u << j);
                    }
                    // rows of a warp are consecutive: 128 contiguous bytes per store
                    p.cmask[((size_t)(tile0 + t) * (2 * NCH) + (size_t)(half * NCH + rnd * 2 + ch)) * (size_t)p.c_rows_pad + (size_t)grow] = bits;
                    c_total += __popc(bits);
                } else if (MODE == MODE_ROWSUM) {
                    // d = sqrt(max(0, |a|^2 + |b|^2 - 2 a.b)) and a running bound on its error.  The squared distance is
                    // off by at most E = ea + eb (error weights of the two rows: split-fp16 dot, fp32 norms, subnormal low
                    // parts);  |sqrt(x +- E) - sqrt(x)| <= E / sqrt(x) for x > 4E, and <= 2.5 sqrt(E) below (self pairs,
                    // duplicates, padding: rare, handled after the hot loop)
                    const uint4* se4 = st4 + TN / 4;
                    // hot loop, 10 instructions per pair: with rs = rsqrt(d2) the distance is d2 rs and the error term
                    // (ea + eb) rs is accumulated as ea * sum(rs) + sum(eb rs).  Pairs that MAY be in the small regime
                    // (d2 <= 4 (ea + max eb), a superset of d2 <= 4E) contribute no error term here and are revisited below.
                    float ps[4] = {0.f, 0.f, 0.f, 0.f}, srs[4] = {0.f, 0.f, 0.f, 0.f}, seb[4] = {0.f, 0.f, 0.f, 0.f};
                    uint32_t smallbits = 0;
#pragma unroll
                    for (int g = 0; g < 8; g++) {
                        uint4 cc = st4[g], ee = se4[g];
                        uint32_t c4[4] = {cc.x, cc.y, cc.z, cc.w}, e4[4] = {ee.x, ee.y, ee.z, ee.w};
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            float tsum = na + __uint_as_float(c4[u]);
                            float d2 = fmaf(-2.f, __uint_as_float(r[g * 4 + u]), tsum);
                            float ddc = fmaxf(d2, 1e-30f);
                            float rs;
                            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(ddc));
                            ps[u] = fmaf(ddc, rs, ps[u]);
                            bool small = d2 <= thr4;
                            float rse = small ? 0.f : rs;
                            srs[u] += rse;
                            seb[u] = fmaf(__uint_as_float(e4[u]), rse, seb[u]);
                            smallbits |= small ? (1u << (g * 4 + u)) : 0u;
                        }
                    }
                    float psum = (ps[0] + ps[1]) + (ps[2] + ps[3]);
                    float perr = fmaf(ea, (srs[0] + srs[1]) + (srs[2] + srs[3]), (seb[0] + seb[1]) + (seb[2] + seb[3])) + 4e-14f;
                    if (smallbits != 0u) {
                        const uint32_t* sn = reinterpret_cast<const uint32_t*>(st4);
                        while (smallbits != 0u) {
                            const int j = __ffs(smallbits) - 1;
                            smallbits &= smallbits - 1u;
                            const float nb = __uint_as_float(sn[j]);
                            if (nb > -1e29f) {  // a real column (padding columns have d = 0 and no error)
                                const float E = ea + __uint_as_float(sn[TN + j]);
                                const float dd = fmaxf(fmaf(-2.f, __uint_as_float(pick32(r, j)), na + nb), 0.f);
                                perr += dd <= 4.f * E ? 2.5f * sqrtf(E) : E * rsqrtf(dd);
                            }
                        }
                    }
                    // fp32 evaluation of the 32 terms: rsqrt.approx (2 ulp), products and <= 10 additions
                    rs_sum += (double)psum;
                    rs_err += (double)perr * 1.00001 + 1.5e-6 * (double)psum;
                } else {
#pragma unroll
                    for (int g = 0; g < 8; g++) {
                        uint4 cc = st4[g];
                        uint32_t c4[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            float v = __uint_as_float(c4[u]) - __uint_as_float(r[g * 4 + u]);
                            float nbst = fminf(v, best);
                            second = fminf(second, fmaxf(v, best));
                            bidx = v < best ? cbase + g * 4 + u : bidx;
                            best = nbst;
                        }
                    }
                }
            }
            }
        }
        if (MODE == MODE_TOPP) {
            if (valid) {
                for (int e = 0; e < HL; e++) {
                    p.list_s[grow * TC_LIST + half * HL + e] = myls[e * TM];
                    p.list_i[grow * TC_LIST + half * HL + e] = myli[e * TM];
                }
                if (p.blockmax != nullptr) {
                    p.blockmax[(grow * 2 + half) * TC_NBLK + bnext - 1] = bm;
                    for (int b = bnext; b < TC_NBLK; b++) p.blockmax[(grow * 2 + half) * TC_NBLK + b] = -INFINITY;
                }
            }
        } else if (MODE == MODE_COLLECT) {
            if (valid && c_total) atomicAdd(&p.c_count[grow], c_total);
        } else if (MODE == MODE_ROWSUM) {
            double* dm = reinterpret_cast<double*>(merge);
            if (half == 1) { dm[me * 2 + 0] = rs_sum; dm[me * 2 + 1] = rs_err; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0 && valid) {
                p.rowsum[grow] = rs_sum + dm[me * 2 + 0];
                p.rowerr[grow] = rs_err + dm[me * 2 + 1];
            }
        } else {
            asm volatile("bar.sync 1, 256;" ::: "memory");   // the staging tile is free: it doubles as the hand-over buffer
            if (half == 1) { merge[me * 3 + 0] = best; merge[me * 3 + 1] = second; merge[me * 3 + 2] = __int_as_float(bidx); }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0 && valid) {
                float b1 = merge[me * 3 + 0], s1 = merge[me * 3 + 1];
                int i1 = __float_as_int(merge[me * 3 + 2]);
                float sec = fminf(fminf(second, s1), fmaxf(best, b1));
                bool take1 = b1 < best || (b1 == best && i1 < bidx);
                p.best[grow] = take1 ? b1 : best;
                p.second[grow] = sec;
                p.best_idx[grow] = take1 ? i1 : bidx;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // before the next row tile stages its first column tile
        }
        }  // row tiles
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------- operand preparation
// mode 0: unit rows times `scale`; mode 1: rows times `scale`; mode 2: CENTRED unit rows (cosine top-P scans): with
// e = (1, ..., 1)/sqrt(D), alpha = u.e and d = u - alpha e, the operand holds d times `scale` in columns [0, D) and alpha times
// `scale` split three ways (h + l + m, 33 bits) in three extra columns laid out as (hi block, lo block) = (m, h), (l, m), (h, l):
// the three products the issuer forms (hi.hi + lo.hi + hi.lo) then add up to EXACTLY (h+l+m)_a (h+l+m)_b, so the accumulator
// holds scale^2 (d_a.d_b + alpha_a alpha_b) = scale^2 cos(a, b) while the split-fp16 and fp32-accumulation errors are
// proportional to |d_a||d_b| instead of |a||b| = 1.  Rating-like vectors are dominated by their mean fill (alpha ~ 0.98,
// |d| ~ 0.2; single-coin users have d = 0), which is exactly where the similarities crowd: the filter resolves them 5-40x
// finer.  eps_out[row] = the filter's error bound for a query row: 3e-7 (the alpha term meets the accumulator in the last
// MMA of a tile: one fp32 rounding at magnitude 1, lo_first) + 8e-6 |d| (|d_c| <= 1).
template <typename T>
__global__ void tc_prep_rows_kernel(const T* __restrict__ x, int ld, int D, const double* __restrict__ sqn, int64_t n,
                                    int mode, double scale, int nkb, const int32_t* __restrict__ rowmap, __half* __restrict__ out,
                                    float* __restrict__ norm_s, float* __restrict__ errw_s, double* __restrict__ eps_out) {
    int64_t orow = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (orow >= n) return;
    int64_t row = rowmap ? (int64_t)rowmap[orow] : orow;  // operand row orow holds point row
    double s = scale, shift = 0.0, alpha = 0.0;
    if (mode == 0 || mode == 2) {
        double nn = sqn[row];
        s = nn > 0.0 ? scale / sqrt(nn) : 0.0;
    }
    if (mode == 2) {
        double sx = 0.0;
        for (int c = lane; c < D; c += 32) sx += (double)x[row * ld + c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o);
        const double rD = sqrt((double)D);
        alpha = sx * (s / scale) / rD;       // u . e
        shift = alpha / rD * scale;          // alpha e_k, scaled
        if (eps_out && lane == 0) eps_out[orow] = 3e-7 + 8e-6 * sqrt(fmax(0.0, 1.0 - alpha * alpha));
    }
    if (norm_s && lane == 0) {
        // 6.3e-6 n covers the three-product split-fp16 dot with fp32 accumulation (2 x 3e-6 |a||b| <= 3e-6 (na + nb))
        // and the fp32 norms; the square-root term covers low parts that fall below the fp16 normal range
        double nn = sqn[row] * scale * scale;
        norm_s[orow] = (float)nn;
        errw_s[orow] = (float)crx_tc_errw(nn, D);
    }
    int W = nkb * 64;
    __half* o = out + orow * (size_t)(2 * W);
    for (int c = lane; c < W; c += 32) {
        double v = c < D ? (double)x[row * ld + c] * s - (s != 0.0 ? shift : 0.0) : 0.0;
        __half hi = __double2half(v);
        __half lo = __double2half(v - (double)__half2float(hi));
        if (mode == 2 && c >= D && c < D + 3 && s != 0.0) {
            const double ap = alpha * scale;
            const __half h = __double2half(ap);
            const __half l = __double2half(ap - (double)__half2float(h));
            const __half m = __double2half(ap - (double)__half2float(h) - (double)__half2float(l));
            // the large product h_a h_b sits in the LAST of the three columns: the last K = 16 slice the issuer touches
            if (c == D) { hi = m; lo = h; } else if (c == D + 1) { hi = l; lo = m; } else { hi = h; lo = l; }
        }
        o[c] = hi;
        o[W + c] = lo;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map(const TcOperand& op, int box_rows, CUtensorMap* tm) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
            crx_set_error("cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
            return CRX_ERR_CUDA;
        }
        fn = (EncodeTiledFn)ptr;
    }
    cuuint64_t cols = (cuuint64_t)op.nkb * 2 * 64;
    cuuint64_t gdim[2] = {cols, (cuuint64_t)op.rows_pad};
    cuuint64_t gstride[1] = {cols * sizeof(__half)};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, op.data, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        crx_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return CRX_ERR_CUDA;
    }
    return CRX_OK;
}

size_t smem_for(int mode, bool atm = false) {
    if (atm) {   // no A tile; 4-stage ring of 24 KB blocks
        size_t s = (size_t)NS_ATM * TN_ATM * 128 + 2 * TN_ATM * 4 + (8 + 2 * NS_ATM) * 8 + 16;
        if (mode == MODE_TOPP) s += (size_t)2 * HL * TM * 8;
        else s += (size_t)2 * 6 * TM * 4;
        return s;
    }
    size_t s = (size_t)4 * BLK_BYTES + (size_t)NS * BBLK_BYTES + 2 * TN * 4 + (8 + 2 * NS) * 8 + 16;
    if (mode == MODE_TOPP) s += (size_t)2 * HL * TM * 8;
    else if (mode == MODE_COLLECT) s += (size_t)2 * 8 * TM * 4;
    else if (mode == MODE_ROWSUM) s += TM * 2 * 8 + 2 * TN * 4;
    else s += 4 * BLK_BYTES;   // ARGMIN: second A buffer; the hand-over of the halves aliases the staging tile
    return s;
}

int last_steps_of(const TcOperand& op) {
    int rem = op.d - 64 * (op.nkb - 1);
    return (rem + 15) / 16;
}

int alloc_operand(crx_ctx* c, int64_t rows, int D, TcOperand* out) {
    out->d = D;
    out->rows = rows;
    out->rows_pad = (rows + TM - 1) / TM * TM;
    out->nkb = D <= 64 ? 1 : 2;
    size_t bytes = (size_t)out->rows_pad * out->nkb * 2 * 64 * sizeof(__half);
    out->owner = c;
    CRX_TRY(crx_alloc(c, (char**)&out->data, bytes));
    CRX_CUDA(cudaMemsetAsync(out->data, 0, bytes, c->stream));
    return CRX_OK;
}

}  // namespace

int crx_tc_prepare(crx_ctx* c, const crx_points* p, int mode, double scale_log2, TcOperand* out, const int32_t* rowmap, float* norm_s, float* errw_s,
                   double* eps_out) {
    CRX_REQUIRE(p->d <= 128, "tensor path supports D <= 128");
    CRX_REQUIRE(mode != 2 || p->d + 3 <= 128, "centred operands need three spare columns (D <= 125)");
    CRX_TRY(alloc_operand(c, p->n, mode == 2 ? p->d + 3 : p->d, out));
    out->centered = mode == 2;
    out->scale_log2 = scale_log2;
    double scale = ldexp(1.0, (int)scale_log2);
    int g = (int)((p->n + 7) / 8);
    CRX_KERNEL(c, "tc_prep");
    if (p->x64) tc_prep_rows_kernel<double><<<g, 256, 0, c->stream>>>(p->x64, p->ld, p->d, p->sqn, p->n, mode, scale, out->nkb, rowmap, (__half*)out->data, norm_s, errw_s, eps_out);
    else tc_prep_rows_kernel<float><<<g, 256, 0, c->stream>>>(p->x32, p->ld, p->d, p->sqn, p->n, mode, scale, out->nkb, rowmap, (__half*)out->data, norm_s, errw_s, eps_out);
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

int crx_tc_prepare_matrix(crx_ctx* c, const double* m, int K, int D, int ld, double scale_log2, TcOperand* out) {
    CRX_REQUIRE(D <= 128, "tensor path supports D <= 128");
    CRX_TRY(alloc_operand(c, K, D, out));
    out->scale_log2 = scale_log2;
    int g = (K + 7) / 8;
    CRX_KERNEL(c, "tc_prep");
    tc_prep_rows_kernel<double><<<g, 256, 0, c->stream>>>(m, ld, D, nullptr, K, 1, ldexp(1.0, (int)scale_log2), out->nkb, nullptr, (__half*)out->data, nullptr, nullptr, nullptr);
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

// columns per tile of the top-P / collection scans: 256 with both operands in shared memory (default), 192 with the A operand
// in tensor memory (CRX_TC_ATM=1).  Measured at C2 (1M x 1M x 100): 471 ms against 586 ms -- relieving the shared-memory pipe
// does not pay for the N = 192 MMAs (the tensor core appears to step N in units of 128); the variant stays for the record.
bool crx_tc_atm() {
    static const bool on = getenv("CRX_TC_ATM") != nullptr && getenv("CRX_TC_ATM")[0] == '1';
    return on;
}
int crx_tc_tile_cols() { return crx_tc_atm() ? TN_ATM : TN; }

// geometric column blocks: block 0 = the first half of the tiles, block 1 = the next quarter, ... , the last block = one tile
int crx_tc_blocks(int64_t b_rows, int* bt) {
    const int tn = crx_tc_tile_cols();
    const int nt = (int)((b_rows + tn - 1) / tn);
    int nblk = 1;
    bt[0] = 0;
    for (int j = 1; j < TC_NBLK && (nt >> j) >= 1; j++) {
        int b = nt - (nt >> j);
        if (b <= bt[nblk - 1]) continue;
        bt[nblk++] = b;
    }
    bt[nblk] = nt;
    return nblk;
}

namespace {
__global__ void tc_gather_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ rows, int64_t n, int64_t n_pad,
                                      int row_u4, uint4* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (i >= n_pad) return;
    for (int c = lane; c < row_u4; c += 32)
        dst[i * row_u4 + c] = i < n ? src[(int64_t)rows[i] * row_u4 + c] : make_uint4(0u, 0u, 0u, 0u);
}
}  // namespace

int crx_tc_gather(crx_ctx* c, const TcOperand& src, const int32_t* d_rows, int64_t n, TcOperand* out) {
    out->d = src.d;
    out->rows = n;
    out->rows_pad = (n + TM - 1) / TM * TM;
    out->nkb = src.nkb;
    out->scale_log2 = src.scale_log2;
    out->centered = src.centered;
    out->owner = c;
    const int row_u4 = src.nkb * 2 * 64 * (int)sizeof(__half) / 16;
    CRX_TRY(crx_alloc(c, (char**)&out->data, (size_t)out->rows_pad * row_u4 * 16));
    CRX_KERNEL(c, "tc_gather");
    tc_gather_rows_kernel<<<crx_grid(out->rows_pad, 8), 256, 0, c->stream>>>((const uint4*)src.data, d_rows, n, out->rows_pad, row_u4, (uint4*)out->data);
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

int64_t crx_tc_collect_mask_words(int64_t nrows, int64_t b_rows) {
    const int tn = crx_tc_tile_cols();
    const int64_t rows_pad = (nrows + TM - 1) / TM * TM, ntiles = (b_rows + tn - 1) / tn;
    return ntiles * (tn / 32) * rows_pad;
}

int crx_tc_collect(crx_ctx* c, const TcOperand& A, int64_t nrows, const TcOperand& B, const uint32_t* qcode, const int32_t* d_qrow,
                   const uint32_t* ccode, int k, int L, bool dense, const float* d_theta, const int32_t* d_colx, uint32_t* cmask,
                   int32_t* d_count) {
    CRX_REQUIRE(A.nkb == B.nkb && A.centered == B.centered && A.d == B.d, "operand layouts differ");
    CRX_REQUIRE(k * L <= 32 && k >= 1, "packed codes need k*L <= 32");
    if (nrows == 0) return CRX_OK;
    CUtensorMap tmA, tmB;
    CRX_TRY(make_tensor_map(A, TM, &tmA));
    CRX_TRY(make_tensor_map(B, TN, &tmB));
    TcParams p;
    memset(&p, 0, sizeof(p));
    const bool atm = crx_tc_atm();
    const int tn = crx_tc_tile_cols();
    if (atm) CRX_TRY(make_tensor_map(B, tn, &tmB));
    p.q0 = 0; p.nq = nrows; p.nb = B.rows; p.ntiles = (int)((B.rows + tn - 1) / tn); p.nkb = A.nkb;
    p.last_steps = last_steps_of(A);
    p.nprod = 3;
    p.a_words = (const uint32_t*)A.data; p.a_rows_pad = A.rows_pad;
    p.lo_first = A.centered ? 1 : 0;
    p.qcode = qcode; p.ccode = ccode;
    uint32_t low = 0, high = 0;
    for (int l = 0; l < L; l++) { low |= 1u << (l * k); high |= 1u << (l * k + k - 1); }
    p.low_mask = low; p.high_mask = high;
    p.c_theta = d_theta; p.c_colx = d_colx; p.c_qrow = d_qrow;
    p.cmask = cmask; p.c_rows_pad = (nrows + TM - 1) / TM * TM; p.c_count = d_count;
    size_t smem = smem_for(MODE_COLLECT, atm);
    int grid = (int)((nrows + TM - 1) / TM);
    // fewer row tiles than SMs: split the columns so that every SM has work (at least 8 tiles per CTA)
    int splits = 1;
    if (grid < c->sm_count) splits = std::max(1, std::min(c->sm_count / grid, p.ntiles / 8));
    p.c_split = (p.ntiles + splits - 1) / splits;
    splits = (p.ntiles + p.c_split - 1) / p.c_split;
    const dim3 grid2((unsigned)grid, (unsigned)splits);
    CRX_KERNEL(c, "tc_collect_scan");
    if (atm) {   // (the table mask is applied to the set bits only: one variant serves dense and sparse candidate sets)
        CRX_CUDA(cudaFuncSetAttribute(tc_scan_kernel<MODE_COLLECT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_scan_kernel<MODE_COLLECT, true, true><<<grid2, NTHREADS_K, smem, c->stream>>>(tmA, tmB, p);
    } else {
        CRX_CUDA(cudaFuncSetAttribute(tc_scan_kernel<MODE_COLLECT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_scan_kernel<MODE_COLLECT, true><<<grid2, NTHREADS_K, smem, c->stream>>>(tmA, tmB, p);
    }
    CRX_CUDA(cudaGetLastError());
    (void)dense;
    return CRX_OK;
}

int crx_tc_topp(crx_ctx* c, const TcOperand& A, int64_t q0, int64_t nq, const TcOperand& B, const uint32_t* qcode,
                const uint32_t* ccode, int k, int L, bool dense, float* list_s, int32_t* list_i, int nprod, float* blockmax) {
    CRX_REQUIRE(A.nkb == B.nkb && A.centered == B.centered && A.d == B.d, "operand layouts differ");
    CRX_REQUIRE(k * L <= 32 && k >= 1, "packed codes need k*L <= 32");
    CUtensorMap tmA, tmB;
    CRX_TRY(make_tensor_map(A, TM, &tmA));
    CRX_TRY(make_tensor_map(B, TN, &tmB));
    TcParams p;
    memset(&p, 0, sizeof(p));
    const bool atm = crx_tc_atm() && nprod != 1;
    const int tn = atm ? TN_ATM : TN;
    if (atm) CRX_TRY(make_tensor_map(B, tn, &tmB));
    p.q0 = q0; p.nq = nq; p.nb = B.rows; p.ntiles = (int)((B.rows + tn - 1) / tn); p.nkb = A.nkb;
    p.last_steps = last_steps_of(A);
    p.nprod = nprod == 1 ? 1 : 3;
    p.a_words = (const uint32_t*)A.data; p.a_rows_pad = A.rows_pad;
    p.lo_first = A.centered ? 1 : 0;
    p.qcode = qcode; p.ccode = ccode;
    uint32_t low = 0, high = 0;
    for (int l = 0; l < L; l++) { low |= 1u << (l * k); high |= 1u << (l * k + k - 1); }
    p.low_mask = low; p.high_mask = high;
    p.list_s = list_s; p.list_i = list_i;
    p.blockmax = blockmax;
    p.nblk = crx_tc_blocks(B.rows, p.bt);
    size_t smem = smem_for(MODE_TOPP, atm);
    int grid = (int)((nq + TM - 1) / TM);
    CRX_KERNEL(c, "tc_topp_scan");
    if (atm && dense) {
        CRX_CUDA(cudaFuncSetAttribute(tc_scan_kernel<MODE_TOPP, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_scan_kernel<MODE_TOPP, true, true><<<grid, NTHREADS_K, smem, c->stream>>>(tmA, tmB, p);
    } else if (atm) {
        CRX_CUDA(cudaFuncSetAttribute(tc_scan_kernel<MODE_TOPP, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_scan_kernel<MODE_TOPP, false, true><<<grid, NTHREADS_K, smem, c->stream>>>(tmA, tmB, p);
    } else if (dense) {
        CRX_CUDA(cudaFuncSetAttribute(tc_scan_kernel<MODE_TOPP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_scan_kernel<MODE_TOPP, true><<<grid, NTHREADS_K, smem, c->stream>>>(tmA, tmB, p);
    } else {
        CRX_CUDA(cudaFuncSetAttribute(tc_scan_kernel<MODE_TOPP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_scan_kernel<MODE_TOPP, false><<<grid, NTHREADS_K, smem, c->stream>>>(tmA, tmB, p);
    }
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

int crx_tc_argmin(crx_ctx* c, const TcOperand& A, int64_t r0, int64_t nr, const TcOperand& B, const float* half_norm,
                  float* best, float* second, int32_t* best_idx) {
    CRX_REQUIRE(A.nkb == B.nkb, "operand widths differ");
    CUtensorMap tmA, tmB;
    CRX_TRY(make_tensor_map(A, TM, &tmA));
    CRX_TRY(make_tensor_map(B, TN, &tmB));
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.q0 = r0; p.nq = nr; p.nb = B.rows; p.ntiles = (int)((B.rows + TN - 1) / TN); p.nkb = A.nkb;
    p.last_steps = last_steps_of(A);
    p.nprod = 3;
    p.half_norm = half_norm; p.best = best; p.second = second; p.best_idx = best_idx;
    size_t smem = smem_for(MODE_ARGMIN);
    CRX_CUDA(cudaFuncSetAttribute(tc_scan_kernel<MODE_ARGMIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (int)std::min<int64_t>((nr + TM - 1) / TM, (int64_t)c->sm_count);   // persistent: one CTA per SM walks the row tiles
    CRX_KERNEL(c, "tc_argmin_scan");
    tc_scan_kernel<MODE_ARGMIN, false><<<grid, NTHREADS_K, smem, c->stream>>>(tmA, tmB, p);
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}

int crx_tc_rowsum(crx_ctx* c, const TcOperand& A, const int4* d_jobs, int njobs, const float* norm_s, const float* errw_s,
                  float errw_max, double* rowsum, double* rowerr) {
    if (njobs == 0) return CRX_OK;
    CUtensorMap tmA, tmB;
    CRX_TRY(make_tensor_map(A, TM, &tmA));
    CRX_TRY(make_tensor_map(A, TN, &tmB));
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.nkb = A.nkb;
    p.last_steps = last_steps_of(A);
    p.nprod = 3;
    p.jobs = d_jobs; p.norm_s = norm_s; p.errw_s = errw_s; p.errw_max = errw_max; p.rowsum = rowsum; p.rowerr = rowerr;
    size_t smem = smem_for(MODE_ROWSUM);
    CRX_CUDA(cudaFuncSetAttribute(tc_scan_kernel<MODE_ROWSUM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CRX_KERNEL(c, "tc_rowsum_scan");
    tc_scan_kernel<MODE_ROWSUM, false><<<njobs, NTHREADS_K, smem, c->stream>>>(tmA, tmB, p);
    CRX_CUDA(cudaGetLastError());
    return CRX_OK;
}
