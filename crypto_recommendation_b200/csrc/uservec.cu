// uservec.cu -- user rating vectors from (user, coin, score) mentions: the step in front of the hot path
// (tweets_to_user_vectors crypto_rec.hpp:79-140, clusters_to_user_vectors :143-210).
//
// The reference walks the tweets once and does `row[user][coin] += score` (only when score > 0) and
// `known[user][coin] = 1`; then per user, in coin order: sum of the known coins, their count, "useless" when every
// coordinate is 0, mean = sum / count, unknown coins := mean.  Floating-point sums depend on the order of the
// mentions, so the mentions are stably sorted by (user, coin) and each (user, coin) run is added by one thread in
// mention order; the per-user pass adds the coins in index order.  Results are bit-identical to the sequential walk.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

__global__ void mention_keys_kernel(const int32_t* __restrict__ user, const int32_t* __restrict__ coin, int64_t E, int D,
                                    int64_t n_users, uint64_t* __restrict__ key, int32_t* __restrict__ idx, int* __restrict__ bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    int32_t u = user[i], c = coin[i];
    if (u < 0 || u >= n_users || c < 0 || c >= D) { atomicExch(bad, 1); u = 0; c = 0; }
    key[i] = (uint64_t)u * (uint64_t)D + (uint64_t)c;
    idx[i] = (int32_t)i;
}

// one thread per sorted position; the first position of a (user, coin) run adds the whole run in mention order
__global__ void mention_runs_kernel(const uint64_t* __restrict__ key, const int32_t* __restrict__ idx,
                                    const double* __restrict__ score, int64_t E, double* __restrict__ X,
                                    uint8_t* __restrict__ known) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    uint64_t k = key[i];
    if (i > 0 && key[i - 1] == k) return;
    double acc = 0.0;  // crypto_rec.hpp:88 value-initialised vector
    for (int64_t j = i; j < E && key[j] == k; j++) {
        double s = score[idx[j]];
        if (s > 0) acc = __dadd_rn(acc, s);  // :98-99
    }
    X[k] = acc;
    known[k] = 1;  // :101
}

// one warp per user: coordinates staged through shared memory, lane 0 adds the known ones in coin order (:113-125)
__global__ void __launch_bounds__(256)
user_fill_kernel(double* __restrict__ X, const uint8_t* __restrict__ known, int64_t n_users, int D, uint8_t* __restrict__ unknown,
                 double* __restrict__ mean_out, uint8_t* __restrict__ keep) {
    extern __shared__ double sm[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t u = (int64_t)blockIdx.x * 8 + warp;
    if (u >= n_users) return;
    double* v = sm + (size_t)warp * D;
    double* row = X + (size_t)u * D;
    const uint8_t* kn = known + (size_t)u * D;
    bool nz = false;
    for (int j = lane; j < D; j += 32) { double x = row[j]; v[j] = x; nz |= (x != 0); }
    nz = __any_sync(0xffffffffu, nz);
    __syncwarp();
    double sum = 0.0;
    int cnt = 0;
    if (lane == 0)
        for (int j = 0; j < D; j++)
            if (kn[j]) { sum = __dadd_rn(sum, v[j]); cnt++; }
    sum = __shfl_sync(0xffffffffu, sum, 0);
    cnt = __shfl_sync(0xffffffffu, cnt, 0);
    double mean = __ddiv_rn(sum, (double)cnt);  // :130 (0/0 = NaN for a user without mentions, dropped as useless)
    for (int j = lane; j < D; j += 32) {
        bool unk = kn[j] == 0;
        unknown[(size_t)u * D + j] = unk ? 1 : 0;
        if (unk && nz) row[j] = mean;  // :133-134
    }
    if (lane == 0) { mean_out[u] = mean; keep[u] = nz ? 1 : 0; }
}

}  // namespace

extern "C" int crx_user_vectors_build(crx_ctx* c, const int32_t* mention_user, const int32_t* mention_coin,
                                      const double* mention_score, int64_t n_mentions, int64_t n_users, int n_coins, double* X,
                                      uint8_t* unknown, double* known_mean, uint8_t* keep, int mem) {
    CRX_REQUIRE(c && X && unknown && known_mean && keep, "NULL argument");
    CRX_REQUIRE(n_mentions == 0 || (mention_user && mention_coin && mention_score), "NULL mention arrays");
    CRX_REQUIRE(n_users >= 0 && n_coins > 0 && n_coins <= 512 && n_mentions >= 0 && n_mentions < (1ll << 31), "bad sizes (n_coins <= 512)");
    CRX_CUDA(cudaSetDevice(c->device));
    int64_t E = n_mentions;
    size_t cells = (size_t)n_users * n_coins;
    IoBuf<int32_t> mu, mc;
    IoBuf<double> ms, xo, mo;
    IoBuf<uint8_t> uo, ko;
    CRX_TRY(mu.bind(c, mention_user, E, mem, true)); CRX_TRY(mc.bind(c, mention_coin, E, mem, true));
    CRX_TRY(ms.bind(c, mention_score, E, mem, true));
    CRX_TRY(xo.bind(c, X, cells, mem, false)); CRX_TRY(uo.bind(c, unknown, cells, mem, false));
    CRX_TRY(mo.bind(c, known_mean, n_users, mem, false)); CRX_TRY(ko.bind(c, keep, n_users, mem, false));
    DevBuf<uint8_t> known;
    CRX_TRY(known.alloc(c, cells));
    CRX_CUDA(cudaMemsetAsync(known.p, 0, cells, c->stream));
    CRX_CUDA(cudaMemsetAsync(xo.dev, 0, cells * sizeof(double), c->stream));
    if (E > 0) {
        DevBuf<uint64_t> k0, k1;
        DevBuf<int32_t> i0, i1;
        DevBuf<int> bad;
        CRX_TRY(k0.alloc(c, E)); CRX_TRY(k1.alloc(c, E)); CRX_TRY(i0.alloc(c, E)); CRX_TRY(i1.alloc(c, E)); CRX_TRY(bad.alloc(c, 1));
        CRX_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), c->stream));
        {
            CRX_KERNEL(c, "mention_keys");
            mention_keys_kernel<<<crx_grid(E, 256), 256, 0, c->stream>>>(mu.dev, mc.dev, E, n_coins, n_users, k0.p, i0.p, bad.p);
        }
        int bits = 1;
        while (bits < 64 && (cells >> bits) != 0) bits++;
        size_t bytes = 0;
        CRX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, i0.p, i1.p, (int)E, 0, bits, c->stream));
        DevBuf<char> tmp;
        CRX_TRY(tmp.alloc(c, bytes));
        CRX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, k0.p, k1.p, i0.p, i1.p, (int)E, 0, bits, c->stream));  // stable
        {
            CRX_KERNEL(c, "mention_runs");
            mention_runs_kernel<<<crx_grid(E, 256), 256, 0, c->stream>>>(k1.p, i1.p, ms.dev, E, xo.dev, known.p);
        }
        int h_bad = 0;
        CRX_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CRX_CUDA(cudaStreamSynchronize(c->stream));
        CRX_REQUIRE(!h_bad, "mention with user or coin index out of range");
    }
    if (n_users > 0) {
        CRX_KERNEL(c, "user_fill");
        user_fill_kernel<<<crx_grid(n_users, 8), 256, (size_t)8 * n_coins * sizeof(double), c->stream>>>(xo.dev, known.p, n_users, n_coins, uo.dev,
                                                                                                        mo.dev, ko.dev);
    }
    CRX_CUDA(cudaGetLastError());
    CRX_TRY(xo.flush()); CRX_TRY(uo.flush()); CRX_TRY(mo.flush()); CRX_TRY(ko.flush());
    CRX_CUDA(cudaStreamSynchronize(c->stream));
    return CRX_OK;
}
