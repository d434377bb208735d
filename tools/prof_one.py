#!/usr/bin/env python
"""Small drivers for profiling one secondary kernel under ncu:  python tools/prof_one.py kpp|sums|hash|pam"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from crypto_recommendation_b200 import capi
    what = sys.argv[1]
    dev = torch.device("cuda", 0)
    ctx = capi.Context(0)
    g = torch.Generator(device=dev); g.manual_seed(1)

    def gen(n, d, k):
        c = torch.randn((k, d), generator=g, device=dev) * 4.0
        X = torch.empty((n, d), dtype=torch.float32, device=dev)
        for lo in range(0, n, 1 << 20):
            hi = min(n, lo + (1 << 20))
            X[lo:hi] = c[torch.randint(0, k, (hi - lo,), generator=g, device=dev)] + torch.randn((hi - lo, d), generator=g, device=dev)
        return X

    if what == "kpp":
        P = capi.Points(ctx, gen(10_000_000, 128, 1024))
        print(capi.k_means_pp(ctx, P, 6, "euclidean", 5))
    elif what == "kpp1024":
        import time
        P = capi.Points(ctx, gen(10_000_000, 128, 1024))
        capi.k_means_pp(ctx, P, 3, "euclidean", 5)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rows = capi.k_means_pp(ctx, P, 1024, "euclidean", 5)
        torch.cuda.synchronize()
        print("k-means++ K=1024 on 10M x 128: %.1f ms, checksum %d" % ((time.perf_counter() - t0) * 1e3, int(rows.astype(np.int64).sum())))
    elif what == "sums":
        n, K = 10_000_000, 1024
        P = capi.Points(ctx, gen(n, 128, 1024))
        lab = torch.randint(0, K, (n,), dtype=torch.int32, device=dev)
        sums = torch.empty((K, 128), dtype=torch.float64, device=dev); counts = torch.empty(K, dtype=torch.int64, device=dev)
        capi.cluster_sums(ctx, P, lab, K, sums, counts)
        print(float(sums.sum()))
    elif what == "hash":
        P = capi.Points(ctx, gen(10_000_000, 128, 1024))
        cube = capi.Hypercube(ctx, P, "euclidean", 16, 4.0, 9)
        print(int(cube.vertex_ids()[:5].sum()))
    elif what == "pam":
        n, K = 2_000_000, 128
        P = capi.Points(ctx, gen(n, 100, K))
        cidx = capi.k_means_pp(ctx, P, K, "euclidean", 6)
        C = torch.from_numpy(np.zeros((K, 100))).to(dev)
        lab = torch.randint(0, K, (n,), dtype=torch.int32, device=dev).cpu().numpy()
        print(capi.pam_lloyds(ctx, P, lab, cidx, "euclidean")[0])
    elif what == "lloyd_pam":
        # one ARGMIN launch at the C4 shard size, then one ROWSUM launch at a C5-like size
        n, K = 12_500_000, 1024
        X = gen(n, 128, K)
        P = capi.Points(ctx, X)
        C = X[torch.randint(0, n, (K,), generator=g, device=dev)].to(torch.float64).contiguous()
        lab = torch.empty(n, dtype=torch.int32, device=dev); dis = torch.empty(n, dtype=torch.float64, device=dev)
        capi.lloyds_assignment(ctx, P, C, None, "euclidean", lab, dis)
        print(int(lab[:5].sum()))
        P.close(); del X, lab, dis
        n, K = 5_000_000, 256
        P = capi.Points(ctx, gen(n, 100, K))
        cidx = capi.k_means_pp(ctx, P, K, "euclidean", 6)
        lab = torch.randint(0, K, (n,), dtype=torch.int32, device=dev).cpu().numpy()
        print(capi.pam_lloyds(ctx, P, lab, cidx, "euclidean")[0])
    ctx.synchronize()


if __name__ == "__main__":
    main()
