#!/usr/bin/env python
"""Timings of the sharded clustering phases over NCCL (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/sharded_bench.py [--scale 1.0]

  * C4 slice: k-means++ rounds + k-means iteration over rows sharded across the ranks (12.5M x 128 per rank, weak)
  * C3: cube range-search assignment, 10M x 128, K=1024, probes 64 -- centroids split, points replicated (strong)
  * C5: PAM update, 5M x 100, K=256 -- candidate rows split, points replicated (strong)
Device time = max over ranks of the wall time between two barriers + synchronize.  Rank 0 prints one JSON object."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from crypto_recommendation_b200 import capi
    from crypto_recommendation_b200 import dist as cdist
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--kpp-k", type=int, default=17, help="centroids drawn by the sharded k-means++ (C4: 1024)")
    ap.add_argument("--only-c4", action="store_true", help="skip the replicated C3 / C5 sections")
    a = ap.parse_args()
    rank, local_rank, world = cdist.init_process_group()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = capi.Context(local_rank, stream.cuda_stream)
    comm = cdist.Comm()
    out = {"world": world}

    def gen(n, d, k, seed):
        g = torch.Generator(device=dev); g.manual_seed(seed)
        c = torch.randn((k, d), generator=g, device=dev) * 4.0
        X = torch.empty((n, d), dtype=torch.float32, device=dev)
        for lo in range(0, n, 1 << 20):
            hi = min(n, lo + (1 << 20))
            X[lo:hi] = c[torch.randint(0, k, (hi - lo,), generator=g, device=dev)] + torch.randn((hi - lo, d), generator=g, device=dev)
        return X

    def timed(fn, reps=2):
        fn()
        ts = []
        for _ in range(reps):
            cdist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize(); cdist.barrier()
            ts.append(cdist.max_over_ranks((time.perf_counter() - t0) * 1e3))
        return round(min(ts), 3)

    # ---- C4 slice (weak): every rank holds its own 12.5M x 128 rows
    n = int(12_500_000 * a.scale)
    X = gen(n, 128, 1024, 100 + rank)     # different rows on every rank
    P = capi.Points(ctx, X)
    del X
    K = a.kpp_k
    ms = timed(lambda: capi.k_means_pp_sharded(ctx, P, rank * n, world * n, K, "euclidean", 5, comm), reps=1)
    out["kmeanspp_sharded_%dx128" % (world * n)] = {"ms_total": ms, "ms_per_round": round(ms / (K - 1), 3), "rounds": K - 1,
                                                    "points_total": world * n}
    rows, vecs = capi.k_means_pp_sharded(ctx, P, rank * n, world * n, 64, "euclidean", 5, comm)
    C = torch.from_numpy(vecs).to(dev)
    lab = torch.empty(n, dtype=torch.int32, device=dev); dis = torch.empty(n, dtype=torch.float64, device=dev)
    capi.lloyds_assignment(ctx, P, C, None, "euclidean", lab, dis)
    ms = timed(lambda: capi.k_means_sharded(ctx, P, lab, C, "euclidean", 0.05, comm))
    out["kmeans_update_sharded_K64"] = {"ms": ms}
    P.close(); del lab, dis
    if a.only_c4:
        out["collectives"] = dict(comm.calls)
        out["comm_error"] = comm.error
        if rank == 0:
            print(json.dumps(out))
        cdist.barrier()
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
        return
    # ---- C3 (strong): points replicated, centroids split
    n = int(10_000_000 * a.scale)
    X = gen(n, 128, 1024, 2)
    P = capi.Points(ctx, X)
    del X
    cube = capi.Hypercube(ctx, P, "euclidean", 16, 4.0, 9)
    cidx = capi.rand_selection(ctx, P, 1024, 3)
    dev_out = (torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev))
    ms = timed(lambda: capi.cube_range_assignment(ctx, P, cube, cidx, "euclidean", 64, comm=comm, out=dev_out))
    del dev_out
    out["cube_range_assignment_c3_%d_K1024_probes64" % n] = {"ms": ms}
    cube.close(); P.close()
    # ---- C5 (strong): points replicated, candidate rows split
    n = int(5_000_000 * a.scale)
    X = gen(n, 100, 256, 4)
    P = capi.Points(ctx, X)
    del X
    cidx = capi.k_means_pp(ctx, P, 256, "euclidean", 6)
    t = capi.LshTables(ctx, P, "euclidean", 4, 5, 100, 0.4, 11)
    lab, _, _ = capi.lsh_range_assignment(ctx, P, t, cidx, "euclidean", comm=comm)
    dev_out = (torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev))
    ms = timed(lambda: capi.lsh_range_assignment(ctx, P, t, cidx, "euclidean", comm=comm, out=dev_out))
    del dev_out
    out["lsh_range_assignment_c5_%d_K256" % n] = {"ms": ms}
    ms = timed(lambda: capi.pam_lloyds(ctx, P, lab, cidx, "euclidean", comm=comm))
    sizes = np.bincount(lab, minlength=256).astype(np.float64)
    out["pam_update_c5_%d_K256" % n] = {"ms": ms, "pair_distances": float((sizes ** 2).sum()),
                                        "pairs_per_s": float((sizes ** 2).sum()) / (ms * 1e-3)}
    out["collectives"] = dict(comm.calls)
    out["comm_error"] = comm.error
    if rank == 0:
        print(json.dumps(out))
    cdist.barrier()
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
