#!/usr/bin/env python
"""SURVEY.md section 8(d) "CPU reference timed beside it": the reference's own implementation (oracle/_ref, the
reference headers compiled here; the port when that is absent) on ONE host thread, at sizes that finish in seconds,
for every sub-path of the hot path.  TEST/BENCH INFRASTRUCTURE: this is the checker being timed, never the product.

    python tests/cpu_baselines.py          # prints one JSON object; `nproc` of the box is recorded
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives under tests/: only tests may use oracle/)
sys.path.insert(0, ROOT)


def timed(fn):
    t0 = time.perf_counter()
    r = fn()
    return time.perf_counter() - t0, r


def main():
    from oracle import load, COSINE, EUCLIDEAN
    from crypto_recommendation_b200 import synth
    o = load("reference") or load("port")
    out = {"kind": o.kind, "threads": 1, "host_cores": os.cpu_count()}
    rng = np.random.default_rng(0)
    # Lloyd assignment: linear in N*K -> pts*centroids/s
    n, K, D = 20_000, 1024, 128
    C = rng.normal(size=(K, D)) * 4
    X = C[rng.integers(0, K, n)] + rng.normal(size=(n, D))
    dt, _ = timed(lambda: o.lloyds_assignment(X, C, None, EUCLIDEAN))
    out["lloyd_assign_N20k_K1024_D128"] = {"s": dt, "pts_centroids_per_s": n * K / dt}
    # LSH recommendation (rec A): cost ~ N * |cand|
    for n in (2_500, 5_000):
        U, unk, mean = synth.rating_users(n, 100, seed=3)
        U = U.astype(np.float64)
        dt, _ = timed(lambda: o.recommend_lsh(U, unk, mean, COSINE, 4, 5, 100, 0.4, 20, 5, 42))
        out["lsh_recs_N%d" % len(U)] = {"s": dt, "recs_per_s": len(U) / dt}
    # k-means++ (string-keyed distance cache of N*K entries in the reference)
    n, K, D = 10_000, 32, 128
    X = synth.gaussian_mixture(n, D, 64, seed=4).astype(np.float64)
    dt, _ = timed(lambda: o.k_means_pp(X, K, EUCLIDEAN, 5))
    out["kmeanspp_N10k_K32_D128"] = {"s": dt, "ms_per_round": dt / (K - 1) * 1e3, "pts_per_s_per_round": n * (K - 1) / dt}
    # PAM update: clusters of <= 2000 members
    n, K, D = 4_000, 4, 100
    X = synth.gaussian_mixture(n, D, K, seed=5).astype(np.float64)
    cidx = o.rand_selection(X, K, 6)
    lab, _ = o.lloyds_assignment(X, X[cidx], cidx, EUCLIDEAN)
    dt, _ = timed(lambda: o.pam_lloyds(X, lab, cidx, EUCLIDEAN))
    pairs = float((np.bincount(lab, minlength=K).astype(np.float64) ** 2).sum())
    out["pam_N4k_K4_D100"] = {"s": dt, "pair_distances_per_s": pairs / dt}
    # range-search assignment (cube) and k-means update
    n, K, D = 50_000, 64, 128
    X = synth.gaussian_mixture(n, D, 64, seed=6).astype(np.float64)
    cidx = o.rand_selection(X, K, 7)
    dt, r = timed(lambda: o.cube_range_assignment(X, cidx, EUCLIDEAN, 10, 4.0, 32, 8))
    out["cube_range_assignment_N50k_K64"] = {"s": dt, "pts_per_s": n / dt}
    dt, _ = timed(lambda: o.k_means(X, r[0], X[cidx], EUCLIDEAN, 0.05))
    out["kmeans_update_N50k_K64_D128"] = {"s": dt, "pts_per_s": n / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
