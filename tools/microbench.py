#!/usr/bin/env python
"""Per-kernel measurements of the secondary (HBM-bound) kernels of the path on one B200, against the measured
HBM peak: hash projection + table build (C2, C3 shapes), k-means++ round, cluster sums, cube / LSH range-search
assignment, PAM update.  Prints one JSON object; numbers go into DESIGN.md section 8.

    python tools/microbench.py [--scale 1.0]
"""
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--pam-only", action="store_true")
    ap.add_argument("--pam-points", type=int, default=500_000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = capi.Context(0, stream.cuda_stream)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]
    out = {"hbm_peak_gbs": hbm}

    def gen(n, d, k, seed):
        g = torch.Generator(device=dev); g.manual_seed(seed)
        c = torch.randn((k, d), generator=g, device=dev) * 4.0
        X = torch.empty((n, d), dtype=torch.float32, device=dev)
        for lo in range(0, n, 1 << 20):
            hi = min(n, lo + (1 << 20))
            X[lo:hi] = c[torch.randint(0, k, (hi - lo,), generator=g, device=dev)] + torch.randn((hi - lo, d), generator=g, device=dev)
        return X

    def timed(name, fn, prefixes, reps=3):
        fn()
        ctx.profile_reset(); ctx.profile(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        ctx.profile(False)
        ks = {p: round(ctx.kernel_time(p)[0] / reps, 4) for p in prefixes}
        return {"wall_ms": round(wall, 3), "kernel_ms": ks}

    def c2_c3():
        # ---- C2 shape: 1M x 100, cosine L=5 k=4
        n = int(1_000_000 * a.scale)
        X = gen(n, 100, 64, 1)
        P = capi.Points(ctx, X)
        r = timed("lsh_c2", lambda: capi.LshTables(ctx, P, "cosine", 4, 5, 100, 0.4, 7).close(), ["hash_rows", "hash_rows32", "bucket_offsets", "iota"])
        r["hash_gbs"] = n * (4 * 100 + 4 * 5) / (r["kernel_ms"]["hash_rows"] * 1e6)
        r["hash_frac_of_hbm"] = r["hash_gbs"] / hbm
        out["lsh_build_c2_1M_x100_cos_L5k4"] = r
        P.close(); del X
        # ---- C3 shape: 10M x 128 euclidean cube d'=16
        n = int(10_000_000 * a.scale)
        X = gen(n, 128, 1024, 2)
        P = capi.Points(ctx, X)
        del X
        cube = [None]
        def build_cube():
            if cube[0] is not None: cube[0].close()
            cube[0] = capi.Hypercube(ctx, P, "euclidean", 16, 4.0, 9)
        r = timed("cube_c3", build_cube, ["hash_rows", "hash_rows32", "cube_keys", "cube_heads", "cube_minmax", "cube_first", "cube_vertex", "bucket_offsets"], reps=2)
        r["hash_gbs"] = n * (4 * 128 + 4 * 16) / (r["kernel_ms"]["hash_rows"] * 1e6)
        r["hash_frac_of_hbm"] = r["hash_gbs"] / hbm
        out["cube_build_c3_10M_x128_d16"] = r
        K = 1024
        cidx = capi.rand_selection(ctx, P, K, 3)
        dev_out = (torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev))
        r = timed("cube_range", lambda: capi.cube_range_assignment(ctx, P, cube[0], cidx, "euclidean", 64, out=dev_out), ["range_fire", "range_finalize", "range_hist", "min_pair", "fill_key", "lloyd_scan", "compact", "tc_argmin", "lloyd_refine",
                    "gather_rows", "gather_int", "pad_centroids", "merge_remaining", "self_assign", "maxabs", "half_norm", "tc_prep", ""], reps=1)
        r["kernel_ms_total"] = r["kernel_ms"].pop("")
        # how many (centroid, bucket member) pairs the probes evaluate: home vertex + 64 extra vertices per centroid
        vid = cube[0].vertex_ids()
        sizes = np.bincount(vid, minlength=1 << 16)
        pairs = 0
        for cr in cidx:
            home = int(vid[cr])
            seq = [home] + capi.get_num_hamming_dist_from(home, 1, 0, 16) + capi.get_num_hamming_dist_from(home, 2, 0, 16)
            pairs += int(sizes[seq[:65]].sum())
        r["probe_pairs"] = pairs
        r["range_fire_gather_gbs"] = pairs * 512 / (r["kernel_ms"]["range_fire"] * 1e6)
        r["vertex_sizes"] = {"nonempty": int((sizes > 0).sum()), "max": int(sizes.max()), "mean_nonempty": float(sizes[sizes > 0].mean())}
        out["cube_range_assignment_c3_K1024_probes64"] = r
        # ---- k-means++ rounds on the same 10M x 128 points (K = 9: 8 rounds)
        r = timed("kpp", lambda: capi.k_means_pp(ctx, P, 9, "euclidean", 5), ["kpp_update", "kpp_filter", "kpp_prune", "kpp_cdist", "kpp_prob", "kpp_pick"], reps=1)
        names = ["kpp_update", "kpp_filter", "kpp_prune", "kpp_cdist", "kpp_prob", "kpp_pick"]
        r17 = timed("kpp", lambda: capi.k_means_pp(ctx, P, 65, "euclidean", 5), names, reps=1)
        r9 = timed("kpp", lambda: capi.k_means_pp(ctx, P, 33, "euclidean", 5), names, reps=1)
        # rounds 33..64 = difference of the two runs: steady state (prune pass + exact update of the listed rows + prob + pick)
        r["steady_round_ms_kernels"] = (sum(r17["kernel_ms"].values()) - sum(r9["kernel_ms"].values())) / 32
        r["wall_ms_K65"] = r17["wall_ms"]
        r["steady_round_ms_wall"] = (r17["wall_ms"] - r9["wall_ms"]) / 32
        r["steady_gbs"] = n * (4 * 128 + 8) / (r["steady_round_ms_kernels"] * 1e6)
        r["steady_frac_of_hbm"] = r["steady_gbs"] / hbm
        out["kmeanspp_round_10M_x128"] = r
        # ---- cluster sums (k-means update) on 10M x 128, K=1024
        lab = torch.randint(0, K, (n,), dtype=torch.int32, device=dev)
        sums = torch.empty((K, 128), dtype=torch.float64, device=dev); counts = torch.empty(K, dtype=torch.int64, device=dev)
        r = timed("sums", lambda: capi.cluster_sums(ctx, P, lab, K, sums, counts), ["chunk_sums", "combine_sums", "bucket_offsets", "iota"])
        r["gbs"] = n * (4 * 128 + 4) / (r["wall_ms"] * 1e6)
        r["frac_of_hbm"] = r["gbs"] / hbm
        out["cluster_sums_10M_x128_K1024"] = r
        cube[0].close(); P.close()

    if not a.pam_only:
        c2_c3()
    # ---- C5 shape (scaled): PAM update + LSH range assignment, 5M x 100, K=256 -> here 500k x 100, K=256
    n = int(a.pam_points * a.scale)
    X = gen(n, 100, 256, 4)
    P = capi.Points(ctx, X)
    del X
    K = 256
    r = timed("lsh_c5", lambda: capi.LshTables(ctx, P, "euclidean", 4, 5, 100, 0.4, 11).close(), ["hash_rows", "hash_rows32", "gather_h", "tuple_flag", "scatter_rank", "bucket_offsets", "iota", ""], reps=1)
    r["kernel_ms_total"] = r["kernel_ms"].pop("")
    out["lsh_build_c5_%d_x100_euc_L5k4" % n] = r
    t = capi.LshTables(ctx, P, "euclidean", 4, 5, 100, 0.4, 11)
    cidx = capi.k_means_pp(ctx, P, K, "euclidean", 6)
    r = timed("lsh_range", lambda: capi.lsh_range_assignment(ctx, P, t, cidx, "euclidean"), ["range_fire", "range_finalize", "lloyd_scan", "tc_argmin", "lloyd_refine"], reps=1)
    out["lsh_range_assignment_%d_x100_K256" % n] = r
    lab, _, _ = capi.lsh_range_assignment(ctx, P, t, cidx, "euclidean")
    r = timed("pam", lambda: capi.pam_lloyds(ctx, P, lab, cidx, "euclidean"), ["pam_rowsum", "pam_pick", "tc_rowsum_scan", "tc_prep", "pam_bounds", "pam_exact", "pam_final"], reps=1)
    sizes = np.bincount(lab, minlength=K).astype(np.float64)
    r["pair_distances"] = float((sizes ** 2).sum())
    r["pairs_per_s"] = r["pair_distances"] / (max(r["kernel_ms"].get("pam_rowsum", 0), r["kernel_ms"].get("tc_rowsum_scan", 0), 1e-9) * 1e-3)
    r["pam_exact_resums"] = ctx.counters()["pam_exact"]
    out["pam_update_%d_x100_K256" % n] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
