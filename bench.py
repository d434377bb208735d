#!/usr/bin/env python
"""bench.py -- the reference's headline loop on B200: recs/sec of the cosine-LSH top-P recommendation
(BASELINE.json configs[1]: 1M users x 100 coins, L=5, k=4, P=20, top-5 coins), plus the Lloyd
assignment rate (pts*centroids/s) on one GPU's shard of configs[3] (100M x 128, K=1024 over 8 GPUs).

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A step = one pass of the hot path over one batch: create_LSH_hashtables over the resident user
vectors + the per-user loop of main.cpp:159-170 for every user (tables, candidates, top-P cosine
neighbours, rating prediction, top-5 coins).  `value` times it with the inputs resident in HBM;
`e2e` times the same through the C ABI with HOST buffers (upload + results back) inside the region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_HASH, L_TABLES, P_NEIGH, N_REC = 4, 5, 20, 5
LSH_BUCKET_DIV, EUCLID_W = 100, 0.4
SEED = 0xC0FFEE + 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="crx", choices=["crx", "reference"])
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--coins", type=int, default=100)
    ap.add_argument("--lloyd-points", type=int, default=12_500_000)
    ap.add_argument("--lloyd-k", type=int, default=1024)
    ap.add_argument("--lloyd-d", type=int, default=128)
    ap.add_argument("--no-lloyd", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--min-warmup", type=int, default=3, help="profiling runs only: allow fewer than 3 warm-up steps")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "which": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "which": "fallback"}


def make_users(n, d, seed):
    """exactly n users (about 5% of the generated ones are all-zero and dropped, crypto_rec.hpp:127)"""
    from crypto_recommendation_b200 import synth
    X, u, m = synth.rating_users_fast(int(n * 1.08) + 64, d, seed)
    assert X.shape[0] >= n
    return np.ascontiguousarray(X[:n]), np.ascontiguousarray(u[:n]), np.ascontiguousarray(m[:n])


# ------------------------------------------------------------------------------------------------
# the reference's CPU path (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(U, unk, mean, threads, seconds, steps=1, warmup=0):
    """recs/s of the reference's own loop on `threads` host threads sharing one set of tables.
    Bounded sample: each step runs `threads * per` consecutive users against the FULL table; the table
    build is timed once and charged pro rata (build_s * sample / N)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import load, RecHandle, COSINE
    n = U.shape[0]
    ref = load("reference")
    # the reference build copies every user into CustVector/std::set objects (~5 KB/user, ~20 s per 1M users)
    o = ref if (ref is not None and n <= 1_100_000) else load("port")
    U64 = np.ascontiguousarray(U, dtype=np.float64)
    t0 = time.perf_counter()
    h = RecHandle(o, U64, unk, mean, COSINE, K_HASH, L_TABLES, LSH_BUCKET_DIV, EUCLID_W, SEED)
    build_s = time.perf_counter() - t0
    # calibrate the per-query cost on one thread
    t0 = time.perf_counter()
    h.query(0, 2, P_NEIGH, N_REC)
    per_q = max(1e-4, (time.perf_counter() - t0) / 2)
    per = max(1, int(seconds / max(1, steps + warmup) / per_q))
    sample = per * threads
    rates = []
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for it in range(warmup + steps):
            base = (it * sample) % max(1, n - sample)
            t0 = time.perf_counter()
            list(ex.map(lambda t: h.query(base + t * per, base + (t + 1) * per, P_NEIGH, N_REC), range(threads)))
            dt = time.perf_counter() - t0
            if it >= warmup:
                rates.append(sample / (dt + build_s * sample / n))
    h.close()
    return {"value": float(np.mean(rates)), "unit": "recs/s", "cores": threads, "kind": o.kind,
            "sample": "%d users/step (x%d steps) of the %d-user rec-A loop against the full table; table build %.1fs charged pro rata"
                      % (sample, steps, n, build_s)}, sample


def cpu_lloyd_rate(dd, kk, seconds):
    """pts*centroids/s of the reference's own lloyds_assignment (assignment.hpp:55-80) on one host core: a bounded
    sample of points of the same mixture against all K centroids."""
    from oracle import load, EUCLIDEAN
    o = load("reference") or load("port")
    rng = np.random.default_rng(SEED)
    C = rng.normal(size=(kk, dd)) * 4.0
    probe = 200
    X = C[rng.integers(0, kk, probe)] + rng.normal(size=(probe, dd))
    t0 = time.perf_counter()
    o.lloyds_assignment(X, C, None, EUCLIDEAN)
    per_pt = max(1e-6, (time.perf_counter() - t0) / probe)
    n = int(min(200_000, max(probe, seconds / per_pt)))
    X = C[rng.integers(0, kk, n)] + rng.normal(size=(n, dd))
    t0 = time.perf_counter()
    o.lloyds_assignment(X, C, None, EUCLIDEAN)
    dt = time.perf_counter() - t0
    return {"value": n * kk / dt, "unit": "pts*centroids/s", "cores": 1, "kind": o.kind,
            "sample": "%d points x %d centroids x %d dims, lloyds_assignment on one core" % (n, kk, dd)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    U, unk, mean = make_users(args.users, args.coins, SEED)
    t0 = time.perf_counter()
    cb, sample = cpu_reference_rate(U, unk, mean, threads, max(20.0, args.cpu_seconds * 3), steps=args.steps, warmup=args.warmup)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": "recs/sec (cosine LSH top-P recommendation)", "value": cb["value"], "unit": "recs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / cb["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (x87 long double dot products)",
        "data": "synthetic", "config": workload_config(U.shape[0], args.coins),
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "recs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))


def dram_traffic(kernel, units):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed `ncu --set full` capture
    (profiles/traffic.json, written next to the summaries); None when no capture exists for this workload size."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(kernel)
        if t and int(t.get("units", -1)) == int(units):
            return t["dram_bytes_per_launch"]
    except (OSError, ValueError):
        pass
    return None


def workload_config(n, d):
    return {"workload": "C2: %d users x %d coins, rating-like vectors (2-9 known coins/user), cosine LSH L=%d k=%d, top-P=%d neighbours, top-%d coins; "
                        "step = create_LSH_hashtables + rec-A loop of main.cpp:159-170 over every user" % (n, d, L_TABLES, K_HASH, P_NEIGH, N_REC),
            "users": n, "coins": d, "L": L_TABLES, "k": K_HASH, "P": P_NEIGH, "N_rec": N_REC,
            "l2": ("inputs (%.0f MB of user vectors + %.0f MB of split-fp16 operands) exceed the 126 MB L2; no explicit flush"
                   % (n * d * 4 / 1e6, n * 512 / 1e6)) if n * d * 4 > 126e6 else
                  ("inputs (%.0f MB) FIT the 126 MB L2 at this --users: not a valid bench size, use the default" % (n * d * 4 / 1e6))}


# ------------------------------------------------------------------------------------------------
# this engine
# ------------------------------------------------------------------------------------------------
def run_crx(args):
    import torch
    from crypto_recommendation_b200 import capi, synth
    from crypto_recommendation_b200 import dist as cdist
    rank, local_rank, world = cdist.init_process_group()
    assert torch.cuda.is_available(), "bench.py --impl crx needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(dev)  # one explicit stream shared by torch (events, generators) and the engine
    torch.cuda.set_stream(stream)
    ctx = capi.Context(local_rank, stream.cuda_stream)
    peaks = measured_peaks()

    # ---------------- C2: recommendation ----------------
    U, unk, mean = make_users(args.users, args.coins, SEED)  # every rank holds the replicated table
    n, d = U.shape
    U_pin = torch.from_numpy(U).pin_memory()
    unk_pin = torch.from_numpy(unk).pin_memory()
    mean_pin = torch.from_numpy(mean).pin_memory()
    P = capi.Points(ctx, U_pin.numpy(), unk_pin.numpy(), mean_pin.numpy())
    out_dev = {"recs": torch.empty((n, N_REC), dtype=torch.int32, device=dev), "ncand": torch.empty(n, dtype=torch.int32, device=dev)}

    def step_resident():
        t = capi.LshTables(ctx, P, "cosine", K_HASH, L_TABLES, LSH_BUCKET_DIV, EUCLID_W, SEED)
        capi.recommend_lsh(ctx, t, P_NEIGH, N_REC, out=out_dev)
        t.close()

    recs_host = torch.empty((n, N_REC), dtype=torch.int32).pin_memory()

    def step_e2e():
        p = capi.Points(ctx, U_pin.numpy(), unk_pin.numpy(), mean_pin.numpy())
        t = capi.LshTables(ctx, p, "cosine", K_HASH, L_TABLES, LSH_BUCKET_DIV, EUCLID_W, SEED)
        capi.recommend_lsh(ctx, t, P_NEIGH, N_REC, out={"recs": recs_host.numpy()})
        t.close()
        p.close()

    def timed(fn, steps, warmup, sample_clocks=False, profile=False):
        for _ in range(warmup):
            fn()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        cdist.barrier()
        torch.cuda.synchronize(dev)
        if profile:
            ctx.profile_reset(); ctx.profile(True)
            ctx.counters(reset=True)
        l0 = ctx.launch_count()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        cdist.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        launches = ctx.launch_count() - l0
        if profile:
            ctx.profile(False)
        return cdist.max_over_ranks(ms) / steps, launches, clocks

    ms_step, launches, clocks = timed(step_resident, args.steps, max(args.min_warmup, args.warmup), sample_clocks=True, profile=True)
    scan_name = "tc_topp_scan" if ctx.kernel_time("tc_topp_scan")[1] > 0 else "topp_scan"
    scan_ms, scan_launches = ctx.kernel_time(scan_name)
    kernel_ms = {k: round(ctx.kernel_time(k)[0] / args.steps, 3) for k in (
        "tc_topp_scan", "topp_scan", "rec_finalize", "rec_topn", "hash_rows", "tc_prep", "pack_codes", "subset_hist", "subset_count",
        "iota", "bucket_offsets", "fill_lists", "sq_sizes")}
    ncand_total = float(out_dev["ncand"].to(torch.float64).sum().item())
    counters = {k: v / args.steps for k, v in ctx.counters().items()}   # per step of 1M queries (near-tie / fallback counts)
    value = world * n / (ms_step / 1e3)
    flops_per_launch = 2.0 * d * ncand_total * args.steps / max(1, scan_launches)
    achieved_tf = flops_per_launch / (scan_ms / max(1, scan_launches) * 1e-3) / 1e12 if scan_ms > 0 else 0.0
    tc = scan_name == "tc_topp_scan"
    # tensor path: 3 fp16 products (hi*hi, lo*hi, hi*lo) over D rounded up to a multiple of 16 columns
    tc_factor = 3.0 * (16 * ((d + 15) // 16)) / d
    roofline = {"kernel": ("tc_scan_kernel<TOPP> (tcgen05 split-fp16 cosine scan, per-row top-64 in the epilogue)" if tc else
                           "topp_scan_kernel (FP64 SIMT masked cosine scan + per-query top-32 list)"), "bound": "tensor",
                "achieved": achieved_tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved_tf / peaks["bf16_tflops"],
                "traffic": dram_traffic("tc_topp_scan" if tc else "topp_scan", n), "peak_source": peaks["which"] + " bf16 dense GEMM (burst)",
                "pipe": ("tcgen05.mma kind::f16 M128 N256 K16, fp32 accumulate in TMEM; executed tensor flops = %.2fx algorithmic "
                         "(3 split-fp16 products, D padded to a multiple of 16)" % tc_factor) if tc else "FP64 SIMT FMA",
                "executed_tensor_tflops": achieved_tf * tc_factor if tc else None,
                "executed_frac_of_peak": achieved_tf * tc_factor / peaks["bf16_tflops"] if tc else None,
                "algorithmic_flops": "2*D*sum|cand(u)|", "kernel_ms_all": kernel_ms,
                "kernel_ms_per_step": scan_ms / args.steps, "share_of_step": scan_ms / args.steps / ms_step}

    e2e_ms, _, _ = timed(step_e2e, max(1, min(2, args.steps)), 2)
    h2d = U_pin.numel() * 4 + unk_pin.numel() + mean_pin.numel() * 8
    d2h = recs_host.numel() * 4
    e2e = {"value": world * n / (e2e_ms / 1e3), "unit": "recs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": e2e_ms}
    P.close()

    # ---------------- C4 shard: Lloyd assignment + k-means update ----------------
    lloyd = None
    if not args.no_lloyd:
        npts, dd, kk = args.lloyd_points, args.lloyd_d, args.lloyd_k
        g = torch.Generator(device=dev); g.manual_seed(1234)
        centres = torch.randn((kk, dd), generator=g, device=dev) * 4.0   # ONE mixture for the whole job ...
        g.manual_seed(1235 + rank)                                       # ... every rank draws its own shard of points from it
        X = torch.empty((npts, dd), dtype=torch.float32, device=dev)
        CH = 1 << 20
        for lo in range(0, npts, CH):  # mixture of K unit Gaussians, generated on the device in chunks
            hi = min(npts, lo + CH)
            which = torch.randint(0, kk, (hi - lo,), generator=g, device=dev)
            X[lo:hi] = centres[which] + torch.randn((hi - lo, dd), generator=g, device=dev)
        Q = capi.Points(ctx, X)
        del X
        C = centres.to(torch.float64).contiguous()
        if world > 1:
            torch.distributed.broadcast(C, 0)
        labels = torch.empty(npts, dtype=torch.int32, device=dev)
        dists = torch.empty(npts, dtype=torch.float64, device=dev)
        sums = torch.empty((kk, dd), dtype=torch.float64, device=dev)
        counts = torch.empty(kk, dtype=torch.int64, device=dev)
        newC = torch.empty_like(C)

        def step_assign():
            capi.lloyds_assignment(ctx, Q, C, None, "euclidean", labels, dists)

        def step_kmeans():
            capi.lloyds_assignment(ctx, Q, C, None, "euclidean", labels, dists)
            capi.cluster_sums(ctx, Q, labels, kk, sums, counts)
            cdist.allreduce_cluster_sums(sums, counts)
            capi.k_means_finish(ctx, sums, counts, C, "euclidean", 0.05, newC)

        def step_assign_labels():   # dists = NULL: what the k-means loop needs (main.cpp:96-103 never reads the stored distance)
            capi.lloyds_assignment(ctx, Q, C, None, "euclidean", labels, want_dists=False)

        def step_kmeans_labels():
            capi.lloyds_assignment(ctx, Q, C, None, "euclidean", labels, want_dists=False)
            capi.cluster_sums(ctx, Q, labels, kk, sums, counts)
            cdist.allreduce_cluster_sums(sums, counts)
            capi.k_means_finish(ctx, sums, counts, C, "euclidean", 0.05, newC)

        a_ms, a_launch, _ = timed(step_assign, args.steps, args.min_warmup)
        k_ms, _, _ = timed(step_kmeans, args.steps, 3)   # the first all-reduces set up NCCL channels
        al_ms, _, _ = timed(step_assign_labels, args.steps, 1)
        kl_ms, _, _ = timed(step_kmeans_labels, args.steps, 2)
        ctx.profile_reset(); ctx.profile(True)
        step_kmeans(); torch.cuda.synchronize(dev)
        ctx.profile(False)
        breakdown = {k: round(ctx.kernel_time(k)[0], 3) for k in ("tc_argmin_scan", "lloyd_refine", "lloyd_scan", "tc_prep", "maxabs", "half_norm", "chunk_sums",
                                                                   "combine_sums", "kmeans_finish", "select_centroids", "bucket_offsets", "iota", "pad_centroids")}
        ctx.profile_reset(); ctx.profile(True)
        for _ in range(args.steps):
            step_assign()
        torch.cuda.synchronize(dev)
        ctx.profile(False)
        ltc = ctx.kernel_time("tc_argmin_scan")[1] > 0
        lk_ms, lk_n = ctx.kernel_time("tc_argmin_scan" if ltc else "lloyd_scan")
        lr_ms, _ = ctx.kernel_time("lloyd_refine")
        lfac = 3.0 * (16 * ((dd + 15) // 16)) / dd
        fl = 2.0 * dd * npts * kk
        lloyd = {"metric": "Lloyd assign pts*centroids/s", "value": world * npts * kk / (a_ms / 1e3), "unit": "pts*centroids/s",
                 "ms_per_step": a_ms, "kmeans_iteration_ms": k_ms, "kmeans_kernel_ms": breakdown, "scaling": "weak",
                 "labels_only": {"assign_ms": al_ms, "kmeans_iteration_ms": kl_ms, "value": world * npts * kk / (al_ms / 1e3),
                                 "note": "crx_lloyds_assignment with dists = NULL; NOT the headline (the reference stores the distance)"},
                 "config": {"workload": "C4 shard: %d x %d fp32 points per GPU, K=%d, euclidean; bit-exact FP64 distances" % (npts, dd, kk)},
                 "roofline": {"kernel": "tc_scan_kernel<ARGMIN> (tcgen05 split-fp16 filter) + lloyd_refine_kernel (exact FP64 distance of the winner)" if ltc else "lloyd_scan_kernel (FP64 SIMT)",
                              "bound": "tensor", "achieved": fl / (lk_ms / max(1, lk_n) * 1e-3) / 1e12,
                              "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                              "frac": fl / (lk_ms / max(1, lk_n) * 1e-3) / 1e12 / peaks["bf16_tflops"],
                              "traffic": dram_traffic("tc_argmin_scan" if ltc else "lloyd_scan", npts),
                              "executed_tensor_tflops": fl * lfac / (lk_ms / max(1, lk_n) * 1e-3) / 1e12 if ltc else None,
                              "pipe": ("tcgen05.mma kind::f16, 3 split-fp16 products (%.2fx algorithmic flops), then exact FP64 refine" % lfac) if ltc
                                      else "FP64 SIMT (sub, mul, add per element: 3 FP64 ops for the 2 algorithmic flops)",
                              "scan_ms": lk_ms / max(1, lk_n), "refine_ms": lr_ms / max(1, lk_n),
                              "hbm_floor_ms": npts * (4 * dd + 12) / (peaks["hbm_gbs"] * 1e6)}}
        Q.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_rate(U, unk, mean, 1, args.cpu_seconds)
        if lloyd is not None:
            lloyd["cpu_baseline"] = cpu_lloyd_rate(args.lloyd_d, args.lloyd_k, min(8.0, args.cpu_seconds))

    if rank == 0:
        line = {"metric": "recs/sec (cosine LSH top-P recommendation)", "value": value, "unit": "recs/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.min_warmup, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(n, d), "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "lloyd": lloyd,
                "mean_candidates_per_user": ncand_total / n, "exactness_counters_per_step": counters}
        line["config"]["parallelism"] = "queries: %d independent replicas of the full batch (tables replicated); Lloyd: rows sharded, NCCL all-reduce of sums" % world
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_crx(a)
