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
    ap.add_argument("--only", default="", help="comma-separated blocks to run: c2,c2_normal,lloyd,kmeanspp,cube_range,pam,c1_main,lloyd_100m (default: all)")
    ap.add_argument("--c1-users", type=int, default=6000, help="tweeting users of the c1_main block (the reference's own scale)")
    ap.add_argument("--cube-points", type=int, default=10_000_000)
    ap.add_argument("--pam-points", type=int, default=5_000_000)
    ap.add_argument("--kpp-k", type=int, default=1024, help="centroids drawn by the k-means++ block (C4: 1024)")
    ap.add_argument("--lloyd-full", type=int, default=100_000_000, help="rows of the one-GPU run of the whole C4 config (0 = skip; N = 1 only)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML in a thread of this process
    (nvidia_ml_py): a polling `nvidia-smi -lms 200` child cost the C2 step 130 ms (15%) through the driver lock; the
    nvidia-smi loop remains as the fallback, at one sample per second."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.sm, self.reasons, self.sm_max = [], set(), None

    def _nvml_loop(self):
        n, h = self.nvml
        names = (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap))
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in names:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber the devices: match NVML's handle through the UUID torch reports when possible
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                for i in range(n.nvmlDeviceGetCount()):
                    hi = n.nvmlDeviceGetHandleByIndex(i)
                    u = n.nvmlDeviceGetUUID(hi)
                    u = u.decode() if isinstance(u, bytes) else u
                    if uuid in u:
                        h = hi
            except Exception:
                h = None
            if h is None:
                h = n.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            self.nvml = (n, h)
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "1000"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


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
    per = max(1, min(n // (2 * threads), int(seconds / max(1, steps + warmup) / per_q)))   # the sample must fit the table
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
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64 (x87 long double dot products)",
        "data": "synthetic", "config": workload_config(U.shape[0], args.coins, args.gpus),
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


def workload_config(n, d, world):
    if world > 1:
        par = ("queries of the ONE %d-user batch sharded over %d GPUs (tables replicated on every GPU, rank r answers users [r n/N, (r+1) n/N)), "
               "recs all-gathered over NCCL inside the timed region" % (n, world))
    else:
        par = "one GPU holds the tables and answers every query"
    return {"workload": "C2: %d users x %d coins, rating-like vectors (every coin known with probability nk/%d, nk ~ U{2..9}: 5.5%% of the users "
                        "know ONE coin and have constant vectors = one clique of ~%dk mutually tied neighbours), cosine LSH L=%d k=%d, top-P=%d "
                        "neighbours, top-%d coins; step = create_LSH_hashtables + rec-A loop of main.cpp:159-170 over every user"
                        % (n, d, d, round(0.055 * n / 1e3), L_TABLES, K_HASH, P_NEIGH, N_REC),
            "users": n, "coins": d, "L": L_TABLES, "k": K_HASH, "P": P_NEIGH, "N_rec": N_REC,
            "l2": ("inputs (%.0f MB of user vectors + %.0f MB of split-fp16 operands) exceed the 126 MB L2; no explicit flush"
                   % (n * d * 4 / 1e6, n * 512 / 1e6)) if n * d * 4 > 126e6 else
                  ("inputs (%.0f MB) FIT the 126 MB L2 at this --users: not a valid bench size, use the default" % (n * d * 4 / 1e6)),
            "parallelism": par}


# ------------------------------------------------------------------------------------------------
# this engine
# ------------------------------------------------------------------------------------------------
class Rig:
    """device, stream, engine context, communicator and the timing rule shared by every block"""

    def __init__(self, args):
        import torch
        from crypto_recommendation_b200 import capi
        from crypto_recommendation_b200 import dist as cdist
        self.torch, self.capi, self.cdist, self.args = torch, capi, cdist, args
        self.rank, self.local_rank, self.world = cdist.init_process_group()
        assert torch.cuda.is_available(), "bench.py --impl crx needs a CUDA device (there is no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.stream = torch.cuda.Stream(self.dev)  # one explicit stream shared by torch (events, generators) and the engine
        torch.cuda.set_stream(self.stream)
        self.ctx = capi.Context(self.local_rank, self.stream.cuda_stream)
        self.comm = cdist.NcclComm(self.ctx) if self.world > 1 else None   # collectives issued by libcrx.so itself
        self.peaks = measured_peaks()

    def timed(self, fn, steps, warmup, sample_clocks=False, profile=False):
        torch, ctx, cdist = self.torch, self.ctx, self.cdist
        for _ in range(warmup):
            fn()
        sampler = ClockSampler(self.local_rank) if sample_clocks else None
        cdist.barrier()
        torch.cuda.synchronize(self.dev)
        if profile:
            ctx.profile_reset(); ctx.profile(True)
            ctx.counters(reset=True)
        l0 = ctx.launch_count()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        torch.cuda.synchronize(self.dev)
        cdist.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        launches = ctx.launch_count() - l0
        if profile:
            ctx.profile(False)
        return cdist.max_over_ranks(ms) / steps, launches, clocks

    def kernel_ms(self, names, per):
        return {k: round(self.ctx.kernel_time(k)[0] / per, 3) for k in names}

    def gen_mixture(self, n, d, k, seed, centre_seed=None):
        """n x d fp32 points of a mixture of k unit Gaussians (centres N(0, 4^2)), generated on the device in chunks"""
        torch, dev = self.torch, self.dev
        g = torch.Generator(device=dev); g.manual_seed(centre_seed if centre_seed is not None else seed)
        c = torch.randn((k, d), generator=g, device=dev) * 4.0
        g.manual_seed(seed + 7919)
        X = torch.empty((n, d), dtype=torch.float32, device=dev)
        for lo in range(0, n, 1 << 20):
            hi = min(n, lo + (1 << 20))
            X[lo:hi] = c[torch.randint(0, k, (hi - lo,), generator=g, device=dev)] + torch.randn((hi - lo, d), generator=g, device=dev)
        return X, c

    def release(self):
        self.ctx.trim()
        self.torch.cuda.empty_cache()


P2_KERNELS = ("uniform_rows", "p2u_flag", "p2u_gather", "p2u_select", "p2u_sort", "p2u_compact", "p2_prepare", "tc_gather", "tc_collect_scan",
              "p2_all_count", "p2_sizes", "p2_trim", "p2_expand", "p2_fill_all", "p2_exact", "p2_resolve", "p2_collect_simt", "p2_leftover")


def bench_c2(rig, args):
    """BASELINE.json configs[1]: the headline.  N > 1: strong scaling -- the same 1M-user batch, queries sharded."""
    torch, capi, cdist, ctx, dev = rig.torch, rig.capi, rig.cdist, rig.ctx, rig.dev
    world, rank = rig.world, rig.rank
    U, unk, mean = make_users(args.users, args.coins, SEED)  # every rank holds the replicated table
    n, d = U.shape
    lo, hi = cdist.shard_range(n, rank, world)
    per = -(-n // world)     # rows per rank in the gathered buffer (the last shard may be shorter)
    U_pin = torch.from_numpy(U).pin_memory()
    unk_pin = torch.from_numpy(unk).pin_memory()
    mean_pin = torch.from_numpy(mean).pin_memory()
    P = capi.Points(ctx, U_pin.numpy(), unk_pin.numpy(), mean_pin.numpy())
    mine = {"recs": torch.zeros((per, N_REC), dtype=torch.int32, device=dev), "ncand": torch.zeros(per, dtype=torch.int32, device=dev)}
    gathered = torch.empty((world * per, N_REC), dtype=torch.int32, device=dev) if world > 1 else None

    def answer(points, out):
        t = capi.LshTables(ctx, points, "cosine", K_HASH, L_TABLES, LSH_BUCKET_DIV, EUCLID_W, SEED)
        view = {k: v[:hi - lo] for k, v in out.items()}
        capi.recommend_lsh(ctx, t, P_NEIGH, N_REC, q_begin=lo, q_end=hi, out=view)
        t.close()
        if world > 1:
            rig.comm.allgather(out["recs"], gathered)

    def step_resident():
        answer(P, mine)

    recs_host = torch.empty((world * per if world > 1 else n, N_REC), dtype=torch.int32).pin_memory()

    def step_e2e():
        p = capi.Points(ctx, U_pin.numpy(), unk_pin.numpy(), mean_pin.numpy())
        answer(p, {"recs": mine["recs"]})
        if rank == 0:   # the caller's answer: every user's coins in host memory
            recs_host.copy_(gathered if world > 1 else mine["recs"][:n], non_blocking=True)
        p.close()

    warm = max(args.min_warmup, args.warmup)
    ms_step, launches, clocks = rig.timed(step_resident, args.steps, warm, sample_clocks=True, profile=True)
    scan_name = "tc_topp_scan" if ctx.kernel_time("tc_topp_scan")[1] > 0 else "topp_scan"
    scan_ms, scan_launches = ctx.kernel_time(scan_name)
    kernel_ms = rig.kernel_ms(("tc_topp_scan", "topp_scan", "rec_finalize", "hash_rows", "tc_prep", "pack_codes", "subset_hist", "subset_count",
                               "bucket_offsets") + P2_KERNELS, args.steps)
    ncand_local = float(mine["ncand"][:hi - lo].to(torch.float64).sum().item())
    counters = {k: v / args.steps for k, v in ctx.counters().items()}   # this rank's share of the queries, per step
    value = n / (ms_step / 1e3)
    flops_per_launch = 2.0 * d * ncand_local * args.steps / max(1, scan_launches)
    achieved_tf = flops_per_launch / (scan_ms / max(1, scan_launches) * 1e-3) / 1e12 if scan_ms > 0 else 0.0
    tc = scan_name == "tc_topp_scan"
    tc_factor = 3.0 * (16 * ((d + 15) // 16)) / d   # 3 fp16 products (hi*hi, lo*hi, hi*lo) over D rounded up to a multiple of 16
    peak = rig.peaks["bf16_tflops_sustained"] or rig.peaks["bf16_tflops"]
    roofline = {"kernel": ("tc_scan_kernel<TOPP> (tcgen05 split-fp16 cosine scan, per-row top-64 in the epilogue)" if tc else
                           "topp_scan_kernel (FP64 SIMT masked cosine scan + per-query top-32 list)"), "bound": "tensor",
                "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tf / peak,
                "traffic": dram_traffic("tc_topp_scan" if tc else "topp_scan", n) if world == 1 else None,
                "peak_source": rig.peaks["which"] + " bf16 dense GEMM, sustained (the kernel is timed inside a multi-second region); burst %.1f" % rig.peaks["bf16_tflops"],
                "pipe": ("tcgen05.mma kind::f16 M128 N256 K16, fp32 accumulate in TMEM; executed tensor flops = %.2fx algorithmic "
                         "(3 split-fp16 products, D padded to a multiple of 16)" % tc_factor) if tc else "FP64 SIMT FMA",
                "executed_tensor_tflops": achieved_tf * tc_factor if tc else None,
                "executed_frac_of_peak": achieved_tf * tc_factor / peak if tc else None,
                "algorithmic_flops": "2*D*sum|cand(u)| over this rank's queries", "kernel_ms_all": kernel_ms,
                "kernel_ms_per_step": scan_ms / args.steps, "share_of_step": scan_ms / args.steps / ms_step,
                "second_pass_ms_per_step": round(sum(kernel_ms[k] for k in P2_KERNELS), 3)}
    e2e_ms, _, _ = rig.timed(step_e2e, args.steps, 2)
    h2d = U_pin.numel() * 4 + unk_pin.numel() + mean_pin.numel() * 8
    d2h = recs_host.numel() * 4
    e2e = {"value": n / (e2e_ms / 1e3), "unit": "recs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
           "note": "every rank uploads the replicated table (h2d is per rank); rank 0 reads all recs back"}
    P.close()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_rate(U, unk, mean, 1, args.cpu_seconds)
    return {"metric": "recs/sec (cosine LSH top-P recommendation)", "value": value, "unit": "recs/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(n, d, world), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "mean_candidates_per_user": ncand_local / max(1, hi - lo), "exactness_counters_per_step": counters}


def bench_c2_normal(rig, args):
    """C2-(ii) of SURVEY 8d: the same call on i.i.d. N(0,1) vectors -- 27% of the rows are candidates of a query, so the scan runs
    its masked (non-dense) epilogue and nothing ties: the second pass has nothing to do.  Resident timing only."""
    torch, capi, cdist, ctx, dev = rig.torch, rig.capi, rig.cdist, rig.ctx, rig.dev
    from crypto_recommendation_b200 import synth
    n, d = args.users, args.coins
    lo, hi = cdist.shard_range(n, rig.rank, rig.world)
    X = synth.normal_points(n, d, seed=2, dtype=np.float32)
    rng = np.random.default_rng(2)
    unk = (rng.random((n, d), dtype=np.float32) < 0.9).astype(np.uint8)
    known = unk == 0
    mean = (np.where(known, X, 0).sum(1, dtype=np.float64) / np.maximum(1, known.sum(1))).astype(np.float32).astype(np.float64)
    P = capi.Points(ctx, X, unk, mean)
    out = {"recs": torch.zeros((hi - lo, N_REC), dtype=torch.int32, device=dev), "ncand": torch.zeros(hi - lo, dtype=torch.int32, device=dev)}

    def step():
        t = capi.LshTables(ctx, P, "cosine", K_HASH, L_TABLES, LSH_BUCKET_DIV, EUCLID_W, SEED)
        capi.recommend_lsh(ctx, t, P_NEIGH, N_REC, q_begin=lo, q_end=hi, out=out)
        t.close()

    ms, launches, _ = rig.timed(step, args.steps, 2, profile=True)
    scan_ms, scan_n = ctx.kernel_time("tc_topp_scan")
    ncand = float(out["ncand"].to(torch.float64).sum().item())
    fl = 2.0 * d * ncand
    peak = rig.peaks["bf16_tflops_sustained"] or rig.peaks["bf16_tflops"]
    tf = fl / (scan_ms / max(1, scan_n) * 1e-3) / 1e12 if scan_ms > 0 else 0.0
    res = {"metric": "recs/sec (cosine LSH top-P recommendation)", "value": n / (ms / 1e3), "unit": "recs/s", "ms_per_step": ms, "gpu_launches": int(launches),
           "scaling": "strong" if rig.world > 1 else "weak", "mean_candidates_per_user": ncand / max(1, hi - lo),
           "kernel_ms": rig.kernel_ms(("tc_topp_scan", "rec_finalize", "hash_rows", "subset_hist") + P2_KERNELS, args.steps),
           "exactness_counters_per_step": {k: v / args.steps for k, v in ctx.counters().items()},
           "config": {"workload": "C2-(ii): %d x %d i.i.d. N(0,1) vectors, 90%% of the coins unknown, cosine LSH L=%d k=%d, top-P=%d, top-%d coins (E|cand| = 0.276 N)"
                                  % (n, d, L_TABLES, K_HASH, P_NEIGH, N_REC)},
           "roofline": {"kernel": "tc_scan_kernel<TOPP, masked epilogue>", "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                        "note": "algorithmic flops count candidate pairs only (2 D |cand|); the scan computes every pair and masks 72% of them away",
                        "scan_ms": scan_ms / max(1, scan_n), "traffic": None}}
    P.close()
    rig.release()
    return res


def bench_lloyd(rig, args, npts, tag):
    """C4: Lloyd assignment + k-means update, rows sharded (weak: `npts` rows per GPU), NCCL all-reduce issued by libcrx.so"""
    torch, capi, ctx, dev = rig.torch, rig.capi, rig.ctx, rig.dev
    world, rank = rig.world, rig.rank
    dd, kk = args.lloyd_d, args.lloyd_k
    X, centres = rig.gen_mixture(npts, dd, kk, 1235 + rank, centre_seed=1234)   # ONE mixture, every rank its own shard of points
    Q = capi.Points(ctx, X)
    del X
    torch.cuda.empty_cache()
    C = centres.to(torch.float64).contiguous()
    labels = torch.empty(npts, dtype=torch.int32, device=dev)
    dists = torch.empty(npts, dtype=torch.float64, device=dev)

    def step_assign():
        capi.lloyds_assignment(ctx, Q, C, None, "euclidean", labels, dists)

    def step_kmeans():
        capi.lloyds_assignment(ctx, Q, C, None, "euclidean", labels, dists)
        capi.k_means_sharded(ctx, Q, labels, C, "euclidean", 0.05, rig.comm)

    def step_assign_labels():   # dists = NULL: what the k-means loop needs (main.cpp:96-103 never reads the stored distance)
        capi.lloyds_assignment(ctx, Q, C, None, "euclidean", labels, want_dists=False)

    a_ms, _, _ = rig.timed(step_assign, args.steps, args.min_warmup)
    k_ms, _, _ = rig.timed(step_kmeans, args.steps, 3)   # the first all-reduces set up NCCL channels
    al_ms, _, _ = rig.timed(step_assign_labels, args.steps, 1)
    ctx.profile_reset(); ctx.profile(True)
    step_kmeans(); torch.cuda.synchronize(dev)
    ctx.profile(False)
    breakdown = rig.kernel_ms(("tc_argmin_scan", "lloyd_refine", "lloyd_scan", "lloyd_label", "tc_prep", "maxabs", "half_norm", "chunk_sums", "combine_sums",
                               "pack_counts", "kmeans_finish", "select_centroids", "bucket_offsets", "iota", "pad_centroids"), 1)
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
    peak = rig.peaks["bf16_tflops"]
    out = {"metric": "Lloyd assign pts*centroids/s", "value": world * npts * kk / (a_ms / 1e3), "unit": "pts*centroids/s",
           "ms_per_step": a_ms, "kmeans_iteration_ms": k_ms, "kmeans_kernel_ms": breakdown, "scaling": "weak",
           "labels_only": {"assign_ms": al_ms, "value": world * npts * kk / (al_ms / 1e3),
                           "note": "crx_lloyds_assignment with dists = NULL; NOT the headline (the reference stores the distance)"},
           "config": {"workload": "%s: %d x %d fp32 points per GPU (%d in total), K=%d, euclidean; labels + bit-exact FP64 distances; k-means iteration = "
                                  "assign + crx_k_means_sharded (one ncclAllReduce of K*D sums + K counts)" % (tag, npts, dd, world * npts, kk)},
           "roofline": {"kernel": "tc_scan_kernel<ARGMIN> (tcgen05 split-fp16 filter) + lloyd_refine_kernel (exact FP64 distance of the winner)" if ltc else "lloyd_scan_kernel (FP64 SIMT)",
                        "bound": "tensor", "achieved": fl / (lk_ms / max(1, lk_n) * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": fl / (lk_ms / max(1, lk_n) * 1e-3) / 1e12 / peak, "peak_source": rig.peaks["which"] + " bf16 dense GEMM (burst: the kernel runs for milliseconds)",
                        "traffic": dram_traffic("tc_argmin_scan" if ltc else "lloyd_scan", npts),
                        "executed_tensor_tflops": fl * lfac / (lk_ms / max(1, lk_n) * 1e-3) / 1e12 if ltc else None,
                        "pipe": ("tcgen05.mma kind::f16, 3 split-fp16 products (%.2fx algorithmic flops), then exact FP64 refine" % lfac) if ltc
                                else "FP64 SIMT (sub, mul, add per element: 3 FP64 ops for the 2 algorithmic flops)",
                        "scan_ms": lk_ms / max(1, lk_n), "refine_ms": lr_ms / max(1, lk_n),
                        "hbm_floor_ms": npts * (4 * dd + 12) / (rig.peaks["hbm_gbs"] * 1e6)}}
    if rig.comm is not None:
        out["collectives_issued"] = rig.comm.calls
    Q.close()
    del labels, dists
    rig.release()
    if rank == 0 and world == 1 and not args.no_cpu_baseline and tag.startswith("C4 shard"):
        out["cpu_baseline"] = cpu_lloyd_rate(dd, kk, min(8.0, args.cpu_seconds))
    return out


def bench_kmeanspp(rig, args):
    """C4 initialisation: k-means++ rounds over rows sharded across the ranks (weak: --lloyd-points rows per GPU)"""
    torch, capi, ctx = rig.torch, rig.capi, rig.ctx
    world, rank = rig.world, rig.rank
    npts, dd = args.lloyd_points, args.lloyd_d
    X, _ = rig.gen_mixture(npts, dd, args.lloyd_k, 2235 + rank, centre_seed=1234)
    Q = capi.Points(ctx, X)
    del X
    torch.cuda.empty_cache()
    K = args.kpp_k

    def run():
        capi.k_means_pp_sharded(ctx, Q, rank * npts, world * npts, K, "euclidean", 5, rig.comm)

    capi.k_means_pp_sharded(ctx, Q, rank * npts, world * npts, 9, "euclidean", 5, rig.comm)   # warm-up: buffers, NCCL channels
    ms, launches, _ = rig.timed(run, 1, 0, profile=True)
    names = ("kpp_update", "kpp_filter", "kpp_prune", "kpp_cdist", "kpp_prob", "kpp_total", "kpp_pick", "kpp_share")
    km = rig.kernel_ms(names, K - 1)
    per_round = ms / (K - 1)
    bytes_round = npts * (4 * dd + 16)
    gbs = bytes_round / (per_round * 1e6)
    out = {"metric": "k-means++ ms per round", "value": per_round, "unit": "ms/round", "higher_is_better": False, "scaling": "weak",
           "rounds": K - 1, "ms_total": ms, "points_per_s_per_round": world * npts / (per_round * 1e-3), "kernel_ms_per_round": km,
           "gpu_launches": int(launches),
           "config": {"workload": "C4 initialisation: k_means_pp with K=%d over %d x %d fp32 points per GPU (%d in total): all %d rounds (triangle-inequality "
                                  "pruning + fp32 filter + exact update), draws and prefix search on the device, %s"
                                  % (K, npts, dd, world * npts, K - 1, "3 NCCL collectives per round issued by libcrx.so" if world > 1 else "no host round trip per round")},
           "roofline": {"kernel": "kpp_prune + kpp_filter + kpp_update + kpp_prob + cub scan (one round, average over the run)", "bound": "hbm", "achieved": gbs,
                        "peak": rig.peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / rig.peaks["hbm_gbs"],
                        "algorithmic_bytes": "(4D + 16) B per point and round (SURVEY 8d); the pruning skips rows, so the figure is algorithmic, not measured traffic",
                        "traffic": None}}
    if rig.comm is not None:
        out["collectives_issued"] = rig.comm.calls
    Q.close()
    rig.release()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import load, EUCLIDEAN
        from crypto_recommendation_b200 import synth
        o = load("reference") or load("port")
        n, K = 10_000, 17
        Xc = synth.gaussian_mixture(n, dd, 64, seed=4).astype(np.float64)
        t0 = time.perf_counter(); o.k_means_pp(Xc, K, EUCLIDEAN, 5); dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": dt / (K - 1) * 1e3 * (npts / n), "unit": "ms/round (scaled linearly from N=%d to N=%d)" % (n, npts), "cores": 1, "kind": o.kind,
                               "sample": "k_means_pp N=%d K=%d D=%d on one core: %.1f ms/round measured" % (n, K, dd, dt / (K - 1) * 1e3)}
    return out


def bench_cube_range(rig, args):
    """C3: Euclidean hypercube (d'=16) build + range-search assignment with 64 probes, K=1024; N > 1: centroids split (strong)"""
    torch, capi, ctx, dev = rig.torch, rig.capi, rig.ctx, rig.dev
    world, rank = rig.world, rig.rank
    n, dd, K, probes, dprime, w = args.cube_points, 128, 1024, 64, 16, 4.0
    X, _ = rig.gen_mixture(n, dd, 1024, 3001)    # replicated: the same points on every rank
    P = capi.Points(ctx, X)
    del X
    torch.cuda.empty_cache()
    cube = [None]

    def build():
        if cube[0] is not None:
            cube[0].close()
        cube[0] = capi.Hypercube(ctx, P, "euclidean", dprime, w, 9)

    b_ms, _, _ = rig.timed(build, args.steps, 1, profile=True)
    hash_ms = ctx.kernel_time("hash_rows")[0] / args.steps
    cidx = capi.rand_selection(ctx, P, K, 3)
    outs = (torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev))

    def step():
        capi.cube_range_assignment(ctx, P, cube[0], cidx, "euclidean", probes, comm=rig.comm, out=outs)

    ms, launches, _ = rig.timed(step, args.steps, 2, profile=True)
    km = rig.kernel_ms(("range_fire", "range_hist", "range_finalize", "tc_argmin_scan", "lloyd_refine", "lloyd_scan", "tc_prep", "gather_rows", "compact", "min_pair"), args.steps)
    assigned = int((outs[2] >= 0).sum().item())
    remainder = n - assigned
    # algorithmic work: the probes gather |bucket| rows of 4D bytes per centroid; the Lloyd pass over the remainder is 2 D K flop per point
    fl = 2.0 * dd * remainder * K
    tms = km["tc_argmin_scan"] if km["tc_argmin_scan"] > 0 else km["lloyd_scan"]
    out = {"metric": "cube range-search assignment pts/s", "value": n / (ms * 1e-3), "unit": "pts/s", "ms_per_step": ms, "scaling": "strong" if world > 1 else "weak",
           "cube_build_ms": b_ms, "cube_build_hash_ms": hash_ms, "assigned_by_range_search": assigned, "kernel_ms": km, "gpu_launches": int(launches),
           "config": {"workload": "C3: %d x %d fp32 points (mixture of 1024 Gaussians), Euclidean hypercube d'=%d w=%.1f, K=%d random centroids, probes=%d, "
                                  "range search + lloyds_for_remaining; results left on the device%s"
                                  % (n, dd, dprime, w, K, probes, "; points replicated, centroids / remainder rows split over the ranks" if world > 1 else "")},
           "roofline": {"kernel": "tc_scan_kernel<ARGMIN> over the remainder (the probes themselves take %.2f ms)" % km["range_fire"], "bound": "tensor",
                        "achieved": fl / (tms * 1e-3) / 1e12 * (1.0 / world if world > 1 else 1.0) if tms > 0 else 0.0, "peak": rig.peaks["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": (fl / (tms * 1e-3) / 1e12 * (1.0 / world if world > 1 else 1.0) / rig.peaks["bf16_tflops"]) if tms > 0 else 0.0,
                        "hash_pass": {"bound": "hbm", "achieved": n * (4 * dd + 4 * dprime) / (hash_ms * 1e6) if hash_ms > 0 else None, "peak": rig.peaks["hbm_gbs"], "unit": "GB/s",
                                      "frac": n * (4 * dd + 4 * dprime) / (hash_ms * 1e6) / rig.peaks["hbm_gbs"] if hash_ms > 0 else None},
                        "traffic": None}}
    cube[0].close(); P.close()
    del outs
    rig.release()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import load, EUCLIDEAN
        from crypto_recommendation_b200 import synth
        o = load("reference") or load("port")
        nc, Kc = 50_000, 64
        Xc = synth.gaussian_mixture(nc, dd, 64, seed=6).astype(np.float64)
        ci = o.rand_selection(Xc, Kc, 7)
        t0 = time.perf_counter(); o.cube_range_assignment(Xc, ci, EUCLIDEAN, 10, 4.0, 32, 8); dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nc / dt, "unit": "pts/s", "cores": 1, "kind": o.kind,
                               "sample": "cube_range_assignment N=%d K=%d d'=10 probes=32 on one core (cost grows with K: the bench config has K=1024)" % (nc, Kc)}
    return out


def bench_pam(rig, args):
    """C5: Euclidean LSH tables, LSH range-search assignment and the PAM medoid update, K=256; N > 1: work split (strong)"""
    torch, capi, ctx, dev = rig.torch, rig.capi, rig.ctx, rig.dev
    world, rank = rig.world, rig.rank
    n, dd, K = args.pam_points, 100, 256
    X, _ = rig.gen_mixture(n, dd, 256, 4001)
    P = capi.Points(ctx, X)
    del X
    torch.cuda.empty_cache()
    tab = [None]

    def build():
        if tab[0] is not None:
            tab[0].close()
        tab[0] = capi.LshTables(ctx, P, "euclidean", K_HASH, L_TABLES, LSH_BUCKET_DIV, EUCLID_W, 11)

    b_ms, _, _ = rig.timed(build, args.steps, 1)
    cidx = capi.rand_selection(ctx, P, K, 6)
    outs = (torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev))

    def assign():
        capi.lsh_range_assignment(ctx, P, tab[0], cidx, "euclidean", comm=rig.comm, out=outs)

    a_ms, _, _ = rig.timed(assign, args.steps, 2)
    labels = outs[0]
    sizes = torch.bincount(labels.to(torch.int64), minlength=K).to(torch.float64)
    pairs = float((sizes * sizes).sum().item())

    def update():
        capi.pam_lloyds(ctx, P, labels, cidx, "euclidean", comm=rig.comm)

    ctx.counters(reset=True)
    u_ms, launches, _ = rig.timed(update, args.steps, 2, profile=True)
    km = rig.kernel_ms(("tc_rowsum_scan", "pam_rowsum", "tc_prep", "pam_bounds", "pam_exact", "pam_final", "bucket_offsets"), args.steps)
    scan = km["tc_rowsum_scan"] if km["tc_rowsum_scan"] > 0 else km["pam_rowsum"]
    fl = 2.0 * dd * pairs / world
    fac = 3.0 * (16 * ((dd + 15) // 16)) / dd
    out = {"metric": "PAM update pair distances/s", "value": pairs / (u_ms * 1e-3), "unit": "pair distances/s", "ms_per_step": u_ms, "scaling": "strong" if world > 1 else "weak",
           "lsh_build_ms": b_ms, "lsh_range_assignment_ms": a_ms, "pair_distances": pairs, "kernel_ms": km, "gpu_launches": int(launches),
           "exact_resums_per_step": ctx.counters()["pam_exact"] / args.steps,
           "config": {"workload": "C5: %d x %d fp32 points (mixture of 256 Gaussians), Euclidean LSH L=%d k=%d w=%.1f, K=%d: table build, lsh_range_assignment, pam_lloyds "
                                  "(medoid = member with the smallest sum of distances to its cluster)%s"
                                  % (n, dd, L_TABLES, K_HASH, EUCLID_W, K, "; points replicated, candidate rows split over the ranks, one all-reduce of the row sums" if world > 1 else "")},
           "roofline": {"kernel": "tc_scan_kernel<ROWSUM> (tcgen05 split-fp16 pair distances + error bars, 10-instruction epilogue per pair)", "bound": "tensor",
                        "achieved": fl / (scan * 1e-3) / 1e12 if scan > 0 else 0.0, "peak": rig.peaks["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": fl / (scan * 1e-3) / 1e12 / rig.peaks["bf16_tflops"] if scan > 0 else 0.0,
                        "executed_tensor_tflops": fl * fac / (scan * 1e-3) / 1e12 if scan > 0 else None,
                        "algorithmic_flops": "2 D per ordered pair of co-members (this rank's share)", "traffic": dram_traffic("tc_rowsum_scan", n)}}
    tab[0].close(); P.close()
    del outs
    rig.release()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import load, EUCLIDEAN
        from crypto_recommendation_b200 import synth
        o = load("reference") or load("port")
        nc, Kc = 4_000, 4
        Xc = synth.gaussian_mixture(nc, dd, Kc, seed=5).astype(np.float64)
        ci = o.rand_selection(Xc, Kc, 6)
        lab, _ = o.lloyds_assignment(Xc, Xc[ci], ci, EUCLIDEAN)
        t0 = time.perf_counter(); o.pam_lloyds(Xc, lab, ci, EUCLIDEAN); dt = time.perf_counter() - t0
        pc = float((np.bincount(lab, minlength=Kc).astype(np.float64) ** 2).sum())
        out["cpu_baseline"] = {"value": pc / dt, "unit": "pair distances/s", "cores": 1, "kind": o.kind, "sample": "pam_lloyds N=%d K=%d D=%d on one core" % (nc, Kc, dd)}
    return out


def bench_c1_main(rig, args):
    """BASELINE config 1 (the reference's own run): the reference's UNCHANGED main.cpp, built once over its own headers
    (oracle/_ref/recommendation_ref, CPU) and once over the drop-in headers + libcrx.so (oracle/_ref/recommendation_crx), on the
    same synthetic input files (tools/make_main_inputs.py, 6000 tweeting users -> ~4.4k kept).  Stage times are the ones
    main.cpp prints itself (main.cpp:181,227,271,375); the outputs must be identical."""
    import shutil
    import subprocess
    import tempfile
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "recommendation_ref")
    crx_bin = os.path.join(ROOT, "oracle", "_ref", "recommendation_crx")
    if not (os.path.exists(ref_bin) and os.path.exists(crx_bin)):
        return {"skipped": "oracle/_ref/recommendation_{ref,crx} not built (tools/build_main_dropin.sh needs the reference sources)"}
    work = tempfile.mkdtemp(prefix="crx_c1_")
    try:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_main_inputs.py"), work, "--users", str(args.c1_users)],
                       check=True, stdout=subprocess.DEVNULL)
        out = {"workload": "C1: main.cpp end to end, %d tweeting users x 100 coins, P=20, cosine LSH L=5 k=4 + two clustering stages" % args.c1_users,
               "stages": ["cosine LSH real users (main.cpp:146-186)", "cosine LSH cluster users (:196-232)",
                          "clustering recommendation A (:242-275)", "clustering recommendation B (:344-380)"]}
        texts = {}
        for tag, exe in (("reference_cpu", ref_bin), ("dropin_gpu", crx_bin), ("dropin_gpu_columnar_input", crx_bin)):
            if tag == "dropin_gpu_columnar_input":   # row f-4: the tweet-vector file converted once to the binary columnar form
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import csv_to_columnar
                src = os.path.join(work, "proj2_input.csv")
                csv_to_columnar.convert(src, src + ".crxcol")
            env = dict(os.environ, CRX_FAKE_SEED="5", CRX_DEVICE=str(rig.local_rank), CRX_SHIM_PROFILE="1")
            t0 = time.perf_counter()
            run = subprocess.run([exe, "-d", "./tweets.tsv", "-o", "./out_%s.txt" % tag], cwd=work, env=env, check=True,
                                 stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=900)
            wall = time.perf_counter() - t0
            lines = open(os.path.join(work, "out_%s.txt" % tag)).read().splitlines()
            stage_ms = [int(l.split(":")[1]) for l in lines if l.startswith("Execution Time:")]
            texts[tag] = [l for l in lines if not l.startswith("Execution Time:")]
            out[tag] = {"stage_ms": stage_ms, "stages_total_ms": sum(stage_ms), "process_wall_ms": round(wall * 1e3)}
            for l in run.stderr.decode().splitlines():   # the drop-in headers' own clock (CRX_SHIM_PROFILE)
                if "VectorReader::read" in l:
                    out[tag]["tweet_vectors_read_ms"] = float(l.split(")")[1].split("ms")[0])
                elif "CUDA start-up" in l:
                    out[tag]["cuda_startup_ms"] = float(l.split(")")[1].split("ms")[0])
        out["output_lines"] = len(texts["reference_cpu"])
        out["outputs_identical"] = texts["reference_cpu"] == texts["dropin_gpu"] == texts["dropin_gpu_columnar_input"]
        out["speedup_stages_total"] = round(out["reference_cpu"]["stages_total_ms"] / max(1, out["dropin_gpu"]["stages_total_ms"]), 2)
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


def run_crx(args):
    rig = Rig(args)
    torch = rig.torch
    only = set(args.only.split(",")) if args.only else None
    want = lambda name: only is None or name in only

    def guarded(name, fn):
        if not want(name):
            return None
        try:
            return fn()
        except Exception as e:   # a secondary block must not take the headline down with it
            rig.release()
            return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    line = bench_c2(rig, args) if want("c2") else {"metric": "recs/sec (cosine LSH top-P recommendation)", "value": None, "n_gpus": rig.world, "skipped": "--only"}
    rig.release()
    line["c2_normal"] = guarded("c2_normal", lambda: bench_c2_normal(rig, args))
    line["lloyd"] = None if args.no_lloyd else guarded("lloyd", lambda: bench_lloyd(rig, args, args.lloyd_points, "C4 shard (1/8 of the 100M config)"))
    line["kmeanspp"] = guarded("kmeanspp", lambda: bench_kmeanspp(rig, args))
    line["cube_range"] = guarded("cube_range", lambda: bench_cube_range(rig, args))
    line["pam"] = guarded("pam", lambda: bench_pam(rig, args))
    if rig.world == 1:
        line["c1_main"] = guarded("c1_main", lambda: bench_c1_main(rig, args))
    if rig.world == 1 and args.lloyd_full > 0:
        line["lloyd_100m"] = guarded("lloyd_100m", lambda: bench_lloyd(rig, args, args.lloyd_full, "C4 whole config on ONE GPU"))
    if rig.rank == 0:
        print(json.dumps(line))
    if rig.comm is not None:
        rig.comm.close()
    rig.ctx.close()
    if rig.world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_crx(a)
