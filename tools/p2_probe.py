"""Per-kernel times of one C2 step (second pass included).  usage: python tools/p2_probe.py [n_users]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from crypto_recommendation_b200 import capi
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = capi.Context(0)
U, unk, mean = bench.make_users(n, 100, bench.SEED)
P = ctx.points(U, unk, mean)
t = capi.LshTables(ctx, P, "cosine", 4, 5, 100, 0.4, bench.SEED)
capi.recommend_lsh(ctx, t, 20, 5)
ctx.counters(reset=True)
ctx.profile(True); ctx.profile_reset()
import time
t0 = time.perf_counter()
capi.recommend_lsh(ctx, t, 20, 5)
wall = time.perf_counter() - t0
names = ["tc_topp_scan", "rec_finalize", "tc_prep", "subset_hist", "uniform_rows", "p2u_flag", "p2u_gather", "p2u_select", "p2u_sort", "p2u_compact",
         "p2_prepare", "tc_gather", "tc_collect_scan", "p2_all_count", "p2_sizes", "p2_trim", "p2_expand", "p2_fill_all",
         "p2_exact", "p2_resolve", "p2_collect_simt", "p2_leftover", ""]
print("wall %.1f ms" % (wall * 1e3))
for k in names:
    ms, cnt = ctx.kernel_time(k)
    print("%-18s %9.2f ms  %3d launches" % (k, ms, cnt))
print(ctx.counters())
