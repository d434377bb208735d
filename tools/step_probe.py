"""Where a C2 bench step goes: table build and the recommendation call timed separately (CUDA events, device outputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from crypto_recommendation_b200 import capi
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
n = 1_000_000
U, unk, mean = bench.make_users(n, 100, bench.SEED)
P = capi.Points(ctx, U, unk, mean)
out = {"recs": torch.zeros((n, 5), dtype=torch.int32, device=dev), "ncand": torch.zeros(n, dtype=torch.int32, device=dev)}
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tabs = []
def build():
    for t in tabs: t.close()
    tabs.clear()
    tabs.append(capi.LshTables(ctx, P, "cosine", 4, 5, 100, 0.4, bench.SEED))
print("tables            %.1f ms" % timed(build))
print("recommend only    %.1f ms" % timed(lambda: capi.recommend_lsh(ctx, tabs[0], 20, 5, out=out)))
def both():
    t = capi.LshTables(ctx, P, "cosine", 4, 5, 100, 0.4, bench.SEED)
    capi.recommend_lsh(ctx, t, 20, 5, out=out)
    t.close()
print("tables+recommend  %.1f ms" % timed(both))
import time
t0 = time.perf_counter(); both(); torch.cuda.synchronize(); print("one step, wall    %.1f ms" % ((time.perf_counter() - t0) * 1e3))
