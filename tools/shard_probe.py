"""One rank's share of a C2 step when the queries of the batch are split over `world` ranks (no collective): step time by
CUDA events against the sum of the kernel times, to see what does not shrink with the shard.
    python tools/shard_probe.py [world] [rank]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from crypto_recommendation_b200 import capi, dist as cdist
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
n = 1_000_000
U, unk, mean = bench.make_users(n, 100, bench.SEED)
P = capi.Points(ctx, U, unk, mean)
lo, hi = cdist.shard_range(n, rank, world)
out = {"recs": torch.zeros((hi - lo, 5), dtype=torch.int32, device=dev), "ncand": torch.zeros(hi - lo, dtype=torch.int32, device=dev)}
def step():
    t = capi.LshTables(ctx, P, "cosine", 4, 5, 100, 0.4, bench.SEED)
    capi.recommend_lsh(ctx, t, 20, 5, q_begin=lo, q_end=hi, out=out)
    t.close()
def tables_only():
    t = capi.LshTables(ctx, P, "cosine", 4, 5, 100, 0.4, bench.SEED)
    t.close()
for _ in range(3): step()
torch.cuda.synchronize()
reps = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps): step()
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
e0.record(stream)
for _ in range(reps): tables_only()
e1.record(stream); torch.cuda.synchronize()
ms_tab = e0.elapsed_time(e1) / reps
ctx.profile_reset(); ctx.profile(True)
for _ in range(reps): step()
torch.cuda.synchronize()
ctx.profile(False)
total, launches = ctx.kernel_time("")
names = ("tc_topp_scan", "rec_finalize", "hash_rows", "tc_prep", "pack_codes", "subset_hist", "subset_count", "bucket_offsets", "radix", "segments") + bench.P2_KERNELS
rows = [(k, ctx.kernel_time(k)[0] / reps) for k in names]
print("world %d rank %d: step %.1f ms (tables alone %.1f ms); kernels %.1f ms in %d launches per step" % (world, rank, ms, ms_tab, total / reps, launches // reps))
print("  " + ", ".join("%s %.1f" % (k, v) for k, v in rows if v >= 0.05))
