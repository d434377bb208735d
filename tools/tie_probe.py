"""Time rec_finalize on the C2 workload and print the tie counters (run once with CRX_TIE_ORDER=0, once without).
usage: python tools/tie_probe.py [n_users]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crypto_recommendation_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = capi.Context(0)
import numpy as np
X, u, m = synth.rating_users_fast(int(n * 1.08) + 64, 100, 2024)   # the bench.py users (make_users)
U, unk, mean = (np.ascontiguousarray(a[:n]) for a in (X, u, m))
P = ctx.points(U, unk, mean)
t = capi.LshTables(ctx, P, "cosine", 4, 5, 100, 0.4, 7)
capi.recommend_lsh(ctx, t, 20, 5)
ctx.counters(reset=True)
ctx.profile(True); ctx.profile_reset()
capi.recommend_lsh(ctx, t, 20, 5)
ms, k = ctx.kernel_time("rec_finalize")
ms2, _ = ctx.kernel_time("tc_topp_scan")
print("CRX_TIE_ORDER=%s n=%d rec_finalize %.2f ms, tc_topp_scan %.2f ms, counters %s" % (os.environ.get("CRX_TIE_ORDER", "1"), n, ms, ms2, ctx.counters()))
