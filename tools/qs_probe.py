"""How fast one warp runs the literal top-`need` sort (warp_qs_topn_big) on long lists: kernel time of crx_parallel_quickSort_topn
for a few key distributions.  python tools/qs_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from crypto_recommendation_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(1)
def run(name, s, need=20):
    ids = np.arange(len(s), dtype=np.int32)
    ctx.parallel_quickSort_topn(s.copy(), ids.copy(), need)
    ctx.profile_reset(); ctx.profile(True)
    ctx.parallel_quickSort_topn(s.copy(), ids.copy(), need)
    ctx.profile(False)
    ms, _ = ctx.kernel_time("warp_qs")
    print("%-58s n=%8d  %8.2f ms  (%.2f ns per entry)" % (name, len(s), ms, ms * 1e6 / len(s)))
for n in (100_000, 1_000_000):
    run("distinct random keys", rng.random(n))
    run("plateau of 5% at the maximum, last element in it", np.r_[np.where(rng.random(n - 1) < 0.05, 1.0, rng.random(n - 1) * 0.9), 1.0])
    run("three levels 1-ulp, 1, 1+ulp (80%) over random (20%)", np.where(rng.random(n) < 0.8, 1.0 + rng.integers(-1, 2, n) * 2.0 ** -52, rng.random(n)))
    run("all equal", np.ones(n))
# (no monotone runs here: a descending list makes every pivot a minimum of its range, which costs the literal algorithm --
#  and this closed form of it -- one walk of the range per ELEMENT, i.e. minutes at 10^6 entries)
