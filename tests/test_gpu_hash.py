"""GPU parity (through the C ABI): hash generators, LSH tables, hypercube vs the CPU oracle."""
import numpy as np
import pytest

from oracle import EUCLIDEAN, COSINE
from crypto_recommendation_b200 import capi, synth

pytestmark = pytest.mark.gpu


def test_kats_on_device(ctx, golden, port):
    s, d = ctx.parallel_quickSort(golden["kat_qs_in"], np.arange(8))
    assert d.tolist() == [6, 1, 3, 2, 5, 0, 7, 4] and np.array_equal(s, golden["kat_qs_sims"])
    s, d = ctx.parallel_quickSort(golden["kat_qs2_in"], np.arange(97))
    assert np.array_equal(d, golden["kat_qs2_ids"]) and np.array_equal(s, golden["kat_qs2_sims"])
    s, d = ctx.parallel_quickSort(np.ones(100), np.arange(100))
    assert d.tolist() == list(range(100))
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 17, 100, 128):
        v = rng.integers(0, 4, n) / 3.0
        s, d = ctx.parallel_quickSort(v, np.arange(n))
        ps, pd_ = port.quicksort(v, np.arange(n))
        assert np.array_equal(d, pd_) and np.array_equal(s, ps)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_vector_math(ctx, port, golden, dtype):
    a = golden["vm_a"].astype(dtype); b = golden["vm_b"].astype(dtype)
    pa, pb = ctx.points(a), ctx.points(b)
    idx = np.arange(6)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    eu = capi.pair_op(ctx, pa, idx, pb, idx, 1)
    assert np.array_equal(eu, [port.euclidean_distance(a64[i], b64[i]) for i in range(6)])  # bit-exact
    for op, f in ((0, port.inner_product), (2, port.cosine_distance), (3, port.cosine_similarity)):
        got = capi.pair_op(ctx, pa, idx, pb, idx, op)
        want = np.array([f(a64[i], b64[i]) for i in range(6)])
        assert np.array_equal(got, want), (op, got - want)   # the x87 accumulation and the two roundings of the quotient, bit for bit


CASES = [(COSINE, 4, 5, 100, 0.4), (EUCLIDEAN, 4, 5, 10, 4.0), (EUCLIDEAN, 4, 5, 100, 0.4), (COSINE, 7, 3, 1, 1.0),
         (EUCLIDEAN, 2, 1, 7, 2.5), (COSINE, 1, 1, 1, 1.0), (COSINE, 16, 2, 1, 1.0)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_lsh_tables_golden(ctx, golden, dtype):
    X = golden["hash_X"]
    if dtype == np.float32:
        pytest.skip("golden inputs are float64 (gaussian_mixture cast), float32 path covered by test_lsh_tables_oracle")
    P = ctx.points(X)
    t = capi.LshTables(ctx, P, "cosine", 4, 5, 100, 0.4, 7001)
    assert np.array_equal(t.bucket_ids(), golden["lsh_cos_ids"])
    for q in (0, 123, 499):
        assert np.array_equal(t.combined_buckets(q, 1), golden["lsh_cos_cand_%d" % q])
    t = capi.LshTables(ctx, P, "euclidean", 4, 5, 10, 4.0, 7002)
    assert np.array_equal(t.bucket_ids(), golden["lsh_euc_ids"])
    assert np.array_equal(t.detailed_hashes(), golden["lsh_euc_det"])
    for q in (0, 123, 499):
        assert np.array_equal(t.combined_buckets(q, 1), golden["lsh_euc_cand_f_%d" % q])
        assert np.array_equal(t.combined_buckets(q, 0), golden["lsh_euc_cand_u_%d" % q])
    c = capi.Hypercube(ctx, P, "cosine", 5, 0.4, 7003)
    assert np.array_equal(c.vertex_ids(), golden["cube_cos_ids"])
    for p in (1, 2, 7, 40):
        assert np.array_equal(c.combined_buckets(17, p), golden["cube_cos_cand_p%d" % p])
    c = capi.Hypercube(ctx, P, "euclidean", 5, 4.0, 7004)
    assert np.array_equal(c.vertex_ids(), golden["cube_euc_ids"])
    for p in (1, 2, 7, 40):
        assert np.array_equal(c.combined_buckets(17, p), golden["cube_euc_cand_p%d" % p])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", CASES)
def test_lsh_tables_oracle(ctx, port, dtype, case):
    metric, k, L, div, w = case
    X = synth.gaussian_mixture(3001, 37 if dtype == np.float64 else 100, 8, seed=21, dtype=dtype)
    X64 = X.astype(np.float64)
    P = ctx.points(X)
    t = capi.LshTables(ctx, P, metric, k, L, div, w, 555)
    ids, det = port.lsh_hash(X64, metric, k, L, div, w, 555)
    assert np.array_equal(t.bucket_ids(), ids)
    if metric == EUCLIDEAN:
        assert np.array_equal(t.detailed_hashes(), det)
        pr = t.params()
        assert pr["r"].min() >= 0 and pr["r"].max() <= 100
    for q in (0, 1500, 3000):
        for filt in (0, 1):
            assert np.array_equal(t.combined_buckets(q, filt), port.lsh_candidates(X64, metric, k, L, div, w, 555, q, filt))
    # a whole point set hashed like queries (crx_lsh_hash_points): the stored rows get their stored ids, other vectors
    # what the one-vector call gives
    hb, hd = t.hash_points(P)
    assert np.array_equal(hb, ids) and (metric != EUCLIDEAN or np.array_equal(hd, det))
    Y = X64[::-1][:40] * 1.25 + 0.5
    Q = ctx.points(Y)
    hb, hd = t.hash_points(Q)
    for i in (0, 7, 39):
        b1, d1 = t.hash_vector(Y[i])
        assert np.array_equal(hb[:, i], b1) and (metric != EUCLIDEAN or np.array_equal(hd[:, i], d1))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", [(COSINE, 6, 0.4), (EUCLIDEAN, 6, 0.4), (EUCLIDEAN, 10, 4.0), (COSINE, 16, 1.0), (EUCLIDEAN, 16, 8.0)])
def test_hypercube_oracle(ctx, port, dtype, case):
    metric, k, w = case
    X = synth.gaussian_mixture(4000, 128 if dtype == np.float32 else 33, 16, seed=22, dtype=dtype)
    X64 = X.astype(np.float64)
    P = ctx.points(X)
    c = capi.Hypercube(ctx, P, metric, k, w, 777)
    assert np.array_equal(c.vertex_ids(), port.cube_hash(X64, metric, k, w, 777))
    for probes in (1, 2, 5, 64, 5000):
        assert np.array_equal(c.combined_buckets(9, probes), port.cube_candidates(X64, metric, k, w, 777, 9, probes))


def test_zero_vector_and_boundaries(ctx, port):
    # all-zero rows hash to bit 1 under cosine (r.x = 0 >= 0) and to floor(t/w) under euclidean; rows that
    # are exact multiples of each other and tiny / huge magnitudes exercise the certified-boundary fallback
    rng = np.random.default_rng(5)
    X = rng.normal(size=(512, 24))
    X[0] = 0.0
    X[1] = 1e-300
    X[2] = X[3] * 1e-9
    X[4] = X[5] * 1e9
    P = ctx.points(X)
    for metric, k, L, div, w in [(COSINE, 4, 5, 100, 0.4), (EUCLIDEAN, 4, 3, 8, 0.4)]:
        t = capi.LshTables(ctx, P, metric, k, L, div, w, 99)
        ids, det = port.lsh_hash(X, metric, k, L, div, w, 99)
        assert np.array_equal(t.bucket_ids(), ids)
    assert ctx.counters()["hash_dd"] >= 1  # the zero row needs the exact path


def test_errors(ctx):
    X = np.zeros((10, 4))
    P = ctx.points(X)
    with pytest.raises(capi.CrxError):
        capi.LshTables(ctx, P, "euclidean", 4, 5, 100, 0.4, 1)  # N / div == 0 buckets (lsh_cube.hpp:60)
    with pytest.raises(capi.CrxError):
        capi.LshTables(ctx, P, "cosine", 40, 5, 100, 0.4, 1)
    with pytest.raises(capi.CrxError):
        ctx.points(np.zeros((4, 600)))   # rows wider than 512
