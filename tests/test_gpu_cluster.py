"""GPU parity (through the C ABI): clustering phases vs the CPU oracle."""
import os

import numpy as np
import pytest

from oracle import EUCLIDEAN, COSINE
from crypto_recommendation_b200 import capi, synth
from helpers import assert_float_close, assert_labels

pytestmark = pytest.mark.gpu
METRICS = [EUCLIDEAN, COSINE]
# the A/B switches of DESIGN.md section 5 change which kernels run, never the results: the suite passes under them too
NO_TC = os.environ.get("CRX_NO_TC") == "1"
NO_PRUNE = os.environ.get("CRX_KPP_NOPRUNE") == "1"


def dist_fn(port, X64, C, metric):
    f = port.euclidean_distance if metric == EUCLIDEAN else port.cosine_distance
    return lambda v, c: f(X64[v], C[c])


@pytest.mark.parametrize("metric", METRICS)
def test_golden_clustering(ctx, port, golden, metric):
    g, name = golden, ("euc" if metric == EUCLIDEAN else "cos")
    X = g["cl_X"]
    P = ctx.points(X)
    assert np.array_equal(capi.rand_selection(ctx, P, 9, 8001), g["cl_rand_sel"])
    cidx = capi.k_means_pp(ctx, P, 7, metric, 8002)
    assert np.array_equal(cidx, g["cl_kpp_%s" % name])
    lab, dist = capi.lloyds_assignment(ctx, P, X[cidx], cidx, metric)
    assert_labels(lab, g["cl_lloyd_lab_%s" % name], dist_fn(port, X, X[cidx], metric))
    assert np.array_equal(dist, g["cl_lloyd_dist_%s" % name])  # bit-exact, both metrics (cosine: x87 emulation)
    ret, newc = capi.k_means(ctx, P, g["cl_lloyd_lab_%s" % name], X[cidx], metric, 0.05)
    assert ret == bool(g["cl_kmeans_ret_%s" % name])
    assert_float_close(newc, g["cl_kmeans_C_%s" % name], 1e-12)
    lab2, dist2 = capi.lloyds_assignment(ctx, P, g["cl_kmeans_C_%s" % name], None, metric)
    assert_labels(lab2, g["cl_lloyd2_lab_%s" % name], dist_fn(port, X, g["cl_kmeans_C_%s" % name], metric))
    assert np.array_equal(dist2, g["cl_lloyd2_dist_%s" % name])
    t = capi.LshTables(ctx, P, metric, 4, 5, 10, 4.0, 8003)
    l, d, b = capi.lsh_range_assignment(ctx, P, t, cidx, metric)
    assert np.array_equal(b, g["cl_lshrange_before_%s" % name])
    assert np.array_equal(l, g["cl_lshrange_lab_%s" % name]) and np.array_equal(d, g["cl_lshrange_dist_%s" % name])
    cube = capi.Hypercube(ctx, P, metric, 5, 4.0, 8004)
    l, d, b = capi.cube_range_assignment(ctx, P, cube, cidx, metric, 6)
    assert np.array_equal(b, g["cl_cuberange_before_%s" % name])
    assert np.array_equal(l, g["cl_cuberange_lab_%s" % name]) and np.array_equal(d, g["cl_cuberange_dist_%s" % name])
    sw, new = capi.pam_lloyds(ctx, P, g["cl_lloyd_lab_%s" % name], cidx, metric)
    assert sw == bool(g["cl_pam_sw_%s" % name]) and np.array_equal(new, g["cl_pam_new_%s" % name])
    assert_float_close(capi.silhouette_cluster(ctx, P, g["cl_lloyd_lab_%s" % name], X[cidx], metric), g["cl_sil_%s" % name], 1e-10)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cosine_lloyd_near_ties_follow_the_reference(ctx, port, dtype):
    """Centroids that are scalar multiples of each other (and rating-like users, whose vectors repeat directions) have
    cosine distances that differ only through the rounding of the reference's x87 accumulation: the plain FP64 scan cannot
    rank them, so such rows are decided again with the reference's arithmetic (lowest index wins exact ties)."""
    U, _, _ = synth.rating_users_fast(5000, 100, seed=77, min_known=1, max_known=3, dtype=dtype)
    X64 = U.astype(np.float64)
    n = U.shape[0]
    rng = np.random.default_rng(5)
    base = X64[rng.choice(n, 12, replace=False)]
    scales = np.array([1.0, 3.0, 0.37, 2.0, 1.0 / 3.0])
    C = np.concatenate([base * s for s in scales])          # 60 centroids, 12 directions
    C = C[rng.permutation(len(C))]
    P = ctx.points(U)
    lab, dist = capi.lloyds_assignment(ctx, P, C, None, COSINE)
    rl, rd = port.lloyds_assignment(X64, C, None, COSINE)
    assert np.array_equal(lab, rl), "labels differ for %d rows" % int((lab != rl).sum())
    assert np.array_equal(dist, rd)


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype,n,d,k", [(np.float32, 20000, 128, 64), (np.float64, 5000, 100, 30), (np.float32, 4097, 7, 3),
                                         (np.float32, 1000, 128, 1), (np.float64, 777, 1, 5), (np.float32, 3000, 100, 129)])
def test_lloyd_kmeans_oracle(ctx, port, metric, dtype, n, d, k):
    X = synth.gaussian_mixture(n, d, max(2, k // 2), seed=31, dtype=dtype)
    if d == 1 and metric == COSINE:
        X = np.abs(X) + 0.5  # 1-d cosine distances are all 0 or 2: keep them tie-free per sign
    X64 = X.astype(np.float64)
    P = ctx.points(X)
    cidx = capi.rand_selection(ctx, P, k, 4000 + k)
    assert np.array_equal(cidx, port.rand_selection(X64, k, 4000 + k))
    C = X64[cidx]
    lab, dist = capi.lloyds_assignment(ctx, P, C, cidx, metric)
    rl, rd = port.lloyds_assignment(X64, C, cidx, metric)
    near = assert_labels(lab, rl, dist_fn(port, X64, C, metric), max_near=0)
    assert np.array_equal(lab, rl) and np.array_equal(dist, rd)   # bit-exact for both metrics
    ret, newc = capi.k_means(ctx, P, rl, C, metric, 0.05)
    pret, pC = port.k_means(X64, rl, C, metric, 0.05)
    assert ret == pret
    assert_float_close(newc, pC, 1e-12)
    # second iteration from k-means centres (not data points)
    lab2, dist2 = capi.lloyds_assignment(ctx, P, pC, None, metric)
    rl2, rd2 = port.lloyds_assignment(X64, pC, None, metric)
    assert_labels(lab2, rl2, dist_fn(port, X64, pC, metric), max_near=0 if metric == EUCLIDEAN else 3)
    # convergence: a huge min_dist stops, centres unchanged
    ret, same_c = capi.k_means(ctx, P, rl, C, metric, 1e9)
    assert ret is False and np.array_equal(same_c, C)


def test_lloyd_ties_and_duplicates(ctx, port):
    # duplicate centroids and points equidistant from two centroids: lowest index must win (assignment.hpp:67)
    X = np.zeros((600, 8))
    X[:, 0] = np.arange(600) % 3 - 1.0  # -1, 0, 1
    C = np.zeros((4, 8)); C[0, 0] = -1; C[1, 0] = 1; C[2, 0] = 1; C[3, 0] = -1
    P = ctx.points(X)
    lab, dist = capi.lloyds_assignment(ctx, P, C, None, EUCLIDEAN)
    rl, rd = port.lloyds_assignment(X, C, None, EUCLIDEAN)
    assert np.array_equal(lab, rl) and np.array_equal(dist, rd)
    assert set(lab.tolist()) == {0, 1}
    # empty clusters become the zero vector (cust_vector.hpp:189)
    ret, newc = capi.k_means(ctx, P, rl, C, EUCLIDEAN, 0.0)
    pret, pC = port.k_means(X, rl, C, EUCLIDEAN, 0.0)
    assert ret == pret and np.array_equal(newc, pC)


def test_centroid_self_assignment(ctx, port):
    X = synth.gaussian_mixture(500, 16, 4, seed=8).astype(np.float64)
    cidx = np.array([5, 5, 17, 300], np.int32)  # the same row twice: the later centroid wins the overwrite
    P = ctx.points(X)
    lab, dist = capi.lloyds_assignment(ctx, P, X[cidx], cidx, EUCLIDEAN)
    rl, rd = port.lloyds_assignment(X, X[cidx], cidx, EUCLIDEAN)
    assert np.array_equal(lab, rl) and np.array_equal(dist, rd)
    assert lab[5] == 1 and dist[5] == 0


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype,n,d,k", [(np.float32, 12000, 128, 24), (np.float64, 3000, 20, 9)])
def test_kmeanspp_oracle(ctx, port, metric, dtype, n, d, k):
    X = synth.gaussian_mixture(n, d, 8, seed=41, dtype=dtype)
    P = ctx.points(X)
    ctx.counters(reset=True)
    got = capi.k_means_pp(ctx, P, k, metric, 909)
    want = port.k_means_pp(X.astype(np.float64), k, metric, 909)
    near = ctx.counters()["kpp_near"]
    if near == 0:
        assert np.array_equal(got, want)
    else:  # a draw landed within rounding of a prefix boundary: the two sequences may part from there
        first = int(np.flatnonzero(got != want)[0]) if not np.array_equal(got, want) else k
        assert first >= 1


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_range_assignment_oracle(ctx, port, metric, dtype):
    X = synth.gaussian_mixture(6000, 32, 10, seed=51, dtype=dtype)
    X64 = X.astype(np.float64)
    P = ctx.points(X)
    cidx = port.k_means_pp(X64, 12, metric, 111)
    fn = dist_fn(port, X64, X64[cidx], metric)
    for (k, L, div, w) in [(4, 5, 100, 0.4), (3, 4, 50, 6.0)]:
        t = capi.LshTables(ctx, P, metric, k, L, div, w, 112)
        l, d, b = capi.lsh_range_assignment(ctx, P, t, cidx, metric)
        rl, rd, rb = port.lsh_range_assignment(X64, cidx, metric, k, L, div, w, 112)
        assert np.array_equal(b, rb), (np.sum(b != rb), (rb >= 0).sum())
        assert np.array_equal(l, rl) and np.array_equal(d, rd)   # labels and distances bit-exact, both metrics
    for (k, w, probes) in [(8, 6.0, 20), (6, 0.4, 1), (10, 3.0, 64)]:
        cube = capi.Hypercube(ctx, P, metric, k, w, 113)
        l, d, b = capi.cube_range_assignment(ctx, P, cube, cidx, metric, probes)
        rl, rd, rb = port.cube_range_assignment(X64, cidx, metric, k, w, probes, 113)
        assert np.array_equal(b, rb), (np.sum(b != rb), (rb >= 0).sum())
        assert np.array_equal(l, rl) and np.array_equal(d, rd)   # labels and distances bit-exact, both metrics


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_range_assignment_with_heap_centroids(ctx, port, metric, dtype):
    """lsh_range_assignment with centroids that are not stored rows (the centres k_means leaves behind), with unique ids and
    with the shared "k_means_center" id whose distance-cache collisions the reference has from the second iteration on
    (SURVEY App. A-2): labels before and after the Lloyd pass and distances bit-exact."""
    X = synth.gaussian_mixture(6000, 32, 10, seed=52, dtype=dtype)
    X64 = X.astype(np.float64)
    P = ctx.points(X)
    rng = np.random.default_rng(9)
    cidx = port.k_means_pp(X64, 12, metric, 111)
    lab, _ = port.lloyds_assignment(X64, X64[cidx], cidx, metric)
    _, C = port.k_means(X64, lab, X64[cidx], metric, 0.0)          # real k-means centres
    C2 = np.stack([X64[rng.choice(len(X64), 30)].mean(0) for _ in range(12)])   # and centres of random subsets
    assigned = 0
    differs = 0
    for (k, L, div, w) in [(4, 5, 100, 0.4), (3, 4, 50, 6.0)]:
        t = capi.LshTables(ctx, P, metric, k, L, div, w, 112)
        for cen in (C, C2):
            per_mode = []
            for shared in (False, True):
                l, d, b = capi.lsh_range_assignment_vectors(ctx, P, t, cen, metric, shared_ids=shared)
                rl, rd, rb = port.lsh_range_assignment_vectors(X64, cen, None, shared, metric, k, L, div, w, 112)
                assert np.array_equal(b, rb), (shared, int(np.sum(b != rb)), int((rb >= 0).sum()))
                assert np.array_equal(l, rl) and np.array_equal(d, rd)
                assigned += int((rb >= 0).sum())
                per_mode.append(rb)
            differs += int(np.sum(per_mode[0] != per_mode[1]))
        # some centroids are stored rows (their own unique ids, assigned to their own cluster at the end)
        rows = np.full(12, -1, np.int32)
        rows[[2, 7]] = cidx[[2, 7]]
        cen = C.copy()
        cen[[2, 7]] = X64[cidx[[2, 7]]]
        l, d, b = capi.lsh_range_assignment_vectors(ctx, P, t, cen, metric, shared_ids=False, centroid_rows=rows)
        rl, rd, rb = port.lsh_range_assignment_vectors(X64, cen, rows, 0, metric, k, L, div, w, 112)
        assert np.array_equal(b, rb) and np.array_equal(l, rl) and np.array_equal(d, rd)
        t.close()
    assert assigned > 0 and differs > 0   # the range search assigned rows, and the id collision changed some of them


@pytest.mark.parametrize("metric", METRICS)
def test_iterated_range_assignment_and_k_means(ctx, port, metric):
    """The loop the reference's library allows (and round 1 aborted in): {lsh_range_assignment, k_means} repeated; from the
    second iteration the centroids are heap vectors that all carry the id "k_means_center"."""
    X = synth.gaussian_mixture(5000, 24, 8, seed=53, dtype=np.float64)
    P = ctx.points(X)
    k, L, div, w = 4, 4, 50, 2.0
    t = capi.LshTables(ctx, P, metric, k, L, div, w, 77)
    cidx = port.k_means_pp(X, 9, metric, 78)
    lab, dist, _ = capi.lsh_range_assignment(ctx, P, t, cidx, metric)
    rlab, rdist, _ = port.lsh_range_assignment(X, cidx, metric, k, L, div, w, 77)
    assert np.array_equal(lab, rlab) and np.array_equal(dist, rdist)
    C = X[cidx]
    for it in range(3):
        cont, C_new = capi.k_means(ctx, P, lab, C, metric, 0.0)
        rcont, rC = port.k_means(X, rlab, C, metric, 0.0)
        assert cont == rcont
        if not rcont:
            break
        assert_float_close(C_new, rC, 1e-12)
        C = rC   # both sides continue from the oracle's centres (the engine's sums may differ in the last bits)
        lab, dist, bef = capi.lsh_range_assignment_vectors(ctx, P, t, C, metric, shared_ids=True)
        rlab, rdist, rbef = port.lsh_range_assignment_vectors(X, C, None, 1, metric, k, L, div, w, 77)
        assert np.array_equal(bef, rbef), (it, int(np.sum(bef != rbef)))
        # (an emptied cluster leaves a zero centre: cosine distances to it are NaN on both sides)
        assert np.array_equal(lab, rlab) and np.array_equal(dist, rdist, equal_nan=True), it
        if it == 0:
            assert len(set(rlab.tolist())) > 1
    t.close()


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pam_silhouette_oracle(ctx, port, metric, dtype):
    X = synth.gaussian_mixture(2500, 24, 6, seed=61, dtype=dtype)
    X64 = X.astype(np.float64)
    P = ctx.points(X)
    cidx = port.rand_selection(X64, 7, 5)
    rl, _ = port.lloyds_assignment(X64, X64[cidx], cidx, metric)
    sw, new = capi.pam_lloyds(ctx, P, rl, cidx, metric)
    psw, pnew = port.pam_lloyds(X64, rl, cidx, metric)
    assert sw == psw and np.array_equal(new, pnew)
    # a second PAM step from the new medoids is a fixed point or keeps matching
    rl2, _ = port.lloyds_assignment(X64, X64[pnew], pnew, metric)
    sw2, new2 = capi.pam_lloyds(ctx, P, rl2, pnew, metric)
    psw2, pnew2 = port.pam_lloyds(X64, rl2, pnew, metric)
    assert sw2 == psw2 and np.array_equal(new2, pnew2)
    s = capi.silhouette_cluster(ctx, P, rl, X64[cidx], metric)
    assert_float_close(s, port.silhouette(X64, rl, X64[cidx], metric), 1e-10)


def test_pam_exact_duplicates(ctx, port):
    # identical rows inside a cluster give identical row sums: the first member must win (update.hpp:125)
    rng = np.random.default_rng(2)
    base = rng.normal(size=(40, 6))
    X = np.concatenate([base, base, base + 10.0])
    lab = np.concatenate([np.zeros(80, np.int32), np.ones(40, np.int32)])
    cidx = np.array([79, 100], np.int32)
    P = ctx.points(X)
    sw, new = capi.pam_lloyds(ctx, P, lab, cidx, EUCLIDEAN)
    psw, pnew = port.pam_lloyds(X, lab, cidx, EUCLIDEAN)
    assert sw == psw and np.array_equal(new, pnew)
    assert ctx.counters()["pam_exact"] >= 2


@pytest.mark.parametrize("dtype,d", [(np.float32, 24), (np.float64, 100)])
def test_pam_tensor_path(ctx, port, dtype, d):
    # N >= 4096 Euclidean goes through the tcgen05 row-sum scan + exact re-sum of the rows it cannot separate;
    # ragged clusters: an empty one, a singleton, sizes that are not multiples of the 128-row tile
    n, K = 9000, 6
    X = synth.gaussian_mixture(n, d, 4, seed=67, dtype=dtype)
    X64 = X.astype(np.float64)
    cidx = port.rand_selection(X64, 4, 9)
    lab4, _ = port.lloyds_assignment(X64, X64[cidx], cidx, EUCLIDEAN)
    lab = lab4.copy()
    lab[17] = 5                                   # singleton cluster 5, cluster 4 stays empty
    cidx = np.concatenate([cidx, [3, 17]]).astype(np.int32)
    P = ctx.points(X)
    ctx.profile(True)
    ctx.profile_reset()
    sw, new = capi.pam_lloyds(ctx, P, lab, cidx, EUCLIDEAN)
    _, launches = ctx.kernel_time("tc_rowsum_scan")
    ctx.profile(False)
    psw, pnew = port.pam_lloyds(X64, lab, cidx, EUCLIDEAN)
    assert sw == psw and np.array_equal(new, pnew)
    assert launches == (0 if NO_TC else 1), "the tensor path ran"


def test_pam_tensor_path_duplicates_first_wins(ctx, port):
    # every row appears twice in its cluster: the two copies have identical sums, the earlier member must win, and
    # the tensor bounds cannot separate them -> exact sequential re-sum
    rng = np.random.default_rng(12)
    base = rng.normal(size=(2600, 16)).astype(np.float32)
    base[1300:] += 6.0
    X = np.concatenate([base, base])
    lab = np.concatenate([np.zeros(1300, np.int32), np.ones(1300, np.int32)] * 2)
    cidx = np.array([5200 - 1, 1300], np.int32)
    P = ctx.points(X)
    ctx.counters(reset=True)
    sw, new = capi.pam_lloyds(ctx, P, lab, cidx, EUCLIDEAN)
    psw, pnew = port.pam_lloyds(X.astype(np.float64), lab, cidx, EUCLIDEAN)
    assert sw == psw and np.array_equal(new, pnew)
    assert (new < 2600).all()
    assert ctx.counters()["pam_exact"] >= 4


def test_cluster_sums_chunking(ctx, port):
    # one cluster larger than the 16384-row chunk with too few clusters to fill the GPU otherwise: it is cut into chunks
    # (deterministic, equal to the sequential sum to ~1e-15); clusters of up to 16384 rows are one sequential sum
    n = 40_000
    X = synth.gaussian_mixture(n, 24, 3, seed=71, dtype=np.float32)
    lab = (np.arange(n) % 10 == 0).astype(np.int32)  # cluster 0: 36000 rows, cluster 1: 4000 rows
    P = ctx.points(X)
    s1, c1 = capi.cluster_sums(ctx, P, lab, 3)
    s2, c2 = capi.cluster_sums(ctx, P, lab, 3)
    assert np.array_equal(s1, s2) and c1.tolist() == [36000, 4000, 0]
    X64 = X.astype(np.float64)
    seq = np.zeros((3, 24))
    for v in range(n):
        seq[lab[v]] += X64[v]
    assert np.array_equal(s1[1], seq[1])          # fits one chunk: the reference's own sequential sum
    assert_float_close(s1[0], seq[0], 1e-13)
    assert not s1[2].any()


@pytest.mark.parametrize("K,metric", [(64, EUCLIDEAN), (7, EUCLIDEAN), (40, COSINE)])
def test_lloyd_labels_only_equals_full_call(ctx, port, K, metric):
    # dists == NULL: same labels (tensor filter + exact re-scan of ambiguous rows at K >= 32, exact scan otherwise),
    # including duplicated centroids (exact ties -> lowest index) and the centroid self-assignment
    X = synth.gaussian_mixture(9000, 48, 10, seed=91, dtype=np.float32)
    X64 = X.astype(np.float64)
    cidx = port.rand_selection(X64, K, 4)
    C = X64[cidx].copy()
    C[K - 1] = C[0]
    P = ctx.points(X)
    lab, dist = capi.lloyds_assignment(ctx, P, C, cidx, metric)
    lab2, none = capi.lloyds_assignment(ctx, P, C, cidx, metric, want_dists=False)
    assert none is None and np.array_equal(lab, lab2)
    rl, _ = port.lloyds_assignment(X64, C, cidx, metric)
    assert np.array_equal(lab2, rl)


@pytest.mark.parametrize("metric", METRICS)
def test_kmeanspp_filter_with_duplicates_and_zero_rows(ctx, port, metric):
    # fp32 data, N >= 4096: rounds after the first go through the fp32 filter + exact update of the listed rows.
    # Duplicated rows (distance exactly 0 to a chosen centroid), tightly clustered rows (filter margin) and, for
    # cosine, zero vectors (NaN distances) must not change the chosen rows.
    rng = np.random.default_rng(77)
    base = synth.gaussian_mixture(3000, 40, 6, seed=43, dtype=np.float32)
    X = np.concatenate([base, base[:1500], base[:700] * np.float32(1.0000001), (base[:300] * 1e-3).astype(np.float32)])
    if metric == COSINE:
        X[[17, 4000]] = 0
    X = X[rng.permutation(len(X))]
    P = ctx.points(X)
    ctx.profile(True); ctx.profile_reset(); ctx.counters(reset=True)
    got = capi.k_means_pp(ctx, P, 14, metric, 31)
    _, nfilter = ctx.kernel_time("kpp_filter")
    _, nprune = ctx.kernel_time("kpp_prune")
    ctx.profile(False)
    want = port.k_means_pp(X.astype(np.float64), 14, metric, 31)
    assert nfilter == 12, "rounds 2..13 use the fp32 filter"
    assert nprune == (12 if metric == EUCLIDEAN and not NO_PRUNE else 0), "Euclidean rounds are pruned by the triangle inequality first"
    if ctx.counters()["kpp_near"] == 0:
        assert np.array_equal(got, want)


def test_kmeans_sums_are_sequential_when_clusters_are_many(ctx, port):
    # 300 clusters of ~2000 members (> the 1024-row base chunk): with at least two chunks per SM left, every cluster is
    # summed as ONE sequential chain in input order, so the new centres equal the reference's bit for bit
    rng = np.random.default_rng(5)
    n, d, K = 600_000, 16, 300
    X = rng.normal(size=(n, d)).astype(np.float32)
    lab = rng.integers(0, K, n).astype(np.int32)
    C = rng.normal(size=(K, d))
    P = ctx.points(X)
    cont, newc = capi.k_means(ctx, P, lab, C, EUCLIDEAN, 0.05)
    pcont, pC = port.k_means(X.astype(np.float64), lab, C, EUCLIDEAN, 0.05)
    assert cont == pcont and np.array_equal(newc, pC)


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_wide_rows_clustering_loop(ctx, port, metric, dtype):
    # D = 203 (> 128): the tweet vectors main.cpp clusters in front of the recommendation (main.cpp:78-110).  The
    # clustering core takes rows up to 512 wide: k-means++, Lloyd assignment (+ for remaining), k-means, pair_op --
    # all bit-exact against the oracle; the other entry points refuse such rows with a clear error.
    n, d, K = 1500, 203, 11
    rng = np.random.default_rng(203)
    X = (rng.gamma(2.0, 1.0, size=(n, d)) * (rng.random((n, d)) < 0.3)).astype(dtype) + dtype(0.01)
    X64 = X.astype(np.float64)
    P = ctx.points(X)
    ctx.counters(reset=True)
    cidx = capi.k_means_pp(ctx, P, K, metric, 77)
    want = port.k_means_pp(X64, K, metric, 77)
    if ctx.counters()["kpp_near"] == 0:
        assert np.array_equal(cidx, want)
    cidx = want
    C = X64[cidx]; rows = cidx
    for it in range(3):
        lab, dist = capi.lloyds_assignment(ctx, P, C, rows, metric)
        rl, rd = port.lloyds_assignment(X64, C, rows, metric)
        assert np.array_equal(lab, rl) and np.array_equal(dist, rd)
        cont, C2 = capi.k_means(ctx, P, lab, C, metric, 0.05)
        pcont, pC = port.k_means(X64, rl, C, metric, 0.05)
        assert cont == pcont and np.array_equal(C2, pC)
        C = pC; rows = None
    lab[::7] = -1
    rl2 = lab.copy(); rd2 = dist.copy()
    l3, d3 = capi.lloyds_for_remaining(ctx, P, C, metric, lab.copy(), dist.copy())
    full_l, full_d = port.lloyds_assignment(X64, C, None, metric)
    assert np.array_equal(l3[::7], full_l[::7]) and np.array_equal(d3[::7], full_d[::7])
    idx = np.arange(0, 40)
    for op, f in ((0, port.inner_product), (1, port.euclidean_distance), (2, port.cosine_distance), (3, port.cosine_similarity)):
        got = capi.pair_op(ctx, P, idx, P, idx[::-1].copy(), op)
        assert np.array_equal(got, np.array([f(X64[i], X64[j]) for i, j in zip(idx, idx[::-1])]))
    for bad in (lambda: capi.LshTables(ctx, P, "cosine", 4, 5, 100, 0.4, 1), lambda: capi.Hypercube(ctx, P, "cosine", 4, 1.0, 1),
                lambda: capi.pam_lloyds(ctx, P, lab.clip(0), cidx, metric), lambda: capi.silhouette_cluster(ctx, P, lab.clip(0), C, metric)):
        with pytest.raises(capi.CrxError):
            bad()


def test_full_clustering_loop_matches_oracle(ctx, port):
    # the loop of main.cpp:96-103 / 246-254: init -> (assign, update) x iters
    X = synth.gaussian_mixture(8000, 100, 12, seed=81, dtype=np.float32)
    X64 = X.astype(np.float64)
    P = ctx.points(X)
    for metric in METRICS:
        cidx = capi.k_means_pp(ctx, P, 10, metric, 1234)
        assert np.array_equal(cidx, port.k_means_pp(X64, 10, metric, 1234))
        C = X64[cidx]; Cr = C.copy(); rows = cidx
        for it in range(4):
            lab, _ = capi.lloyds_assignment(ctx, P, C, rows, metric)
            rl, _ = port.lloyds_assignment(X64, Cr, rows, metric)
            assert_labels(lab, rl, dist_fn(port, X64, Cr, metric), max_near=2)
            cont, C = capi.k_means(ctx, P, lab, C, metric, 0.05)
            rcont, Cr = port.k_means(X64, rl, Cr, metric, 0.05)
            assert cont == rcont
            assert_float_close(C, Cr, 1e-9)
            rows = None
            if not cont:
                break


def test_lloyd_tensor_path_ties_and_margins(ctx, port):
    # K >= 32 and N >= 1024 route the Euclidean assignment through the tcgen05 filter + exact refine.
    # Duplicate centroids, points equidistant from two centroids and points within 1e-7 of a bisector must
    # still come out bit-identical (the filter flags them as ambiguous and the exact FP64 scan decides).
    rng = np.random.default_rng(9)
    K, D, N = 48, 20, 6000
    C = rng.normal(size=(K, D)) * 3
    C[7] = C[3]                      # exact duplicate: the lower index must win
    C[40] = C[41] + 1e-9             # near duplicate
    X = C[rng.integers(0, K, N)] + rng.normal(size=(N, D)) * 0.7
    X[:200] = 0.5 * (C[1] + C[2])    # exactly on a bisector
    X[200:400] = 0.5 * (C[5] + C[6]) + rng.normal(size=(200, D)) * 1e-7
    X[400:600] = C[3]                # exactly on a duplicated centroid
    P = ctx.points(X)
    ctx.counters(reset=True)
    lab, dist = capi.lloyds_assignment(ctx, P, C, None, EUCLIDEAN)
    rl, rd = port.lloyds_assignment(X, C, None, EUCLIDEAN)
    assert np.array_equal(lab, rl), np.flatnonzero(lab != rl)[:10]
    assert np.array_equal(dist, rd)
    assert NO_TC or ctx.counters()["lloyd_exact"] >= 600
    # float32 points, K not a multiple of the tile width, larger magnitudes
    X32 = (X * 1000).astype(np.float32)
    C2 = C[:37] * 1000
    P32 = ctx.points(X32)
    lab, dist = capi.lloyds_assignment(ctx, P32, C2, None, EUCLIDEAN)
    rl, rd = port.lloyds_assignment(X32.astype(np.float64), C2, None, EUCLIDEAN)
    assert np.array_equal(lab, rl) and np.array_equal(dist, rd)
