"""CPU: pins the restatement (oracle/crx_oracle.cpp) against (1) the committed golden vectors made
from the reference's own code and (2) the reference build itself when it is available here."""
import numpy as np
import pytest

from oracle import EUCLIDEAN, COSINE
from crypto_recommendation_b200 import synth


def same(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


# ---------------- golden vectors (always) ----------------
def test_golden_kats(port, golden):
    g = golden
    assert [port.mod_ii(-7, 5), port.mod_li(-7, 2147483647), port.mod_iz(-1, 16), port.mod_iz(-3, 50000),
            port.mod_ii(5, 1), port.mod_ii(-5, 2)] == g["kat_mod"].tolist() == [3, 2147483640, 15, 1613, 0, 1]
    i = 0
    while "kat_hamming_args_%d" % i in g:
        assert port.hamming(*g["kat_hamming_args_%d" % i].tolist()) == g["kat_hamming_out_%d" % i].tolist()
        i += 1
    assert i == 7
    assert g["kat_hamming_out_0"].tolist() == [4, 7, 1, 13] and g["kat_hamming_out_1"].tolist() == [6, 0, 12, 3, 15, 9]
    s, d = port.quicksort(g["kat_qs_in"], np.arange(8))
    assert same(d, g["kat_qs_ids"]) and d.tolist() == [6, 1, 3, 2, 5, 0, 7, 4] and same(s, g["kat_qs_sims"])
    s, d = port.quicksort(g["kat_qs2_in"], np.arange(97))
    assert same(d, g["kat_qs2_ids"]) and same(s, g["kat_qs2_sims"])
    s, d = port.quicksort(np.ones(10), np.arange(10))
    assert d.tolist() == list(range(10))
    nd, nf, uf, ui, u12 = port.rng_kat(42)
    assert same(nd, g["kat_rng_nd"]) and same(nf, g["kat_rng_nf"]) and same(uf, g["kat_rng_uf"])
    assert same(ui, g["kat_rng_ui"]) and same(u12, g["kat_rng_u12"])
    assert nd[0] == -1.7141127395876619 and ui[0] == 19 and u12.tolist() == [2, 2, 2, 1, 1, 2]


def test_golden_vector_math(port, golden):
    g = golden
    for i in range(6):
        assert port.inner_product(g["vm_a"][i], g["vm_b"][i]) == g["vm_ip"][i]
        assert port.euclidean_distance(g["vm_a"][i], g["vm_b"][i]) == g["vm_eu"][i]
        assert port.cosine_distance(g["vm_a"][i], g["vm_b"][i]) == g["vm_cd"][i]
        assert port.cosine_similarity(g["vm_a"][i], g["vm_b"][i]) == g["vm_cs"][i]


def test_golden_hash(port, golden):
    g = golden
    X = g["hash_X"]
    assert same(port.lsh_hash(X, COSINE, 4, 5, 100, 0.4, 7001)[0], g["lsh_cos_ids"])
    ids, det = port.lsh_hash(X, EUCLIDEAN, 4, 5, 10, 4.0, 7002)
    assert same(ids, g["lsh_euc_ids"]) and same(det, g["lsh_euc_det"])
    for q in (0, 123, 499):
        assert same(port.lsh_candidates(X, COSINE, 4, 5, 100, 0.4, 7001, q, 1), g["lsh_cos_cand_%d" % q])
        assert same(port.lsh_candidates(X, EUCLIDEAN, 4, 5, 10, 4.0, 7002, q, 1), g["lsh_euc_cand_f_%d" % q])
        assert same(port.lsh_candidates(X, EUCLIDEAN, 4, 5, 10, 4.0, 7002, q, 0), g["lsh_euc_cand_u_%d" % q])
    assert same(port.cube_hash(X, COSINE, 5, 0.4, 7003), g["cube_cos_ids"])
    assert same(port.cube_hash(X, EUCLIDEAN, 5, 4.0, 7004), g["cube_euc_ids"])
    for p in (1, 2, 7, 40):
        assert same(port.cube_candidates(X, COSINE, 5, 0.4, 7003, 17, p), g["cube_cos_cand_p%d" % p])
        assert same(port.cube_candidates(X, EUCLIDEAN, 5, 4.0, 7004, 17, p), g["cube_euc_cand_p%d" % p])


@pytest.mark.parametrize("m,name", [(EUCLIDEAN, "euc"), (COSINE, "cos")])
def test_golden_clustering(port, golden, m, name):
    g = golden
    X = g["cl_X"]
    assert same(port.rand_selection(X, 9, 8001), g["cl_rand_sel"])
    cidx = port.k_means_pp(X, 7, m, 8002)
    assert same(cidx, g["cl_kpp_%s" % name])
    lab, dist = port.lloyds_assignment(X, X[cidx], cidx, m)
    assert same(lab, g["cl_lloyd_lab_%s" % name]) and same(dist, g["cl_lloyd_dist_%s" % name])
    ret, newc = port.k_means(X, lab, X[cidx], m, 0.05)
    assert ret == bool(g["cl_kmeans_ret_%s" % name]) and same(newc, g["cl_kmeans_C_%s" % name])
    lab2, dist2 = port.lloyds_assignment(X, newc, None, m)
    assert same(lab2, g["cl_lloyd2_lab_%s" % name]) and same(dist2, g["cl_lloyd2_dist_%s" % name])
    l, d, b = port.lsh_range_assignment(X, cidx, m, 4, 5, 10, 4.0, 8003)
    assert same(l, g["cl_lshrange_lab_%s" % name]) and same(d, g["cl_lshrange_dist_%s" % name]) and same(b, g["cl_lshrange_before_%s" % name])
    l, d, b = port.cube_range_assignment(X, cidx, m, 5, 4.0, 6, 8004)
    assert same(l, g["cl_cuberange_lab_%s" % name]) and same(d, g["cl_cuberange_dist_%s" % name]) and same(b, g["cl_cuberange_before_%s" % name])
    sw, new = port.pam_lloyds(X, lab, cidx, m)
    assert sw == bool(g["cl_pam_sw_%s" % name]) and same(new, g["cl_pam_new_%s" % name])
    assert same(port.silhouette(X, lab, X[cidx], m), g["cl_sil_%s" % name])


def test_golden_recommend(port, golden):
    g = golden
    U, unk, mean = g["rec_U"], g["rec_unk"], g["rec_mean"]
    r = port.recommend_lsh(U, unk, mean, COSINE, 4, 5, 100, 0.4, 20, 5, 9001)
    for got, key in zip(r, ("rec_A_recs", "rec_A_nbr", "rec_A_sim", "rec_A_ncand")):
        assert same(got, g[key]), key
    r = port.recommend_lsh(U[:25], unk[:25], mean[:25], COSINE, 4, 5, 100, 0.4, 20, 2, 9002, Xq=U[25:], unknown_q=unk[25:], mean_q=mean[25:])
    for got, key in zip(r, ("rec_B_recs", "rec_B_nbr", "rec_B_sim", "rec_B_ncand")):
        assert same(got, g[key]), key
    assert same(port.recommend_cluster(U, unk, mean, g["rec_C_lab"], 12, 5), g["rec_C_recs"])


# ---------------- live differential against the reference build (where present) ----------------
@pytest.fixture(scope="module")
def both(port, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built here (no /root/reference): golden vectors pin the port instead")
    return ref, port


@pytest.mark.parametrize("seed", [1, 2])
def test_live_hash_and_tables(both, seed):
    ref, port = both
    X = synth.gaussian_mixture(1500, 20, 8, seed=seed).astype(np.float64)
    for metric, k, L, div, w in [(COSINE, 4, 5, 100, 0.4), (EUCLIDEAN, 4, 5, 100, 0.4), (EUCLIDEAN, 3, 2, 10, 4.0), (COSINE, 7, 3, 1, 1.0)]:
        r = ref.lsh_hash(X, metric, k, L, div, w, 70 + seed); p = port.lsh_hash(X, metric, k, L, div, w, 70 + seed)
        assert same(r[0], p[0])
        if metric == EUCLIDEAN:
            assert same(r[1], p[1])
        for q in (0, 17, 1499):
            for filt in (0, 1):
                assert same(ref.lsh_candidates(X, metric, k, L, div, w, 70 + seed, q, filt), port.lsh_candidates(X, metric, k, L, div, w, 70 + seed, q, filt))
    for metric, k, w in [(COSINE, 6, 0.4), (EUCLIDEAN, 6, 0.4), (EUCLIDEAN, 8, 4.0)]:
        assert same(ref.cube_hash(X, metric, k, w, 90 + seed), port.cube_hash(X, metric, k, w, 90 + seed))
        for probes in (1, 2, 5, 10, 300):
            assert same(ref.cube_candidates(X, metric, k, w, 90 + seed, 5, probes), port.cube_candidates(X, metric, k, w, 90 + seed, 5, probes))


@pytest.mark.parametrize("metric", [EUCLIDEAN, COSINE])
def test_live_clustering(both, metric):
    ref, port = both
    X = synth.gaussian_mixture(2000, 24, 8, seed=5).astype(np.float64)
    assert same(ref.rand_selection(X, 30, 123), port.rand_selection(X, 30, 123))
    cidx = ref.k_means_pp(X, 12, metric, 321)
    assert same(cidx, port.k_means_pp(X, 12, metric, 321))
    C = X[cidx]
    rl, rd = ref.lloyds_assignment(X, C, cidx, metric); pl, pd_ = port.lloyds_assignment(X, C, cidx, metric)
    assert same(rl, pl) and same(rd, pd_)
    rr, rC = ref.k_means(X, rl, C, metric, 0.05); pr, pC = port.k_means(X, rl, C, metric, 0.05)
    assert rr == pr and same(rC, pC)
    for a, b in zip(ref.lsh_range_assignment(X, cidx, metric, 4, 5, 100, 0.4, 55), port.lsh_range_assignment(X, cidx, metric, 4, 5, 100, 0.4, 55)):
        assert same(a, b)
    for a, b in zip(ref.cube_range_assignment(X, cidx, metric, 6, 0.4, 10, 56), port.cube_range_assignment(X, cidx, metric, 6, 0.4, 10, 56)):
        assert same(a, b)
    # heap centroids (k_means centres): unique ids, and the shared "k_means_center" id of SURVEY App. A-2 -- the port keeps
    # the reference's string-keyed distance cache literally
    rng = np.random.default_rng(3)
    Cm = np.stack([X[rng.choice(len(X), 40)].mean(0) for _ in range(9)])
    mixed = np.array([-1, int(cidx[1]), -1, -1, int(cidx[4]), -1, -1, -1, -1], np.int32)
    for shared, rows in ((0, None), (1, None), (0, mixed), (1, mixed)):
        ra = ref.lsh_range_assignment_vectors(X, Cm, rows, shared, metric, 4, 5, 100, 0.4, 55)
        pa = port.lsh_range_assignment_vectors(X, Cm, rows, shared, metric, 4, 5, 100, 0.4, 55)
        for a, b in zip(ra, pa):
            assert same(a, b)
    Xs = X[:500]
    cs = ref.rand_selection(Xs, 6, 9); ls, _ = ref.lloyds_assignment(Xs, Xs[cs], cs, metric)
    a, b = ref.pam_lloyds(Xs, ls, cs, metric), port.pam_lloyds(Xs, ls, cs, metric)
    assert a[0] == b[0] and same(a[1], b[1])
    assert same(ref.silhouette(Xs, ls, Xs[cs], metric), port.silhouette(Xs, ls, Xs[cs], metric))


def test_live_recommend(both):
    ref, port = both
    U, unk, mean = synth.rating_users(900, 100, seed=3)
    for a, b in zip(ref.recommend_lsh(U, unk, mean, COSINE, 4, 5, 100, 0.4, 20, 5, 4242), port.recommend_lsh(U, unk, mean, COSINE, 4, 5, 100, 0.4, 20, 5, 4242)):
        assert same(a, b)
    for a, b in zip(ref.recommend_lsh(U, unk, mean, EUCLIDEAN, 4, 5, 100, 0.4, 20, 5, 4244), port.recommend_lsh(U, unk, mean, EUCLIDEAN, 4, 5, 100, 0.4, 20, 5, 4244)):
        assert same(a, b)
    cs = ref.rand_selection(U, 30, 9); ls, _ = ref.lloyds_assignment(U, U[cs], cs, EUCLIDEAN)
    assert same(ref.recommend_cluster(U, unk, mean, ls, 30, 5), port.recommend_cluster(U, unk, mean, ls, 30, 5))
