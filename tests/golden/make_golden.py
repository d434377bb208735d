"""Generates tests/golden/golden_v1.npz from the REFERENCE build (oracle/_ref/libcrx_ref.so, i.e. the
reference's own headers compiled from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are stored next to the outputs so the fixtures do not depend on numpy's generators.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import load, EUCLIDEAN, COSINE  # noqa: E402
from crypto_recommendation_b200 import synth  # noqa: E402


def main():
    ref = load("reference")
    assert ref is not None and ref.kind == "reference", "reference build unavailable"
    g = {}
    # ---- known answers
    g["kat_mod"] = np.array([ref.mod_ii(-7, 5), ref.mod_li(-7, 2147483647), ref.mod_iz(-1, 16), ref.mod_iz(-3, 50000),
                             ref.mod_ii(5, 1), ref.mod_ii(-5, 2)], np.int64)
    for i, a in enumerate([(5, 1, 0, 4), (5, 2, 0, 4), (5, 3, 0, 4), (0, 2, 0, 4), (1234, 3, 0, 16), (9, 4, 0, 4), (9, 5, 0, 4)]):
        g["kat_hamming_args_%d" % i] = np.array(a, np.int32)
        g["kat_hamming_out_%d" % i] = np.array(ref.hamming(*a), np.int32)
    sims = np.array([.5, .9, .5, .9, .1, .5, 1.0, .5])
    g["kat_qs_in"] = sims
    g["kat_qs_sims"], g["kat_qs_ids"] = ref.quicksort(sims, np.arange(8))
    rng = np.random.default_rng(11)
    s2 = rng.integers(0, 6, 97) / 5.0
    g["kat_qs2_in"] = s2
    g["kat_qs2_sims"], g["kat_qs2_ids"] = ref.quicksort(s2, np.arange(97))
    nd, nf, uf, ui, u12 = ref.rng_kat(42)
    g["kat_rng_nd"], g["kat_rng_nf"], g["kat_rng_uf"], g["kat_rng_ui"], g["kat_rng_u12"] = nd, nf, uf, ui, u12
    a = rng.normal(size=(6, 100)); b = rng.normal(size=(6, 100))
    g["vm_a"], g["vm_b"] = a, b
    g["vm_ip"] = np.array([ref.inner_product(a[i], b[i]) for i in range(6)])
    g["vm_eu"] = np.array([ref.euclidean_distance(a[i], b[i]) for i in range(6)])
    g["vm_cd"] = np.array([ref.cosine_distance(a[i], b[i]) for i in range(6)])
    g["vm_cs"] = np.array([ref.cosine_similarity(a[i], b[i]) for i in range(6)])

    # ---- hashing / tables
    X = synth.gaussian_mixture(500, 16, 6, seed=101).astype(np.float64)
    g["hash_X"] = X
    g["lsh_cos_ids"], _ = ref.lsh_hash(X, COSINE, 4, 5, 100, 0.4, 7001)
    g["lsh_euc_ids"], g["lsh_euc_det"] = ref.lsh_hash(X, EUCLIDEAN, 4, 5, 10, 4.0, 7002)
    for q in (0, 123, 499):
        g["lsh_cos_cand_%d" % q] = ref.lsh_candidates(X, COSINE, 4, 5, 100, 0.4, 7001, q, 1)
        g["lsh_euc_cand_f_%d" % q] = ref.lsh_candidates(X, EUCLIDEAN, 4, 5, 10, 4.0, 7002, q, 1)
        g["lsh_euc_cand_u_%d" % q] = ref.lsh_candidates(X, EUCLIDEAN, 4, 5, 10, 4.0, 7002, q, 0)
    g["cube_cos_ids"] = ref.cube_hash(X, COSINE, 5, 0.4, 7003)
    g["cube_euc_ids"] = ref.cube_hash(X, EUCLIDEAN, 5, 4.0, 7004)
    for probes in (1, 2, 7, 40):
        g["cube_cos_cand_p%d" % probes] = ref.cube_candidates(X, COSINE, 5, 0.4, 7003, 17, probes)
        g["cube_euc_cand_p%d" % probes] = ref.cube_candidates(X, EUCLIDEAN, 5, 4.0, 7004, 17, probes)

    # ---- clustering
    Xc = synth.gaussian_mixture(700, 12, 5, seed=202).astype(np.float64)
    g["cl_X"] = Xc
    for m, name in ((EUCLIDEAN, "euc"), (COSINE, "cos")):
        g["cl_rand_sel"] = ref.rand_selection(Xc, 9, 8001)
        cidx = ref.k_means_pp(Xc, 7, m, 8002)
        g["cl_kpp_%s" % name] = cidx
        lab, dist = ref.lloyds_assignment(Xc, Xc[cidx], cidx, m)
        g["cl_lloyd_lab_%s" % name], g["cl_lloyd_dist_%s" % name] = lab, dist
        ret, newc = ref.k_means(Xc, lab, Xc[cidx], m, 0.05)
        g["cl_kmeans_ret_%s" % name], g["cl_kmeans_C_%s" % name] = np.array(ret), newc
        lab2, dist2 = ref.lloyds_assignment(Xc, newc, None, m)
        g["cl_lloyd2_lab_%s" % name], g["cl_lloyd2_dist_%s" % name] = lab2, dist2
        l, d, bef = ref.lsh_range_assignment(Xc, cidx, m, 4, 5, 10, 4.0, 8003)
        g["cl_lshrange_lab_%s" % name], g["cl_lshrange_dist_%s" % name], g["cl_lshrange_before_%s" % name] = l, d, bef
        l, d, bef = ref.cube_range_assignment(Xc, cidx, m, 5, 4.0, 6, 8004)
        g["cl_cuberange_lab_%s" % name], g["cl_cuberange_dist_%s" % name], g["cl_cuberange_before_%s" % name] = l, d, bef
        sw, new = ref.pam_lloyds(Xc, lab, cidx, m)
        g["cl_pam_sw_%s" % name], g["cl_pam_new_%s" % name] = np.array(sw), new
        g["cl_sil_%s" % name] = ref.silhouette(Xc, lab, Xc[cidx], m)

    # ---- recommendation
    U, unk, mean = synth.rating_users(320, 100, seed=303)
    g["rec_U"], g["rec_unk"], g["rec_mean"] = U, unk, mean
    r = ref.recommend_lsh(U, unk, mean, COSINE, 4, 5, 100, 0.4, 20, 5, 9001)
    g["rec_A_recs"], g["rec_A_nbr"], g["rec_A_sim"], g["rec_A_ncand"] = r
    V = U[:25]
    r = ref.recommend_lsh(V, unk[:25], mean[:25], COSINE, 4, 5, 100, 0.4, 20, 2, 9002, Xq=U[25:], unknown_q=unk[25:], mean_q=mean[25:])
    g["rec_B_recs"], g["rec_B_nbr"], g["rec_B_sim"], g["rec_B_ncand"] = r
    cs = ref.rand_selection(U, 12, 9003)
    lab, _ = ref.lloyds_assignment(U, U[cs], cs, EUCLIDEAN)
    g["rec_C_cs"], g["rec_C_lab"] = cs, lab
    g["rec_C_recs"] = ref.recommend_cluster(U, unk, mean, lab, 12, 5)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
