"""crx_user_vectors_build (tweets_to_user_vectors / clusters_to_user_vectors, crypto_rec.hpp:79-210) against a
sequential restatement of the reference's walk: bit-exact, including the order-dependent double sums."""
import numpy as np
import pytest

from crypto_recommendation_b200 import capi

pytestmark = pytest.mark.gpu


def sequential_walk(mu, mc, ms, n_users, D):
    """crypto_rec.hpp:84-104 then :108-137, one mention at a time."""
    X = np.zeros((n_users, D)); known = np.zeros((n_users, D), bool)
    for u, c, s in zip(mu, mc, ms):
        if s > 0:
            X[u, c] = X[u, c] + s
        known[u, c] = True
    mean = np.zeros(n_users); keep = np.zeros(n_users, np.uint8)
    for u in range(n_users):
        tot = 0.0; cnt = 0
        for j in range(D):
            if known[u, j]:
                tot = tot + X[u, j]; cnt += 1
        with np.errstate(invalid="ignore", divide="ignore"):
            mean[u] = np.float64(tot) / np.float64(cnt)
        if (X[u] != 0).any():
            keep[u] = 1
            X[u, ~known[u]] = mean[u]
    return X, (~known).astype(np.uint8), mean, keep


@pytest.mark.parametrize("n_users,D,E,seed", [(1, 3, 0, 0), (50, 7, 400, 1), (3000, 100, 40000, 2), (257, 33, 5000, 3)])
def test_user_vectors_match_sequential_walk(ctx, n_users, D, E, seed):
    rng = np.random.default_rng(seed)
    mu = rng.integers(0, n_users, E).astype(np.int32)
    # a few heavy (user, coin) runs so that the summation order matters
    mc = np.where(rng.random(E) < 0.3, 0, rng.integers(0, D, E)).astype(np.int32)
    t = rng.normal(0, 2, E)
    ms = t / np.sqrt(t * t + 15)       # tweet.cpp:40, negative scores are mentions that add nothing
    ms[rng.random(E) < 0.05] = 0.0
    X, unk, mean, keep = capi.user_vectors_build(ctx, mu, mc, ms, n_users, D)
    eX, eunk, emean, ekeep = sequential_walk(mu, mc, ms, n_users, D)
    assert np.array_equal(keep, ekeep)
    assert np.array_equal(unk, eunk)
    k = keep.astype(bool)
    assert np.array_equal(mean[k].view(np.int64), emean[k].view(np.int64))
    assert np.array_equal(X.view(np.int64), eX.view(np.int64))


def test_user_vectors_rejects_bad_indices(ctx):
    with pytest.raises(capi.CrxError):
        capi.user_vectors_build(ctx, [0, 5], [0, 0], [0.1, 0.2], 3, 4)
