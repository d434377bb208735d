"""Seeded synthetic inputs of the shapes SURVEY.md §8(d) names (numpy only; no torch, no oracle).

All generators are counter-based on (seed, config) through numpy's Philox so that a shard
[lo, hi) of a dataset can be generated without materialising the rest.
"""
import numpy as np


def _rng(seed, stream=0):
    return np.random.Generator(np.random.Philox(key=[int(seed) & (2**64 - 1), int(stream)]))


def gaussian_mixture(n, d, n_centres, seed, centre_std=4.0, std=1.0, dtype=np.float32, lo=0):
    """C3/C4/C5 points: mixture of `n_centres` unit Gaussians, centres ~ N(0, centre_std^2).
    Rows [lo, lo+n) of the infinite dataset defined by `seed` (blocks of 65536 rows)."""
    centres = _rng(seed, 1).normal(0.0, centre_std, size=(n_centres, d))
    out = np.empty((n, d), dtype=dtype)
    BLK = 65536
    b0 = lo // BLK
    b1 = (lo + n + BLK - 1) // BLK
    pos = 0
    for b in range(b0, b1):
        g = _rng(seed, 1000 + b)
        which = g.integers(0, n_centres, size=BLK)
        blk = centres[which] + g.normal(0.0, std, size=(BLK, d))
        s = max(lo, b * BLK) - b * BLK
        e = min(lo + n, (b + 1) * BLK) - b * BLK
        out[pos:pos + (e - s)] = blk[s:e].astype(dtype)
        pos += e - s
    return out


def normal_points(n, d, seed, dtype=np.float32):
    """C2-(ii): i.i.d. N(0,1) coordinates."""
    return _rng(seed, 2).normal(size=(n, d)).astype(dtype)


def rating_users(n, d, seed, min_known=2, max_known=9, dtype=np.float64):
    """C1/C2-(i): rating-like user vectors mimicking crypto_rec.hpp:97-132.

    Each user knows `min_known..max_known` coins; a known coin's rating is the sum of the
    positive sentiment scores s = t/sqrt(t^2+15), t ~ N(0,2) over 1..3 tweets (tweet.cpp:40-41),
    or 0 when none is positive; unknown coins hold the mean over the known ones.  Users whose
    vector is all zero are dropped (crypto_rec.hpp:127).  Returns (X[n',d], unknown[n',d] u8, mean[n']).
    With dtype=float32 the ratings are rounded to fp32 first and the mean recomputed from them.
    """
    g = _rng(seed, 3)
    X = np.zeros((n, d), dtype=np.float64)
    known = np.zeros((n, d), dtype=bool)
    nk = g.integers(min_known, max_known + 1, size=n)
    for i in range(n):
        coins = g.choice(d, size=nk[i], replace=False)
        known[i, coins] = True
        for c in coins:
            t = g.normal(0.0, 2.0, size=g.integers(1, 4))
            s = t / np.sqrt(t * t + 15.0)
            X[i, c] = s[s > 0].sum()
    keep = (X != 0).any(axis=1)
    X, known = X[keep], known[keep]
    if dtype == np.float32:
        X = X.astype(np.float32).astype(np.float64)
    cnt = known.sum(axis=1)
    mean = (X * known).sum(axis=1) / cnt
    if dtype == np.float32:
        mean = mean.astype(np.float32).astype(np.float64)
    X = np.where(known, X, mean[:, None])
    return X.astype(dtype), (~known).astype(np.uint8), mean


def rating_users_fast(n, d, seed, min_known=2, max_known=9, dtype=np.float32, chunk=100_000):
    """Vectorised variant of rating_users for bench-scale n (1M users): same distribution family
    (each coin known with probability nk/d, nk ~ U{min_known..max_known}; a known coin's rating is the
    sum of the positive scores of two tweets), generated in chunks.  Not bit-identical to rating_users."""
    Xs, Us, Ms = [], [], []
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        g = _rng(seed, 4000 + c0 // chunk)
        nk = g.integers(min_known, max_known + 1, size=m)
        known = g.random((m, d), dtype=np.float32) < (nk[:, None] / np.float32(d))
        t = g.standard_normal((m, d, 2), dtype=np.float32) * np.float32(2.0)
        s = t / np.sqrt(t * t + np.float32(15.0))
        r = np.where(s > 0, s, np.float32(0)).sum(axis=2)
        X = np.where(known, r, np.float32(0)).astype(np.float32)
        keep = (X != 0).any(axis=1)
        X, known = X[keep], known[keep]
        X64 = X.astype(np.float64)
        mean = ((X64 * known).sum(axis=1) / known.sum(axis=1)).astype(np.float32)
        X = np.where(known, X, mean[:, None]).astype(np.float32)
        Xs.append(X); Us.append((~known).astype(np.uint8)); Ms.append(mean.astype(np.float64))
    X = np.ascontiguousarray(np.concatenate(Xs)).astype(dtype)
    return X, np.ascontiguousarray(np.concatenate(Us)), np.ascontiguousarray(np.concatenate(Ms))
