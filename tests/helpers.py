"""Comparison helpers for the parity tests.

Integer outputs must be bit-exact.  The only tolerated disagreement is a near-tie: the oracle's own
values for the two alternatives differ by less than 1e-6 relative (BASELINE.json north_star); those
are counted and reported, and every test bounds their number.
"""
import numpy as np

REL_TIE = 1e-6
REL_FLOAT = 1e-4


def rel_close(a, b, rtol=REL_FLOAT, atol=1e-12):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    ok = np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b))
    return ok | both_nan


def assert_float_close(a, b, rtol=REL_FLOAT, what=""):
    ok = rel_close(a, b, rtol)
    assert ok.all(), "%s: %d of %d values differ by more than %g relative (first at %s)" % (
        what, (~ok).sum(), ok.size, rtol, np.argwhere(~ok)[:3].tolist())


def label_mismatch_report(labels, labels_ref, dist_fn):
    """dist_fn(row, centroid) -> the oracle's distance.  Returns (n_mismatch, n_near_tie)."""
    bad = np.flatnonzero(labels != labels_ref)
    near = 0
    for v in bad:
        a, b = dist_fn(v, labels[v]), dist_fn(v, labels_ref[v])
        if abs(a - b) <= REL_TIE * max(abs(a), abs(b)):
            near += 1
    return len(bad), near


def assert_labels(labels, labels_ref, dist_fn, max_near=0, what="labels"):
    n_bad, n_near = label_mismatch_report(labels, labels_ref, dist_fn)
    assert n_bad == n_near, "%s: %d hard mismatches (%d near-ties)" % (what, n_bad - n_near, n_near)
    assert n_near <= max_near, "%s: %d near-tie mismatches (allowed %d)" % (what, n_near, max_near)
    return n_near


def topp_compare(nbr, sim, nbr_ref, sim_ref):
    """Per query: neighbour lists must hold the same rows in the same order; a difference is a
    near-tie only if the similarities at the differing positions agree within REL_TIE.
    Returns (n_queries_with_hard_mismatch, n_queries_with_near_tie_only)."""
    hard = soft = 0
    for q in range(nbr.shape[0]):
        if np.array_equal(nbr[q], nbr_ref[q]):
            continue
        pos = np.flatnonzero(nbr[q] != nbr_ref[q])
        a, b = sim[q][pos], sim_ref[q][pos]
        if (np.abs(a - b) <= REL_TIE * np.maximum(np.abs(a), np.abs(b))).all() and \
                sorted(nbr[q][nbr[q] >= 0].tolist()) == sorted(nbr_ref[q][nbr_ref[q] >= 0].tolist()):
            soft += 1
        elif (np.abs(a - b) <= REL_TIE * np.maximum(np.abs(a), np.abs(b))).all():
            soft += 1
        else:
            hard += 1
    return hard, soft
