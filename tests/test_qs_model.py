"""CPU check of the closed form of the Lomuto partition that `warp_qs_topn_big` (csrc/recommend_pass2.cuh) runs.

The reference's quicksort (crypto_rec.hpp:235-277: pivot = last, `>=` goes left, swaps) permutes the elements below the
pivot; the kernel reproduces that permutation without the sequential swap loop:
  * the `>=` elements keep their order in front, the pivot follows them;
  * with q1 = the first `<` element and rho = #{`>=` behind q1} + 1, every `<` element at a position >= q1 + rho stays
    where it is, and the holes left there by `>=` elements (plus the pivot's slot) receive the `<` elements of
    [q1, q1 + rho): hole number r receives the content of position q1 + r - 1, which is the content of hole r'' when that
    position holds the r''-th `>=` element itself;
  * ranges that start at or beyond `need` are skipped; a pivot that is a minimum of its range stays in place together
    with its equals behind the last larger element.
`model` is that algorithm, element for element what the warp does; the literal sort is the oracle's."""
import numpy as np
import pytest


def model(keys, need, passes=None, prefix_min_jump=False):
    """prefix_min_jump: the remedy DESIGN.md section 9 proposes for lists with long non-increasing tails (NOT in the kernel yet):
    when every element of the range is >= the pivot, jump straight to the last element that is not a minimum of what precedes
    it -- one walk instead of one walk per trailing minimum."""
    key = list(keys); val = list(range(len(keys))); n = len(key)
    stack = [(0, n - 1)]
    touched = 0
    while stack:
        lo, hi = stack.pop()
        while lo < hi and lo < need:
            pivot = key[hi]; m = hi - lo
            rng = range(lo, hi)
            touched += m
            cnt = sum(1 for j in rng if key[j] >= pivot)
            if all(key[j] == pivot for j in rng):
                break
            if cnt == m:
                if prefix_min_jump:
                    touched += m
                    run_min, last = key[lo], None
                    for j in range(lo + 1, hi):
                        if key[j] > run_min:
                            last = j
                        run_min = min(run_min, key[j])
                    if last is None:
                        break          # non-increasing: every remaining step leaves everything where it is
                    hi = last
                    continue
                hi = max(j for j in rng if key[j] > pivot)
                continue
            p = lo + cnt
            pv = val[hi]
            literal = p + 1 < need and p < hi
            if literal:
                q1 = min(j for j in rng if not key[j] >= pivot)
                posge = [None] + [j for j in range(q1 + 1, hi) if key[j] >= pivot]
                rho = len(posge)
                posge.append(hi)
                assert q1 + rho - 1 == p
                rank_of = {posge[r]: r for r in range(1, rho + 1)}
                content = [None] * (rho + 1)
                for r in range(1, rho + 1):
                    src = q1 + r - 1
                    content[r] = (key[src], val[src]) if not key[src] >= pivot else content[rank_of[src]]
            ge = [(key[j], val[j]) for j in rng if key[j] >= pivot]
            for i, (kk, vv) in enumerate(ge):
                key[lo + i], val[lo + i] = kk, vv
            key[p], val[p] = pivot, pv
            if literal:
                for r in range(1, rho + 1):
                    if posge[r] > p:
                        key[posge[r]], val[posge[r]] = content[r]
                stack.append((p + 1, hi))
            hi = p - 1
    if passes is not None:
        passes.append(touched)
    return val[:min(need, n)]


def sequences(rng, count, nmax):
    for it in range(count):
        n = int(rng.integers(1, nmax))
        levels = int(rng.choice([1, 2, 3, 5, 10, 50, 1000, 100000]))
        s = rng.integers(0, levels + 1, n) / levels
        if rng.random() < 0.25:   # a plateau with a few larger values in it (rating-like users)
            s = np.where(rng.random(n) < 0.9, 0.5, s)
        yield s


@pytest.mark.parametrize("which", ["reference", "port"])
def test_closed_form_partition_matches_literal_quicksort(which, port, ref):
    oracle = ref if which == "reference" else port
    if oracle is None:
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(3)
    for s in sequences(rng, 3000, 300):
        n = len(s)
        _, full = oracle.quicksort(s, np.arange(n, dtype=np.int32))
        for need in (int(rng.integers(1, 33)), n):
            assert model(s.tolist(), need) == full[:min(need, n)].tolist(), (n, need)


def test_prefix_minimum_jump_gives_the_same_order_and_is_linear_on_descending_tails(port):
    """The literal quicksort (and the kernel's closed form) walks a descending list once per element; with the jump over
    trailing prefix minima the same prefix of the order comes out of a few walks.  Specification of the next kernel step."""
    rng = np.random.default_rng(6)
    for s in sequences(rng, 1500, 200):
        n = len(s)
        if rng.random() < 0.5:   # a non-increasing tail behind a random head
            s = np.r_[s, np.sort(rng.integers(0, 6, int(rng.integers(1, 80))) / 5.0)[::-1]]
            n = len(s)
        _, full = port.quicksort(s, np.arange(n, dtype=np.int32))
        for need in (int(rng.integers(1, 33)), n):
            assert model(s.tolist(), need, prefix_min_jump=True) == full[:min(need, n)].tolist(), (n, need)
    n = 3000
    s = np.arange(n, 0, -1, dtype=np.float64)
    _, full = port.quicksort(s, np.arange(n, dtype=np.int32))
    slow, fast = [], []
    assert model(s.tolist(), 20, slow) == full[:20].tolist()
    assert model(s.tolist(), 20, fast, prefix_min_jump=True) == full[:20].tolist()
    assert slow[0] > n * n // 4 and fast[0] <= 3 * n, (slow, fast)


def test_plateau_costs_a_few_passes(port):
    """8000 equal similarities with 15 larger ones in between: the literal sort walks ~n^2/2 elements, the closed form with
    the minimum-pivot shortcut a few times n."""
    rng = np.random.default_rng(4)
    s = np.full(8000, 0.5)
    s[rng.integers(0, 8000, 15)] = 0.5 + rng.random(15) * 0.1
    _, full = port.quicksort(s, np.arange(8000, dtype=np.int32))
    passes = []
    assert model(s.tolist(), 20, passes) == full[:20].tolist()
    assert passes[0] < 40 * 8000, passes
