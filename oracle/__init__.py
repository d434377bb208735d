"""ctypes loader for the two CPU checkers (TEST INFRASTRUCTURE ONLY -- see oracle/oracle_api.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package never does.

    from oracle import load
    port = load("port")        # oracle/_build/libcrx_oracle.so  (restatement; built on demand)
    ref  = load("reference")   # oracle/_ref/libcrx_ref.so       (reference headers; None if absent)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_build", "libcrx_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libcrx_ref.so")
REF_DIR = os.environ.get("CRX_REF_DIR", "/root/reference")

EUCLIDEAN, COSINE = 0, 1


def build(verbose=False):
    """Compile the port always, the reference build only where the reference tree exists."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=out)
    if os.path.isdir(os.path.join(REF_DIR, "lib")):
        subprocess.check_call(["make", "-C", _HERE, "ref", "CRX_REF_DIR=" + REF_DIR], stdout=out)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


class Oracle:
    """numpy front-end over oracle_api.h; identical for both shared objects."""

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.orc_kind.restype = ctypes.c_char_p
        for n in ("orc_inner_product", "orc_euclidean_distance", "orc_cosine_distance", "orc_cosine_similarity"):
            getattr(L, n).restype = ctypes.c_double
        L.orc_lsh_candidates.restype = ctypes.c_int64
        L.orc_cube_candidates.restype = ctypes.c_int64
        L.orc_rec_handle_create.restype = ctypes.c_void_p
        self.kind = L.orc_kind().decode()

    # ---- KATs
    def mod_ii(self, x, n): return self.lib.orc_mod_ii(ctypes.c_int(x), ctypes.c_int(n))
    def mod_li(self, x, n): return self.lib.orc_mod_li(ctypes.c_long(x), ctypes.c_int(n))
    def mod_iz(self, x, n): return self.lib.orc_mod_iz(ctypes.c_int(x), ctypes.c_size_t(n))
    def mod_ui(self, x, n): return self.lib.orc_mod_ui(ctypes.c_uint(x), ctypes.c_int(n))

    def hamming(self, num, dist, min_bit, bits):
        cap = 1 << 16
        out = np.zeros(cap, np.int32)
        n = self.lib.orc_hamming(num, dist, min_bit, bits, _p(out), cap)
        return out[:n].tolist()

    def quicksort(self, sims, ids):
        s = _f64(sims).copy()
        d = _i32(ids).copy()
        self.lib.orc_quicksort(_p(s), _p(d), ctypes.c_int(len(s)))
        return s, d

    def rng_kat(self, seed):
        nd = np.zeros(3); nf = np.zeros(2, np.float32); uf = np.zeros(1, np.float32)
        ui = np.zeros(1, np.int32); u12 = np.zeros(6, np.int32)
        self.lib.orc_rng_kat(ctypes.c_uint64(seed), _p(nd), _p(nf), _p(uf), _p(ui), _p(u12))
        return nd, nf, uf, ui, u12

    # ---- vector math
    def _vv(self, name, a, b):
        a = _f64(a); b = _f64(b)
        return getattr(self.lib, name)(_p(a), _p(b), ctypes.c_int(len(a)))

    def inner_product(self, a, b): return self._vv("orc_inner_product", a, b)
    def euclidean_distance(self, a, b): return self._vv("orc_euclidean_distance", a, b)
    def cosine_distance(self, a, b): return self._vv("orc_cosine_distance", a, b)
    def cosine_similarity(self, a, b): return self._vv("orc_cosine_similarity", a, b)

    # ---- LSH / cube
    def lsh_hash(self, X, metric, k, L, div, w, seed):
        X = _f64(X); N, D = X.shape
        ids = np.zeros((L, N), np.int32)
        det = np.zeros((L, N, k), np.int32) if metric == EUCLIDEAN else None
        self.lib.orc_lsh_hash(_p(X), ctypes.c_int64(N), D, metric, k, L, div, ctypes.c_double(w),
                              ctypes.c_uint64(seed), _p(ids), _p(det))
        return ids, det

    def lsh_candidates(self, X, metric, k, L, div, w, seed, q, filtered):
        X = _f64(X); N, D = X.shape
        out = np.zeros(N, np.int32)
        n = self.lib.orc_lsh_candidates(_p(X), ctypes.c_int64(N), D, metric, k, L, div, ctypes.c_double(w),
                                        ctypes.c_uint64(seed), ctypes.c_int64(q), int(filtered), _p(out),
                                        ctypes.c_int64(N))
        return out[:n].copy()

    def cube_hash(self, X, metric, k, w, seed):
        X = _f64(X); N, D = X.shape
        ids = np.zeros(N, np.int32)
        self.lib.orc_cube_hash(_p(X), ctypes.c_int64(N), D, metric, k, ctypes.c_double(w), ctypes.c_uint64(seed), _p(ids))
        return ids

    def cube_candidates(self, X, metric, k, w, seed, q, probes):
        X = _f64(X); N, D = X.shape
        out = np.zeros(N, np.int32)
        n = self.lib.orc_cube_candidates(_p(X), ctypes.c_int64(N), D, metric, k, ctypes.c_double(w),
                                         ctypes.c_uint64(seed), ctypes.c_int64(q), probes, _p(out), ctypes.c_int64(N))
        return out[:n].copy()

    # ---- clustering
    def rand_selection(self, X, K, seed):
        X = _f64(X); N, D = X.shape
        idx = np.zeros(K, np.int32)
        self.lib.orc_rand_selection(_p(X), ctypes.c_int64(N), D, K, ctypes.c_uint64(seed), _p(idx))
        return idx

    def k_means_pp(self, X, K, metric, seed):
        X = _f64(X); N, D = X.shape
        idx = np.zeros(K, np.int32)
        self.lib.orc_k_means_pp(_p(X), ctypes.c_int64(N), D, K, metric, ctypes.c_uint64(seed), _p(idx))
        return idx

    def lloyds_assignment(self, X, C, cidx, metric):
        X = _f64(X); C = _f64(C); N, D = X.shape; K = C.shape[0]
        cidx = _i32(cidx) if cidx is not None else None
        labels = np.zeros(N, np.int32); dists = np.zeros(N)
        self.lib.orc_lloyds_assignment(_p(X), ctypes.c_int64(N), D, _p(C), K, _p(cidx), metric, _p(labels), _p(dists))
        return labels, dists

    def lsh_range_assignment(self, X, cidx, metric, k, L, div, w, seed):
        X = _f64(X); N, D = X.shape; cidx = _i32(cidx); K = len(cidx)
        labels = np.zeros(N, np.int32); dists = np.zeros(N); before = np.zeros(N, np.int32)
        self.lib.orc_lsh_range_assignment(_p(X), ctypes.c_int64(N), D, _p(cidx), K, metric, k, L, div,
                                          ctypes.c_double(w), ctypes.c_uint64(seed), _p(labels), _p(dists), _p(before))
        return labels, dists, before

    def lsh_range_assignment_vectors(self, X, C, cidx, shared_ids, metric, k, L, div, w, seed):
        X = _f64(X); C = _f64(C); N, D = X.shape; K = C.shape[0]
        cidx = _i32(cidx) if cidx is not None else None
        labels = np.zeros(N, np.int32); dists = np.zeros(N); before = np.zeros(N, np.int32)
        self.lib.orc_lsh_range_assignment_vectors(_p(X), ctypes.c_int64(N), D, _p(C), _p(cidx), K, int(bool(shared_ids)), metric, k, L, div,
                                                  ctypes.c_double(w), ctypes.c_uint64(seed), _p(labels), _p(dists), _p(before))
        return labels, dists, before

    def cube_range_assignment(self, X, cidx, metric, k, w, probes, seed):
        X = _f64(X); N, D = X.shape; cidx = _i32(cidx); K = len(cidx)
        labels = np.zeros(N, np.int32); dists = np.zeros(N); before = np.zeros(N, np.int32)
        self.lib.orc_cube_range_assignment(_p(X), ctypes.c_int64(N), D, _p(cidx), K, metric, k, ctypes.c_double(w),
                                           probes, ctypes.c_uint64(seed), _p(labels), _p(dists), _p(before))
        return labels, dists, before

    def k_means(self, X, labels, C, metric, min_dist):
        X = _f64(X); C = _f64(C); labels = _i32(labels); N, D = X.shape; K = C.shape[0]
        newC = np.zeros_like(C)
        r = self.lib.orc_k_means(_p(X), ctypes.c_int64(N), D, _p(labels), _p(C), K, metric, ctypes.c_double(min_dist),
                                 _p(newC))
        return bool(r), newC

    def pam_lloyds(self, X, labels, cidx, metric):
        X = _f64(X); labels = _i32(labels); cidx = _i32(cidx); N, D = X.shape; K = len(cidx)
        new = np.zeros(K, np.int32)
        r = self.lib.orc_pam_lloyds(_p(X), ctypes.c_int64(N), D, _p(labels), _p(cidx), K, metric, _p(new))
        return bool(r), new

    def silhouette(self, X, labels, C, metric):
        X = _f64(X); C = _f64(C); labels = _i32(labels); N, D = X.shape; K = C.shape[0]
        s = np.zeros(K + 1)
        self.lib.orc_silhouette(_p(X), ctypes.c_int64(N), D, _p(labels), _p(C), K, metric, _p(s))
        return s

    # ---- recommendation
    def recommend_lsh(self, X, unknown, mean, metric, k, L, div, w, P, Nrec, seed, Xq=None, unknown_q=None, mean_q=None):
        X = _f64(X); unknown = _u8(unknown); mean = _f64(mean); N, D = X.shape
        Xq = _f64(Xq); unknown_q = _u8(unknown_q); mean_q = _f64(mean_q)
        Nq = N if Xq is None else Xq.shape[0]
        recs = np.zeros((Nq, Nrec), np.int32); nidx = np.zeros((Nq, P), np.int32)
        nsim = np.zeros((Nq, P)); ncand = np.zeros(Nq, np.int32)
        self.lib.orc_recommend_lsh(_p(X), _p(unknown), _p(mean), ctypes.c_int64(N), D, _p(Xq), _p(unknown_q), _p(mean_q),
                                   ctypes.c_int64(Nq), metric, k, L, div, ctypes.c_double(w), P, Nrec,
                                   ctypes.c_uint64(seed), _p(recs), _p(nidx), _p(nsim), _p(ncand))
        return recs, nidx, nsim, ncand

    def recommend_cluster(self, X, unknown, mean, labels, K, Nrec, Xq=None, unknown_q=None, mean_q=None, qlabels=None):
        X = _f64(X); unknown = _u8(unknown); mean = _f64(mean); labels = _i32(labels); N, D = X.shape
        Xq = _f64(Xq); unknown_q = _u8(unknown_q); mean_q = _f64(mean_q); qlabels = _i32(qlabels)
        Nq = N if Xq is None else Xq.shape[0]
        recs = np.zeros((Nq, Nrec), np.int32)
        self.lib.orc_recommend_cluster(_p(X), _p(unknown), _p(mean), _p(labels), ctypes.c_int64(N), D, K, _p(Xq),
                                       _p(unknown_q), _p(mean_q), _p(qlabels), ctypes.c_int64(Nq), Nrec, _p(recs))
        return recs


class RecHandle:
    """Tables built once (rec A of main.cpp:155), then slices of the per-user loop (main.cpp:159-170);
    query() releases the GIL, so several threads can share one handle (cosine tables are read-only)."""

    def __init__(self, oracle, X, unknown, mean, metric, k, L, div, w, seed):
        self.o = oracle
        self.X = _f64(X); self.unknown = _u8(unknown); self.mean = _f64(mean)  # kept alive for the port
        N, D = self.X.shape
        self.h = ctypes.c_void_p(oracle.lib.orc_rec_handle_create(_p(self.X), _p(self.unknown), _p(self.mean), ctypes.c_int64(N), D,
                                                                  metric, k, L, div, ctypes.c_double(w), ctypes.c_uint64(seed)))

    def query(self, q_begin, q_end, P, Nrec):
        nq = q_end - q_begin
        recs = np.zeros((nq, Nrec), np.int32); ncand = np.zeros(nq, np.int32)
        self.o.lib.orc_rec_handle_query(self.h, ctypes.c_int64(q_begin), ctypes.c_int64(q_end), P, Nrec, _p(recs), _p(ncand))
        return recs, ncand

    def query_nbr(self, q_begin, q_end, P, Nrec):
        """query() plus the neighbours get_P_closest kept: (recs, ncand, nbr_rows[nq][P], nbr_sims[nq][P])."""
        nq = q_end - q_begin
        recs = np.zeros((nq, Nrec), np.int32); ncand = np.zeros(nq, np.int32)
        nidx = np.zeros((nq, P), np.int32); nsim = np.zeros((nq, P))
        self.o.lib.orc_rec_handle_query_nbr(self.h, ctypes.c_int64(q_begin), ctypes.c_int64(q_end), P, Nrec, _p(recs), _p(ncand), _p(nidx), _p(nsim))
        return recs, ncand, nidx, nsim

    def close(self):
        if self.h:
            self.o.lib.orc_rec_handle_destroy(self.h)
            self.h = None


_cache = {}


def load(kind="port"):
    """kind: 'port' (always available; built on demand) or 'reference' (None when no prebuilt/_ref)."""
    if kind in _cache:
        return _cache[kind]
    if kind == "port":
        if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(_HERE, "crx_oracle.cpp")):
            subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
        o = Oracle(PORT_SO)
    elif kind == "reference":
        if not os.path.exists(REF_SO) and os.path.isdir(os.path.join(REF_DIR, "lib")):
            subprocess.check_call(["make", "-C", _HERE, "ref", "CRX_REF_DIR=" + REF_DIR], stdout=subprocess.DEVNULL)
        o = Oracle(REF_SO) if os.path.exists(REF_SO) else None
    else:
        raise ValueError(kind)
    _cache[kind] = o
    return o
