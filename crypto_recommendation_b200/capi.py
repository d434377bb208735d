"""ctypes binding of include/crx.h (libcrx.so) and a thin numpy/torch host mirror of the reference's
free functions (same names, same argument meaning).

The library is the product path: there is no CPU fallback.  Loading fails loudly when libcrx.so has
not been built, and every compute call fails with CrxError when no sm_100a device is present.
Nothing in this module imports or calls oracle/.
"""
import ctypes
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrx.so")

F32, F64 = 0, 1
HOST, DEVICE = 0, 1
EUCLIDEAN, COSINE = 0, 1
METRICS = {"euclidean": EUCLIDEAN, "cosine": COSINE, EUCLIDEAN: EUCLIDEAN, COSINE: COSINE}

# every symbol include/crx.h declares (tests check the .so exports each of them)
SYMBOLS = [
    "crx_version", "crx_last_error", "crx_ctx_create", "crx_ctx_destroy", "crx_ctx_synchronize",
    "crx_ctx_launch_count", "crx_ctx_trim", "crx_ctx_profile", "crx_ctx_profile_reset", "crx_ctx_kernel_time", "crx_ctx_counters",
    "crx_points_create", "crx_points_set_ratings", "crx_points_destroy", "crx_points_n", "crx_points_d",
    "crx_pair_op", "crx_create_LSH_hashtables", "crx_lsh_destroy", "crx_lsh_bucket_ids", "crx_lsh_detailed_hashes",
    "crx_get_LSH_combined_buckets", "crx_lsh_params", "crx_create_hypercube", "crx_cube_destroy",
    "crx_cube_vertex_ids", "crx_get_hypercube_combined_buckets", "crx_get_num_hamming_dist_from",
    "crx_rand_selection", "crx_k_means_pp", "crx_lloyds_assignment", "crx_lloyds_for_remaining",
    "crx_columnar_write", "crx_columnar_open", "crx_columnar_n", "crx_columnar_d", "crx_columnar_id", "crx_columnar_points",
    "crx_columnar_rows", "crx_columnar_close",
    "crx_lsh_range_assignment", "crx_lsh_range_assignment_vectors", "crx_cube_range_assignment", "crx_cluster_sums", "crx_k_means_finish",
    "crx_k_means", "crx_pam_lloyds", "crx_silhouette_cluster", "crx_recommend_lsh", "crx_recommend_lsh_status", "crx_recommend_cluster",
    "crx_parallel_quickSort", "crx_parallel_quickSort_topn", "crx_get_P_closest", "crx_get_top_N_recom", "crx_lsh_hash_vector", "crx_lsh_hash_points",
    "crx_user_vectors_build",
    "crx_k_means_pp_sharded", "crx_k_means_sharded", "crx_pam_lloyds_sharded", "crx_silhouette_cluster_sharded",
    "crx_lsh_range_assignment_sharded", "crx_cube_range_assignment_sharded",
    "crx_comm_nccl_version", "crx_comm_nccl_unique_id", "crx_comm_nccl_create", "crx_comm_nccl_calls", "crx_comm_nccl_destroy",
]


class CrxError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libcrx.so (once).  Raises if it was never built -- there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CrxError("libcrx.so is missing: run `python -m crypto_recommendation_b200.build` "
                           "(this engine has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.crx_last_error.restype = ctypes.c_char_p
        L.crx_ctx_launch_count.restype = ctypes.c_int64
        L.crx_points_n.restype = ctypes.c_int64
        _lib = L
    return _lib


def _check(status):
    if status != 0:
        raise CrxError("crx error %d: %s" % (status, lib().crx_last_error().decode()))


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _ptr(a):
    """(pointer, mem) of a numpy array (HOST) or a CUDA torch tensor (DEVICE); None -> (NULL, HOST)."""
    if a is None:
        return None, HOST
    if _is_torch(a):
        assert a.is_contiguous()
        return ctypes.c_void_p(a.data_ptr()), (DEVICE if a.is_cuda else HOST)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p), HOST


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Context:
    """One GPU + one stream (crx_ctx).  stream: a raw cudaStream_t (int) or None for an own stream."""

    def __init__(self, device=0, stream=None):
        """stream=None: the context creates its own stream.  stream=0 (torch's default stream) selects the
        legacy default stream (cudaStreamLegacy, handle 0x1); any other int is a cudaStream_t."""
        self.h = ctypes.c_void_p()
        sp = None if stream is None else ctypes.c_void_p(stream if stream != 0 else 1)
        _check(lib().crx_ctx_create(int(device), sp, ctypes.byref(self.h)))
        self.device = device
        self._children = weakref.WeakSet()  # handles that must be destroyed before the context

    def close(self):
        if self.h:
            for ch in list(self._children):
                ch.close()
            lib().crx_ctx_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(lib().crx_ctx_synchronize(self.h))

    def launch_count(self):
        return int(lib().crx_ctx_launch_count(self.h))

    def trim(self):
        """give the memory the engine's stream-ordered pool holds back to the driver (between workloads of different shape)"""
        _check(lib().crx_ctx_trim(self.h))

    def profile(self, enable=True):
        _check(lib().crx_ctx_profile(self.h, int(enable)))

    def profile_reset(self):
        _check(lib().crx_ctx_profile_reset(self.h))

    def kernel_time(self, prefix):
        ms = ctypes.c_double()
        n = ctypes.c_int64()
        _check(lib().crx_ctx_kernel_time(self.h, prefix.encode(), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def counters(self, reset=False):
        out = (ctypes.c_int64 * 8)()
        _check(lib().crx_ctx_counters(self.h, out, int(reset)))
        v = list(out)
        return {"hash_dd": v[0], "topp_uncertified": v[1], "kpp_near": v[2], "pam_exact": v[3], "lloyd_exact": v[4],
                "topp_tie_order": v[5], "topp_tied": v[6], "topp_pass2": v[7]}

    # ---- vector<CustVector<T>> -------------------------------------------------------------
    def points(self, X, unknown=None, known_mean=None):
        return Points(self, X, unknown, known_mean)

    # ---- known-answer helpers ----------------------------------------------------------------
    def parallel_quickSort(self, sims, ids):
        s = _np(sims, np.float64).copy()
        d = _np(ids, np.int32).copy()
        _check(lib().crx_parallel_quickSort(self.h, _ptr(s)[0], _ptr(d)[0], len(s)))
        return s, d

    def parallel_quickSort_topn(self, sims, ids, need):
        s = _np(sims, np.float64).copy()
        d = _np(ids, np.int32).copy()
        _check(lib().crx_parallel_quickSort_topn(self.h, _ptr(s)[0], _ptr(d)[0], len(s), int(need)))
        return s[:need], d[:need]


def get_num_hamming_dist_from(num, dist, min_bit, bits):
    out = np.zeros(1 << 16, np.int32)
    n = lib().crx_get_num_hamming_dist_from(num, dist, min_bit, bits, _ptr(out)[0], len(out))
    return out[:n].tolist()


class Points:
    def __init__(self, ctx, X, unknown=None, known_mean=None):
        self.ctx = ctx
        self.h = ctypes.c_void_p()
        if _is_torch(X):
            import torch
            dt = {torch.float32: F32, torch.float64: F64}[X.dtype]
            n, d = X.shape
        else:
            if X.dtype not in (np.float32, np.float64):
                X = X.astype(np.float64)
            X = np.ascontiguousarray(X)
            dt = F32 if X.dtype == np.float32 else F64
            n, d = X.shape
        p, mem = _ptr(X)
        _check(lib().crx_points_create(ctx.h, p, dt, ctypes.c_int64(n), int(d), mem, ctypes.byref(self.h)))
        self.n, self.d = int(n), int(d)
        ctx._children.add(self)
        if unknown is not None:
            self.set_ratings(unknown, known_mean)

    def set_ratings(self, unknown, known_mean):
        if not _is_torch(unknown):
            unknown = _np(unknown, np.uint8)
            known_mean = _np(known_mean, np.float64)
        pu, mem = _ptr(unknown)
        pm, mem2 = _ptr(known_mean)
        assert mem == mem2
        _check(lib().crx_points_set_ratings(self.h, pu, pm, mem))

    def close(self):
        if self.h:
            lib().crx_points_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Columnar:
    """A binary columnar vector file (include/crx.h, "input ingest"): ids + float64 columns, opened with mmap."""

    def __init__(self, path):
        L = lib()
        L.crx_columnar_n.restype = ctypes.c_int64
        L.crx_columnar_id.restype = ctypes.c_char_p
        self.h = ctypes.c_void_p()
        _check(L.crx_columnar_open(os.fsencode(path), ctypes.byref(self.h)))
        self.n, self.d = int(L.crx_columnar_n(self.h)), int(L.crx_columnar_d(self.h))

    @staticmethod
    def write(path, ids, rows):
        rows = np.ascontiguousarray(rows, np.float64)
        arr = (ctypes.c_char_p * len(ids))(*[str(i).encode() for i in ids])
        _check(lib().crx_columnar_write(os.fsencode(path), arr, _ptr(rows)[0], ctypes.c_int64(rows.shape[0]), int(rows.shape[1])))

    def ids(self):
        return [lib().crx_columnar_id(self.h, ctypes.c_int64(i)).decode() for i in range(self.n)]

    def points(self, ctx):
        """The vectors as a point set on the GPU (upload of the columns + transpose kernel)."""
        P = Points.__new__(Points)
        P.ctx, P.h, P.n, P.d = ctx, ctypes.c_void_p(), self.n, self.d
        _check(lib().crx_columnar_points(ctx.h, self.h, ctypes.byref(P.h)))
        ctx._children.add(P)
        return P

    def rows(self, ctx):
        out = np.zeros((self.n, self.d))
        _check(lib().crx_columnar_rows(ctx.h, self.h, _ptr(out)[0]))
        return out

    def close(self):
        if self.h:
            lib().crx_columnar_close(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pair_op(ctx, pa, a, pb, b, op):
    a = _np(a, np.int32); b = _np(b, np.int32)
    out = np.zeros(len(a))
    _check(lib().crx_pair_op(ctx.h, pa.h, _ptr(a)[0], pb.h, _ptr(b)[0], ctypes.c_int64(len(a)), int(op), _ptr(out)[0]))
    return out


class LshTables:
    """create_LSH_hashtables (lsh_cube.hpp:45)."""

    def __init__(self, ctx, input_vectors, metric_type, k, L, lsh_bucket_div, euclidean_h_w, seed):
        self.ctx, self.pts = ctx, input_vectors
        self.metric, self.k, self.L = METRICS[metric_type], k, L
        self.h = ctypes.c_void_p()
        _check(lib().crx_create_LSH_hashtables(ctx.h, input_vectors.h, self.metric, k, L, int(lsh_bucket_div),
                                               ctypes.c_double(euclidean_h_w), ctypes.c_uint64(seed), ctypes.byref(self.h)))
        ctx._children.add(self)

    def close(self):
        if self.h:
            lib().crx_lsh_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bucket_ids(self):
        out = np.zeros((self.L, self.pts.n), np.int32)
        _check(lib().crx_lsh_bucket_ids(self.h, _ptr(out)[0], HOST))
        return out

    def detailed_hashes(self):
        out = np.zeros((self.L, self.pts.n, self.k), np.int32)
        _check(lib().crx_lsh_detailed_hashes(self.h, _ptr(out)[0], HOST))
        return out

    def combined_buckets(self, query_row, filtered):
        out = np.zeros(self.pts.n, np.int32)
        cnt = ctypes.c_int64()
        _check(lib().crx_get_LSH_combined_buckets(self.h, ctypes.c_int64(query_row), int(filtered), _ptr(out)[0],
                                                  ctypes.c_int64(len(out)), ctypes.byref(cnt)))
        return out[:cnt.value].copy()

    def hash_vector(self, x):
        """CustHashtable::getHash of a vector that is not stored: (bucket_ids[L], detailed[L][k] or None)."""
        x = _np(x, np.float64)
        ids = np.zeros(self.L, np.int32)
        det = np.zeros((self.L, self.k), np.int32) if self.metric == EUCLIDEAN else None
        _check(lib().crx_lsh_hash_vector(self.h, _ptr(x)[0], _ptr(ids)[0], _ptr(det)[0]))
        return ids, det

    def hash_points(self, pts):
        """getHash of every row of another point set: (bucket_ids[L][n], detailed[L][n][k] or None)."""
        ids = np.zeros((self.L, pts.n), np.int32)
        det = np.zeros((self.L, pts.n, self.k), np.int32) if self.metric == EUCLIDEAN else None
        _check(lib().crx_lsh_hash_points(self.h, pts.h, _ptr(ids)[0], None if det is None else _ptr(det)[0], HOST))
        return ids, det

    def params(self):
        H, D = self.L * self.k, self.pts.d
        if self.metric == COSINE:
            r = np.zeros((self.L, self.k, D))
            _check(lib().crx_lsh_params(self.h, _ptr(r)[0], None, None, None))
            return {"r": r}
        v = np.zeros((self.L, self.k, D), np.float32); t = np.zeros((self.L, self.k), np.float32)
        ri = np.zeros((self.L, self.k), np.int32)
        _check(lib().crx_lsh_params(self.h, None, _ptr(v)[0], _ptr(t)[0], _ptr(ri)[0]))
        return {"v": v, "t": t, "r": ri}


class Hypercube:
    """create_hypercube (lsh_cube.hpp:109)."""

    def __init__(self, ctx, input_vectors, metric_type, k, euclidean_h_w, seed):
        self.ctx, self.pts, self.k = ctx, input_vectors, k
        self.h = ctypes.c_void_p()
        _check(lib().crx_create_hypercube(ctx.h, input_vectors.h, METRICS[metric_type], k, ctypes.c_double(euclidean_h_w),
                                          ctypes.c_uint64(seed), ctypes.byref(self.h)))
        ctx._children.add(self)

    def close(self):
        if self.h:
            lib().crx_cube_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def vertex_ids(self):
        out = np.zeros(self.pts.n, np.int32)
        _check(lib().crx_cube_vertex_ids(self.h, _ptr(out)[0], HOST))
        return out

    def combined_buckets(self, query_row, probes):
        out = np.zeros(self.pts.n, np.int32)
        cnt = ctypes.c_int64()
        _check(lib().crx_get_hypercube_combined_buckets(self.h, ctypes.c_int64(query_row), int(probes), _ptr(out)[0],
                                                        ctypes.c_int64(len(out)), ctypes.byref(cnt)))
        return out[:cnt.value].copy()


# ---- clustering phases (initialization.hpp / assignment.hpp / update.hpp / silhouette.hpp) --------
def rand_selection(ctx, pts, cluster_num, seed):
    out = np.zeros(cluster_num, np.int32)
    _check(lib().crx_rand_selection(ctx.h, pts.h, cluster_num, ctypes.c_uint64(seed), _ptr(out)[0]))
    return out


def k_means_pp(ctx, pts, cluster_num, metric_type, seed):
    out = np.zeros(cluster_num, np.int32)
    _check(lib().crx_k_means_pp(ctx.h, pts.h, cluster_num, METRICS[metric_type], ctypes.c_uint64(seed), _ptr(out)[0]))
    return out


def _comm_ptr(comm):
    return comm.ptr() if comm is not None else None


def k_means_pp_sharded(ctx, local_pts, row_offset, n_global, cluster_num, metric_type, seed, comm):
    """k_means_pp over rows sharded by contiguous range (crx_k_means_pp_sharded): (global rows[K], vectors[K][D])."""
    rows = np.zeros(cluster_num, np.int64)
    vecs = np.zeros((cluster_num, local_pts.d))
    _check(lib().crx_k_means_pp_sharded(ctx.h, local_pts.h, ctypes.c_int64(row_offset), ctypes.c_int64(n_global), cluster_num,
                                        METRICS[metric_type], ctypes.c_uint64(seed), _comm_ptr(comm), _ptr(rows)[0], _ptr(vecs)[0]))
    return rows, vecs


def k_means_sharded(ctx, local_pts, labels, centroids, metric_type, min_dist, comm):
    """k_means with the rows sharded: local sums, all-reduce, redundant finish.  labels / centroids: numpy or torch."""
    K = centroids.shape[0]
    pl, lmem = _ptr(labels if _is_torch(labels) else _np(labels, np.int32))
    if not _is_torch(centroids):
        centroids = _np(centroids, np.float64)
    pc, cmem = _ptr(centroids)
    if _is_torch(centroids):
        import torch
        newc = torch.empty_like(centroids)
    else:
        newc = np.zeros_like(centroids)
    cont = ctypes.c_int()
    _check(lib().crx_k_means_sharded(ctx.h, local_pts.h, pl, lmem, pc, K, METRICS[metric_type], ctypes.c_double(min_dist),
                                     _comm_ptr(comm), _ptr(newc)[0], cmem, ctypes.byref(cont)))
    return bool(cont.value), newc


def _out_pair(pts, labels, dists):
    if labels is None:
        labels = np.zeros(pts.n, np.int32)
        dists = np.zeros(pts.n, np.float64)
    return labels, dists


def lloyds_assignment(ctx, pts, centroids, centroid_rows, metric_type, labels=None, dists=None, want_dists=True):
    """centroids: [K][D] float64 (numpy or CUDA tensor).  labels/dists: optional preallocated outputs
    (numpy -> host, CUDA tensors -> written in place on the device).  want_dists=False: labels only (dists = None)."""
    if not _is_torch(centroids):
        centroids = _np(centroids, np.float64)
    K = centroids.shape[0]
    cr = None if centroid_rows is None else _np(centroid_rows, np.int32)
    if want_dists:
        labels, dists = _out_pair(pts, labels, dists)
    else:
        dists = None
        if labels is None:
            labels = np.zeros(pts.n, np.int32)
    pc, cmem = _ptr(centroids)
    pl, mem = _ptr(labels)
    _check(lib().crx_lloyds_assignment(ctx.h, pts.h, pc, cmem, K, _ptr(cr)[0], METRICS[metric_type], pl, _ptr(dists)[0], mem))
    return labels, dists


def lloyds_for_remaining(ctx, pts, centroids, metric_type, labels, dists):
    if not _is_torch(centroids):
        centroids = _np(centroids, np.float64)
    pc, cmem = _ptr(centroids)
    pl, mem = _ptr(labels)
    _check(lib().crx_lloyds_for_remaining(ctx.h, pts.h, pc, cmem, centroids.shape[0], METRICS[metric_type], pl, _ptr(dists)[0], mem))
    return labels, dists


def _range_outputs(pts, out):
    """(labels, dists, before): fresh host arrays, or the caller's (all numpy or all CUDA tensors -> results stay on the device)."""
    if out is not None:
        return out
    return np.zeros(pts.n, np.int32), np.zeros(pts.n), np.zeros(pts.n, np.int32)


def lsh_range_assignment(ctx, pts, tables, centroid_rows, metric_type, comm=None, out=None):
    """comm: dist.Comm -- the centroids (and the Lloyd pass over the remainder) are split over its ranks.
    out: optional preallocated (labels int32, dists float64, labels_before_lloyd int32)."""
    cr = _np(centroid_rows, np.int32)
    labels, dists, before = _range_outputs(pts, out)
    pl, mem = _ptr(labels)
    _check(lib().crx_lsh_range_assignment_sharded(ctx.h, pts.h, tables.h, _ptr(cr)[0], len(cr), METRICS[metric_type], _comm_ptr(comm),
                                                  pl, _ptr(dists)[0], mem, _ptr(before)[0]))
    return labels, dists, before


def lsh_range_assignment_vectors(ctx, pts, tables, centroids, metric_type, shared_ids=False, centroid_rows=None, out=None):
    """lsh_range_assignment for centroids that are any vectors (the k_means centres of the second and later iterations):
    centroids [K][D] float64 on the host.  shared_ids: all centroid ids equal ("k_means_center")."""
    C = np.ascontiguousarray(centroids, np.float64)
    cr = None if centroid_rows is None else _np(centroid_rows, np.int32)
    labels, dists, before = _range_outputs(pts, out)
    pl, mem = _ptr(labels)
    _check(lib().crx_lsh_range_assignment_vectors(ctx.h, pts.h, tables.h, _ptr(C)[0], None if cr is None else _ptr(cr)[0], C.shape[0],
                                                  METRICS[metric_type], int(bool(shared_ids)), pl, _ptr(dists)[0], mem, _ptr(before)[0]))
    return labels, dists, before


def cube_range_assignment(ctx, pts, cube, centroid_rows, metric_type, probes, comm=None, out=None):
    cr = _np(centroid_rows, np.int32)
    labels, dists, before = _range_outputs(pts, out)
    pl, mem = _ptr(labels)
    _check(lib().crx_cube_range_assignment_sharded(ctx.h, pts.h, cube.h, _ptr(cr)[0], len(cr), METRICS[metric_type], int(probes),
                                                   _comm_ptr(comm), pl, _ptr(dists)[0], mem, _ptr(before)[0]))
    return labels, dists, before


def cluster_sums(ctx, pts, labels, K, sums=None, counts=None):
    if sums is None:
        sums = np.zeros((K, pts.d)); counts = np.zeros(K, np.int64)
    pl, lmem = _ptr(labels if _is_torch(labels) else _np(labels, np.int32))
    ps, mem = _ptr(sums)
    _check(lib().crx_cluster_sums(ctx.h, pts.h, pl, lmem, K, ps, _ptr(counts)[0], mem))
    return sums, counts


def k_means_finish(ctx, sums, counts, old_centroids, metric_type, min_dist, new_centroids=None):
    K, D = old_centroids.shape
    if new_centroids is None:
        new_centroids = np.zeros((K, D))
    cont = ctypes.c_int()
    ps, mem = _ptr(sums)
    _check(lib().crx_k_means_finish(ctx.h, ps, _ptr(counts)[0], _ptr(old_centroids)[0], K, D, METRICS[metric_type],
                                    ctypes.c_double(min_dist), _ptr(new_centroids)[0], mem, ctypes.byref(cont)))
    return bool(cont.value), new_centroids


def k_means(ctx, pts, labels, centroids, metric_type, min_dist):
    """k_means (update.hpp:38): returns (continue_clustering, centres after the call)."""
    centroids = _np(centroids, np.float64)
    labels = _np(labels, np.int32)
    newc = np.zeros_like(centroids)
    cont = ctypes.c_int()
    _check(lib().crx_k_means(ctx.h, pts.h, _ptr(labels)[0], HOST, _ptr(centroids)[0], centroids.shape[0], METRICS[metric_type],
                             ctypes.c_double(min_dist), _ptr(newc)[0], HOST, ctypes.byref(cont)))
    return bool(cont.value), newc


def pam_lloyds(ctx, pts, labels, centroid_rows, metric_type, comm=None):
    """comm: dist.Comm -- the candidate medoid rows are split over its ranks (points replicated)."""
    if not _is_torch(labels):
        labels = _np(labels, np.int32)
    cr = _np(centroid_rows, np.int32)
    new = np.zeros(len(cr), np.int32)
    sw = ctypes.c_int()
    pl, lmem = _ptr(labels)
    _check(lib().crx_pam_lloyds_sharded(ctx.h, pts.h, pl, lmem, _ptr(cr)[0], len(cr), METRICS[metric_type], _comm_ptr(comm),
                                        _ptr(new)[0], ctypes.byref(sw)))
    return bool(sw.value), new


def silhouette_cluster(ctx, pts, labels, centroids, metric_type, comm=None):
    labels = _np(labels, np.int32); centroids = _np(centroids, np.float64)
    K = centroids.shape[0]
    s = np.zeros(K + 1)
    _check(lib().crx_silhouette_cluster_sharded(ctx.h, pts.h, _ptr(labels)[0], HOST, _ptr(centroids)[0], HOST, K, METRICS[metric_type],
                                                _comm_ptr(comm), _ptr(s)[0]))
    return s


# ---- recommendation (crypto_rec.hpp) ------------------------------------------------------------------
Q_EXACT, Q_PLATEAU, Q_TIE_ORDER = 0, 1, 2


def recommend_lsh(ctx, tables, P, Nrec, queries=None, q_begin=0, q_end=None, out=None, want=("recs", "nbr_rows", "nbr_sims", "ncand")):
    """The per-user loop of main.cpp:159-170 (queries=None) / 205-216 (queries = other users).
    out: optional dict of preallocated outputs (numpy or CUDA tensors, all in the same memory).
    "status" in want (or in out): per-query Q_EXACT / Q_PLATEAU / Q_TIE_ORDER through crx_recommend_lsh_status."""
    nqtot = tables.pts.n if queries is None else queries.n
    if q_end is None:
        q_end = nqtot
    nq = q_end - q_begin
    if out is None:
        out = {}
        if "recs" in want: out["recs"] = np.zeros((nq, Nrec), np.int32)
        if "nbr_rows" in want: out["nbr_rows"] = np.zeros((nq, P), np.int32)
        if "nbr_sims" in want: out["nbr_sims"] = np.zeros((nq, P))
        if "ncand" in want: out["ncand"] = np.zeros(nq, np.int32)
        if "status" in want: out["status"] = np.zeros(nq, np.int32)
    mems = {_ptr(v)[1] for v in out.values()}
    assert len(mems) == 1
    if "status" in out:
        _check(lib().crx_recommend_lsh_status(ctx.h, tables.h, queries.h if queries is not None else None, ctypes.c_int64(q_begin),
                                              ctypes.c_int64(q_end), int(P), int(Nrec), _ptr(out.get("recs"))[0],
                                              _ptr(out.get("nbr_rows"))[0], _ptr(out.get("nbr_sims"))[0], _ptr(out.get("ncand"))[0],
                                              _ptr(out["status"])[0], mems.pop()))
        return out
    _check(lib().crx_recommend_lsh(ctx.h, tables.h, queries.h if queries is not None else None, ctypes.c_int64(q_begin),
                                   ctypes.c_int64(q_end), int(P), int(Nrec), _ptr(out.get("recs"))[0], _ptr(out.get("nbr_rows"))[0],
                                   _ptr(out.get("nbr_sims"))[0], _ptr(out.get("ncand"))[0], mems.pop()))
    return out


def recommend_cluster(ctx, users, labels, K, Nrec, queries=None, qlabels=None):
    labels = _np(labels, np.int32)
    nq = users.n if queries is None else queries.n
    ql = None if qlabels is None else _np(qlabels, np.int32)
    recs = np.zeros((nq, Nrec), np.int32)
    _check(lib().crx_recommend_cluster(ctx.h, users.h, _ptr(labels)[0], HOST, K, queries.h if queries is not None else None,
                                       _ptr(ql)[0], int(Nrec), _ptr(recs)[0], HOST))
    return recs


def get_P_closest(ctx, users, neighbor_rows, query_set, query_row, P):
    """get_P_closest (crypto_rec.hpp:214): returns (kept rows, similarities)."""
    nb = _np(neighbor_rows, np.int32).copy()
    sims = np.zeros(max(1, min(len(nb), P)))
    kept = ctypes.c_int64()
    _check(lib().crx_get_P_closest(ctx.h, users.h, _ptr(nb)[0], ctypes.c_int64(len(nb)), query_set.h, ctypes.c_int64(query_row), int(P),
                                   _ptr(sims)[0], ctypes.byref(kept)))
    return nb[:kept.value], sims[:kept.value]


def get_top_N_recom(ctx, users, neighbor_rows, similarities, query_set, query_row, N):
    """get_top_N_recom (crypto_rec.hpp:310 / :328 when similarities is None): returns (recs[N], predicted[D])."""
    nb = _np(neighbor_rows, np.int32)
    sm = None if similarities is None else _np(similarities, np.float64)
    recs = np.zeros(N, np.int32); pred = np.zeros(users.d)
    _check(lib().crx_get_top_N_recom(ctx.h, users.h, _ptr(nb)[0], _ptr(sm)[0], ctypes.c_int64(len(nb)), query_set.h,
                                     ctypes.c_int64(query_row), int(N), _ptr(pred)[0], _ptr(recs)[0]))
    return recs, pred


def user_vectors_build(ctx, mention_user, mention_coin, mention_score, n_users, n_coins):
    """tweets_to_user_vectors / clusters_to_user_vectors (crypto_rec.hpp:79-210) after string resolution.
    Returns (X[n_users][n_coins], unknown u8, known_mean, keep u8); rows with keep == 0 are the users the reference drops."""
    mu = _np(mention_user, np.int32); mc = _np(mention_coin, np.int32); ms = _np(mention_score, np.float64)
    assert len(mu) == len(mc) == len(ms)
    X = np.zeros((n_users, n_coins)); unk = np.zeros((n_users, n_coins), np.uint8)
    mean = np.zeros(n_users); keep = np.zeros(n_users, np.uint8)
    _check(lib().crx_user_vectors_build(ctx.h, _ptr(mu)[0], _ptr(mc)[0], _ptr(ms)[0], ctypes.c_int64(len(mu)), ctypes.c_int64(n_users),
                                        int(n_coins), _ptr(X)[0], _ptr(unk)[0], _ptr(mean)[0], _ptr(keep)[0], HOST))
    return X, unk, mean, keep
