"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Exchanges on this path:
  * k-means update: all-reduce(sum) of the per-cluster coordinate sums [K][D] and counts [K];
  * k-means++ over sharded rows: all-reduce(max), all-gather of shard totals, broadcast of the pick, per round;
  * PAM / silhouette / range-search assignment with split work: one all-reduce of an N-long device array;
  * top-P with sharded CANDIDATES: all-gather of the per-shard (similarity, row) lists and a P-way merge.
Queries, projections and Lloyd assignment shard with no data-path collective.

`Comm` is the crx_comm of include/crx.h: the engine (libcrx.so) drives the sharded algorithms and calls back into
this module for the collectives, which run on torch.distributed (NCCL for device buffers; with the gloo backend
device buffers are staged through the host).
"""
import ctypes
import os

import numpy as np

_ALLREDUCE = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int)
_ALLGATHER = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int)
_BROADCAST = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int)


class CrxComm(ctypes.Structure):
    """struct crx_comm (include/crx.h)."""
    _fields_ = [("rank", ctypes.c_int), ("world", ctypes.c_int), ("user", ctypes.c_void_p),
                ("allreduce", _ALLREDUCE), ("allgather", _ALLGATHER), ("broadcast", _BROADCAST), ("stream_ordered", ctypes.c_int)]


_NP_DTYPES = {0: np.float32, 1: np.float64, 2: np.int32, 3: np.int64}   # CRX_F32, CRX_F64, CRX_I32, CRX_I64


class _DevView:
    """A raw device pointer as something torch.as_tensor understands (no copy)."""
    def __init__(self, ptr, count, dt):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": np.dtype(dt).str, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class Comm:
    """crx_comm over torch.distributed.  Keep the object alive while the engine may call it (callbacks are owned here)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.on = dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1
        self.backend = dist.get_backend(group) if self.on else None
        self.calls = {"allreduce": 0, "allgather": 0, "broadcast": 0}
        self.error = None
        self._cbs = (_ALLREDUCE(self._allreduce), _ALLGATHER(self._allgather), _BROADCAST(self._broadcast))
        self.struct = CrxComm(self.rank, self.world, None, *self._cbs, 0)

    def ptr(self):
        return ctypes.byref(self.struct)

    # a (tensor the collective can run on, write-back function) pair for a raw buffer
    def _tensor(self, ptr, count, dtype, mem):
        import torch
        dt = _NP_DTYPES[dtype]
        if mem == 0:  # host
            arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(int(count),))
            t = torch.from_numpy(arr)
            if self.backend == "nccl":
                d = t.cuda()
                return d, lambda: t.copy_(d.cpu())
            return t, lambda: None
        t = torch.as_tensor(_DevView(ptr, count, dt), device=torch.device("cuda", torch.cuda.current_device()))
        if self.backend != "nccl":   # gloo: stage through the host
            h = t.cpu()
            return h, lambda: t.copy_(h)
        return t, lambda: None

    def _finish(self, mem):
        if mem == 1 or self.backend == "nccl":
            import torch
            torch.cuda.synchronize()

    def _guard(self, name, fn):
        try:
            self.calls[name] += 1
            fn()
            return 0
        except Exception as e:  # the engine reports CRX_ERR_COMM; the text is kept here
            self.error = "%s: %r" % (name, e)
            return 1

    def _allreduce(self, user, buf, count, dtype, op, mem):
        def run():
            t, back = self._tensor(buf, count, dtype, mem)
            ops = {0: self.dist.ReduceOp.SUM, 1: self.dist.ReduceOp.MAX, 2: self.dist.ReduceOp.MIN}
            self.dist.all_reduce(t, op=ops[op], group=self.group)
            self._finish(mem)
            back()
            self._finish(mem)
        return self._guard("allreduce", run)

    def _allgather(self, user, send, recv, count, dtype, mem):
        def run():
            import torch
            s, _ = self._tensor(send, count, dtype, mem)
            r, back = self._tensor(recv, count * self.world, dtype, mem)
            parts = [torch.empty_like(s) for _ in range(self.world)]
            self.dist.all_gather(parts, s.clone(), group=self.group)
            r.copy_(torch.cat(parts))
            self._finish(mem)
            back()
            self._finish(mem)
        return self._guard("allgather", run)

    def _broadcast(self, user, buf, count, dtype, root, mem):
        def run():
            t, back = self._tensor(buf, count, dtype, mem)
            src = root if self.group is None else self.dist.get_global_rank(self.group, root)
            self.dist.broadcast(t, src=src, group=self.group)
            self._finish(mem)
            back()
            self._finish(mem)
        return self._guard("broadcast", run)


class NcclComm:
    """crx_comm over NCCL inside libcrx.so (crx_comm_nccl_create, csrc/comm_nccl.cu): the collectives of the sharded entry
    points are enqueued on the context's stream by the engine itself -- no Python on the data path, no host synchronisation
    around device buffers.  torch.distributed only carries the 128-byte unique id from rank 0 to the others."""

    def __init__(self, ctx, group=None):
        import torch
        import torch.distributed as dist
        from . import capi
        self.lib = capi.lib()
        self.on = dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1
        ident = (ctypes.c_uint8 * 128)()
        if self.rank == 0:
            capi._check(self.lib.crx_comm_nccl_unique_id(ident))
        if self.on:
            t = torch.tensor(list(ident), dtype=torch.uint8)
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0 if group is None else dist.get_global_rank(group, 0), group=group)
            ident = (ctypes.c_uint8 * 128)(*t.cpu().tolist())
        self.h = ctypes.POINTER(CrxComm)()
        capi._check(self.lib.crx_comm_nccl_create(ctx.h, ident, self.rank, self.world, ctypes.byref(self.h)))
        self.ctx = ctx   # the communicator enqueues on this context's stream: keep it alive

    def ptr(self):
        return self.h

    _CODES = {"torch.float32": 0, "torch.float64": 1, "torch.int32": 2, "torch.int64": 3}

    def allgather(self, send, recv):
        """all-gather of CUDA tensors on the context's stream (recv holds world * send.numel() elements, rank order)"""
        assert send.is_cuda and recv.is_cuda and send.is_contiguous() and recv.is_contiguous() and recv.numel() == self.world * send.numel()
        c = self.h.contents
        rc = c.allgather(c.user, ctypes.c_void_p(send.data_ptr()), ctypes.c_void_p(recv.data_ptr()), send.numel(), self._CODES[str(send.dtype)], 1)
        if rc != 0:
            raise RuntimeError("ncclAllGather failed: %s" % self.lib.crx_last_error().decode())

    def allreduce(self, t, op="sum"):
        """in-place all-reduce of a CUDA tensor on the context's stream"""
        assert t.is_cuda and t.is_contiguous()
        c = self.h.contents
        rc = c.allreduce(c.user, ctypes.c_void_p(t.data_ptr()), t.numel(), self._CODES[str(t.dtype)], {"sum": 0, "max": 1, "min": 2}[op], 1)
        if rc != 0:
            raise RuntimeError("ncclAllReduce failed: %s" % self.lib.crx_last_error().decode())

    @property
    def calls(self):
        out = (ctypes.c_int64 * 3)()
        self.lib.crx_comm_nccl_calls(self.h, out)
        return {"allreduce": out[0], "allgather": out[1], "broadcast": out[2]}

    def close(self):
        if self.h:
            self.lib.crx_comm_nccl_destroy(self.h)
            self.h = ctypes.POINTER(CrxComm)()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_range(n, rank, world):
    """Contiguous row range [lo, hi) of shard `rank` (first n % world shards get one extra row)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend=None):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def allreduce_cluster_sums(sums, counts):
    """In-place sum over ranks of the k_means partial sums (torch tensors, any device)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return sums, counts


def merge_topP(sims, rows, P):
    """P-way merge of per-shard descending lists.  sims/rows: [world][nq][P] (rows global, -1 padded).
    Order: similarity descending, then row ascending (the engine's tie rule)."""
    sims = np.concatenate(list(sims), axis=1)
    rows = np.concatenate(list(rows), axis=1)
    key_s = np.where(rows >= 0, sims, -np.inf)
    order = np.lexsort((rows, -key_s), axis=1)[:, :P]
    out_s = np.take_along_axis(sims, order, axis=1)
    out_r = np.take_along_axis(rows, order, axis=1)
    out_s = np.where(out_r >= 0, out_s, 0.0)
    return out_s, out_r


def allgather_topP(sims, rows, P):
    """sims/rows: this rank's [nq][P] torch tensors; returns the merged global top-P (numpy)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return merge_topP([sims.cpu().numpy()], [rows.cpu().numpy()], P)
    world = dist.get_world_size()
    gs = [torch.empty_like(sims) for _ in range(world)]
    gr = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(gs, sims)
    dist.all_gather(gr, rows)
    return merge_topP([t.cpu().numpy() for t in gs], [t.cpu().numpy() for t in gr], P)


def max_over_ranks(value):
    """max of a python float over ranks (device timing rule: report the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
