"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Only two exchanges exist on this path:
  * k-means update: all-reduce(sum) of the per-cluster coordinate sums [K][D] and counts [K];
  * top-P with sharded CANDIDATES: all-gather of the per-shard (similarity, row) lists and a P-way merge.
Queries, projections and Lloyd assignment shard with no data-path collective.
"""
import os

import numpy as np


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
