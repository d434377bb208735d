"""CPU: the N>1 host logic under gloo, world_size 2 (sharding, the k-means all-reduce, the top-P merge)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crypto_recommendation_b200 import dist as cdist


def test_shard_range_partitions():
    for n in (0, 1, 7, 100, 1000003):
        for world in (1, 2, 3, 8):
            r = [cdist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_merge_topP_matches_global_sort():
    rng = np.random.default_rng(0)
    nq, P, world, n = 50, 20, 4, 400
    sims_all = rng.random((nq, n))
    sims_all[:, ::7] = 0.5  # ties
    shards = np.array_split(np.arange(n), world)
    S, R = [], []
    for sh in shards:
        s = sims_all[:, sh]
        order = np.lexsort((np.broadcast_to(sh, s.shape), -s), axis=1)[:, :P]
        S.append(np.take_along_axis(s, order, axis=1)); R.append(sh[order])
    ms, mr = cdist.merge_topP(S, R, P)
    order = np.lexsort((np.broadcast_to(np.arange(n), sims_all.shape), -sims_all), axis=1)[:, :P]
    assert np.array_equal(mr, order) and np.array_equal(ms, np.take_along_axis(sims_all, order, axis=1))
    # padded lists (-1) never win
    S2 = [np.zeros((3, 4)), np.ones((3, 4))]; R2 = [np.full((3, 4), -1), np.arange(12).reshape(3, 4)]
    ms, mr = cdist.merge_topP(S2, R2, 4)
    assert (mr >= 0).all()


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    cdist.init_process_group("gloo")
    rng = np.random.default_rng(1)
    N, D, K = 1001, 16, 5
    X = rng.normal(size=(N, D)); lab = rng.integers(0, K, N)
    lo, hi = cdist.shard_range(N, rank, world)
    sums = np.zeros((K, D)); counts = np.zeros(K, np.int64)
    for v in range(lo, hi):  # stand-in for crx_cluster_sums on this rank's rows
        sums[lab[v]] += X[v]; counts[lab[v]] += 1
    ts, tc = torch.from_numpy(sums), torch.from_numpy(counts)
    cdist.allreduce_cluster_sums(ts, tc)
    full = np.zeros((K, D)); fc = np.zeros(K, np.int64)
    for v in range(N):
        full[lab[v]] += X[v]; fc[lab[v]] += 1
    ok1 = np.allclose(ts.numpy(), full, rtol=1e-13) and np.array_equal(tc.numpy(), fc)
    # top-P with sharded candidates
    nq, P = 30, 8
    sims_all = rng.random((nq, N))
    s = sims_all[:, lo:hi]; rows = np.arange(lo, hi)
    order = np.lexsort((np.broadcast_to(rows, s.shape), -s), axis=1)[:, :P]
    ms, mr = cdist.allgather_topP(torch.from_numpy(np.take_along_axis(s, order, axis=1)), torch.from_numpy(rows[order]), P)
    want = np.lexsort((np.broadcast_to(np.arange(N), sims_all.shape), -sims_all), axis=1)[:, :P]
    ok2 = np.array_equal(mr, want)
    ok3 = cdist.max_over_ranks(float(rank + 1)) == float(world)
    # the crx_comm callbacks exactly as libcrx.so invokes them (host buffers): all-reduce sum/max/min, all-gather, broadcast
    import ctypes
    comm = cdist.Comm()
    cs = comm.struct
    ok4 = comm.world == world and comm.rank == rank
    a = np.arange(6, dtype=np.float64) * (rank + 1)
    ok4 &= cs.allreduce(None, a.ctypes.data, 6, 1, 0, 0) == 0 and np.array_equal(a, np.arange(6) * 3.0)
    m = np.array([rank, 10 - rank, 5], np.int32)
    ok4 &= cs.allreduce(None, m.ctypes.data, 3, 2, 1, 0) == 0 and m.tolist() == [1, 10, 5]
    m = np.array([rank, 10 - rank, 5], np.int32)
    ok4 &= cs.allreduce(None, m.ctypes.data, 3, 2, 2, 0) == 0 and m.tolist() == [0, 9, 5]
    send = np.array([100 + rank, 200 + rank], np.int64); recv = np.zeros(4, np.int64)
    ok4 &= cs.allgather(None, send.ctypes.data, recv.ctypes.data, 2, 3, 0) == 0 and recv.tolist() == [100, 200, 101, 201]
    b = np.full(4, float(rank + 7))
    ok4 &= cs.broadcast(None, b.ctypes.data, 4, 1, 1, 0) == 0 and (b == 8.0).all()
    ok4 &= comm.error is None and comm.calls == {"allreduce": 3, "allgather": 1, "broadcast": 1}
    cdist.barrier()
    q.put((rank, ok1, ok2, ok3 and bool(ok4)))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res
