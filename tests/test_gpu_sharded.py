"""SURVEY.md section 8e on one box: two processes (ranks) drive the sharded entry points of libcrx.so through
dist.Comm and must reproduce the single-GPU results -- bit for bit where the split does not change the arithmetic
(PAM, silhouette, range-search assignment), and up to the documented summation order for the k-means sums.  The two
ranks share GPU 0 and talk over gloo (device buffers staged through the host), so the test needs one GPU only; the
same code runs over NCCL with one GPU per rank (bench.py, tools/sharded_bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        os.environ.update(RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import torch.distributed as dist
        from crypto_recommendation_b200 import capi, synth
        from crypto_recommendation_b200 import dist as cdist
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
        ctx = capi.Context(0)
        comm = cdist.Comm()
        res = {}
        N, D, K = 6001, 24, 9
        X = synth.gaussian_mixture(N, D, 8, seed=3, dtype=np.float32)
        X64 = X.astype(np.float64)
        lo, hi = cdist.shard_range(N, rank, world)
        Pf = ctx.points(X)
        Pl = ctx.points(X[lo:hi])
        for metric in ("euclidean", "cosine"):
            # ---- k-means++ over sharded rows
            rows, vecs = capi.k_means_pp_sharded(ctx, Pl, lo, N, K, metric, 11, comm)
            single = capi.k_means_pp(ctx, Pf, K, metric, 11)
            res["kpp_" + metric] = bool(np.array_equal(rows, single) and np.array_equal(vecs, X64[single]))
            # ---- k-means over sharded rows
            lab, _ = capi.lloyds_assignment(ctx, Pf, X64[single], single, metric)
            cont_s, C_s = capi.k_means_sharded(ctx, Pl, lab[lo:hi], X64[single], metric, 0.05, comm)
            cont, C = capi.k_means(ctx, Pf, lab, X64[single], metric, 0.05)
            res["kmeans_" + metric] = bool(cont_s == cont and np.allclose(C_s, C, rtol=1e-12, atol=0))
            # ---- PAM / silhouette with split work (euclidean N >= 4096: tensor row sums; cosine: exact row sums)
            sw_s, new_s = capi.pam_lloyds(ctx, Pf, lab, single, metric, comm=comm)
            sw, new = capi.pam_lloyds(ctx, Pf, lab, single, metric)
            res["pam_" + metric] = bool(sw_s == sw and np.array_equal(new_s, new))
            s_s = capi.silhouette_cluster(ctx, Pf, lab, X64[single], metric, comm=comm)
            s_1 = capi.silhouette_cluster(ctx, Pf, lab, X64[single], metric)
            res["sil_" + metric] = bool(np.array_equal(s_s, s_1))
        # ---- range-search assignment with the centroids split (K >= 32 and a large remainder: tensor Lloyd pass, split by rows)
        for K2, tag in ((40, "tc"), (7, "exact")):
            cidx = capi.rand_selection(ctx, Pf, K2, 5)
            cube = capi.Hypercube(ctx, Pf, "euclidean", 8, 4.0, 9)
            a = capi.cube_range_assignment(ctx, Pf, cube, cidx, "euclidean", 12, comm=comm)
            b = capi.cube_range_assignment(ctx, Pf, cube, cidx, "euclidean", 12)
            res["cube_" + tag] = bool(all(np.array_equal(x, y) for x, y in zip(a, b)))
            t = capi.LshTables(ctx, Pf, "euclidean", 3, 4, 50, 4.0, 7)
            a = capi.lsh_range_assignment(ctx, Pf, t, cidx, "euclidean", comm=comm)
            b = capi.lsh_range_assignment(ctx, Pf, t, cidx, "euclidean")
            res["lsh_" + tag] = bool(all(np.array_equal(x, y) for x, y in zip(a, b)))
            res["assigned_by_range_" + tag] = int((b[2] >= 0).sum())
            cube.close(); t.close()
        tcos = capi.LshTables(ctx, Pf, "cosine", 3, 4, 50, 4.0, 7)
        cidx = capi.rand_selection(ctx, Pf, 12, 6)
        a = capi.lsh_range_assignment(ctx, Pf, tcos, cidx, "cosine", comm=comm)
        b = capi.lsh_range_assignment(ctx, Pf, tcos, cidx, "cosine")
        res["lsh_cosine"] = bool(all(np.array_equal(x, y) for x, y in zip(a, b)))
        res["comm_error"] = comm.error
        res["collectives"] = dict(comm.calls)
        dist.barrier()
        q.put((rank, res))
        dist.destroy_process_group()
    except Exception as e:  # surface the failure instead of a hang
        import traceback
        q.put((rank, {"exception": traceback.format_exc()}))


def _nccl_worker(rank, world, port, q):
    """One GPU per rank, collectives issued by libcrx.so itself (crx_comm_nccl_create): the sharded results must equal the
    single-GPU ones, and the communicator must have been used."""
    try:
        os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import torch
        import torch.distributed as dist
        from crypto_recommendation_b200 import capi, synth
        from crypto_recommendation_b200 import dist as cdist
        torch.cuda.set_device(rank)
        dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        ctx = capi.Context(rank)
        comm = cdist.NcclComm(ctx)
        res = {}
        N, D, K = 200_001, 32, 40
        X = synth.gaussian_mixture(N, D, 16, seed=3, dtype=np.float32)
        X64 = X.astype(np.float64)
        lo, hi = cdist.shard_range(N, rank, world)
        Pf = ctx.points(X)
        Pl = ctx.points(X[lo:hi])
        for metric in ("euclidean", "cosine"):
            rows, vecs = capi.k_means_pp_sharded(ctx, Pl, lo, N, K, metric, 11, comm)
            single = capi.k_means_pp(ctx, Pf, K, metric, 11)
            res["kpp_" + metric] = bool(np.array_equal(rows, single) and np.array_equal(vecs, X64[single]))
            lab, _ = capi.lloyds_assignment(ctx, Pf, X64[single], single, metric)
            cont_s, C_s = capi.k_means_sharded(ctx, Pl, lab[lo:hi], X64[single], metric, 0.05, comm)
            cont, C = capi.k_means(ctx, Pf, lab, X64[single], metric, 0.05)
            res["kmeans_" + metric] = bool(cont_s == cont and np.allclose(C_s, C, rtol=1e-12, atol=0))
            sw_s, new_s = capi.pam_lloyds(ctx, Pf, lab, single, metric, comm=comm)
            sw, new = capi.pam_lloyds(ctx, Pf, lab, single, metric)
            res["pam_" + metric] = bool(sw_s == sw and np.array_equal(new_s, new))
        cidx = capi.rand_selection(ctx, Pf, 40, 5)
        cube = capi.Hypercube(ctx, Pf, "euclidean", 8, 4.0, 9)
        a = capi.cube_range_assignment(ctx, Pf, cube, cidx, "euclidean", 12, comm=comm)
        b = capi.cube_range_assignment(ctx, Pf, cube, cidx, "euclidean", 12)
        res["cube"] = bool(all(np.array_equal(x, y) for x, y in zip(a, b)))
        res["collectives"] = comm.calls
        ctx.synchronize()
        dist.barrier()
        q.put((rank, res))
        comm.close()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, {"exception": traceback.format_exc()}))


def test_two_gpus_native_nccl_communicator():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL does not run two ranks on one device); tools/sharded_bench.py covers it under gpurun --gpus 2")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank in (0, 1):
        res = out[rank]
        assert "exception" not in res, res["exception"]
        bad = [k for k, v in res.items() if isinstance(v, bool) and not v]
        assert not bad, "rank %d: sharded over NCCL != single-GPU for %s" % (rank, bad)
        assert res["collectives"]["allreduce"] > 10 and res["collectives"]["allgather"] > 10


def test_two_ranks_reproduce_single_gpu_results():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank in (0, 1):
        res = out[rank]
        assert "exception" not in res, res["exception"]
        assert res["comm_error"] is None
        bad = [k for k, v in res.items() if isinstance(v, bool) and not v]
        assert not bad, "rank %d: sharded != single-GPU for %s" % (rank, bad)
        assert res["collectives"]["allreduce"] > 10 and res["collectives"]["allgather"] > 10
        assert res["assigned_by_range_tc"] > 0 and res["assigned_by_range_exact"] > 0
