"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (subject sharding, seed-row gather,
int64 all-reduce of partial centroid sums) that runs unchanged over NCCL on the GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_parallel():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "leida_parallel", os.path.join(ROOT, "leida-python_b200", "pyleida", "parallel.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _load_kmeans_host():
    """sweep_runs / _draw_inits / _dedupe of clustering/kmeans.py without importing the CUDA-backed package: the
    functions are pure numpy, so their source is executed in a scratch namespace."""
    import ast
    src = open(os.path.join(ROOT, "leida-python_b200", "pyleida", "clustering", "kmeans.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("sweep_runs", "_draw_inits", "_dedupe")]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "kmeans_host", "exec"), ns)
    return ns


def test_shard_bounds_cover_everything():
    P = _load_parallel()
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            b = [P.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert list(P.row_offsets([3, 0, 5])) == [0, 3, 3, 8]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = _load_parallel()
        comm = P.Comm()
        assert (comm.world, comm.rank) == (world, rank)
        # a row-sharded matrix: global row i holds the value i in every column
        M, N = 101, 6
        lo, hi = P.shard_bounds(M, world, rank)
        X = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, N)
        assert comm.all_gather_int(hi - lo) == [P.shard_bounds(M, world, r)[1] - P.shard_bounds(M, world, r)[0] for r in range(world)]
        seeds = np.array([100, 0, 50, 51, 7], dtype=np.int64)
        got = P.gather_seed_rows(X, N, seeds, lo, comm)
        assert torch.equal(got, torch.from_numpy(seeds).float()[:, None].repeat(1, N))
        # fixed-point partial sums: the int64 all-reduce is exact and order independent
        rng = np.random.default_rng(5)
        full = rng.standard_normal((M, N)).astype(np.float32)
        q40 = np.rint(full.astype(np.float64) * 2.0 ** 40).astype(np.int64)
        part = torch.from_numpy(q40[lo:hi].sum(0))
        comm.all_reduce_sum_(part)
        assert np.array_equal(part.numpy(), q40.sum(0))
        # rank 0's values reach every rank; max-reduce
        mine = np.arange(7, dtype=np.int64) * (rank + 1) + 100 * rank
        assert np.array_equal(comm.bcast0_int64(mine), np.arange(7, dtype=np.int64))
        assert float(comm.all_reduce_max_(torch.tensor([float(rank)], dtype=torch.float64)).item()) == world - 1
        assert comm.backend == "gloo"
        # k-means initial rows with random_state=None: every process has its own numpy global RNG, so the draws are
        # made on rank 0 and broadcast (ranks that seed different rows reduce unrelated centroids)
        K = _load_kmeans_host()
        np.random.seed(1234 + rank)                      # deliberately different streams
        runs, plan = K["sweep_runs"](40_000, 2, 5, 3, None, None, comm=comm)
        flat = np.concatenate([np.asarray(r[1], dtype=np.int64) for r in runs])
        ref = torch.from_numpy(flat.copy())
        dist.broadcast(ref, src=0)
        assert np.array_equal(flat, ref.numpy()), "ranks drew different initial rows"
        smp = torch.from_numpy(np.asarray(K["sweep_runs"].last_sample, dtype=np.int64).copy())
        ref = smp.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(smp, ref) and smp.numel() == 30_000
        comm.barrier()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
