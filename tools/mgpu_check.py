"""Shard-count invariance on real GPUs: run under torchrun with N ranks.

Every rank owns a block of subjects and runs the sharded pipeline; then EVERY rank also recomputes the whole job alone
(world-1 communicator, same device) and compares its own shard of the sharded results with the matching rows of the
single-GPU results, bit for bit: labels of every K, centroids, inertia, and the three metric tables matched on
(subject_id, condition).  Covers explicit init rows and the random_state=None path (initial rows drawn on rank 0 and
broadcast).  Exit code 0 only if every rank agrees.

    torchrun --nproc-per-node N tools/mgpu_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import pyleida  # noqa: E402
from pyleida import synthetic as syn  # noqa: E402
from pyleida.parallel import Comm, shard_bounds  # noqa: E402


class Solo(Comm):
    """A world-1 communicator inside a multi-rank job."""

    def __init__(self):
        self.world, self.rank, self.group = 1, 0, None

    @property
    def backend(self):
        return "none"

    def all_reduce_sum_(self, t):
        return t

    def all_reduce_max_(self, t):
        return t

    def bcast0_int64(self, a):
        return a

    def all_gather_int(self, v):
        return [int(v)]

    def barrier(self):
        pass


def table_rows(tab, ids):
    t = tab[tab.subject_id.isin(ids)]
    return t.sort_values(["condition", "subject_id"]).iloc[:, 2:].values


def check(S, N, T, K_max, n_init, explicit, extras=False):
    lo, hi = shard_bounds(S, world, rank)
    ts, classes = syn.make_shard(S, N, T, lo, hi, "states")
    M = S * (T - 2)
    inits = syn.init_indices(M, 2, K_max, n_init, seed=1) if explicit else None
    if not explicit:
        np.random.seed(100 + rank)                      # different global RNG state on every rank, on purpose
    # extras: quality scores (the scored sample is gathered) and permutation statistics (the tables are gathered)
    kw = dict(scores=extras, stats=extras, n_perm=300, random_state=5 if (extras and explicit) else None)
    if extras:
        np.random.seed(77)                              # the 30 000-row score sample is drawn from the global stream (rank 0's)
    m = pyleida.Leida.from_arrays(ts, classes, comm=Comm())
    m.fit_predict(TR=0.72, K_min=2, K_max=K_max, n_init=n_init, init_indices=inits, **kw)
    if not explicit:
        # the rows rank 0 drew: replay its stream for the single-GPU run
        np.random.seed(100)
    if extras:
        np.random.seed(77)
    ts_all, cl_all = syn.make_shard(S, N, T, 0, S, "states")
    ref = pyleida.Leida.from_arrays(ts_all, cl_all, comm=Solo())
    ref.fit_predict(TR=0.72, K_min=2, K_max=K_max, n_init=n_init, init_indices=inits, **kw)
    r0, r1 = lo * (T - 2), hi * (T - 2)
    ids = list(ts.keys())
    ok = True
    for k in range(2, K_max + 1):
        ok &= bool(np.array_equal(m.predictions[f"k_{k}"].values, ref.predictions[f"k_{k}"].values[r0:r1]))
        ok &= bool(np.array_equal(m.load_model(k).cluster_centers_, ref.load_model(k).cluster_centers_))
        ok &= m.load_model(k).inertia_ == ref.load_model(k).inertia_
        for name in ("occupancies", "dwell_times", "transitions"):
            a = table_rows(getattr(m, name)(k), ids)
            b = table_rows(getattr(ref, name)(k), ids)
            ok &= a.shape == b.shape and bool(np.array_equal(a, b))
    if extras:
        pa, pb = m._clustering_.performance, ref._clustering_.performance
        for col in ("Dunn_score", "silhouette_score", "Davies-Bouldin_score", "distortion"):
            ok &= bool(np.allclose(pa[col].values, pb[col].values, rtol=1e-12, atol=0, equal_nan=False))
        for k in range(2, K_max + 1):
            for metric in ("occupancies", "dwell_times"):
                a, b = m.stats(k, metric), ref.stats(k, metric)
                ok &= list(a.columns) == list(b.columns) and len(a) == len(b)
                for col in ("statistic", "p-value", "effect_size"):
                    ok &= bool(np.allclose(a[col].values.astype(float), b[col].values.astype(float), rtol=1e-10, atol=1e-14,
                                           equal_nan=True))
    return ok


results = []
for case in (dict(S=12 * world, N=40, T=150, K_max=8, n_init=3, explicit=True),
             dict(S=5 * world + 3, N=116, T=90, K_max=6, n_init=2, explicit=False),       # uneven shards, drawn seeds
             dict(S=6 * world, N=400, T=120, K_max=20, n_init=2, explicit=True),
             dict(S=8 * world, N=30, T=90, K_max=5, n_init=2, explicit=True, extras=True),          # all rows scored
             dict(S=17 * world, N=24, T=1002, K_max=4, n_init=1, explicit=True, extras=True)):      # > 30 000 rows: sampled
    ok = check(**case)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    results.append(int(flag.item()))
    if rank == 0:
        print(f"world={world} exchange={os.environ.get('LEIDA_EXCHANGE', 'auto')} {case}: every shard identical to the "
              f"single-GPU run (labels, centroids, inertia, metric tables): {bool(flag.item())}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if all(results) else 1)
