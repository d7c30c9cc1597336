"""Short k-means sweep (config-2 rows, 3 restarts per K) used as the ncu target for k_tc_assign."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
from pyleida import synthetic as syn
from pyleida.clustering.kmeans import KMeansEngine, sweep_runs
from pyleida.signal_tools.phase import leading_eigenvectors

S, N, T = 100, 116, 1200
n_init = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ts, _ = syn.make_shard(S, N, T, 0, S, "states")
x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
M = S * (T - 2)
inits = syn.init_indices(M, 2, 20, n_init, seed=0)
_, xf = leading_eigenvectors(x, want_f64=False)
eng = KMeansEngine(xf, n_features=N)
runs, plan = sweep_runs(M, 2, 20, n_init, None, inits)
torch.cuda.synchronize(); t0 = time.perf_counter()
res = eng.run(runs)
torch.cuda.synchronize()
print(f"runs {len(runs)} lockstep {res.lockstep_passes} run-passes {res.total_passes()} time {1e3*(time.perf_counter()-t0):.1f} ms")
