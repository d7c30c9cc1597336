"""Short k-means sweep (config-2 rows) used as the ncu target for k_diff_apply (pass 0 = dense, later = sparse)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
from pyleida import synthetic as syn
from pyleida.clustering.kmeans import KMeansEngine, sweep_runs
from pyleida.signal_tools.phase import leading_eigenvectors
S, N, T = 100, 116, 1200
ts, _ = syn.make_shard(S, N, T, 0, S, "states")
x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
M = S * (T - 2)
inits = syn.init_indices(M, 2, 20, 15, seed=0)
_, xf = leading_eigenvectors(x, want_f64=False)
eng = KMeansEngine(xf, n_features=N)
runs, plan = sweep_runs(M, 2, 20, 15, None, inits)
res = eng.run(runs, n_iter=8)
torch.cuda.synchronize()
print("done", res.lockstep_passes)
