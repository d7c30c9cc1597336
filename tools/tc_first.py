"""Times the first lock-step assign passes (all 285 runs active) of the config-1 sweep."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
from pyleida import synthetic as syn
from pyleida.clustering.kmeans import KMeansEngine, sweep_runs
from pyleida.signal_tools.phase import leading_eigenvectors
S, N, T = 100, 90, 1200
ts, _ = syn.make_shard(S, N, T, 0, S, "states")
x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
M = S * (T - 2)
inits = syn.init_indices(M, 2, 20, 15, seed=0)
_, xf = leading_eigenvectors(x, want_f64=False)
eng = KMeansEngine(xf, n_features=N)
runs, plan = sweep_runs(M, 2, 20, 15, None, inits)
for rep in range(2):
    prof = {"assign": []}
    try:
        res = eng.run(runs, n_iter=12, profile=prof)
    except Exception as e:
        print("run failed:", type(e).__name__)
    torch.cuda.synchronize()
    a = [e0.elapsed_time(e1) for e0, e1 in prof["assign"]]
    print("assign first %.3f  p2 %.3f p5 %.3f p10 %.3f" % tuple(a[i] if i < len(a) else -1 for i in (0, 2, 5, 10)))
