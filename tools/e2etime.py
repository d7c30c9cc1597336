"""Mean end-to-end time of Leida.from_arrays(...).fit_predict(...) from host arrays (config 2), no profiler."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
import pyleida
from pyleida import synthetic as syn
S, N, T = 100, 116, 1200
ts, classes = syn.make_shard(S, N, T, 0, S, "states")
M = S * (T - 2)
inits = syn.init_indices(M, 2, 20, 15, seed=0)
times = []
for rep in range(9):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = pyleida.Leida.from_arrays(ts, classes)
    m.fit_predict(TR=0.72, K_min=2, K_max=20, n_init=15, init_indices=inits, keep_eigenvectors=True, scores=False, stats=False)
    _ = m.occupancies(5)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    times.append(1e3 * (t1 - t0)); del m
print("e2e ms:", " ".join(f"{t:.1f}" for t in times), "| mean of last 6: %.1f" % np.mean(times[3:]))
