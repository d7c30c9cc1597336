"""fit_predict end to end at config 2 with the statistics stage on (default) and off."""
import os, sys, time, contextlib, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
import pyleida
from pyleida import synthetic as syn
S, N, T = 100, 116, 1200
ts, classes = syn.make_shard(S, N, T, 0, S, "states")
inits = syn.init_indices(S * (T - 2), 2, 20, 15, seed=0)
for stats in (False, True, False, True, True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        m = pyleida.Leida.from_arrays(ts, classes).fit_predict(TR=0.72, init_indices=inits, scores=False, stats=stats, random_state=None)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"stats={stats}: {1e3 * (t1 - t0):.1f} ms")
print(m.stats(5, "occupancies").to_string())
