import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
import pyleida
from pyleida import synthetic as syn
S, N, T = 100, 116, 1200
ts, classes = syn.make_shard(S, N, T, 0, S, "states")
M = S * (T - 2)
inits = syn.init_indices(M, 2, 20, 15, seed=0)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if rep == 3:
        pr = cProfile.Profile(); pr.enable()
    m = pyleida.Leida.from_arrays(ts, classes)
    m.fit_predict(TR=0.72, K_min=2, K_max=20, n_init=15, init_indices=inits, keep_eigenvectors=True, scores=False, stats=False)
    if rep == 3:
        pr.disable()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"rep{rep} scores={rep % 2 == 1}: fit_predict {1e3*(t1-t0):.1f} ms  h2d {m.h2d_bytes/1e6:.0f} MB d2h {m.d2h_bytes/1e6:.0f} MB")
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
pstats.Stats(pr).sort_stats("cumtime").print_stats(35)
