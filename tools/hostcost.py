"""How long does the HOST take to issue one lock-step pass (no GPU sync inside)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
from pyleida import synthetic as syn, _native as nv
from pyleida.clustering import kmeans as km
from pyleida.signal_tools.phase import leading_eigenvectors

S, N, T = 100, 116, 1200
ts, _ = syn.make_shard(S, N, T, 0, S, "states")
x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
M = S * (T - 2)
inits = syn.init_indices(M, 2, 20, 15, seed=0)
_, xf = leading_eigenvectors(x, want_f64=False)
eng = km.KMeansEngine(xf, n_features=N)
runs, plan = km.sweep_runs(M, 2, 20, 15, None, inits)
# time nv.call itself
orig = nv.call
acc = {}
def timed_call(name, *a):
    t = time.perf_counter(); orig(name, *a); acc.setdefault(name, []).append(time.perf_counter() - t)
nv.call = timed_call
for rep in range(2):
    acc.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = eng.run(runs)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep{rep}: host issue loop {1e3*(t1-t0):.1f} ms, + drain {1e3*(t2-t1):.1f} ms, passes {res.lockstep_passes}")
    for k, v in acc.items():
        print(f"   {k:32s} n={len(v):4d} host total {1e3*sum(v):7.2f} ms  median {1e6*np.median(v):6.1f} us")
