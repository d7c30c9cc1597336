"""Wall-clock breakdown of one device-resident step (synchronising between stages)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
import pyleida
from pyleida import synthetic as syn, _native as nv
from pyleida.clustering.kmeans import KMeansEngine, sweep_runs, select_models
from pyleida.signal_tools.phase import leading_eigenvectors

S, N, T = 100, 116, 1200
ts, classes = syn.make_shard(S, N, T, 0, S, "states")
x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
M = S * (T - 2)
inits = syn.init_indices(M, 2, 20, 15, seed=0)

def sync():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(3):
    t0 = sync()
    lei, xf = leading_eigenvectors(x)
    t1 = sync()
    eng = KMeansEngine(xf, n_features=N)
    t2 = sync()
    runs, plan = sweep_runs(M, 2, 20, 15, None, inits)
    t3 = sync()
    prof = {"*": True}
    l0 = nv.launch_count()
    res = eng.run(runs, profile=prof if rep == 2 else None)
    t4 = sync()
    models = select_models(eng, res, plan, 15, to_host=False)
    t5 = sync()
    print(f"rep{rep}: stage12 {1e3*(t1-t0):.2f} ms | engine init {1e3*(t2-t1):.2f} | sweep_runs {1e3*(t3-t2):.2f} | "
          f"kmeans run {1e3*(t4-t3):.2f} (launches {nv.launch_count()-l0}, lockstep {res.lockstep_passes}, run-passes {res.total_passes()}) | select {1e3*(t5-t4):.2f}")
    print(f"  assign={res.assign} rechecked {res.rechecked_total} of {res.total_passes() * M} row-passes "
          f"({res.rechecked_total / (res.total_passes() * M):.2e}), label changes {res.changed_total}")
    for name, evs in prof.items():
        if name == "*": continue
        t = [a.elapsed_time(b) for a, b in evs]
        print(f"  {name:8s} total {sum(t):8.2f} ms  n={len(t)}  first {t[0]:.3f}  median {np.median(t):.3f}  max {max(t):.3f}")
# active runs per pass histogram
p = np.sort(res._passes)
print("passes per run: min/median/mean/max", p.min(), np.median(p), p.mean(), p.max())
for q in (10, 20, 30, 40, 60, 80, 100, 150):
    print(f"  runs still active after {q} passes: {(p > q).sum()}")

# per-pass series (ms) for the first passes, from the last profiled run
for name, evs in prof.items():
    if name == "*": continue
    t = [a.elapsed_time(b) for a, b in evs]
    print(name, "passes 0-47:", " ".join(f"{v:.2f}" for v in t[:48]))
    print(name, "sum 0-9 %.2f | 10-39 %.2f | 40-99 %.2f | 100+ %.2f" % (sum(t[:10]), sum(t[10:40]), sum(t[40:100]), sum(t[100:])))
