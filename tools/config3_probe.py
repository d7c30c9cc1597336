"""Config-3-shaped probe on one GPU: S subjects x 400 ROIs x 1200 volumes, K=2..20 x 15 restarts."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
from pyleida import synthetic as syn
from pyleida.clustering.kmeans import KMeansEngine, sweep_runs
from pyleida.signal_tools.phase import leading_eigenvectors
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N, T = 400, 1200
t0 = time.perf_counter()
ts, _ = syn.make_shard(S, N, T, 0, S, "states")
x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
print(f"data {time.perf_counter()-t0:.1f} s; input {x.numel()*8/1e9:.2f} GB")
M = S * (T - 2)
def sync():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    a = sync(); lei, xf = leading_eigenvectors(x); b = sync()
    print(f"stage 1+2: {1e3*(b-a):.2f} ms  ({(8*S*N*T + 8*M*N + 4*M*xf.shape[1])/1e9/(b-a):.0f} GB/s algorithmic)")
eng = KMeansEngine(xf, n_features=N)
inits = syn.init_indices(M, 2, 20, 15, seed=0)
runs, plan = sweep_runs(M, 2, 20, 15, None, inits)
for rep in range(2):
    prof = {"*": True} if rep == 1 else None
    a = sync(); res = eng.run(runs, profile=prof); b = sync()
    print(f"k-means sweep rep{rep}: {1e3*(b-a):.1f} ms, lockstep {res.lockstep_passes}, run-passes {res.total_passes()}, "
          f"rechecked {res.rechecked_total/(res.total_passes()*M):.2e}; {M*1.0/(b-a):.0f} volumes/s (k-means only)")
for name, evs in prof.items():
    if name == "*": continue
    t = [p.elapsed_time(q) for p, q in evs]
    print(f"  {name:10s} total {sum(t):8.2f} ms  first {t[0]:.3f} median {np.median(t):.3f}")
alg = (4.0*M*N + 8.0*M) * res.total_passes()
ta = sum(p.elapsed_time(q) for p, q in prof["assign"]) * 1e-3
print(f"assign: SURVEY-metric {alg/ta/1e12:.1f} TB/s")
# cross-check a few runs against the CUDA-core path
sub = [runs[i] for i in (0, 100, 200, 284)]
ra = eng.run(sub, n_iter=12, assign="tc"); rb = eng.run(sub, n_iter=12, assign="simt")
ok = all(bool((ra.labels_device(i) == rb.labels_device(i)).all()) and ra.inertia(i) == rb.inertia(i) for i in range(len(sub)))
print("TC == CUDA-core on 4 runs x 12 iterations:", ok)
