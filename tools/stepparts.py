import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
from pyleida import synthetic as syn, _native as nv
from pyleida.clustering.kmeans import KMeansEngine, sweep_runs, select_models
from pyleida.signal_tools.phase import leading_eigenvectors, padded_pitch
from pyleida.dynamics_metrics.metrics import Segments
S, N, T = 100, 116, 1200
ts, classes = syn.make_shard(S, N, T, 0, S, "states")
ids = list(ts)
x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
M = S * (T - 2)
inits = syn.init_indices(M, 2, 20, 15, seed=0)
seg = Segments.contiguous(np.asarray(ids), np.asarray([classes[s][0] for s in ids]), [T - 2] * S)
seg_off = torch.from_numpy(seg.off).cuda(); seg_rows = torch.from_numpy(seg.rows).cuda()
def sync():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    t = [sync()]
    xp = padded_pitch(N)
    x32 = torch.empty((M, xp), dtype=torch.float32, device="cuda"); lei = torch.empty((M, N), dtype=torch.float64, device="cuda")
    leading_eigenvectors(x, out_f64=lei, out_f32=x32); t.append(sync())
    eng = KMeansEngine(x32, n_features=N); t.append(sync())
    runs, plan = sweep_runs(M, 2, 20, 15, None, inits); t.append(sync())
    res = eng.run(runs); t.append(sync())
    models = select_models(eng, res, plan, 15, to_host=False); t.append(sync())
    ks = list(range(2, 21))
    labels = torch.stack([models[k].labels_device_ for k in ks]); t.append(sync())
    names = ["stage12+alloc", "engine init", "sweep_runs", "run", "select", "stack"]
    print(f"rep{rep}: total {1e3*(t[-1]-t[0]):.1f} ms | " + " | ".join(f"{n} {1e3*(b-a):.2f}" for n, a, b in zip(names, t[:-1], t[1:])))
