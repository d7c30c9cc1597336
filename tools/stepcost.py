"""Where does a full device-resident bench step go? (synchronising between pieces)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
from pyleida import synthetic as syn, _native as nv
from pyleida.clustering.kmeans import KMeansEngine, sweep_runs, select_models
from pyleida.signal_tools.phase import leading_eigenvectors, padded_pitch
from pyleida.dynamics_metrics.metrics import Segments
from pyleida.parallel import Comm, row_offsets
from pyleida import pipeline

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

import cProfile, pstats
for rep in range(3):
    t0 = sync()
    if rep == 2:
        pr = cProfile.Profile(); pr.enable()
    res = pipeline.run_device([x], N, seg_off, seg_rows, init_indices=inits)
    if rep == 2:
        pr.disable()
    t1 = sync()
    print(f"rep{rep}: run_device {1e3*(t1-t0):.1f} ms")
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
