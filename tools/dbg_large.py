import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
import pyleida
from pyleida.clustering import kmeans as km
rng = np.random.default_rng(0)
for M in (3000, 20000, 70000, 200000):
    N, K = 116, 7
    cent = rng.standard_normal((K, N))
    X = (cent[rng.integers(0, K, M)] + 0.8 * rng.standard_normal((M, N))).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    eng = km.KMeansEngine(X)
    idx = rng.choice(M, K, replace=False)
    out = {}
    for mode in ("simt", "tc"):
        res = eng.run([(K, idx)], n_iter=40, assign=mode)
        lab = res.labels_device(0).cpu().numpy()
        out[mode] = (res.passes(0), res.inertia(0), lab, res.counts(0), res.changed_total, res.rechecked_total)
        print(M, mode, "passes", res.passes(0), "inertia", res.inertia(0), "counts", res.counts(0).tolist(), "bincount", np.bincount(lab, minlength=K).tolist(), "changed", res.changed_total, "rechecked", res.rechecked_total)
    print("   labels equal:", np.array_equal(out["simt"][2], out["tc"][2]))
