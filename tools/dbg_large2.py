import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
import pyleida
from pyleida.clustering import kmeans as km
rng = np.random.default_rng(0)
M, N, K = 200_000, 116, 7
cent = rng.standard_normal((K, N))
X = (cent[rng.integers(0, K, M)] + 0.8 * rng.standard_normal((M, N))).astype(np.float32)
X /= np.linalg.norm(X, axis=1, keepdims=True)
eng = km.KMeansEngine(X)
idx = rng.choice(M, K, replace=False)
for mode in ("simt", "tc", "tc"):
    for n_iter in (1000,):
        res = eng.run([(K, idx)], n_iter=n_iter, assign=mode)
        print(mode, "n_iter", n_iter, "passes", res.passes(0), "lockstep", res.lockstep_passes, "inertia", res.inertia(0), "changed", res.changed_total, "rechecked", res.rechecked_total, "counts", res.counts(0).tolist())
