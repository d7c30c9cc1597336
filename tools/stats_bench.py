"""SURVEY 8f-4 measurement: the statistics stage (_compute_stats: permutation tests per state, K = 2..20, occupancies
and dwell times = 418 tests) on the GPU, next to the CPU oracle's restatement of the reference's SciPy calls on a
bounded sample (one K), same tables, n_perm = 5000."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, pandas as pd, torch
import pyleida
import leida_oracle as O

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_perm = 5000
rng = np.random.default_rng(0)
tables = {"occupancies": {}, "dwell_times": {}}
for k in range(2, 21):
    cols = [f"PL_state_{i + 1}" for i in range(k)]
    for metric in tables:
        v = rng.dirichlet(np.ones(k), size=S) if metric == "occupancies" else rng.gamma(4.0, 2.0, size=(S, k))
        v[: S // 2, 0] *= 1.15
        df = pd.DataFrame(v, columns=cols)
        df.insert(0, "condition", ["A"] * (S // 2) + ["B"] * (S - S // 2))
        df.insert(0, "subject_id", [f"sub-{i:04d}" for i in range(S)])
        tables[metric][f"k_{k}"] = df
import contextlib, io
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        res = pyleida.stats._compute_stats(tables, n_perm=n_perm, save_results=False, random_state=rep)
        _ = res["occupancies"]["k_20"]
    torch.cuda.synchronize(); t1 = time.perf_counter()
n_tests = 2 * sum(range(2, 21))
_ = res["occupancies"]["k_20"], res["dwell_times"]["k_20"]          # touched like a user would
gpu_s = t1 - t0
print(f"GPU: {n_tests} tests x {n_perm} permutations, {S} subjects: {1e3 * gpu_s:.1f} ms  ({n_tests * n_perm / gpu_s / 1e6:.1f} M permuted statistics/s)")
t0 = time.perf_counter()
occ = tables["occupancies"]["k_10"]
rows = O.permtest_table(occ.iloc[:, 2:].values, occ.condition.values, list(occ.columns[2:]), n_perm=n_perm,
                        random_state=np.random.RandomState(0))
cpu_s = time.perf_counter() - t0
print(f"CPU oracle (SciPy-style, 1 core): 10 tests x {n_perm} permutations: {1e3 * cpu_s:.0f} ms  ({10 * n_perm / cpu_s / 1e6:.3f} M/s) "
      f"=> x{(n_tests * n_perm / gpu_s) / (10 * n_perm / cpu_s):.0f}")
