"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions (through
oracle/ref_shim.py) on seeded inputs.  Run in the build container only:

    python oracle/make_golden.py

The fixtures travel to the GPU box (where /root/reference does not exist); this script is the
record of how they were made.  Reference functions called: hilbert_phase, phase_coherence,
get_eigenvectors (signal_tools/_signal_tools.py:24-44,137-224), KMeansLeida.fit
(clustering/_clustering.py:82-153), identify_states (:250-374), compute_dynamics_metrics and
fractional_occupancy / dwell_times / transition_probabilities
(dynamics_metrics/_dynamics_metrics.py:11-136,208-276,423-473).
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from ref_shim import load_reference  # noqa: E402

import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location(
    "leida_synthetic", os.path.join(ROOT, "leida-python_b200", "pyleida", "synthetic.py"))
syn = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(syn)

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
ref = load_reference()


def signal_cases():
    """Stage 1+2 on assorted shapes: even / odd / prime T, even N (sign-rule tie reachable)."""
    out = {}
    shapes = [(12, 50, "states"), (9, 45, "iid"), (16, 53, "iid"), (20, 64, "states"),
              (10, 200, "states"), (6, 97, "iid"), (24, 120, "states"), (4, 31, "iid")]
    for i, (N, T, kind) in enumerate(shapes):
        x = syn.subject_series(100 + i, N, T, kind)
        ph = ref.hilbert_phase(x)
        dfc = ref.phase_coherence(ph)
        lei = ref.get_eigenvectors(dfc)
        out[f"x_{i}"] = x
        out[f"phase_{i}"] = ph
        out[f"lei_{i}"] = lei
        if N <= 12:
            out[f"dfc_{i}"] = dfc
    # the exact-half tie branch of the sign rule (_signal_tools.py:219): N=4, two entries of each sign
    th = np.zeros((4, 5))
    th[:, 1:4] = np.array([[0.1, 0.2, -0.3], [0.25, 0.1, -0.1], [2.9, -3.0, 3.0], [-2.8, 2.95, 3.1]])
    out["tie_phase"] = th
    out["tie_lei"] = ref.get_eigenvectors(ref.phase_coherence(th))
    out["n_cases"] = np.int64(len(shapes))
    np.savez_compressed(os.path.join(OUT, "signal_cases.npz"), **out)


def kmeans_cases():
    """KMeansLeida.fit on float32 eigenvector-like rows; int seed (identical restarts) and
    global-RNG restarts (np.random.seed(s) once, then n_init draws)."""
    out = {}
    ts, _ = syn.make_dataset(6, 24, 90, "states")
    lei = np.vstack([ref.get_eigenvectors(ref.phase_coherence(ref.hilbert_phase(v))) for v in ts.values()])
    X = np.array(lei, dtype=np.float32)
    out["X_states"] = X
    Xi = np.array(np.vstack([ref.get_eigenvectors(ref.phase_coherence(ref.hilbert_phase(
        syn.subject_series(50 + s, 17, 70, "iid")))) for s in range(5)]), dtype=np.float32)
    out["X_iid"] = Xi
    idx = 0
    for name, data in (("states", X), ("iid", Xi)):
        for k in (2, 3, 5, 8):
            km = ref.KMeansLeida(k=k, n_init=2, n_iter=1000)
            km.fit(data, random_state=0)
            out[f"case{idx}_meta"] = np.array([k, 2, 0, 1], dtype=np.int64)   # k, n_init, seed, int-seed?
            out[f"case{idx}_data"] = np.array(name)
            out[f"case{idx}_labels"] = km.labels_.astype(np.int64)
            out[f"case{idx}_centers"] = km.cluster_centers_
            out[f"case{idx}_inertia"] = np.float64(km.inertia_)
            idx += 1
        for k in (3, 6):
            np.random.seed(11)
            km = ref.KMeansLeida(k=k, n_init=4, n_iter=1000)
            km.fit(data, random_state=None)
            out[f"case{idx}_meta"] = np.array([k, 4, 11, 0], dtype=np.int64)
            out[f"case{idx}_data"] = np.array(name)
            out[f"case{idx}_labels"] = km.labels_.astype(np.int64)
            out[f"case{idx}_centers"] = km.cluster_centers_
            out[f"case{idx}_inertia"] = np.float64(km.inertia_)
            idx += 1
    out["n_cases"] = np.int64(idx)
    np.savez_compressed(os.path.join(OUT, "kmeans_cases.npz"), **out)


def metric_cases():
    """Single-sequence metrics on random label sequences (some states never visited)."""
    rng = np.random.default_rng(5)
    out = {}
    specs = [(3, 40), (5, 200), (8, 57), (20, 300), (4, 2), (6, 1)]
    for i, (k, n) in enumerate(specs):
        lab = rng.integers(0, k, size=n)
        if k >= 5:
            lab[lab == 1] = 0                      # state 1 never visited
        lab = np.repeat(lab, rng.integers(1, 4, size=n))[:n]
        occ = ref.fractional_occupancy(lab)
        occ_dense = np.zeros(k)
        occ_dense[occ.cluster.values.astype(int)] = occ.occupancy.values
        _, dw = ref.dwell_times(lab, TR=0.72)
        dw_dense = np.zeros(k)
        dw_dense[dw.Cluster.values.astype(int)] = dw.Mean_lifetime.values
        _, dw2 = ref.dwell_times(lab, TR=None)
        dw2_dense = np.zeros(k)
        dw2_dense[dw2.Cluster.values.astype(int)] = dw2.Mean_lifetime.values
        out[f"labels_{i}"] = lab.astype(np.int64)
        out[f"k_{i}"] = np.int64(k)
        out[f"occ_{i}"] = occ_dense
        out[f"dwell_tr_{i}"] = dw_dense
        out[f"dwell_{i}"] = dw2_dense
        out[f"trans_{i}"] = ref.transition_probabilities(lab, k)
        out[f"trans_raw_{i}"] = ref.transition_probabilities(lab, k, norm=False)
    out["n_cases"] = np.int64(len(specs))
    np.savez_compressed(os.path.join(OUT, "metric_cases.npz"), **out)


def pipeline_case():
    """_leida.py:279-331 restated with the reference's own functions: 8 subjects x 20 ROIs x 70
    volumes (two conditions), identify_states K=2..6 random_state=0, compute_dynamics_metrics."""
    S, N, T = 8, 20, 70
    ts, classes = syn.make_dataset(S, N, T, "states")
    eig, subs, conds = [], [], []
    for sid, x in ts.items():
        eig.append(ref.get_eigenvectors(ref.phase_coherence(ref.hilbert_phase(x))))
        for v in range(T - 2):
            subs.append(sid)
            conds.append(classes[sid][0])
    lei = np.vstack(eig)
    df = pd.DataFrame(lei, columns=[f"roi{j}" for j in range(N)])
    df.insert(0, "subject_id", subs)
    df.insert(1, "condition", conds)
    pred, perf, models = ref.identify_states(df, K_min=2, K_max=6, n_init=3, random_state=0,
                                             plot=False, save_results=False)
    dyn = ref.compute_dynamics_metrics(pred, TR=0.72, save_results=False)
    out = dict(S=np.int64(S), N=np.int64(N), T=np.int64(T), lei=lei,
               perf_k=perf["k"].values.astype(np.int64), perf_dunn=perf["Dunn_score"].values.astype(np.float64),
               perf_distortion=perf["distortion"].values.astype(np.float64),
               perf_silhouette=perf["silhouette_score"].values.astype(np.float64),
               perf_db=perf["Davies-Bouldin_score"].values.astype(np.float64),
               subject_id=np.array(subs), condition=np.array(conds))
    for k in range(2, 7):
        out[f"labels_k{k}"] = pred[f"k_{k}"].values.astype(np.int64)
        out[f"centers_k{k}"] = models[f"k_{k}"].cluster_centers_
        out[f"inertia_k{k}"] = np.float64(models[f"k_{k}"].inertia_)
        for m in ("occupancies", "dwell_times", "transitions"):
            t = dyn[m][f"k_{k}"]
            out[f"{m}_k{k}"] = t.iloc[:, 2:].values.astype(np.float64)
            out[f"{m}_k{k}_cols"] = np.array([str(v) for v in t.columns[2:]], dtype="U")
            out[f"{m}_k{k}_subject"] = np.array([str(v) for v in t.subject_id.values], dtype="U")
            out[f"{m}_k{k}_condition"] = np.array([str(v) for v in t.condition.values], dtype="U")
    np.savez_compressed(os.path.join(OUT, "pipeline_case.npz"), **out)


def stability_case():
    """patterns_stability (clustering/_clustering.py:471-605) on the pipeline case's eigenvectors; the global
    RNG is seeded right before each call (it drives the train/test split and the k-means restarts)."""
    import contextlib, io
    g = np.load(os.path.join(OUT, "pipeline_case.npz"))
    X = np.array(g["lei"], dtype=np.float32)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        np.random.seed(5)
        out["ari_y"] = ref.patterns_stability(X, y=g["labels_k3"], folds=4, metric="ari", plot=False)
        np.random.seed(6)
        out["ami_nc"] = ref.patterns_stability(X, n_clusters=4, folds=3, metric="ami", plot=False)
    np.savez_compressed(os.path.join(OUT, "stability_case.npz"), **out)


def stats_case():
    """permtest_ind / permtest_rel / hedges_g (stats/_stats.py:54-269) on the occupancy and dwell-time tables of
    the pipeline case plus a synthetic 2-condition table with a real effect and unequal variances.  The reference
    functions run unmodified except for the ttest_ind signature shim (ref_shim._compat_ttest_ind): statistic,
    Levene decision ('test'), effect size, Bonferroni alpha are the reference's; p-values are random (unseeded in
    the reference) and are stored only as a plausibility band."""
    import contextlib, io
    rng = np.random.default_rng(11)
    n1, n2, nv = 23, 31, 7
    vals = np.vstack([rng.normal(0.30, 1.0, (n1, nv)), rng.normal(0.0, 1.0, (n2, nv))])
    vals[n1:, 2] *= 2.5                      # unequal variances -> Welch
    vals[:n1, 4] += 1.2                      # a clear effect
    cond = np.array(["patients"] * n1 + ["controls"] * n2)
    cols = [f"PL_state_{i + 1}" for i in range(nv)]
    df = pd.DataFrame(vals, columns=cols)
    df.insert(0, "condition", cond)
    with contextlib.redirect_stdout(io.StringIO()):
        np.random.seed(3)
        ind = ref.permtest_ind(df, class_column="condition", n_perm=5000, alternative="two-sided")
        half = min(n1, n2)
        dfp = pd.concat([df.iloc[:half], df.iloc[n1:n1 + half]])
        rel = ref.permtest_rel(dfp, class_column="condition", n_perm=5000, alternative="two-sided")
    out = {"values": vals, "condition": cond.astype("U"), "columns": np.array(cols, dtype="U"), "half": np.int64(half)}
    for name, t in (("ind", ind), ("rel", rel)):
        for c in t.columns:
            v = t[c].values
            v = np.asarray(v)
            out[f"{name}_{c}"] = v if v.dtype.kind in "biuf" else np.array([str(x) for x in v], dtype="U")
    np.savez_compressed(os.path.join(OUT, "stats_case.npz"), **out)


if __name__ == "__main__":
    signal_cases()
    kmeans_cases()
    metric_cases()
    pipeline_case()
    stability_case()
    stats_case()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
