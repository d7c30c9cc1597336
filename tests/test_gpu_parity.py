"""GPU parity tests: the CUDA path (through the C-ABI, via the pyleida host layer) against the
committed golden fixtures (unmodified reference outputs) and against the CPU oracle on seeded
inputs.  Tolerances follow BASELINE.json's north_star / SURVEY 8d:
  phases: wrapped angular error <= 1e-5*pi;  eigenvectors: max|v - v_ref| <= 1e-5 * max|v_ref| per row;
  labels: identical after _remap_labels;  inertia: rel 1e-6;  metric tables: bit-exact.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def P():
    import pyleida
    return pyleida


@pytest.fixture(scope="module")
def O():
    import leida_oracle
    return leida_oracle


@pytest.fixture(scope="module")
def sig():
    return np.load(os.path.join(GOLD, "signal_cases.npz"))


def wrapped(a, b):
    return np.abs(np.angle(np.exp(1j * (a - b))))


def eig_ok(got, ref):
    return np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref).max(axis=1, keepdims=True))


# ---- stage 1 --------------------------------------------------------------------------------
def test_hilbert_phase_golden(P, sig):
    for i in range(int(sig["n_cases"])):
        got = P.signal_tools.hilbert_phase(sig[f"x_{i}"])
        assert got.shape == sig[f"phase_{i}"].shape and got.dtype == np.float64
        err = wrapped(got, sig[f"phase_{i}"]).max()
        assert err <= 1e-5 * np.pi, (i, err)
        assert err <= 1e-10, (i, err)


@pytest.mark.parametrize("T", [3, 4, 5, 7, 16, 50, 59, 77, 128, 200, 211, 1200, 1201, 2000, 2047, 4096])
def test_hilbert_phase_lengths_vs_oracle(P, O, T):
    rng = np.random.default_rng(T)
    x = rng.standard_normal((5, T))
    got = P.signal_tools.hilbert_phase(x)
    assert wrapped(got, O.hilbert_phase(x)).max() <= 1e-9


def test_hilbert_pure_tone_is_a_ramp(P):
    T = 400
    t = np.arange(T)
    x = np.cos(2 * np.pi * 8 * t / T)[None, :].repeat(3, 0)
    ph = P.signal_tools.hilbert_phase(x)
    assert wrapped(ph[0], 2 * np.pi * 8 * t / T).max() < 1e-9


def test_hilbert_unsupported_length_raises(P):
    from pyleida._native import LeidaCudaError
    with pytest.raises(LeidaCudaError):
        P.signal_tools.hilbert_phase(np.zeros((2, 2053 * 2)))     # prime factor > 31 and > 2048


# ---- stage 2 --------------------------------------------------------------------------------
def test_phase_coherence_golden(P, sig):
    for i in range(int(sig["n_cases"])):
        if f"dfc_{i}" in sig:
            got = P.signal_tools.phase_coherence(sig[f"phase_{i}"])
            np.testing.assert_allclose(got, sig[f"dfc_{i}"], atol=1e-14)
    with pytest.raises(Exception):
        P.signal_tools.phase_coherence(np.zeros(5))


def test_rank2_eigenvectors_golden(P, sig):
    for i in range(int(sig["n_cases"])):
        got = P.signal_tools.eigenvectors_from_phases(sig[f"phase_{i}"])
        assert eig_ok(got, sig[f"lei_{i}"]), i
        assert np.abs(got - sig[f"lei_{i}"]).max() < 1e-12
    got = P.signal_tools.eigenvectors_from_phases(sig["tie_phase"])
    np.testing.assert_allclose(got, sig["tie_lei"], atol=1e-13)


def test_fused_leading_eigenvectors_golden(P, sig):
    for i in range(int(sig["n_cases"])):
        lei, xf = P.signal_tools.leading_eigenvectors(sig[f"x_{i}"][None])
        assert eig_ok(lei, sig[f"lei_{i}"]), i
        x32 = xf.cpu().numpy()
        n = sig[f"x_{i}"].shape[0]
        assert np.all(x32[:, n:] == 0)
        np.testing.assert_allclose(x32[:, :n], sig[f"lei_{i}"].astype(np.float32), atol=2e-7)


def test_dense_eigenvectors_golden_and_generic(P, O, sig):
    for i in range(int(sig["n_cases"])):
        dfc = O.phase_coherence(sig[f"phase_{i}"])
        got, info = P.signal_tools.get_eigenvectors(dfc, return_info=True)
        assert info["unconverged"] == 0, info
        assert eig_ok(got, sig[f"lei_{i}"]), i
    # generic symmetric matrices with a clear leading eigenvalue (not rank 2)
    rng = np.random.default_rng(2)
    N, T = 24, 40
    A = rng.standard_normal((N, N, T))
    A = A + A.transpose(1, 0, 2)
    u = rng.standard_normal((N, T))
    A += 6.0 * u[:, None, :] * u[None, :, :] / N
    got, info = P.signal_tools.get_eigenvectors(A, max_iter=400, return_info=True)
    assert info["unconverged"] == 0, info
    ref = O.get_eigenvectors(A)
    assert np.abs(got - ref).max() < 1e-8
    with pytest.raises(Exception):
        P.signal_tools.get_eigenvectors(np.zeros((3, 3)))


# ---- stage 3 --------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["exact", "reference"])
def test_kmeans_golden(P, mode):
    g = np.load(os.path.join(GOLD, "kmeans_cases.npz"))
    for c in range(int(g["n_cases"])):
        k, n_init, seed, is_int = (int(v) for v in g[f"case{c}_meta"])
        X = g["X_" + str(g[f"case{c}_data"])]
        km = P.clustering.KMeansLeida(k=k, n_init=n_init)
        if is_int:
            km.fit(X, random_state=seed, centroid_update=mode)
        else:
            np.random.seed(seed)
            km.fit(X, random_state=None, centroid_update=mode)
        assert km.labels_.dtype == np.int64
        assert np.array_equal(km.labels_, g[f"case{c}_labels"]), (c, mode)
        assert km.inertia_ == pytest.approx(float(g[f"case{c}_inertia"]), rel=1e-6)
        if mode == "reference":
            assert np.array_equal(km.cluster_centers_, g[f"case{c}_centers"]), c        # bit-identical float32 means
        else:
            np.testing.assert_allclose(km.cluster_centers_, g[f"case{c}_centers"], rtol=0, atol=5e-7)


def test_kmeans_predict_transform(P, O):
    g = np.load(os.path.join(GOLD, "kmeans_cases.npz"))
    X = g["X_states"]
    km = P.clustering.KMeansLeida(k=5, n_init=1)
    km.fit(X, random_state=3)
    ref = O.KMeansLeida(k=5, n_init=1)
    ref.cluster_centers_, ref._is_fitted = km.cluster_centers_, True
    np.testing.assert_allclose(km.transform(X), ref.transform(X), atol=1e-14)
    assert np.array_equal(km.predict(X), ref.predict(X))
    assert np.array_equal(km.predict(X), km.labels_)
    np.testing.assert_allclose(km.transform(X, closest=True), ref.transform(X, closest=True), atol=1e-14)


def test_kmeans_constructor_errors(P):
    K = P.clustering.KMeansLeida
    with pytest.raises(TypeError):
        K(metric=3)
    with pytest.raises(ValueError):
        K(metric="manhattan")
    with pytest.raises(TypeError):
        K(k=2.0)
    with pytest.raises(ValueError):
        K(k=1)
    with pytest.raises(Exception):
        K(k=2).predict(np.zeros((3, 3), dtype=np.float32))


@pytest.mark.parametrize("kind,N,S,T", [("states", 90, 20, 200), ("iid", 33, 6, 150)])
def test_kmeans_vs_oracle_config1(P, O, kind, N, S, T):
    """Config-1-sized sweep (20 x 90 x 200) against the oracle with shared explicit initial rows."""
    from pyleida import synthetic
    ts, _ = synthetic.make_dataset(S, N, T, kind)
    lei = np.vstack([O.leading_eigvec_rank2(O.hilbert_phase(v)) for v in ts.values()])
    X = lei.astype(np.float32)
    inits = synthetic.init_indices(X.shape[0], 2, 9, 2, seed=4)
    eng = P.clustering.KMeansEngine(X)
    for k in (2, 3, 5, 9):
        ref = O.KMeansLeida(k=k, n_init=2).fit(X, init_indices=inits[k])
        for mode in ("exact", "reference"):
            km = P.clustering.KMeansLeida(k=k, n_init=2).fit(X, init_indices=inits[k], centroid_update=mode, engine=eng)
            assert np.array_equal(km.labels_, ref.labels_), (k, mode)
            assert km.inertia_ == pytest.approx(ref.inertia_, rel=1e-6 if mode == "exact" else 1e-12)
            assert km.n_passes_ == ref.n_iter_runs_[km.best_init_] + 1
            if mode == "reference":
                assert np.array_equal(km.cluster_centers_, ref.cluster_centers_)


def test_kmeans_large_k_and_wide_rows(P, O):
    rng = np.random.default_rng(8)
    X = rng.standard_normal((3000, 400)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    for k in (24, 40, 64):
        idx = [rng.choice(3000, k, replace=False)]
        ref = O.KMeansLeida(k=k, n_init=1, n_iter=6).fit(X, init_indices=idx)
        km = P.clustering.KMeansLeida(k=k, n_init=1, n_iter=6).fit(X, init_indices=idx, centroid_update="reference")
        assert np.array_equal(km.labels_, ref.labels_), k
        assert km.inertia_ == pytest.approx(ref.inertia_, rel=1e-12)


def test_kmeans_properties_large(P):
    """Size-independent properties at a size the oracle cannot reach in seconds."""
    import torch
    rng = np.random.default_rng(0)
    M, N, K = 200_000, 116, 7
    cent = rng.standard_normal((K, N))
    X = (cent[rng.integers(0, K, M)] + 0.8 * rng.standard_normal((M, N))).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    eng = P.clustering.KMeansEngine(X)
    idx = [rng.choice(M, K, replace=False)]
    a = P.clustering.KMeansLeida(k=K, n_init=1).fit(None, init_indices=idx, engine=eng)
    b = P.clustering.KMeansLeida(k=K, n_init=1).fit(None, init_indices=idx, engine=eng)
    assert np.array_equal(a.labels_, b.labels_) and a.inertia_ == b.inertia_            # run-to-run deterministic
    # fixed point: labels are the argmin of the final centroids, centroids are the member means
    assert np.array_equal(a.predict(eng), a.labels_)
    means = np.vstack([X[a.labels_ == c].astype(np.float64).mean(0) for c in range(K)])
    np.testing.assert_allclose(a.cluster_centers_, means, atol=1e-6)
    # relabelling: sizes are non-increasing, and permuting the initial rows only permutes clusters
    sizes = np.bincount(a.labels_, minlength=K)
    assert np.all(np.diff(sizes) <= 0)
    c = P.clustering.KMeansLeida(k=K, n_init=1).fit(None, init_indices=[idx[0][::-1].copy()], engine=eng)
    assert np.array_equal(c.labels_, a.labels_) and c.inertia_ == a.inertia_
    # inertia equals the fp64 sum of squared cosine distances to the assigned centroid
    Xd = X.astype(np.float64)
    C = a.cluster_centers_.astype(np.float64)[a.labels_]
    d = 1 - (Xd * C).sum(1) / (np.linalg.norm(Xd, axis=1) * np.linalg.norm(C, axis=1))
    assert a.inertia_ == pytest.approx((d ** 2).sum(), rel=1e-10)
    torch.cuda.synchronize()


# ---- stage 4 --------------------------------------------------------------------------------
def test_single_sequence_metrics_golden(P):
    g = np.load(os.path.join(GOLD, "metric_cases.npz"))
    D = P.dynamics_metrics
    for i in range(int(g["n_cases"])):
        lab, k = g[f"labels_{i}"], int(g[f"k_{i}"])
        occ = D.fractional_occupancy(lab)
        dense = np.zeros(k)
        dense[occ.cluster.values] = occ.occupancy.values
        assert np.array_equal(dense, g[f"occ_{i}"])
        for tr, key in ((0.72, "dwell_tr"), (None, "dwell")):
            _, dw = D.dwell_times(lab, TR=tr)
            dense = np.zeros(k)
            dense[dw.Cluster.values] = dw.Mean_lifetime.values
            assert np.array_equal(dense, g[f"{key}_{i}"])
        assert np.array_equal(D.transition_probabilities(lab, k), g[f"trans_{i}"])
        assert np.array_equal(D.transition_probabilities(lab, k, norm=False), g[f"trans_raw_{i}"])


def test_group_metrics_noncontiguous_segments(P, O):
    import pandas as pd
    rng = np.random.default_rng(9)
    subs = np.repeat([f"s{i}" for i in range(5)], 30)
    conds = np.tile(np.repeat(["B", "A", "B"], 10), 5)            # per-volume conditions: groups interleave
    lab = np.repeat(rng.integers(0, 4, size=75), 2)
    meta = pd.DataFrame({"subject_id": subs, "condition": conds})
    t = O.dynamics_tables(subs, conds, lab, TR=2.0)
    D = P.dynamics_metrics
    occ = D.fractional_occupancy_group(meta, lab)
    assert list(zip(occ.subject_id, occ.condition)) == t["rows"]
    assert np.array_equal(occ.iloc[:, 2:].values, t["occupancies"])
    assert np.array_equal(D.dwell_times_group(meta, lab, TR=2.0).iloc[:, 2:].values, t["dwell_times"])
    assert np.array_equal(D.transition_probabilities_group(meta, lab).iloc[:, 2:].values, t["transitions"])
    with pytest.raises(AssertionError):
        D.fractional_occupancy_group(meta, lab[:-1])


# ---- pipeline -------------------------------------------------------------------------------
def test_pipeline_golden(P):
    from pyleida import synthetic
    g = np.load(os.path.join(GOLD, "pipeline_case.npz"))
    S, N, T = int(g["S"]), int(g["N"]), int(g["T"])
    ts, classes = synthetic.make_dataset(S, N, T, "states")
    for mode in ("exact", "reference"):
        m = P.Leida.from_arrays(ts, classes).fit_predict(TR=0.72, random_state=0, K_min=2, K_max=6, n_init=3,
                                                         centroid_update=mode)
        lei = m.eigenvectors.iloc[:, 2:].values
        assert eig_ok(lei, g["lei"])
        assert list(m.eigenvectors.subject_id) == list(g["subject_id"])
        for k in range(2, 7):
            assert np.array_equal(m.predictions[f"k_{k}"].values, g[f"labels_k{k}"]), (k, mode)
            assert m.load_model(k).inertia_ == pytest.approx(float(g[f"inertia_k{k}"]), rel=1e-6)
            for name in ("occupancies", "dwell_times", "transitions"):
                tab = getattr(m._dynamics_, name)[f"k_{k}"]
                assert np.array_equal(tab.iloc[:, 2:].values, g[f"{name}_k{k}"]), (k, name)
                assert list(tab.columns[2:]) == list(g[f"{name}_k{k}_cols"])
                assert list(tab.subject_id) == list(g[f"{name}_k{k}_subject"])
                assert list(tab.condition) == list(g[f"{name}_k{k}_condition"])


def test_pipeline_ragged_and_per_volume_conditions(P, O):
    from pyleida import synthetic
    ts = {f"s{i}": synthetic.subject_series(i, 10, T) for i, T in enumerate([40, 40, 55, 31, 31])}
    classes = {"s0": ["A"], "s1": ["B"], "s2": ["A"], "s3": ["B"], "s4": ["A"] * 10 + ["B"] * 21}
    m = P.Leida.from_arrays(ts, classes).fit_predict(TR=1.0, random_state=1, K_min=2, K_max=3, n_init=1)
    ref = O.execute_all(ts, classes, TR=1.0, K_min=2, K_max=3, n_init=1, random_state=1)
    assert eig_ok(m.eigenvectors.iloc[:, 2:].values, ref["eigenvectors"])
    assert list(m.eigenvectors.condition) == list(ref["condition"])
    for k in (2, 3):
        assert np.array_equal(m.predictions[f"k_{k}"].values, ref["models"][k].labels_)
        tab = m.occupancies(k)
        assert list(zip(tab.subject_id, tab.condition)) == ref["metrics"][k]["rows"]
        assert np.array_equal(tab.iloc[:, 2:].values, ref["metrics"][k]["occupancies"])
        assert np.array_equal(m.transitions(k).iloc[:, 2:].values, ref["metrics"][k]["transitions"])
        assert np.array_equal(m.dwell_times(k).iloc[:, 2:].values, ref["metrics"][k]["dwell_times"])


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


# ---- wider coverage -----------------------------------------------------------------------
def test_config1_full_sweep_vs_oracle(P, O):
    """BASELINE config 1 (20 x 90 x 200), the whole pipeline, K=2..20, random_state=0 (the reference's
    semantics: identical restarts), against the CPU oracle."""
    from pyleida import synthetic
    ts, classes = synthetic.make_dataset(20, 90, 200, "states")
    m = P.Leida.from_arrays(ts, classes).fit_predict(TR=0.72, random_state=0)
    ref = O.execute_all(ts, classes, TR=0.72, K_min=2, K_max=20, n_init=1, random_state=0)
    assert eig_ok(m.eigenvectors.iloc[:, 2:].values, ref["eigenvectors"])
    for k in range(2, 21):
        assert np.array_equal(m.predictions[f"k_{k}"].values, ref["models"][k].labels_), k
        assert m.load_model(k).inertia_ == pytest.approx(ref["models"][k].inertia_, rel=1e-6)
        for name in ("occupancies", "dwell_times", "transitions"):
            assert np.array_equal(getattr(m._dynamics_, name)[f"k_{k}"].iloc[:, 2:].values, ref["metrics"][k][name]), (k, name)


def test_tensor_core_and_cuda_core_assignment_agree(P):
    """The tcgen05 path and the fp32 CUDA-core path must give identical labels, pass counts and
    inertia (both are the fp64 argmin by construction) -- on iid data, the worst case for margins,
    and on wide rows (N=1000 -> 16 k-blocks)."""
    rng = np.random.default_rng(3)
    for M, N, ks in ((30_000, 116, (2, 7, 20)), (5_000, 1000, (3, 20)), (9_000, 64, (40,))):
        X = rng.standard_normal((M, N)).astype(np.float32)
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        eng = P.clustering.KMeansEngine(X)
        runs = [(k, rng.choice(M, k, replace=False)) for k in ks for _ in range(2)]
        a = eng.run(runs, n_iter=25, assign="tc")
        b = eng.run(runs, n_iter=25, assign="simt")
        for i in range(len(runs)):
            assert a.passes(i) == b.passes(i), (M, N, i)
            assert a.inertia(i) == b.inertia(i), (M, N, i)
            assert bool((a.labels_device(i) == b.labels_device(i)).all()), (M, N, i)
        assert a.rechecked_total < 0.05 * a.total_passes() * M      # the fp64 re-check stays a rare path


def test_recheck_queue_overflow_changes_nothing(P, monkeypatch):
    """A re-check queue far too small for the undecided pairs (iid data, 64 slots) must give the same
    labels, pass counts and inertia: the fast pass marks the pairs, the re-check scans for the marks."""
    rng = np.random.default_rng(5)
    M, N = 20_000, 90
    X = rng.standard_normal((M, N)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    eng = P.clustering.KMeansEngine(X)
    runs = [(k, rng.choice(M, k, replace=False)) for k in (2, 5, 12, 20) for _ in range(2)]
    a = eng.run(runs, n_iter=20, assign="tc")
    monkeypatch.setenv("LEIDA_KMEANS_WORKLIST_CAP", "64")
    b = eng.run(runs, n_iter=20, assign="tc")
    assert a.rechecked_total > 64                  # the small queue did overflow
    for i in range(len(runs)):
        assert a.passes(i) == b.passes(i), i
        assert a.inertia(i) == b.inertia(i), i
        assert bool((a.labels_device(i) == b.labels_device(i)).all()), i


def test_wide_rows_rank2_and_dense_eigenvectors(P, O):
    """N = 1000 ROIs (BASELINE config 4 shape, few volumes): closed form and general solver."""
    from pyleida import synthetic
    x = synthetic.subject_series(5, 1000, 30, "states")
    ph = O.hilbert_phase(x)
    ref = O.leading_eigvec_rank2(ph)
    assert eig_ok(P.signal_tools.eigenvectors_from_phases(ph), ref)
    lei, _ = P.signal_tools.leading_eigenvectors(x[None])
    assert eig_ok(lei, ref)
    got, info = P.signal_tools.get_eigenvectors(P.signal_tools.phase_coherence(ph), return_info=True)
    assert info["unconverged"] == 0 and info["iterations"] <= 4
    assert eig_ok(got, ref)


def test_empty_cluster_policy(P):
    """Two identical initial rows: the second centroid never wins a tie (first minimum), so its
    cluster is empty after the first pass.  The reference would carry NaN centroids; here the run is
    flagged and discarded."""
    rng = np.random.default_rng(1)
    X = rng.standard_normal((500, 12)).astype(np.float32)
    X[7] = X[3]
    bad, good = np.array([3, 7, 11]), np.array([1, 2, 5])
    km = P.clustering.KMeansLeida(k=3, n_init=2)
    with pytest.warns(UserWarning):
        km.fit(X, init_indices=[bad, good])
    assert km.best_init_ == 1 and np.isnan(km.inertia_runs_[0]) and np.isfinite(km.inertia_)
    with pytest.raises(RuntimeError):
        P.clustering.KMeansLeida(k=3, n_init=1).fit(X, init_indices=[bad])


def test_kmeans_argument_errors(P):
    X = np.random.default_rng(0).standard_normal((50, 8)).astype(np.float32)
    eng = P.clustering.KMeansEngine(X)
    with pytest.raises(ValueError):
        eng.run([(3, np.array([0, 1]))])                 # k rows needed
    with pytest.raises(ValueError):
        eng.run([(2, np.array([0, 50]))])                # index out of range
    with pytest.raises(ValueError):
        eng.run([(65, np.arange(65) % 50)])              # k <= 64
    with pytest.raises(ValueError):
        eng.run([(2, np.array([0, 1]))], metric="manhattan")
    with pytest.raises(NotImplementedError):
        P.signal_tools.get_eigenvectors(np.zeros((4, 4, 3)), n=2)


def test_leida_input_validation_and_loader(P, tmp_path):
    import pickle
    from pyleida import synthetic
    ts, classes = synthetic.make_dataset(3, 6, 20, "iid")
    with pytest.raises(TypeError):
        P.Leida(123)
    with pytest.raises(ValueError):
        P.Leida(str(tmp_path / "missing"))
    bad = dict(ts)
    bad["x"] = np.zeros((5, 20))
    with pytest.raises(Exception):
        P.Leida.from_arrays(bad, {**classes, "x": ["A"]})
    (tmp_path / "rois_labels.txt").write_text("\n".join(f"roi {i}" for i in range(6)))
    pickle.dump(ts, open(tmp_path / "time_series.pkl", "wb"))
    pickle.dump(classes, open(tmp_path / "metadata.pkl", "wb"))
    m = P.Leida(str(tmp_path))
    assert m.rois_labels[0] == "roi 0" and set(m.time_series) == set(ts)
    with pytest.raises(TypeError):
        m.fit_predict(TR="fast")
    with pytest.raises(ValueError):
        m.fit_predict(n_perm=10)
    m.fit_predict(TR=2, K_min=2, K_max=3, n_init=1, random_state=0)
    assert list(m.predictions.columns) == ["subject_id", "condition", "k_2", "k_3"]
    assert m.load_centroids(3).shape == (6, 3)


# ---- clustering quality scores ----------------------------------------------------
def test_identify_states_scores_golden(P):
    """identify_states' performance table (Dunn, distortion, silhouette, Davies-Bouldin) against the
    table the unmodified reference produced (tests/golden/pipeline_case.npz)."""
    import pandas as pd
    g = np.load(os.path.join(GOLD, "pipeline_case.npz"))
    df = pd.DataFrame(g["lei"], columns=[f"roi{j}" for j in range(int(g["N"]))])
    df.insert(0, "subject_id", g["subject_id"])
    df.insert(1, "condition", g["condition"])
    pred, perf, models = P.clustering.identify_states(df, K_min=2, K_max=6, n_init=3, random_state=0, plot=False)
    assert list(perf.columns) == ["k", "Dunn_score", "distortion", "silhouette_score", "Davies-Bouldin_score"]
    assert np.array_equal(perf["k"].values, g["perf_k"])
    for k in range(2, 7):
        assert np.array_equal(pred[f"k_{k}"].values, g[f"labels_k{k}"])
    np.testing.assert_allclose(perf["distortion"].values, g["perf_distortion"], rtol=1e-6)
    # Dunn is a ratio of single float32 distances (sklearn computes them with BLAS): 1e-4 is the noise floor
    np.testing.assert_allclose(perf["Dunn_score"].values, g["perf_dunn"], rtol=2e-4)
    np.testing.assert_allclose(perf["silhouette_score"].values, g["perf_silhouette"], atol=2e-6)
    np.testing.assert_allclose(perf["Davies-Bouldin_score"].values, g["perf_db"], rtol=1e-5)


def test_scores_subsample_and_many_k_vs_oracle(P, O):
    rng = np.random.default_rng(11)
    M, N = 6000, 33
    cent = rng.standard_normal((5, N))
    X = (cent[rng.integers(0, 5, M)] + 0.9 * rng.standard_normal((M, N))).astype(np.float32)
    eng = P.clustering.KMeansEngine(X)
    ks = [2, 3, 7, 12]
    import torch
    labs = [rng.integers(0, k, M).astype(np.int32) for k in ks]
    for lab, k in zip(labs, ks):
        lab[:k] = np.arange(k)                      # every cluster present
    cents = [np.vstack([X[lab == c].astype(np.float64).mean(0) for c in range(k)]) for lab, k in zip(labs, ks)]
    sel = np.sort(rng.choice(M, 2500, replace=False))
    dev_labels = torch.from_numpy(np.stack(labs)).cuda()
    got = P.clustering.clustering_scores(eng, dev_labels, ks, cents, sample_idx=sel)
    for c, (lab, k) in enumerate(zip(labs, ks)):
        ref = O.clustering_scores(X, lab, sample_idx=sel)
        assert got["Dunn_score"][c] == pytest.approx(ref["Dunn_score"], rel=2e-4)
        assert got["silhouette_score"][c] == pytest.approx(ref["silhouette_score"], abs=2e-6)
        assert got["Davies-Bouldin_score"][c] == pytest.approx(ref["Davies-Bouldin_score"], rel=1e-5)


def test_results_folder_layout(P, tmp_path, monkeypatch):
    """save_results=True writes the reference's LEiDA_results layout and refuses to overwrite it."""
    import pickle
    import pandas as pd
    from pyleida import synthetic
    monkeypatch.chdir(tmp_path)
    ts, classes = synthetic.make_dataset(4, 8, 30, "states")
    m = P.Leida.from_arrays(ts, classes).fit_predict(TR=1.0, save_results=True, random_state=0, K_min=2, K_max=3, n_init=1)
    root = tmp_path / "LEiDA_results"
    eig = pd.read_csv(root / "eigenvectors.csv", sep="\t")
    assert list(eig.columns[:2]) == ["subject_id", "condition"] and eig.shape == (4 * 28, 10)
    np.testing.assert_allclose(eig.iloc[:, 2:].values, m.eigenvectors.iloc[:, 2:].values.astype(float), rtol=0, atol=1e-15)
    pred = pd.read_csv(root / "clustering" / "predictions.csv", sep="\t")
    assert np.array_equal(pred["k_3"].values, m.predictions["k_3"].values)
    perf = pd.read_csv(root / "clustering" / "clustering_performance.csv", sep="\t")
    assert list(perf.columns) == ["k", "Dunn_score", "distortion", "silhouette_score", "Davies-Bouldin_score"]
    model = pickle.load(open(root / "clustering" / "models" / "model_k_2.pkl", "rb"))
    assert np.array_equal(model.cluster_centers_, m.load_model(2).cluster_centers_)
    for k in (2, 3):
        for f in ("occupancies.csv", "dwell_times.csv", "transitions_probabilities.csv"):
            assert (root / "dynamics_metrics" / f"k_{k}" / f).exists()
    occ = pd.read_csv(root / "dynamics_metrics" / "k_3" / "occupancies.csv", sep="\t")
    np.testing.assert_allclose(occ.iloc[:, 2:].values, m.occupancies(3).iloc[:, 2:].values.astype(float), atol=1e-15)
    with pytest.raises(Warning):
        P.Leida.from_arrays(ts, classes).fit_predict(TR=1.0, save_results=True, K_min=2, K_max=2, n_init=1)


# ---- callers of predict / transform ---------------------------------------------------
def test_patterns_stability_golden(P):
    """Same global-RNG seed as the unmodified reference => same split, same restarts, same score matrix."""
    g = np.load(os.path.join(GOLD, "pipeline_case.npz"))
    s = np.load(os.path.join(GOLD, "stability_case.npz"))
    X = np.array(g["lei"], dtype=np.float32)
    np.random.seed(5)
    got = P.clustering.patterns_stability(X, y=g["labels_k3"], folds=4, metric="ari", plot=False)
    np.testing.assert_allclose(got, s["ari_y"], atol=1e-12)
    np.random.seed(6)
    got = P.clustering.patterns_stability(X, n_clusters=4, folds=3, metric="ami", plot=False)
    np.testing.assert_allclose(got, s["ami_nc"], atol=1e-12)
    with pytest.raises(Exception):
        P.clustering.patterns_stability(X, metric="nmi")
    with pytest.raises(Exception):
        P.clustering.patterns_stability(X)


def test_centroids_distances(P, O):
    from pyleida import synthetic
    ts, classes = synthetic.make_dataset(5, 10, 40, "states")
    m = P.Leida.from_arrays(ts, classes).fit_predict(random_state=0, K_min=2, K_max=4, n_init=1)
    d = m.centroids_distances(3)
    assert list(d.columns) == ["subject_id", "condition", "centroid_1", "centroid_2", "centroid_3"]
    X = m.eigenvectors.iloc[:, 2:].values.astype(np.float32)
    ref = O.KMeansLeida(k=3)
    ref.cluster_centers_, ref._is_fitted = m.load_model(3).cluster_centers_, True
    np.testing.assert_allclose(d.iloc[:, 2:].values, ref.transform(X), atol=1e-13)
    assert np.array_equal(np.argmin(d.iloc[:, 2:].values, axis=1), m.predictions["k_3"].values)
    with pytest.raises(ValueError):
        m.centroids_distances(9)


# ---- post-hoc statistics -------------------------------------------------------------------

def _stats_frame(g):
    import pandas as pd
    df = pd.DataFrame(g["values"], columns=[str(c) for c in g["columns"]])
    df.insert(0, "condition", [str(c) for c in g["condition"]])
    return df


def test_permtest_ind_golden_and_identical_permutations(P, O):
    """permtest_ind on the GPU against the fixture made by the unmodified reference (statistic, Levene decision,
    effect size, alpha, table layout) and against the oracle on IDENTICAL permutations: the whole null distribution
    and the p-values must agree (for every alternative)."""
    g = np.load(os.path.join(GOLD, "stats_case.npz"))
    df = _stats_frame(g)
    n = len(df)
    n1 = int((df.condition == "controls").sum())               # np.unique order: 'controls' is group 1
    rng = np.random.RandomState(12)
    perms = np.asarray([rng.permutation(n) for _ in range(3000)])
    res, null = P.stats.permtest_ind(df, class_column="condition", n_perm=3000, perms=perms, return_null=True)
    assert list(res.columns) == ["variable", "group_1", "group_2", "statistic", "p-value", "test", "effect_size",
                                 "alpha_Bonferroni", "reject_null"]
    for j in range(len(res)):
        assert res["variable"][j] == str(g["ind_variable"][j]) and res["test"][j] == str(g["ind_test"][j])
        assert res["group_1"][j] == str(g["ind_group_1"][j]) and res["group_2"][j] == str(g["ind_group_2"][j])
        assert res["statistic"][j] == pytest.approx(g["ind_statistic"][j], rel=1e-12)
        assert res["effect_size"][j] == pytest.approx(g["ind_effect_size"][j], rel=1e-12)
        assert res["alpha_Bonferroni"][j] == g["ind_alpha_Bonferroni"][j]
    vals, cond = g["values"], np.asarray([str(c) for c in g["condition"]])
    for alt in ("two-sided", "less", "greater"):
        got = P.stats.permtest_ind(df, class_column="condition", n_perm=3000, perms=perms, alternative=alt)
        for j, c in enumerate(res["variable"]):
            x1, x2 = vals[cond == "controls", j], vals[cond == "patients", j]
            _, p, ref_null = O.permutation_ttest_ind(x1, x2, 3000, res["test"][j] == "t-test", alt, perms=perms)
            if alt == "two-sided":
                assert np.allclose(null[j], ref_null, rtol=1e-10, atol=1e-12), j
            assert got["p-value"][j] == pytest.approx(p, abs=1.5 / 3001), (alt, j)    # a tie at fp64 rounding may flip one count
    assert n1 == 31


def test_permtest_ind_device_rng_and_exact_branch(P, O):
    """Device-drawn splits: reproducible for a seed, different for another, and statistically equal to the oracle's
    p-values (6 standard errors); groups small enough to enumerate give the exact test, equal to the oracle's."""
    import pandas as pd
    g = np.load(os.path.join(GOLD, "stats_case.npz"))
    df = _stats_frame(g)
    a = P.stats.permtest_ind(df, class_column="condition", n_perm=5000, random_state=7)
    b = P.stats.permtest_ind(df, class_column="condition", n_perm=5000, random_state=7)
    c = P.stats.permtest_ind(df, class_column="condition", n_perm=5000, random_state=8)
    assert np.array_equal(a["p-value"].values, b["p-value"].values)
    assert not np.array_equal(a["p-value"].values, c["p-value"].values)
    for j in range(len(a)):
        p = float(g["ind_p-value"][j])
        assert abs(a["p-value"][j] - p) <= 6 * np.sqrt(max(p * (1 - p), 1e-3) / 5000) * 1.5, j
    rng = np.random.default_rng(3)
    small = pd.DataFrame(rng.normal(size=(9, 3)), columns=["s1", "s2", "s3"])
    small.insert(0, "condition", ["x"] * 5 + ["y"] * 4)
    got = P.stats.permtest_ind(small, class_column="condition", n_perm=5000)        # C(9,5) = 126 splits: exact
    rows = O.permtest_table(small.iloc[:, 1:].values, small.condition.values, ["s1", "s2", "s3"], n_perm=5000)
    for j, r in enumerate(rows):
        assert got["p-value"][j] == pytest.approx(r["p-value"], abs=1e-15)
        assert got["statistic"][j] == pytest.approx(r["statistic"], rel=1e-12)


def test_permtest_rel_and_compute_stats(P, O):
    """Paired test: statistic / effect size against the reference fixture, null distribution against the oracle on
    identical swaps, exact branch for few pairs; _compute_stats returns the reference's nested layout."""
    import pandas as pd
    g = np.load(os.path.join(GOLD, "stats_case.npz"))
    df = _stats_frame(g)
    half = int(g["half"])
    n1 = int((df.condition == "patients").sum())
    dfp = pd.concat([df.iloc[:half], df.iloc[n1:n1 + half]]).reset_index(drop=True)
    rng = np.random.default_rng(4)
    swaps = rng.integers(0, 2, size=(2000, half))
    res, null = P.stats.permtest_rel(dfp, class_column="condition", n_perm=2000, swaps=swaps, return_null=True)
    assert list(res.columns) == ["variable", "group_1", "group_2", "statistic", "p-value", "effect_size",
                                 "alpha_Bonferroni", "reject_null"]
    vals, cond = dfp.iloc[:, 1:].values, dfp.condition.values
    for j in range(len(res)):
        assert res["statistic"][j] == pytest.approx(g["rel_statistic"][j], rel=1e-12, abs=1e-15)
        assert res["effect_size"][j] == pytest.approx(g["rel_effect_size"][j], rel=1e-12)
        x1, x2 = vals[cond == "controls", j], vals[cond == "patients", j]
        _, p, ref_null = O.permutation_paired(x1, x2, 2000, swaps=swaps)
        assert np.allclose(null[j], ref_null, rtol=1e-10, atol=1e-13), j
        assert res["p-value"][j] == pytest.approx(p, abs=1e-15)
    few = pd.concat([dfp.iloc[:6], dfp.iloc[half:half + 6]])
    ex = P.stats.permtest_rel(few, class_column="condition", n_perm=5000)                 # 2^6 = 64 sign patterns: exact
    fv, fc = few.iloc[:, 1:].values, few.condition.values
    for j in range(len(ex)):
        _, p, _ = O.permutation_paired(fv[fc == "controls", j], fv[fc == "patients", j], 5000)
        assert ex["p-value"][j] == pytest.approx(p, abs=1e-15)
    # _compute_stats on a fitted model's tables (two conditions)
    from pyleida import synthetic
    ts, classes = synthetic.make_dataset(12, 20, 90, "states")
    m = P.Leida.from_arrays(ts, classes).fit_predict(TR=0.72, K_min=2, K_max=4, n_init=2, random_state=0, n_perm=200, scores=False)
    t = m.stats(3, "occupancies")
    assert list(t.columns)[:2] == ["k", "variable"] and len(t) == 3 and int(t["k"][0]) == 3
    assert set(m._dynamics_["stats"].keys()) == {"dwell_times", "occupancies"}
    occ = m.occupancies(3)
    mine = O.permtest_table(occ.iloc[:, 2:].values, occ.condition.values, list(occ.columns[2:]), n_perm=200,
                            random_state=np.random.RandomState(0))
    for j, r in enumerate(mine):
        assert t["statistic"][j] == pytest.approx(r["statistic"], rel=1e-11)
        assert t["test"][j] == r["test"] and t["effect_size"][j] == pytest.approx(r["effect_size"], rel=1e-11)


def test_stats_stage_edge_cases(P, O):
    """One condition: the statistics stage is skipped and Leida.stats says so; argument validation follows the
    reference (stats/_stats.py:86-92, 314-318); both randomised p-value rules differ exactly by the +1 adjustment."""
    import pandas as pd
    from pyleida import synthetic
    ts, _ = synthetic.make_dataset(6, 16, 60, "states")
    one = {s: ["A"] for s in ts}
    m = P.Leida.from_arrays(ts, one).fit_predict(TR=1.0, K_min=2, K_max=3, n_init=1, random_state=0, scores=False)
    with pytest.raises(Exception):
        m.stats(2)
    rng = np.random.default_rng(8)
    df = pd.DataFrame(rng.normal(size=(24, 3)), columns=["a", "b", "c"])
    df.insert(0, "condition", ["x"] * 11 + ["y"] * 13)
    with pytest.raises(ValueError):
        P.stats.permtest_ind(df, class_column=None)
    with pytest.raises(ValueError):
        P.stats.permtest_ind(df, class_column="group")
    with pytest.raises(TypeError):
        P.stats.permtest_ind(df, class_column="condition", n_perm=100.0)
    with pytest.raises(ValueError):
        P.stats._compute_stats({"occupancies": {"k_2": df}, "dwell_times": {"k_2": df}}, alternative="both", save_results=False)
    with pytest.raises(ValueError):
        P.stats.permtest_rel(df, class_column="condition")                       # 11 vs 13 rows cannot be paired
    a = P.stats.permtest_ind(df, class_column="condition", n_perm=999, random_state=3)
    b = P.stats.permtest_ind(df, class_column="condition", n_perm=999, random_state=3, adjustment="none")
    assert np.allclose(a["p-value"].values * 1000 - 1, b["p-value"].values * 999, atol=1e-9)
    g = P.stats.hedges_g(df.a.values[:11], df.a.values[11:])
    assert g == O.hedges_g(df.a.values[:11], df.a.values[11:]) and a["effect_size"][0] == pytest.approx(g, rel=1e-12)
