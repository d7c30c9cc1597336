"""GPU parity tests at the sizes the benchmark runs at (BASELINE configs 2 and 3), through the C-ABI:
  * full Lloyd runs to convergence on M = 119 800 x 116 rows against the CPU oracle (labels, pass counts, inertia);
  * every pass's labels on M ~ 200 000 x 400 rows against an independent fp64 brute-force Lloyd (this is what validates
    the tensor-core trust threshold at N = 400);
  * the measured error of the tensor-core score against fp64, which the threshold is derived from;
  * compute_dynamics_metrics called directly on a predictions DataFrame against the golden tables;
  * unnormalised rows (the engine's power-of-two range scaling).
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


def _rows(n_subjects, n_rois, kind):
    """fp32 eigenvector rows of a synthetic data set, made by the CUDA stage 1+2 (device tensor, host copy)."""
    import torch
    from pyleida import synthetic
    from pyleida.signal_tools.phase import leading_eigenvectors
    ts, _ = synthetic.make_shard(n_subjects, n_rois, 1200, 0, n_subjects, kind)
    x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
    _, xf = leading_eigenvectors(x, want_f64=False)
    return xf, xf[:, :n_rois].cpu().numpy()


def _same_partition_fraction(got, ref, k):
    """Fraction of rows on which `got` (any cluster numbering) induces the partition `ref`: clusters are matched through
    the first row of each `got` cluster."""
    u, first = np.unique(got, return_index=True)
    lut = np.full(k, -1, dtype=np.int64)
    lut[u] = ref[first]
    return float((lut[got] == ref).mean())


@pytest.mark.parametrize("kind,ks,n_iter", [("states", (5, 12, 20), 1000), ("iid", (4, 11), 30)])
def test_config2_size_full_lloyd_vs_oracle(P, O, kind, ks, n_iter):
    """M = 119 800 x 116 (the config-2 data set): whole Lloyd runs against oracle.KMeansLeida on the SAME fp32 rows and
    the same initial rows.  centroid_update='reference' (numpy's sequential float32 member sums, bit for bit): identical
    labels, identical iteration counts, inertia to 1e-6.  centroid_update='exact' (the production mode: exactly rounded
    means, shard-count invariant) differs from the reference only through the rounding noise of the reference's own
    float32 sums (~1e-6 relative on a centroid at this size), which can move rows that sit on a cluster boundary to
    within that noise: on clustered rows at least 99.9 % of the rows end in the same cluster.  iid rows (the worst case
    for margins and iteration counts; no cluster structure, so Lloyd trajectories amplify any rounding difference) are
    capped at 30 iterations on both sides and compared strictly in 'reference' mode only."""
    from pyleida.clustering.kmeans import KMeansEngine
    xf, X = _rows(100, 116, kind)
    eng = KMeansEngine(xf, n_features=116)
    rng = np.random.default_rng(5)
    inits = [rng.choice(X.shape[0], size=k, replace=False).astype(np.int64) for k in ks]
    runs = [(k, idx) for k, idx in zip(ks, inits)]
    strict = eng.run(runs, n_iter=n_iter, centroid_update="reference")
    exact = eng.run(runs, n_iter=n_iter)
    assert strict.tau_watch <= 0.5 and exact.tau_watch <= 0.5
    for i, (k, idx) in enumerate(runs):
        ref = O.KMeansLeida(k=k, n_init=1, n_iter=n_iter).fit(X, init_indices=[idx])
        got = strict.labels_device(i).cpu().numpy()
        assert _same_partition_fraction(got, ref.labels_, k) == 1.0, (kind, k)
        assert strict.passes(i) == ref.n_iter_runs_[0] + 1, (kind, k)         # + the initial assignment pass
        assert strict.inertia(i) == pytest.approx(float(ref.inertia_), rel=1e-6)
        if kind == "states":
            frac = _same_partition_fraction(exact.labels_device(i).cpu().numpy(), ref.labels_, k)
            assert frac >= 0.999, (kind, k, frac)
            assert exact.inertia(i) == pytest.approx(float(ref.inertia_), rel=1e-4)
        else:
            assert exact.inertia(i) == pytest.approx(float(ref.inertia_), rel=1e-2)


def _lloyd_fp64(X32, init_rows, n_passes):
    """Independent brute-force Lloyd in torch fp64 (clustering/_clustering.py:105-133): labels after the initial
    assignment and after each of n_passes iterations."""
    import torch
    X = X32.double()
    xn = X.norm(dim=1)
    C = X32[init_rows].clone()                                             # fp32 centroids (:105-108)
    out = []
    for it in range(n_passes + 1):
        Cd = C.double()
        d = 1.0 - (X @ Cd.T) / (xn[:, None] * Cd.norm(dim=1)[None, :])
        lab = d.argmin(dim=1)
        out.append(lab)
        K = C.shape[0]
        sums = torch.zeros((K, X.shape[1]), dtype=torch.float64, device=X.device).index_add_(0, lab, X)
        cnt = torch.bincount(lab, minlength=K).double()
        C = (sums / cnt[:, None]).float()                                  # member means rounded to fp32 (:120-125)
    return out


@pytest.mark.parametrize("kind", ["states", "iid"])
def test_config3_shape_every_pass_vs_fp64_bruteforce(P, kind):
    """M = 203 660 x 400 (config-3 row width), K in {7, 20}: the labels after EVERY one of the first 10 Lloyd passes
    equal an independent fp64 brute force -- the tensor-core path (trust threshold + fp64 re-check of the candidate
    clusters) is the fp64 argmin, pass by pass, at N = 400."""
    import torch
    from pyleida.clustering.kmeans import KMeansEngine
    xf, _ = _rows(170, 400, kind)
    eng = KMeansEngine(xf, n_features=400)
    rng = np.random.default_rng(11)
    runs = [(k, rng.choice(eng.M, size=k, replace=False).astype(np.int64)) for k in (7, 20)]
    X32 = xf[:, :400].contiguous()
    want = [_lloyd_fp64(X32, torch.from_numpy(idx).cuda(), 10) for _, idx in runs]
    for p in range(0, 11):
        res = eng.run(runs, n_iter=p)
        assert res.tau_watch <= 0.5
        for i in range(len(runs)):
            got = res.labels_device(i).long()
            bad = int((got != want[i][p]).sum())
            assert bad == 0, (kind, runs[i][0], p, bad)


def test_tc_score_error_within_bound(P):
    """The trust threshold must cover the measured tensor-core score error with the stated safety factor: observed
    max |S_tc - S_fp64| / |x| over ~5e7 scores per case, times 2 (two scores are compared) times 2 (safety) <= tau."""
    import importlib.util
    import torch
    from pyleida import _native as nv, synthetic
    from pyleida.clustering.kmeans import KMeansEngine, sweep_runs
    from pyleida.signal_tools.phase import leading_eigenvectors
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("tc_score_error", os.path.join(root, "tools", "tc_score_error.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    for N, kind in ((116, "iid"), (400, "states"), (1000, "iid")):
        S = 12
        ts, _ = synthetic.make_shard(S, N, 1200, 0, S, kind)
        x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
        _, xf = leading_eigenvectors(x, want_f64=False)
        eng = KMeansEngine(xf, n_features=N)
        inits = synthetic.init_indices(eng.M, 2, 20, 15, seed=0)
        runs, _ = sweep_runs(eng.M, 2, 20, 15, None, inits)
        res = eng.run(runs, n_iter=2)
        worst, rms, q, pairs = tool.score_errors(eng, res._cent, res._ks.astype(np.int64), res._crow0.astype(np.int64))
        tau = float(nv.lib.leida_kmeans_tc_tau(N))
        assert 4.0 * worst <= tau, (N, kind, worst, tau)


def test_compute_dynamics_metrics_direct_golden(P):
    """compute_dynamics_metrics(predictions DataFrame) (dynamics_metrics/_dynamics_metrics.py:11-105) called directly,
    against the tables the unmodified reference produced for the same labels (tests/golden/pipeline_case.npz)."""
    import pandas as pd
    g = np.load(os.path.join(GOLD, "pipeline_case.npz"), allow_pickle=True)
    ks = [int(k) for k in g["perf_k"]]
    df = pd.DataFrame({"subject_id": g["subject_id"], "condition": g["condition"]})
    for k in ks:
        df[f"k_{k}"] = g[f"labels_k{k}"]
    out = P.dynamics_metrics.compute_dynamics_metrics(df, TR=0.72)
    assert set(out.keys()) == {"dwell_times", "occupancies", "transitions"}
    for k in ks:
        for name in ("occupancies", "dwell_times", "transitions"):
            tab = out[name][f"k_{k}"]
            assert np.array_equal(tab.iloc[:, 2:].values, g[f"{name}_k{k}"]), (k, name)
            assert list(tab.columns[2:]) == list(g[f"{name}_k{k}_cols"])
            assert list(tab.subject_id) == list(g[f"{name}_k{k}_subject"])
            assert list(tab.condition) == list(g[f"{name}_k{k}_condition"])
    # the reference renames the first two columns in place when they are not called subject_id / condition (:58-61)
    df2 = df.rename(columns={"subject_id": "id", "condition": "group"})
    out2 = P.dynamics_metrics.compute_dynamics_metrics(df2)
    assert list(df2.columns[:2]) == ["subject_id", "condition"]
    assert np.array_equal(out2["occupancies"][f"k_{ks[0]}"].iloc[:, 2:].values, g[f"occupancies_k{ks[0]}"])


def test_unnormalised_rows_are_scaled_into_range(P, O):
    """Rows far outside [-1, 1] (the advisor's 'amplitude 1e4' case) and tiny rows: the engine applies one power-of-two
    factor, so labels equal the oracle's and the centroids come back in the caller's units."""
    rng = np.random.default_rng(21)
    base = rng.standard_normal((6000, 24)).astype(np.float32)
    for amp in (1.0e4, 3.0e-7):
        X = (base * np.float32(amp)).astype(np.float32)
        idx = [rng.choice(X.shape[0], size=5, replace=False)]
        km = P.clustering.KMeansLeida(k=5, n_init=1).fit(X, init_indices=idx)
        ref = O.KMeansLeida(k=5, n_init=1).fit(X, init_indices=idx)
        assert np.array_equal(km.labels_, ref.labels_), amp
        # the oracle's centroids carry the rounding noise of numpy's sequential float32 sums
        assert np.allclose(km.cluster_centers_, ref.cluster_centers_, rtol=0, atol=2e-6 * np.abs(ref.cluster_centers_).max()), amp
        assert km.inertia_ == pytest.approx(float(ref.inertia_), rel=1e-6)


@pytest.mark.parametrize("opts", [
    dict(detrend=True, standardize="zscore", filter_type="butterworth", low_pass=0.1, high_pass=0.01, TR=0.72),
    dict(detrend=True, standardize="psc", filter_type="butterworth", low_pass=0.08, high_pass=None, TR=2.0),
    dict(detrend=False, standardize=False, filter_type="butterworth", low_pass=None, high_pass=0.02, TR=1.0),
    dict(detrend=True, standardize="zscore", filter_type=None, low_pass=None, high_pass=None, TR=None),
    dict(detrend=False, standardize="zscore", filter_type=None, low_pass=None, high_pass=None, TR=None),
])
def test_clean_signals_vs_oracle(P, O, opts):
    """clean_signals (signal_tools/_signal_tools.py:66-133 -> nilearn.signal.clean 0.9.1) against the oracle's
    restatement, which runs the real scipy.signal.filtfilt.  Parity with nilearn itself is unpinned (not installed).
    Tolerance: 1e-8 of the output scale, except for the 10th-order band-pass in transfer-function form that nilearn
    uses (butter(5, [lo, hi], output='ba')): that recursion is ill-conditioned -- scipy's own filtfilt moves by 4e-6
    when its input is perturbed by 1e-15, and differs from the second-order-sections form by 3.7e-6 (measured on this
    input) -- so there the comparison is to 2e-5, the reference algorithm's own noise floor."""
    rng = np.random.default_rng(4)
    t = np.arange(400)
    sig = {"a": 100 + 3 * rng.standard_normal((7, 400)) + 0.01 * t + 2 * np.sin(2 * np.pi * 0.03 * t),
           "b": 50 + rng.standard_normal((7, 400)), "c": 10 + rng.standard_normal((5, 150)) - 0.02 * np.arange(150)}
    got = P.signal_tools.clean_signals(sig, **opts)
    ref = O.clean_signals(sig, **opts)
    assert list(got.keys()) == list(sig.keys())
    for s in sig:
        assert got[s].shape == sig[s].shape and got[s].dtype == np.float64
        scale = max(1.0, float(np.abs(ref[s]).max()))
        band = opts["low_pass"] is not None and opts["high_pass"] is not None
        assert np.abs(got[s] - ref[s]).max() <= (2e-5 if band else 1e-8) * scale, (s, np.abs(got[s] - ref[s]).max())


def test_clean_signals_errors_and_pipeline(P, O):
    with pytest.raises(ValueError):
        P.signal_tools.clean_signals([1, 2, 3])
    x = {"s": np.random.default_rng(0).standard_normal((3, 200))}
    with pytest.raises(ValueError):
        P.signal_tools.clean_signals(x, filter_type="butterworth", low_pass=0.01, high_pass=0.1, TR=1.0)
    with pytest.raises(ValueError):
        P.signal_tools.clean_signals(x, filter_type="butterworth", low_pass=0.1, TR=None)
    with pytest.raises(ValueError):
        P.signal_tools.clean_signals({"s": x["s"][:, :20]}, filter_type="butterworth", low_pass=0.1, high_pass=0.01, TR=1.0)
    # band-limited signals through the phase -> eigenvector path: same eigenvectors as the oracle on the oracle-cleaned series
    from pyleida import synthetic
    ts, _ = synthetic.make_dataset(3, 12, 220, "states")
    kw = dict(detrend=True, standardize="zscore", filter_type="butterworth", low_pass=0.1, high_pass=0.01, TR=0.72)
    cg = P.signal_tools.clean_signals(ts, **kw)
    co = O.clean_signals(ts, **kw)
    for s in ts:
        lei, _ = P.signal_tools.leading_eigenvectors(cg[s])
        ref = O.leading_eigvec_rank2(O.hilbert_phase(co[s]))
        assert np.abs(lei - ref).max() <= 1e-4         # band-pass conditioning (see test_clean_signals_vs_oracle)


def test_results_folder_binary_and_dataloader(P, tmp_path, monkeypatch):
    """save_results='binary' writes the reference's folder plus .npy copies of the volume-sized tables; 'fast' writes
    the .npy copies only.  pyleida.DataLoader (the reference's accessor names, _data_loader.py:146-345) reads both."""
    import pickle
    from pyleida import synthetic
    ts, classes = synthetic.make_dataset(4, 8, 30, "states")
    data = tmp_path / "data"
    data.mkdir()
    with open(data / "time_series.pkl", "wb") as f:
        pickle.dump(ts, f)
    with open(data / "metadata.pkl", "wb") as f:
        pickle.dump(classes, f)
    (data / "rois_labels.txt").write_text("\n".join(f"r{i}" for i in range(8)) + "\n")
    for mode in ("binary", "fast"):
        work = tmp_path / mode
        work.mkdir()
        monkeypatch.chdir(work)
        m = P.Leida(str(data)).fit_predict(TR=1.0, save_results=mode, random_state=0, K_min=2, K_max=3, n_init=1)
        root = work / "LEiDA_results"
        assert (root / "eigenvectors.npy").exists() and (root / "clustering" / "predictions.npy").exists()
        assert (root / "eigenvectors.csv").exists() == (mode == "binary")
        dl = P.DataLoader(str(data), str(root))
        assert np.array_equal(dl.eigenvectors.iloc[:, 2:].values, m.eigenvectors.iloc[:, 2:].values)
        assert list(dl.eigenvectors.columns) == list(m.eigenvectors.columns)
        assert np.array_equal(dl.predictions["k_3"].values, m.predictions["k_3"].values)
        assert np.array_equal(dl.load_model(3).cluster_centers_, m.load_model(3).cluster_centers_)
        assert np.allclose(dl.occupancies(2).iloc[:, 2:].values, m.occupancies(2).iloc[:, 2:].values, rtol=1e-12)
        d = dl.centroids_distances(3)                       # KMeansLeida.transform on the GPU from the loaded model
        assert np.allclose(d.iloc[:, 2:].values, m.centroids_distances(3).iloc[:, 2:].values, atol=1e-6)
        assert type(dl.load_model(2)).__module__ == "pyleida.clustering._clustering"


def test_legacy_update_and_distortion_kernels_agree(P, monkeypatch):
    """The two-kernel update (LEIDA_UPDATE_LEGACY=1) and the gather distortion kernel (LEIDA_DISTORTION_LEGACY=1) stay
    available as fallbacks (rows wider than 1024 / 512 columns, > 4096 runs): same labels, pass counts and centroids as
    the fused update, inertia equal to the last bits' worth of a different fp64 summation order."""
    from pyleida.clustering.kmeans import KMeansEngine
    rng = np.random.default_rng(17)
    X = rng.standard_normal((20_000, 150)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    eng = KMeansEngine(X)
    runs = [(k, rng.choice(X.shape[0], size=k, replace=False).astype(np.int64)) for k in (3, 9, 20, 33)]
    a = eng.run(runs, n_iter=25)
    monkeypatch.setenv("LEIDA_UPDATE_LEGACY", "1")
    monkeypatch.setenv("LEIDA_DISTORTION_LEGACY", "1")
    b = eng.run(runs, n_iter=25)
    for i in range(len(runs)):
        assert bool((a.labels_device(i) == b.labels_device(i)).all()) and a.passes(i) == b.passes(i)
        assert bool((a.centers_device(i) == b.centers_device(i)).all())
        assert a.inertia(i) == pytest.approx(b.inertia(i), rel=1e-12)


def test_dunn_fast_by_name(P, O):
    """dunn_fast(points, labels) (clustering/_clustering.py:934-973) is exported under the reference's name."""
    rng = np.random.default_rng(2)
    cent = rng.standard_normal((4, 21))
    lab = rng.integers(0, 4, 1500)
    pts = (cent[lab] + 0.7 * rng.standard_normal((1500, 21))).astype(np.float32)
    names = np.array([7, 3, 11, 20])[lab]                      # arbitrary label values
    assert P.clustering.dunn_fast(pts, names) == pytest.approx(O.dunn_fast(pts, names), rel=2e-4)


@pytest.mark.parametrize("amp", [1.0, 250.0])
def test_kmeans_euclidean_metric_vs_oracle(P, O, amp):
    """KMeansLeida(metric='euclidean') (accepted by the reference, clustering/_clustering.py:61-75; cdist 'euclidean' at
    :111,127,183,228): labels, centroids, inertia, transform and predict against the oracle (scipy cdist) -- on rows
    inside [-1, 1] and on rows that need the engine's range scaling."""
    rng = np.random.default_rng(31)
    cent = rng.standard_normal((6, 40))
    lab = rng.integers(0, 6, 8000)
    X = ((cent[lab] + 0.8 * rng.standard_normal((8000, 40))) * 0.1 * amp).astype(np.float32)
    for k in (3, 6, 11):
        idx = [rng.choice(X.shape[0], size=k, replace=False) for _ in range(2)]
        km = P.clustering.KMeansLeida(k=k, metric="euclidean", n_init=2).fit(X, init_indices=idx)
        ref = O.KMeansLeida(k=k, metric="euclidean", n_init=2).fit(X, init_indices=idx)
        assert np.array_equal(km.labels_, ref.labels_), (amp, k)
        assert np.allclose(km.cluster_centers_, ref.cluster_centers_, rtol=0, atol=3e-6 * np.abs(ref.cluster_centers_).max())
        assert km.inertia_ == pytest.approx(float(ref.inertia_), rel=1e-6)
        Y = X[:500] + np.float32(0.01 * amp)
        assert np.allclose(km.transform(Y), ref.transform(Y), rtol=1e-6, atol=1e-6 * amp)
        assert np.array_equal(km.predict(Y), ref.predict(Y))
    with pytest.raises(ValueError):
        P.clustering.KMeansLeida(k=3, metric="manhattan")


@pytest.mark.parametrize("N", [40, 116, 150, 270, 400, 500])
def test_distortion_forms_bit_identical(P, monkeypatch, N):
    """The three forms of the distortion kernel (kmeans.cu: pipelined bulk-copy ring, cp.async ring, gather) use the same
    per-row fp64 arithmetic: the exact-integer sums -- hence inertia_ -- are bit-identical between the two ring forms
    and equal to the gather form's.  Rows with persistent labels (quads with one label / one transition) and rows with
    unrelated labels (general path), a row count that is not a multiple of the tile, every column-group count class."""
    from pyleida.clustering.kmeans import KMeansEngine
    rng = np.random.default_rng(N)
    M = 12_345
    cent = rng.standard_normal((7, N))
    seq = np.repeat(rng.integers(0, 7, M // 5 + 1), 5)[:M]               # dwell of 5 volumes: AAAB / AABB / ABBB quads
    seq[M // 2:] = rng.integers(0, 7, M - M // 2)                         # second half: unrelated labels
    X = (cent[seq] + 0.8 * rng.standard_normal((M, N))).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    eng = KMeansEngine(X)
    runs = [(k, rng.choice(M, size=k, replace=False).astype(np.int64)) for k in (2, 5, 7, 12, 20, 3, 9, 16, 20)]
    out = {}
    for form in ("piped", "ring", "gather"):
        monkeypatch.delenv("LEIDA_DISTORTION_LEGACY", raising=False)
        if form == "gather":
            monkeypatch.setenv("LEIDA_DISTORTION_LEGACY", "1")
        else:
            monkeypatch.setenv("LEIDA_DISTORTION_FORM", form)
        res = eng.run(runs, n_iter=4)
        out[form] = [res.inertia(i) for i in range(len(runs))]
    # (a run that ends with an empty cluster has a NaN centroid and a NaN inertia in every form, like the reference)
    assert np.array_equal(np.array(out["piped"]), np.array(out["ring"]), equal_nan=True)
    np.testing.assert_allclose(out["piped"], out["gather"], rtol=1e-12, atol=0, equal_nan=True)
    assert np.isfinite(out["piped"]).sum() >= 5
