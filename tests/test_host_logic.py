"""CPU tests of the host logic that decides what the kernels are asked to do: restart selection and relabelling
(clustering/_clustering.py:139-146, 234-247), the sweep's draw order and de-duplication (:101-105, 318-329), metric row
groups (dynamics_metrics/_dynamics_metrics.py:180-182) and the filter design handed to the clean kernel.  The package
imports on CPU (the library loads; only kernel launches need a GPU)."""
import numpy as np
import pytest

import leida_oracle as O


@pytest.fixture(scope="module")
def K():
    from pyleida.clustering import kmeans
    return kmeans


def test_remap_permutation_matches_the_reference_rule(K):
    rng = np.random.default_rng(0)
    for k in (2, 3, 7, 20):
        labels = rng.integers(0, k, 500)
        labels[:k] = np.arange(k)
        counts = np.bincount(labels, minlength=k)
        lut, order = K._remap_permutation(counts)
        km = O.KMeansLeida(k=k)
        km.labels_, km.cluster_centers_ = labels.copy(), np.arange(k, dtype=np.float32)[:, None] * np.ones((1, 3), np.float32)
        km._remap_labels()
        assert np.array_equal(lut[labels], km.labels_)
        assert np.array_equal(np.arange(k, dtype=np.float32)[order], km.cluster_centers_[:, 0])
    # ties between equally populated clusters resolve like numpy's argsort in the reference
    lut, order = K._remap_permutation(np.array([5, 9, 5, 9]))
    assert sorted(lut.tolist()) == [0, 1, 2, 3] and sorted(order.tolist()) == [0, 1, 2, 3]


def test_select_best_is_first_strictly_smaller(K):
    assert K._select_best([3.0, 2.0, 2.0, 5.0]) == 1
    assert K._select_best([1.0, 1.0]) == 0
    with pytest.warns(UserWarning):
        assert K._select_best([np.nan, 4.0, 3.0]) == 2
    with pytest.raises(RuntimeError):
        K._select_best([np.nan, np.nan])


def test_sweep_runs_draw_order_and_dedup(K):
    # random_state=int re-seeds before every restart: all restarts of a K are identical and run once
    runs, plan = K.sweep_runs(1000, 2, 5, 4, random_state=3)
    assert len(runs) == 4 and all(plan[k] == [plan[k][0]] * 4 for k in range(2, 6))
    np.random.seed(3)
    assert np.array_equal(runs[0][1], np.random.choice(1000, 2, replace=False))
    # random_state=None consumes the global stream K-major; above 30 000 rows the score sample is drawn first (:318-320)
    np.random.seed(11)
    runs, plan = K.sweep_runs(40_000, 2, 3, 2, random_state=None)
    np.random.seed(11)
    sample = np.random.choice(np.arange(40_000), size=30_000, replace=False)
    first = np.random.choice(40_000, 2, replace=False)
    assert np.array_equal(K.sweep_runs.last_sample, sample) and np.array_equal(runs[0][1], first)
    assert [len(plan[k]) for k in (2, 3)] == [2, 2] and len(runs) == 4
    # explicit indices: no draw, no sample
    runs, plan = K.sweep_runs(100, 2, 2, 2, init_indices={2: [np.array([1, 2]), np.array([1, 2])]})
    assert len(runs) == 1 and plan[2] == [0, 0] and K.sweep_runs.last_sample is None


def test_segments_order_and_integer_ids():
    from pyleida.dynamics_metrics.metrics import Segments
    sub = np.array([10, 10, 2, 2, 2, 33, 10])
    cond = np.array(["b", "b", "a", "a", "b", "a", "a"])
    seg = Segments.from_arrays(sub, cond)
    # sorted condition, then sorted subject -- integer ids sort numerically (2 < 10 < 33), as np.unique does in the reference
    assert list(zip(seg.conditions, seg.subjects)) == [("a", 2), ("a", 10), ("a", 33), ("b", 2), ("b", 10)]
    groups = [seg.rows[seg.off[g]:seg.off[g + 1]].tolist() for g in range(len(seg))]
    assert groups == [[2, 3], [6], [5], [4], [0, 1]]
    c = Segments.contiguous(np.array([7, 3]), np.array(["x", "x"]), [4, 2])
    assert c.subjects == [3, 7] and c.rows.tolist() == [4, 5, 0, 1, 2, 3] and c.off.tolist() == [0, 2, 6]


def test_butterworth_design_matches_the_oracle():
    from pyleida.signal_tools.clean import butterworth_design
    from scipy import signal
    for kw in (dict(low_pass=0.1, high_pass=0.01), dict(low_pass=0.08, high_pass=None), dict(low_pass=None, high_pass=0.02)):
        b, a, zi = butterworth_design(0.72, **kw)
        rb, ra = O.butterworth_ba(0.72, **kw)
        assert np.array_equal(b, rb) and np.array_equal(a, ra) and np.array_equal(zi, signal.lfilter_zi(rb, ra))
    assert butterworth_design(1.0) is None
    with pytest.raises(ValueError):
        butterworth_design(1.0, low_pass=0.01, high_pass=0.1)


def test_results_folder_binary_roundtrip_cpu(tmp_path):
    """results_io writes the reference's layout (+ the binary copies) and pyleida.DataLoader reads it back -- file
    access only, no GPU and no reference tree needed."""
    import pickle
    import pandas as pd
    import pyleida
    from pyleida import results_io
    rng = np.random.default_rng(0)
    M, N = 60, 5
    subs = np.repeat(["s1", "s2", "s3"], 20)
    conds = np.repeat(["A", "B", "A"], 20)
    eig = pd.DataFrame(rng.standard_normal((M, N)), columns=[f"r{i}" for i in range(N)])
    eig.insert(0, "subject_id", subs)
    eig.insert(1, "condition", conds)
    pred = pd.DataFrame({"subject_id": subs, "condition": conds, "k_2": rng.integers(0, 2, M), "k_3": rng.integers(0, 3, M)})
    models = {}
    for k in (2, 3):
        km = pyleida.clustering.KMeansLeida(k=k)
        km.cluster_centers_, km.labels_, km.inertia_, km._is_fitted = rng.standard_normal((k, N)).astype(np.float32), pred[f"k_{k}"].values, 1.0, True
        models[f"k_{k}"] = km
    perf = pd.DataFrame({"k": [2, 3], "Dunn_score": [0.1, 0.2], "distortion": [1.0, 0.5], "silhouette_score": [0.3, 0.2],
                         "Davies-Bouldin_score": [1.0, 1.1]})
    ids = pd.DataFrame({"subject_id": ["s1", "s3", "s2"], "condition": ["A", "A", "B"]})
    dyn = {kind: {f"k_{k}": pd.concat([ids, pd.DataFrame(rng.random((3, k)), columns=[f"PL_state_{i + 1}" for i in range(k)])], axis=1)
                  for k in (2, 3)} for kind in ("occupancies", "dwell_times")}
    dyn["transitions"] = {f"k_{k}": pd.concat([ids, pd.DataFrame(rng.random((3, k * k)))], axis=1) for k in (2, 3)}
    data = tmp_path / "data"
    data.mkdir()
    with open(data / "time_series.pkl", "wb") as f:
        pickle.dump({s: rng.standard_normal((N, 22)) for s in ("s1", "s2", "s3")}, f)
    with open(data / "metadata.pkl", "wb") as f:
        pickle.dump({"s1": ["A"], "s2": ["B"], "s3": ["A"]}, f)
    (data / "rois_labels.txt").write_text("\n".join(f"r{i}" for i in range(N)) + "\n")
    for mode, kw in (("csv", {}), ("binary", dict(binary=True)), ("fast", dict(binary=True, csv=False))):
        out = str(tmp_path / mode)
        results_io.save_eigenvectors(out, eig, **kw)
        results_io.save_clustering(out, pred, perf, models, **kw)
        results_io.save_dynamics(out, dyn)
        dl = pyleida.DataLoader(str(data), out)
        got = dl.eigenvectors.iloc[:, 2:].values
        assert np.allclose(got, eig.iloc[:, 2:].values, rtol=0, atol=0 if mode != "csv" else 1e-12)
        assert list(dl.eigenvectors.columns) == list(eig.columns) and list(dl.eigenvectors.subject_id) == list(subs)
        assert np.array_equal(dl.predictions["k_3"].values, pred["k_3"].values) and (dl._K_min_, dl._K_max_) == (2, 3)
        assert np.array_equal(dl.load_model(2).cluster_centers_, models["k_2"].cluster_centers_)
        assert dl.load_centroids(3).shape == (3, N) and dl.occupancies(2).shape == (3, 4) and dl.transitions(3).shape == (3, 11)
        assert dl._classes_lst_ == ["A", "B"] and len(dl.time_series()) == 3
        with pytest.raises(ValueError):
            dl.load_model(7)
    with pytest.raises(ValueError):
        pyleida.DataLoader(str(data), str(tmp_path / "missing"))
