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
