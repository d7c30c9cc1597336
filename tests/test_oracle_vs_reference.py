"""CPU, build container only: the oracle restatement against the UNMODIFIED reference imported
through oracle/ref_shim.py.  Skipped where /root/reference does not exist (the GPU box)."""
import numpy as np
import pytest

import leida_oracle as O
import ref_shim
import synthetic_loader

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree absent")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference()


@pytest.fixture(scope="module")
def syn():
    return synthetic_loader.load()


@pytest.mark.parametrize("N,T,kind", [(14, 40, "states"), (7, 61, "iid"), (30, 36, "iid")])
def test_stage12(ref, syn, N, T, kind):
    x = syn.subject_series(3, N, T, kind)
    ph_r = ref.hilbert_phase(x)
    assert np.abs(np.angle(np.exp(1j * (O.hilbert_phase(x) - ph_r)))).max() < 1e-11
    d_r = ref.phase_coherence(ph_r)
    np.testing.assert_allclose(O.phase_coherence(ph_r), d_r, atol=1e-15)
    le_r = ref.get_eigenvectors(d_r)
    assert np.abs(O.get_eigenvectors(d_r) - le_r).max() < 1e-12
    assert np.abs(O.leading_eigvec_rank2(ph_r) - le_r).max() < 1e-12


@pytest.mark.parametrize("k,seed", [(2, 0), (4, 1), (7, 2)])
def test_kmeans(ref, syn, k, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((500, 18)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    a = ref.KMeansLeida(k=k, n_init=3)
    np.random.seed(seed)
    a.fit(X)
    b = O.KMeansLeida(k=k, n_init=3)
    np.random.seed(seed)
    b.fit(X)
    assert np.array_equal(a.labels_, b.labels_)
    assert np.array_equal(a.cluster_centers_, b.cluster_centers_)
    assert a.inertia_ == pytest.approx(b.inertia_, rel=1e-13)
    assert np.array_equal(a.predict(X[:50]), b.predict(X[:50]))


def test_metrics_group_tables(ref, syn):
    import pandas as pd
    rng = np.random.default_rng(9)
    subs = np.repeat([f"s{i}" for i in range(5)], 30)
    conds = np.repeat(["B", "A", "B", "A", "A"], 30)
    lab = np.repeat(rng.integers(0, 4, size=75), 2)
    meta = pd.DataFrame({"subject_id": subs, "condition": conds})
    t = O.dynamics_tables(subs, conds, lab, TR=2.0)
    assert np.array_equal(ref.fractional_occupancy_group(meta, lab).iloc[:, 2:].values, t["occupancies"])
    assert np.array_equal(ref.dwell_times_group(meta, lab, TR=2.0).iloc[:, 2:].values, t["dwell_times"])
    assert np.array_equal(ref.transition_probabilities_group(meta, lab).iloc[:, 2:].values, t["transitions"])
