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


def _two_group_frame(rng, n1=17, n2=22, nv=5):
    import pandas as pd
    vals = np.vstack([rng.normal(0.2, 1.0, (n1, nv)), rng.normal(0.0, 1.6, (n2, nv))])
    df = pd.DataFrame(vals, columns=[f"PL_state_{i + 1}" for i in range(nv)])
    df.insert(0, "condition", ["b"] * n1 + ["a"] * n2)
    return df


def test_stats_row_against_reference(ref):
    """hedges_g, the Levene decision, the observed statistics and the table layout of the oracle's
    permtest_table equal the unmodified reference functions (permtest_ind runs through the ttest_ind signature
    shim, see ref_shim._compat_ttest_ind); permutation p-values are random in the reference and are compared
    within 6 standard errors of a 5000-permutation estimate."""
    import contextlib, io
    if ref.permtest_rel is None:
        pytest.skip("reference stats module did not import")
    rng = np.random.default_rng(21)
    df = _two_group_frame(rng)
    x1, x2 = df[df.condition == "a"].iloc[:, 1].values, df[df.condition == "b"].iloc[:, 1].values
    for paired in (False, True):
        assert O.hedges_g(x1[:15], x2[:15], paired) == ref.hedges_g(x1[:15], x2[:15], paired)
    cols = [c for c in df.columns if c != "condition"]
    with contextlib.redirect_stdout(io.StringIO()):
        got_ref = ref.permtest_ind(df, class_column="condition", n_perm=5000)
    mine = O.permtest_table(df[cols].values, df.condition.values, cols, n_perm=5000, random_state=np.random.RandomState(4))
    assert list(got_ref.columns) == list(mine[0].keys())
    for j, r in enumerate(mine):
        for c in ("variable", "group_1", "group_2", "test"):
            assert got_ref[c].values[j] == r[c], (j, c)
        assert got_ref["statistic"].values[j] == pytest.approx(r["statistic"], rel=1e-13)
        assert got_ref["effect_size"].values[j] == r["effect_size"]
        assert got_ref["alpha_Bonferroni"].values[j] == r["alpha_Bonferroni"]
        p = r["p-value"]
        assert abs(got_ref["p-value"].values[j] - p) <= 6 * np.sqrt(max(p * (1 - p), 1e-3) / 5000)
    # paired
    dfp = df.iloc[list(range(15)) + list(range(17, 32))]
    with contextlib.redirect_stdout(io.StringIO()):
        rel_ref = ref.permtest_rel(dfp, class_column="condition", n_perm=5000)
    rel = O.permtest_table(dfp[cols].values, dfp.condition.values, cols, n_perm=5000, paired=True,
                           random_state=np.random.default_rng(5))
    assert list(rel_ref.columns) == list(rel[0].keys())
    for j, r in enumerate(rel):
        assert rel_ref["statistic"].values[j] == pytest.approx(r["statistic"], rel=1e-13, abs=1e-16)
        assert rel_ref["effect_size"].values[j] == r["effect_size"]
        p = r["p-value"]
        assert abs(rel_ref["p-value"].values[j] - p) <= 6 * 2 * np.sqrt(max(p * (1 - p), 1e-3) / 5000)
