"""CPU: the oracle restatement (oracle/leida_oracle.py) against the committed golden fixtures,
which are outputs of the UNMODIFIED reference functions (see oracle/make_golden.py)."""
import os

import numpy as np
import pytest

import leida_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def wrapped(a, b):
    return np.abs(np.angle(np.exp(1j * (a - b))))


@pytest.fixture(scope="module")
def sig():
    return np.load(os.path.join(GOLD, "signal_cases.npz"))


def test_phase_matches_reference(sig):
    for i in range(int(sig["n_cases"])):
        err = wrapped(O.hilbert_phase(sig[f"x_{i}"]), sig[f"phase_{i}"]).max()
        assert err <= 1e-5 * np.pi, (i, err)
        assert err <= 1e-11      # in practice


def test_phase_coherence_matches_reference(sig):
    for i in range(int(sig["n_cases"])):
        if f"dfc_{i}" in sig:
            np.testing.assert_allclose(O.phase_coherence(sig[f"phase_{i}"]), sig[f"dfc_{i}"], atol=1e-15)


def test_eigenvectors_dense_and_closed_form(sig):
    for i in range(int(sig["n_cases"])):
        ph, ref = sig[f"phase_{i}"], sig[f"lei_{i}"]
        tol = 1e-5 * np.abs(ref).max(axis=1, keepdims=True)
        assert np.all(np.abs(O.leading_eigvec_rank2(ph) - ref) <= tol)
        assert np.abs(O.leading_eigvec_rank2(ph) - ref).max() < 1e-12
        if ph.shape[0] <= 12:
            assert np.abs(O.get_eigenvectors(O.phase_coherence(ph)) - ref).max() < 1e-12


def test_sign_rule_tie_branch(sig):
    got = O.leading_eigvec_rank2(sig["tie_phase"])
    np.testing.assert_allclose(got, sig["tie_lei"], atol=1e-13)
    # the fixture really exercises the exact-half branch
    assert any((row > 0).sum() * 2 == row.size for row in sig["tie_lei"])


def test_kmeans_matches_reference():
    g = np.load(os.path.join(GOLD, "kmeans_cases.npz"))
    for c in range(int(g["n_cases"])):
        k, n_init, seed, is_int = (int(v) for v in g[f"case{c}_meta"])
        X = g["X_" + str(g[f"case{c}_data"])]
        km = O.KMeansLeida(k=k, n_init=n_init)
        if is_int:
            km.fit(X, random_state=seed)
        else:
            np.random.seed(seed)
            km.fit(X, random_state=None)
        assert np.array_equal(km.labels_, g[f"case{c}_labels"]), c
        assert np.array_equal(km.cluster_centers_, g[f"case{c}_centers"]), c
        assert km.inertia_ == pytest.approx(float(g[f"case{c}_inertia"]), rel=1e-12)


def test_sequential_restatements_of_native_arithmetic():
    from scipy.spatial.distance import cdist
    rng = np.random.default_rng(3)
    A = rng.standard_normal((200, 37)).astype(np.float32)
    B = rng.standard_normal((6, 37)).astype(np.float32)
    # scipy's build may vectorise the dot product (2-wide partial sums): equal to the last ulp or two
    np.testing.assert_allclose(O.cdist_cosine_sequential(A, B), cdist(A, B, "cosine"), rtol=0, atol=1e-15)
    R = (rng.standard_normal((30000, 16)) * 0.05 + 0.01).astype(np.float32)
    sel = R[rng.random(30000) < 0.4]
    assert np.array_equal(sel.mean(axis=0), O.mean_rows_f32_sequential(sel))      # bit-exact


def test_single_sequence_metrics_bit_exact():
    g = np.load(os.path.join(GOLD, "metric_cases.npz"))
    for i in range(int(g["n_cases"])):
        lab, k = g[f"labels_{i}"], int(g[f"k_{i}"])
        assert np.array_equal(O.fractional_occupancy_table(lab, k), g[f"occ_{i}"])
        assert np.array_equal(O.dwell_time_table(lab, k, TR=0.72), g[f"dwell_tr_{i}"])
        assert np.array_equal(O.dwell_time_table(lab, k, TR=None), g[f"dwell_{i}"])
        assert np.array_equal(O.transition_probability_table(lab, k), g[f"trans_{i}"])
        assert np.array_equal(O.transition_counts(lab, k).astype(float), g[f"trans_raw_{i}"])


def test_pipeline_matches_reference():
    import sys
    import synthetic_loader
    syn = synthetic_loader.load()
    g = np.load(os.path.join(GOLD, "pipeline_case.npz"))
    S, N, T = int(g["S"]), int(g["N"]), int(g["T"])
    ts, classes = syn.make_dataset(S, N, T, "states")
    res = O.execute_all(ts, classes, TR=0.72, K_min=2, K_max=6, n_init=3, random_state=0)
    assert np.abs(res["eigenvectors"] - g["lei"]).max() < 1e-12
    for k in range(2, 7):
        assert np.array_equal(res["models"][k].labels_, g[f"labels_k{k}"])
        assert np.array_equal(res["models"][k].cluster_centers_, g[f"centers_k{k}"])
        for m in ("occupancies", "dwell_times", "transitions"):
            assert np.array_equal(res["metrics"][k][m], g[f"{m}_k{k}"]), (k, m)
        assert [s for s, _ in res["metrics"][k]["rows"]] == list(g[f"occupancies_k{k}_subject"])


def test_clustering_scores_match_reference_table():
    g = np.load(os.path.join(GOLD, "pipeline_case.npz"))
    X = g["lei"].astype(np.float32)
    for i, k in enumerate(g["perf_k"]):
        sc = O.clustering_scores(X, g[f"labels_k{k}"])
        assert sc["Dunn_score"] == pytest.approx(g["perf_dunn"][i], rel=1e-6)
        assert sc["silhouette_score"] == pytest.approx(g["perf_silhouette"][i], abs=1e-6)
        assert sc["Davies-Bouldin_score"] == pytest.approx(g["perf_db"][i], rel=1e-9)


def test_stats_fixture_matches_oracle():
    """tests/golden/stats_case.npz (made by the unmodified reference's permtest_ind / permtest_rel, see
    oracle/make_golden.py: stats_case): statistic, Levene decision, effect size, Bonferroni alpha are reproduced by
    the oracle; the stored p-values (random in the reference) lie within 6 standard errors of the oracle's."""
    g = np.load(os.path.join(GOLD, "stats_case.npz"))
    vals, cond, cols = g["values"], g["condition"], [str(c) for c in g["columns"]]
    rows = O.permtest_table(vals, cond, cols, n_perm=5000, random_state=np.random.RandomState(9))
    for j, r in enumerate(rows):
        assert str(g["ind_variable"][j]) == r["variable"] and str(g["ind_test"][j]) == r["test"]
        assert str(g["ind_group_1"][j]) == r["group_1"] and str(g["ind_group_2"][j]) == r["group_2"]
        assert g["ind_statistic"][j] == pytest.approx(r["statistic"], rel=1e-13)
        assert g["ind_effect_size"][j] == r["effect_size"]
        assert g["ind_alpha_Bonferroni"][j] == r["alpha_Bonferroni"]
        p = r["p-value"]
        assert abs(g["ind_p-value"][j] - p) <= 6 * np.sqrt(max(p * (1 - p), 1e-3) / 5000)
    half = int(g["half"])
    n1 = int((cond == "patients").sum())
    sel = np.r_[0:half, n1:n1 + half]
    rel = O.permtest_table(vals[sel], cond[sel], cols, n_perm=5000, paired=True, random_state=np.random.default_rng(1))
    for j, r in enumerate(rel):
        assert g["rel_statistic"][j] == pytest.approx(r["statistic"], rel=1e-13, abs=1e-16)
        assert g["rel_effect_size"][j] == r["effect_size"]
        p = r["p-value"]
        assert abs(g["rel_p-value"][j] - p) <= 12 * np.sqrt(max(p * (1 - p), 1e-3) / 5000)


def test_permutation_tests_exact_branch_equals_scipy():
    """Small groups: every split is enumerated; the oracle's p-values equal the installed SciPy's exact tests
    (ttest_ind with PermutationMethod; permutation_test permutation_type='samples')."""
    from scipy import stats
    rng = np.random.default_rng(2)
    a, b = rng.normal(size=5), rng.normal(0.5, 1.0, size=4)
    for alt in ("two-sided", "less", "greater"):
        mine = O.permutation_ttest_ind(a, b, 5000, True, alt)[1]
        ref = stats.ttest_ind(a, b, alternative=alt, method=stats.PermutationMethod(n_resamples=5000)).pvalue
        if alt != "two-sided":          # SciPy's current two-sided p is 2 min(less, greater); the 1.9-1.14 one is |t| >= |t_obs|
            assert mine == pytest.approx(ref, abs=1e-15)
    e1, e2 = rng.normal(size=8), rng.normal(size=8)
    for alt in ("two-sided", "less", "greater"):
        mine = O.permutation_paired(e1, e2, 5000, alt)[1]
        ref = stats.permutation_test((e1, e2), lambda x, y, axis: np.mean(x, axis=axis) - np.mean(y, axis=axis), n_resamples=5000,
                                     vectorized=True, permutation_type="samples", alternative=alt).pvalue
        assert mine == pytest.approx(ref, abs=1e-15)
