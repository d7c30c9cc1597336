"""Import shim that lets the UNMODIFIED reference hot-path functions run in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (fixture generation) and by
``tests/test_oracle_vs_reference.py`` (which skips when ``/root/reference`` is absent, as on the
GPU box).  Nothing in the product path (``leida-python_b200/``) may import this module.

Why a shim: every hot module of the reference imports plotting /
neuro-imaging packages at top level (``pyleida/signal_tools/_signal_tools.py:2-3,6``,
``pyleida/clustering/_clustering.py:4-5``, ``pyleida/dynamics_metrics/_dynamics_metrics.py:4-5``)
that are not installed here, and ``pyleida/dynamics_metrics/_dynamics_metrics.py:250`` uses
``np.float_`` which NumPy 2 removed.  The shim stubs those imports and restores the alias; no
reference file is copied or edited.
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("LEIDA_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "matplotlib.patches",
    "matplotlib.lines", "matplotlib.gridspec", "matplotlib.ticker", "matplotlib.animation",
    "matplotlib.collections", "matplotlib.figure", "matplotlib.axes",
    "seaborn", "imageio", "nibabel",
    "nilearn", "nilearn.signal", "nilearn.plotting", "nilearn.datasets", "nilearn.surface",
    "nilearn.image", "nilearn.maskers", "nilearn.input_data",
    "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1",
]


def _compat_ttest_ind(a, b, alternative="two-sided", permutations=None, equal_var=True, **kwargs):
    """Stand-in for the `ttest_ind(..., permutations=n)` signature of SciPy 1.7 - 1.14 that the reference calls
    (stats/_stats.py:105-111) and the installed SciPy removed.  Observed statistic: the installed
    scipy.stats.ttest_ind.  Permutation p-value: the restatement of that SciPy's `_permutation_ttest` in
    oracle/leida_oracle.py -- so for THAT number the reference run is not an independent check (the oracle
    header says "parity unpinned" for it); everything else permtest_ind returns is the reference's own."""
    from types import SimpleNamespace
    from scipy import stats
    import leida_oracle
    res = stats.ttest_ind(a, b, equal_var=equal_var, alternative=alternative)
    if not permutations:
        return res
    _, p, _ = leida_oracle.permutation_ttest_ind(a, b, int(permutations), equal_var, alternative)
    return SimpleNamespace(statistic=res.statistic, pvalue=p)


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyleida"))


def load_reference():
    """Return a namespace holding the reference's hot-path callables (unmodified code)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy as np
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = MagicMock(name=name)
    # Load the reference under a private top-level name so that it can never shadow (or be
    # shadowed by) the product package, which is also called ``pyleida``.
    saved = {k: v for k, v in sys.modules.items() if k == "pyleida" or k.startswith("pyleida.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        st = importlib.import_module("pyleida.signal_tools._signal_tools")
        cl = importlib.import_module("pyleida.clustering._clustering")
        dm = importlib.import_module("pyleida.dynamics_metrics._dynamics_metrics")
        try:
            # stats/_stats.py imports `ttest_ind, levene, ks_2samp, permutation_test` from scipy.stats; all exist, but
            # permtest_ind passes `permutations=` to ttest_ind, which the installed SciPy no longer accepts.
            sm = importlib.import_module("pyleida.stats._stats")
            sm.ttest_ind = _compat_ttest_ind
        except Exception:                         # noqa: BLE001 - the stats row is optional for the hot-path tests
            sm = None
    finally:
        sys.path.remove(REFERENCE_ROOT)
        ref_mods = {k: v for k, v in sys.modules.items() if k == "pyleida" or k.startswith("pyleida.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
    ns = types.SimpleNamespace(
        hilbert_phase=st.hilbert_phase,
        phase_coherence=st.phase_coherence,
        get_eigenvectors=st.get_eigenvectors,
        KMeansLeida=cl.KMeansLeida,
        identify_states=cl.identify_states,
        patterns_stability=cl.patterns_stability,
        fractional_occupancy=dm.fractional_occupancy,
        fractional_occupancy_group=dm.fractional_occupancy_group,
        dwell_times=dm.dwell_times,
        dwell_times_group=dm.dwell_times_group,
        transition_probabilities=dm.transition_probabilities,
        transition_probabilities_group=dm.transition_probabilities_group,
        compute_dynamics_metrics=dm.compute_dynamics_metrics,
        hedges_g=getattr(sm, "hedges_g", None),
        permtest_ind=getattr(sm, "permtest_ind", None),
        permtest_rel=getattr(sm, "permtest_rel", None),
        compute_stats=getattr(sm, "_compute_stats", None),
        _modules=(st, cl, dm),
    )
    return ns
