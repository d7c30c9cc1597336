from ._stats import _compute_stats, hedges_g, permtest_ind, permtest_rel  # noqa: F401
