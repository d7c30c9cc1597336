"""Post-hoc statistics of the LEiDA pipeline, same interface as the reference's
pyleida/stats/_stats.py: permtest_ind (:54-133), hedges_g (:135-192), permtest_rel (:194-269),
_compute_stats (:275-364).  Every feature column of a two-group table is tested in one batch on the GPU
(csrc/stats.cu: moments + Levene statistic, permutation null distributions); the host only turns the
Levene statistic into its p-value (F survival function), picks pooled / Welch per column like the reference
(p_levene > 0.05), converts counts to p-values and assembles the DataFrame.

Randomness.  The reference leaves the permutations unseeded (scipy draws them from the global RandomState /
a fresh default_rng), so its p-values are not reproducible; here `random_state` (int) seeds a counter-based
generator on the device, and `perms` / `swaps` accept explicit permutations (tests compare the null
distribution with the CPU oracle on identical permutations).  When the number of distinct splits is at most
n_perm the test is exact, as in SciPy: all splits are enumerated and p = #extreme / #splits.
"""
import os
from itertools import combinations
from math import comb

import numpy as np

from .. import _native as nv


def hedges_g(x1, x2, paired=False):
    """Hedges's g of two samples (stats/_stats.py:135-192; host arithmetic, two tiny vectors)."""
    x1, x2 = np.asarray(x1, dtype=np.float64), np.asarray(x2, dtype=np.float64)
    n1, n2 = x1.size, x2.size
    var1, var2 = np.var(x1), np.var(x2)
    diff_mean = np.abs(np.mean(x1) - np.mean(x2))
    if not paired:
        return diff_mean / np.sqrt((((n1 - 1) * var1) + ((n2 - 1) * var2)) / (n1 + n2 - 2))
    return diff_mean / np.sqrt((var1 + var2) / 2)


def _validate(data, class_column, n_perm):
    if class_column is None:
        raise ValueError("You must specify the 'class_column'.")
    if class_column not in data.columns:
        raise ValueError(f"The 'class_column' '{class_column}' not founded in 'data'!")
    if not isinstance(n_perm, int):
        raise TypeError("'n_perm' must be an integer!")


def _seed(random_state):
    if random_state is None:
        return int.from_bytes(os.urandom(8), "little")
    return int(random_state) & 0xFFFFFFFFFFFFFFFF


def _to_device(vals, cls):
    torch = nv.require_cuda()
    groups = np.unique(cls)
    idx1 = np.flatnonzero(cls == groups[0]).astype(np.int32)
    idx2 = np.flatnonzero(cls == groups[1]).astype(np.int32)
    dev = torch.device("cuda", torch.cuda.current_device())
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    return groups, torch.from_numpy(vals).to(dev), torch.from_numpy(idx1).to(dev), torch.from_numpy(idx2).to(dev), dev


def _moments(d_vals, d1, d2, n_vars, dev):
    torch = nv.require_cuda()
    out = torch.empty((n_vars, 8), dtype=torch.float64, device=dev)
    nv.call("leida_stats_moments_f64", nv.ptr(d_vals), d_vals.shape[0], n_vars, nv.ptr(d1), d1.numel(), nv.ptr(d2),
            d2.numel(), nv.ptr(out), nv.stream())
    return out.cpu().numpy()


def _levene_p(W, n1, n2):
    from scipy.special import betainc
    dfd = n1 + n2 - 2.0
    return betainc(dfd / 2.0, 0.5, dfd / (dfd + W))            # F(1, dfd) survival function


def _ind_core(vals, cls, n_perm, alternative, random_state, perms=None, return_null=False, adjustment="biased"):
    """All columns of `vals` (rows, n_vars) in one batch: arrays of statistic, p-value, equal_var flag, effect size."""
    torch = nv.require_cuda()
    groups, d_vals, d1, d2, dev = _to_device(vals, cls)
    n1, n2, nvars = int(d1.numel()), int(d2.numel()), int(d_vals.shape[1])
    mom = _moments(d_vals, d1, d2, nvars, dev)
    equal_var = _levene_p(mom[:, 4], n1, n2) > 0.05                              # stats/_stats.py:109
    exact = perms is None and n_perm >= comb(n1 + n2, n1)
    if exact:
        perms = np.asarray([list(c) + sorted(set(range(n1 + n2)) - set(c)) for c in combinations(range(n1 + n2), n1)])
    d_perms = None
    if perms is not None:
        perms = np.ascontiguousarray(np.asarray(perms), dtype=np.int32)
        if perms.ndim != 2 or perms.shape[1] != n1 + n2:
            raise ValueError("'perms' must have shape (n_permutations, n1 + n2)")
        d_perms = torch.from_numpy(perms).to(dev)
    n_used = int(perms.shape[0]) if perms is not None else int(n_perm)
    t_obs = torch.empty(nvars, dtype=torch.float64, device=dev)
    counts = torch.empty((nvars, 3), dtype=torch.int32, device=dev)
    null = torch.empty((nvars, n_used), dtype=torch.float64, device=dev) if return_null else None
    d_ev = torch.from_numpy(equal_var.astype(np.uint8)).to(dev)
    nv.call("leida_perm_ttest_ind_f64", nv.ptr(d_vals), d_vals.shape[0], nvars, nv.ptr(d1), n1, nv.ptr(d2), n2, nv.ptr(d_ev),
            n_used, nv.ptr(d_perms), _seed(random_state), nv.ptr(t_obs), nv.ptr(counts), nv.ptr(null), nv.stream())
    t_obs, counts = t_obs.cpu().numpy(), counts.cpu().numpy()
    col = {"less": 0, "greater": 1}.get(alternative, 2)
    adj = 0 if (exact or adjustment == "none") else 1         # randomised test: biased estimate (SciPy >= 1.9)
    pvals = (counts[:, col] + adj) / (n_used + adj)
    var1, var2 = mom[:, 2] / n1, mom[:, 3] / n2                 # np.var (ddof = 0), as hedges_g uses
    eff = np.abs(mom[:, 0] - mom[:, 1]) / np.sqrt(((n1 - 1) * var1 + (n2 - 1) * var2) / (n1 + n2 - 2))
    return dict(groups=groups, statistic=t_obs, pvalue=pvals, equal_var=equal_var, effect=eff,
                null=None if null is None else null.cpu().numpy())


def _rel_core(vals, cls, n_perm, alternative, random_state, swaps=None, return_null=False):
    torch = nv.require_cuda()
    groups, d_vals, d1, d2, dev = _to_device(vals, cls)
    n, nvars = int(d1.numel()), int(d_vals.shape[1])
    if int(d2.numel()) != n:
        raise ValueError("paired test: the two groups must have the same number of rows")
    mom = _moments(d_vals, d1, d2, nvars, dev)
    exact = swaps is None and n < 62 and n_perm >= 2 ** n
    if exact:
        swaps = (np.arange(2 ** n)[:, None] >> np.arange(n)[None, :]) & 1
    d_sw = None
    if swaps is not None:
        swaps = np.ascontiguousarray(np.asarray(swaps), dtype=np.uint8)
        if swaps.ndim != 2 or swaps.shape[1] != n:
            raise ValueError("'swaps' must have shape (n_permutations, n_pairs)")
        d_sw = torch.from_numpy(swaps).to(dev)
    n_used = int(swaps.shape[0]) if swaps is not None else int(n_perm)
    obs = torch.empty(nvars, dtype=torch.float64, device=dev)
    counts = torch.empty((nvars, 2), dtype=torch.int32, device=dev)
    null = torch.empty((nvars, n_used), dtype=torch.float64, device=dev) if return_null else None
    nv.call("leida_perm_paired_f64", nv.ptr(d_vals), d_vals.shape[0], nvars, nv.ptr(d1), nv.ptr(d2), n, n_used, nv.ptr(d_sw),
            _seed(random_state), nv.ptr(obs), nv.ptr(counts), nv.ptr(null), nv.stream())
    obs, counts = obs.cpu().numpy(), counts.cpu().numpy()
    adj = 0 if exact else 1
    p_less = (counts[:, 0] + adj) / (n_used + adj)
    p_greater = (counts[:, 1] + adj) / (n_used + adj)
    pvals = np.clip({"less": p_less, "greater": p_greater}.get(alternative, np.minimum(p_less, p_greater) * 2), 0.0, 1.0)
    eff = np.abs(mom[:, 0] - mom[:, 1]) / np.sqrt((mom[:, 2] / n + mom[:, 3] / n) / 2)
    return dict(groups=groups, statistic=obs, pvalue=pvals, effect=eff, null=None if null is None else null.cpu().numpy())


def _table(core, features, lo, hi, paired):
    """Rows lo..hi of a batched result as the reference's DataFrame (one table = one Bonferroni family)."""
    import pandas as pd
    g = core["groups"]
    rows = []
    for j in range(lo, hi):
        r = {"variable": features[j - lo], "group_1": g[0], "group_2": g[-1], "statistic": core["statistic"][j],
             "p-value": core["pvalue"][j]}
        if not paired:
            r["test"] = "t-test" if core["equal_var"][j] else "welch"
        r["effect_size"] = core["effect"][j]
        rows.append(r)
    results = pd.DataFrame(rows)
    n_tests = hi - lo
    results["alpha_Bonferroni"] = 0.05 / n_tests
    results["reject_null"] = [bool(p < (0.05 / n_tests)) for p in results["p-value"].values]
    return results


def permtest_ind(data, class_column=None, n_perm=5_000, alternative="two-sided", random_state=None, perms=None,
                 return_null=False, adjustment="biased"):
    """Permutation t-test (pooled or Welch, chosen per column by Levene's test) of two independent groups for
    every feature column of `data`, with Hedges's g and a Bonferroni-corrected alpha (stats/_stats.py:54-133).
    perms: optional (n, n1+n2) index permutations of the pooled sample (group 1 first) used for every column.
    adjustment: "biased" -> p = (b + 1) / (m + 1) for a randomised test (SciPy >= 1.9); "none" -> b / m (SciPy 1.8,
    the version the reference's environment.yml pins)."""
    _validate(data, class_column, n_perm)
    features = [c for c in data if c != class_column]
    core = _ind_core(data[features].to_numpy(dtype=np.float64), np.asarray(data[class_column].values), n_perm, alternative,
                     random_state, perms, return_null, adjustment)
    res = _table(core, features, 0, len(features), paired=False)
    return (res, core["null"]) if return_null else res


def permtest_rel(data, class_column=None, n_perm=5_000, alternative="two-sided", random_state=None, swaps=None,
                 return_null=False):
    """Permutation test of the mean difference of two related (paired) groups for every feature column, with
    Hedges's g (paired) and a Bonferroni-corrected alpha (stats/_stats.py:194-269).  Rows are paired in order
    of appearance within each group, like the reference's boolean selections."""
    _validate(data, class_column, n_perm)
    features = [c for c in data if c != class_column]
    core = _rel_core(data[features].to_numpy(dtype=np.float64), np.asarray(data[class_column].values), n_perm, alternative,
                     random_state, swaps, return_null)
    res = _table(core, features, 0, len(features), paired=True)
    return (res, core["null"]) if return_null else res


def _compute_stats(dynamics_data, paired_tests=False, n_perm=5_000, alternative="two-sided", save_results=True, path=None,
                   random_state=None):
    """Statistical analysis of dwell times and occupancies for every state of every K (stats/_stats.py:275-364):
    for each metric and each K, every pair of conditions is tested column by column.  All tables of the pipeline
    share their rows (one per subject and condition), so for a pair of conditions every column of every table is one
    batch on the GPU: two kernel launches for the whole stage."""
    import pandas as pd
    alternative_options = ["two-sided", "greater", "less"]
    if alternative not in alternative_options:
        raise ValueError(f"Valid 'alternative' options are {alternative_options}.")
    ks = list(dynamics_data["occupancies"].keys())
    metrics = ["dwell_times", "occupancies"]
    seed = None if random_state is None else int(random_state)
    fn_core = _rel_core if paired_tests else _ind_core
    raw = all(getattr(dynamics_data[m], "raw", None) is not None for m in metrics)
    if raw:                                 # tables of this package: take the numbers, not the (lazily built) DataFrames
        cond0 = np.asarray(dynamics_data[metrics[0]].conditions)
        frames = [(m, k, None) for m in metrics for k in ks]
        parts = [dynamics_data[m].raw[k]() for m, k, _ in frames]
        feats = [list(c) for _, c in parts]
        big = np.hstack([np.asarray(v, dtype=np.float64) for v, _ in parts])
        shared = True
    else:
        frames = [(m, k, dynamics_data[m][k].iloc[:, 1:]) for m in metrics for k in ks]       # drop the subject ids
        cond0 = np.asarray(frames[0][2].condition.values)
        shared = all(len(f) == len(cond0) and np.array_equal(np.asarray(f.condition.values), cond0) for _, _, f in frames)
        if shared:
            feats = [[c for c in f if c != "condition"] for _, _, f in frames]
            big = np.hstack([f[c].to_numpy(dtype=np.float64) for (_, _, f), c in zip(frames, feats)])
    pieces = {(m, k): [] for m, k, _ in frames}            # per table: callables that build its DataFrame(s)
    if shared:
        offs = np.cumsum([0] + [len(c) for c in feats])
        for conds in combinations(np.unique(cond0), 2):
            rows = np.isin(cond0, conds)
            core = fn_core(big[rows], cond0[rows], n_perm, alternative, seed)
            seed = None if seed is None else seed + 1
            for i, (m, k, _) in enumerate(frames):
                pieces[(m, k)].append(lambda core=core, i=i: _table(core, feats[i], int(offs[i]), int(offs[i + 1]), paired_tests))
    else:                                   # tables with different rows: one batch per table, like the reference's loop
        for m, k, f in frames:
            for conds in combinations(np.unique(f.condition), 2):
                data_ = f[f.condition.isin(conds)]
                fn = permtest_rel if paired_tests else permtest_ind
                res = fn(data_, class_column="condition", alternative=alternative, n_perm=n_perm, random_state=seed)
                pieces[(m, k)].append(lambda res=res: res)
                seed = None if seed is None else seed + 1

    def build(m, k):
        stacked = [make() for make in pieces[(m, k)]]
        metric_results = pd.concat(stacked, axis=0).reset_index(drop=True) if len(stacked) > 1 else stacked[0]
        metric_results.insert(0, "k", int([i for i in k.replace("_", " ").split() if i.isdigit()][0]))
        return metric_results
    # the arithmetic is done; the DataFrames are assembled on first access (pandas construction is the slow part)
    from ..dynamics_metrics.metrics import LazyFrames
    stats_all = {m: LazyFrames({k: (lambda m=m, k=k: build(m, k)) for k in ks}) for m in metrics}
    if save_results:
        for m in metrics:
            for k in ks:
                try:
                    stats_all[m][k].to_csv(f"{path}/dynamics_metrics/{k}/{m}_stats.csv", sep="\t", index=False)
                except Exception:                                   # noqa: BLE001 - the reference only warns (:360-361)
                    print("Warning: there was a problem when saving the results to local folder.")
    print("*The statistical analysis has finished.")
    return stats_all
