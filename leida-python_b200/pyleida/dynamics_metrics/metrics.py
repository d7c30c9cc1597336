"""Host side of stage 4.  Reference: pyleida/dynamics_metrics/_dynamics_metrics.py:11-136,208-342,423-542.

The GPU produces exact integer tables per (condition, subject) group -- state counts, run counts
and transition-pair counts -- in one launch for all K columns; the fp64 ratios are then formed
here with the same operations as the reference (count / T', count / runs * TR, pair / row-sum with
0/0 -> 0), so every number is bit-identical to the reference's given identical labels.
"""
import numpy as np

from .. import _native as nv


class Segments:
    """Row groups in the reference's table order (:180-182): sorted condition, then sorted subject.

    rows: concatenated row indices (ascending inside a group); off: group boundaries.
    """

    def __init__(self, subjects, conditions, rows, off):
        self.subjects, self.conditions = list(subjects), list(conditions)
        self.rows = np.ascontiguousarray(rows, dtype=np.int64)
        self.off = np.ascontiguousarray(off, dtype=np.int64)

    def __len__(self):
        return len(self.subjects)

    def subjects_array(self):
        if getattr(self, "_sa", None) is None:
            self._sa = np.asarray(self.subjects, dtype=object)
        return self._sa

    def conditions_array(self):
        if getattr(self, "_ca", None) is None:
            self._ca = np.asarray(self.conditions, dtype=object)
        return self._ca

    @classmethod
    def from_metadata(cls, metadata):
        sub = np.asarray(metadata["subject_id"].values if hasattr(metadata, "columns") else metadata[0])
        cond = np.asarray(metadata["condition"].values if hasattr(metadata, "columns") else metadata[1])
        return cls.from_arrays(sub, cond)

    @classmethod
    def from_arrays(cls, sub, cond):
        sub, cond = np.asarray(sub), np.asarray(cond)
        # np.unique order == the reference's nested np.unique loops
        cu, ci = np.unique(cond, return_inverse=True)
        su, si = np.unique(sub, return_inverse=True)
        key = ci.astype(np.int64) * len(su) + si
        order = np.argsort(key, kind="stable")
        ks = key[order]
        starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]]) if len(ks) else np.zeros(0, dtype=np.int64)
        off = np.r_[starts, len(ks)].astype(np.int64)
        gk = ks[starts] if len(ks) else np.zeros(0, dtype=np.int64)
        return cls(su[gk % len(su)] if len(su) else [], cu[gk // max(len(su), 1)] if len(cu) else [], order, off)

    @classmethod
    def contiguous(cls, subjects, conditions, lengths):
        """Subjects stored back to back (the Leida pipeline): group g = rows of one subject.
        Table order is still (sorted condition, sorted subject)."""
        subjects, conditions = np.asarray(subjects), np.asarray(conditions)
        lengths = np.asarray(lengths, dtype=np.int64)
        start = np.r_[0, np.cumsum(lengths)[:-1]]
        cu, ci = np.unique(conditions, return_inverse=True)
        su, si = np.unique(subjects, return_inverse=True)
        order = np.argsort(ci.astype(np.int64) * len(su) + si, kind="stable")
        rows = np.concatenate([np.arange(start[g], start[g] + lengths[g]) for g in order]) if len(order) else np.zeros(0)
        off = np.r_[0, np.cumsum(lengths[order])]
        return cls(subjects[order], conditions[order], rows, off)


def dynamics_counts(labels, col_k, segments):
    """GPU integer tables.  labels: (n_cols, M) int32 CUDA tensor or numpy; col_k[c] = number of
    states of column c.  Returns numpy int32 arrays occ (n_cols, G, kmax), runs (n_cols, G, kmax),
    trans (n_cols, G, kmax, kmax)."""
    torch = nv.require_cuda()
    if isinstance(labels, np.ndarray):
        labels = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32)).cuda()
    if labels.dim() == 1:
        labels = labels.unsqueeze(0)
    labels = labels.to(torch.int32).contiguous()
    n_cols, M = labels.shape
    col_k = np.asarray(col_k, dtype=np.int32).reshape(-1)
    assert len(col_k) == n_cols
    kmax = int(col_k.max()) if n_cols else 1
    G = len(segments)
    dev = labels.device
    occ = torch.empty((n_cols, G, kmax), dtype=torch.int32, device=dev)
    runs = torch.empty((n_cols, G, kmax), dtype=torch.int32, device=dev)
    trans = torch.empty((n_cols, G, kmax, kmax), dtype=torch.int32, device=dev)
    if G > 0:
        if segments.rows.size and (segments.rows.min() < 0 or segments.rows.max() >= M):
            raise ValueError("segment row index out of range")
        d_k = torch.from_numpy(col_k).to(dev)
        d_off = torch.from_numpy(segments.off).to(dev)
        d_rows = torch.from_numpy(segments.rows).to(dev)
        nv.call("leida_dynamics_counts_i32", nv.ptr(labels), M, n_cols, nv.ptr(d_k), nv.ptr(d_off), nv.ptr(d_rows), G,
                kmax, nv.ptr(occ), nv.ptr(runs), nv.ptr(trans), nv.stream())
    return occ.cpu().numpy(), runs.cpu().numpy(), trans.cpu().numpy()


# ---- exact ratios, same operations as the reference -------------------------------------------

def _occupancy(occ, lengths):
    return occ.astype(np.float64) / np.asarray(lengths, dtype=np.float64)[:, None]       # :131-134


def _dwell(occ, runs, TR):
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = occ.astype(np.float64) / runs.astype(np.float64)                            # np.mean of run lengths (:456-459)
    if TR is not None:
        mean = mean * TR
    return np.where(runs > 0, mean, 0.0)                                                   # fillna(0.0) (:530)


def _transitions(trans):
    tc = trans.astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        tp = tc / tc.sum(axis=-1, keepdims=True)                                           # :247-252
    return np.nan_to_num(tp, copy=True)                                                    # :256


def _single(labels, k=None):
    labels = np.asarray(labels)
    if labels.ndim != 1:
        raise ValueError("labels must be 1D")
    kk = int(k) if k is not None else (int(labels.max()) + 1 if labels.size else 1)
    seg = Segments(["_"], ["_"], np.arange(labels.size), [0, labels.size])
    occ, runs, trans = dynamics_counts(labels.astype(np.int32)[None, :], [kk], seg)
    return occ[0, 0], runs[0, 0], trans[0, 0]


def fractional_occupancy(cluster_assignement):
    """:110-136 -- DataFrame(cluster, n_volumes, occupancy) of the states that occur, sorted by cluster."""
    import pandas as pd
    lab = np.asarray(cluster_assignement)
    occ, _, _ = _single(lab)
    present = np.nonzero(occ)[0]
    n = lab.shape[0]
    return pd.DataFrame({"cluster": present, "n_volumes": occ[present].astype(np.int64),
                         "occupancy": [i / n for i in occ[present].astype(np.int64)]})


def dwell_times(labels, TR=None, plot=False):
    """:423-473 -- (dict state -> array of run lengths, DataFrame(Cluster, Mean_lifetime))."""
    import pandas as pd
    lab = np.asarray(labels)
    occ, runs, _ = _single(lab)
    present = np.nonzero(occ)[0]
    # run lengths per state (returned for API parity; the means come from the GPU counts)
    edges = np.flatnonzero(np.r_[True, lab[1:] != lab[:-1], True]) if lab.size else np.zeros(1, dtype=np.int64)
    lens, keys = np.diff(edges), lab[edges[:-1]] if lab.size else lab
    dwell = {c: lens[keys == c] for c in present}
    mean = _dwell(occ[present][None, :], runs[present][None, :], TR)[0]
    return dwell, pd.DataFrame({"Cluster": present, "Mean_lifetime": mean})


def transition_probabilities(labels, k, norm=True, plot=False):
    """:208-276 -- (k, k) transition matrix of one sequence."""
    _, _, trans = _single(np.asarray(labels), k)
    return _transitions(trans) if norm else trans.astype(np.float64)


def _group_tables(metadata, labels, segments=None):
    assert metadata.shape[0] == np.asarray(labels).shape[0], \
        "The number of rows in 'metadata' must be the same as the number of provided 'labels'."
    seg = segments if segments is not None else Segments.from_metadata(metadata)
    lab = np.asarray(labels)
    uniq = np.unique(lab)
    k = len(uniq)                                                                          # :316
    if k and (uniq[0] != 0 or uniq[-1] != k - 1):
        raise ValueError("labels must be 0..K-1 with every state present (the reference's tables break otherwise)")
    occ, runs, trans = dynamics_counts(lab.astype(np.int32)[None, :], [k], seg)
    return seg, k, occ[0], runs[0], trans[0]


def _frame(seg, values, cols):
    """DataFrame(subject_id, condition, *cols): one numeric block + the two id columns."""
    import pandas as pd
    meta = pd.DataFrame({"subject_id": pd.Series(seg.subjects_array(), dtype=object),
                         "condition": pd.Series(seg.conditions_array(), dtype=object)})
    return pd.concat([meta, pd.DataFrame(np.ascontiguousarray(values), columns=cols)], axis=1)


class LazyFrames(dict):
    """{'k_K': DataFrame} whose DataFrames are assembled on first access from the exact integer
    tables (the arithmetic is already done; pandas object construction is the expensive part)."""

    def __init__(self, builders, raw=None, conditions=None):
        """raw: optional {key: callable -> (values ndarray, column names)} and conditions: the per-row condition
        labels -- the numbers behind the frames, for consumers that do not need pandas objects (the statistics
        stage)."""
        super().__init__()
        self.raw = raw
        self.conditions = conditions
        self._builders = dict(builders)
        for k in self._builders:
            dict.__setitem__(self, k, None)

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if v is None and key in self._builders:
            v = self._builders.pop(key)()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]


def _state_frame(seg, values, k):
    return _frame(seg, values, [f"PL_state_{i + 1}" for i in range(k)])


def _transition_frame(seg, values, k):
    return _frame(seg, values.reshape(len(seg), k * k), [f"From_{i + 1}_to_{j + 1}" for i in range(k) for j in range(k)])


def _save_one(save_results, path, table, fname):
    if save_results:
        if path is None:
            raise ValueError("You must provide a path to save the results")
        import os
        os.makedirs(path, exist_ok=True)
        table.to_csv(os.path.join(path, fname), sep="\t", index=False)
    return table


def fractional_occupancy_group(metadata, labels, save_results=False, path=None, _segments=None):
    """:138-206."""
    seg, k, occ, _, _ = _group_tables(metadata, labels, _segments)
    return _save_one(save_results, path, _state_frame(seg, _occupancy(occ, np.diff(seg.off)), k), "occupancies.csv")


def dwell_times_group(metadata, labels, TR=None, save_results=False, path=None, _segments=None):
    """:475-542."""
    seg, k, occ, runs, _ = _group_tables(metadata, labels, _segments)
    return _save_one(save_results, path, _state_frame(seg, _dwell(occ, runs, TR), k), "dwell_times.csv")


def transition_probabilities_group(metadata, labels, save_results=False, path=None, _segments=None):
    """:278-342."""
    seg, k, _, _, trans = _group_tables(metadata, labels, _segments)
    return _save_one(save_results, path, _transition_frame(seg, _transitions(trans), k), "transitions_probabilities.csv")


def compute_dynamics_metrics(clusters_labels, TR=None, save_results=False, path=None, _segments=None,
                             _device_labels=None, verbose=False):
    """:11-105 -- occupancies, dwell times and transition probabilities for every 'k_*' column,
    from ONE GPU launch over all columns.  Returns {'dwell_times','occupancies','transitions'} ->
    {'k_K': DataFrame}.  _device_labels: optional (n_cols, M) CUDA int32 tensor holding the same
    labels as the DataFrame columns (avoids the host round trip inside the pipeline)."""
    if save_results and path is None:
        raise ValueError("You must provide a path to save the results")
    if "subject_id" not in clusters_labels.columns:                                       # :58-61
        clusters_labels.rename(columns={clusters_labels.columns[0]: "subject_id"}, inplace=True)
    if "condition" not in clusters_labels.columns:
        clusters_labels.rename(columns={clusters_labels.columns[1]: "condition"}, inplace=True)
    Ks = [c for c in clusters_labels.columns if str(c).startswith("k_")]
    seg = _segments if _segments is not None else Segments.from_metadata(clusters_labels)
    ys = clusters_labels[Ks].values
    col_k = []
    for i in range(len(Ks)):
        u = np.unique(ys[:, i])
        if len(u) and (u[0] != 0 or u[-1] != len(u) - 1):
            raise ValueError(f"column {Ks[i]}: labels must be 0..K-1 with every state present")
        col_k.append(len(u))
    lab = _device_labels if _device_labels is not None else np.ascontiguousarray(ys.T, dtype=np.int32)
    occ, runs, trans = dynamics_counts(lab, col_k, seg)
    out = tables_from_counts(Ks, col_k, seg, occ, runs, trans, TR)
    if save_results:
        from ..results_io import save_dynamics
        save_dynamics(path, out)
    return out


def tables_from_counts(names, col_k, seg, occ, runs, trans, TR=None):
    """The reference's three DataFrames per K column from the integer tables of dynamics_counts()."""
    lengths = np.diff(seg.off)
    b_occ, b_dw, b_tr, r_occ, r_dw = {}, {}, {}, {}, {}
    for i, name in enumerate(names):
        k = int(col_k[i])
        cols = [f"PL_state_{j + 1}" for j in range(k)]
        r_occ[name] = lambda i=i, k=k, cols=cols: (_occupancy(occ[i][:, :k], lengths), cols)
        r_dw[name] = lambda i=i, k=k, cols=cols: (_dwell(occ[i][:, :k], runs[i][:, :k], TR), cols)
        b_occ[name] = lambda name=name, k=k: _state_frame(seg, r_occ[name]()[0], k)
        b_dw[name] = lambda name=name, k=k: _state_frame(seg, r_dw[name]()[0], k)
        b_tr[name] = lambda i=i, k=k: _transition_frame(seg, _transitions(trans[i][:, :k, :k]), k)
    cond = seg.conditions_array()
    return {"dwell_times": LazyFrames(b_dw, r_dw, cond), "occupancies": LazyFrames(b_occ, r_occ, cond),
            "transitions": LazyFrames(b_tr)}
