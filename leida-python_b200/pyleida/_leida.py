"""The Leida pipeline facade over the GPU hot path.

Reference: pyleida/_leida.py:57-401 (class Leida; __init__ :97-109, fit_predict :111-192,
_execute_all :194-401).  Only the orchestration of the hot path is reproduced:
phase -> leading eigenvector (:287-289) -> K sweep (:314-321) -> dynamics metrics (:326-331).
The permutation statistics (:336-342), plots and the results folder are outside the path.
"""
import os
import pickle
import warnings

import numpy as np

from . import _native as nv
from .dynamics_metrics.metrics import Segments
from .parallel import Comm
from .pipeline import run_device


_POOL = None


def _staging_pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        # several ranks on one host share its cores (torchrun exports LOCAL_WORLD_SIZE)
        per_rank = (os.cpu_count() or 4) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        _POOL = ThreadPoolExecutor(max_workers=max(2, min(16, per_rank)))
    return _POOL


class Bunch(dict):
    """dict with attribute access (the reference returns sklearn.utils.Bunch objects, _leida.py:377-389)."""
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def _load_inputs(data_path):
    """Minimal loaders for the reference's input folder (data_utils/_data_utils.py:140-268):
    time_series.pkl or time_series/<subject>/*.csv, metadata.pkl|csv, rois_labels.txt."""
    import pandas as pd
    ts_pkl = os.path.join(data_path, "time_series.pkl")
    if os.path.exists(ts_pkl):
        with open(ts_pkl, "rb") as f:
            tseries = pickle.load(f)
    else:
        root = os.path.join(data_path, "time_series")
        tseries = {}
        for sub in sorted(d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d))):
            for fn in os.listdir(os.path.join(root, sub)):
                if fn.endswith(".csv"):
                    tseries[sub] = np.array(pd.read_csv(os.path.join(root, sub, fn), sep=",", header=None))
    meta_pkl = os.path.join(data_path, "metadata.pkl")
    if os.path.exists(meta_pkl):
        with open(meta_pkl, "rb") as f:
            classes = pickle.load(f)
    else:
        md = pd.read_csv(os.path.join(data_path, "metadata.csv"), sep=",")
        classes = {s: list(md[md.subject_id == s].condition) for s in np.unique(md.subject_id)}
    with open(os.path.join(data_path, "rois_labels.txt")) as f:
        rois = [ln.strip() for ln in f if ln.strip()]
    return tseries, classes, rois


class Leida:
    """Run the LEiDA pipeline on the GPU.

    Leida(data_path) loads the reference's input folder; Leida.from_arrays(...) takes the same
    objects already in memory (time_series: dict subject -> (N_ROIs, N_volumes) array; classes:
    dict subject -> list of condition labels, one per subject or one per volume).
    """

    def __init__(self, data_path=None, *, time_series=None, classes=None, rois_labels=None, comm=None):
        if data_path is not None:
            if not isinstance(data_path, str):
                raise TypeError("'data_path' must be a string!")
            if not os.path.exists(data_path):
                raise ValueError("The specified path could't be founded.")
            time_series, classes, rois_labels = _load_inputs(data_path)
        if time_series is None or classes is None:
            raise ValueError("provide 'data_path' or time_series= and classes=")
        self.time_series = time_series
        self.classes = classes
        n_rois = {np.asarray(v).shape[0] for v in time_series.values()}
        if len(n_rois) != 1:                                                               # _leida.py:848-851
            raise Exception("All the subjects must have the same number of ROIs/parcels.")
        self._N = n_rois.pop()
        self.rois_labels = list(rois_labels) if rois_labels is not None else [f"ROI_{i + 1}" for i in range(self._N)]
        if len(self.rois_labels) != self._N:
            raise Exception("The number of ROIs labels must match the number of rows of the time series.")
        for s in time_series:
            if s not in classes:
                raise Exception(f"subject '{s}' has no condition label in 'classes'")
        self.rois_coordinates = None
        self._comm = comm

    @classmethod
    def from_arrays(cls, time_series, classes, rois_labels=None, comm=None):
        return cls(None, time_series=time_series, classes=classes, rois_labels=rois_labels, comm=comm)

    # ------------------------------------------------------------------------------------------
    def fit_predict(self, TR=None, paired_tests=False, n_perm=5_000, save_results=False, random_state=None,
                    K_min=2, K_max=20, n_init=15, init_indices=None, centroid_update="exact",
                    keep_eigenvectors=True, scores=True, stats=True, _profile=None):
        """_leida.py:111-192.  K range and n_init are arguments here (the reference hard-codes
        2..20 and 15, :156-157 / _clustering.py:250); stats=True runs the permutation tests on the occupancy and
        dwell-time tables like the reference (_leida.py:336-342; on the GPU, stats/_stats.py) when there are at least
        two conditions with two rows each (on multi-GPU runs the per-subject tables are gathered first); save_results=True writes the reference's
        'LEiDA_results' folder layout (results_io.py) and, like the reference, refuses to overwrite it.  scores=True
        fills the Dunn / silhouette / Davies-Bouldin columns of the performance table like the reference
        (they are not part of the benchmarked hot path; multi-GPU runs gather the scored 30 000-row sample)."""
        if TR is not None and not isinstance(TR, (int, float)):
            raise TypeError("'TR' must be 'None', and integer or a floating number!")
        if not isinstance(paired_tests, bool):
            raise TypeError("'paired_tests' must be either True or False.")
        if not isinstance(n_perm, int):
            raise TypeError("'n_perm' must be an integer!")
        if n_perm < 100:
            raise ValueError("The number of permutations cannot be lower than 100.")
        if not isinstance(save_results, bool) and save_results not in ("binary", "fast"):
            raise TypeError("'save_results' must be a boolean value (True or False)!")
        # "binary": the reference's folder plus .npy copies of the two volume-sized tables; "fast": the .npy copies
        # INSTEAD of the two big csv files (see results_io.py)
        save_binary, save_csv = save_results in ("binary", "fast"), save_results != "fast"
        save_results = bool(save_results)
        self._results_path_ = "LEiDA_results"
        if save_results:
            if os.path.exists(self._results_path_):                                        # _leida.py:256-260
                raise Warning("EXECUTION ABORTED: The folder 'LEiDA_results' already exists. If you have results from "
                              "earlier executions of the analysis, consider changing the folder's name or moving the "
                              "folder to another location.")
            if not keep_eigenvectors:
                raise ValueError("save_results=True needs keep_eigenvectors=True")
        self._K_min_, self._K_max_ = int(K_min), int(K_max)
        self._execute_all(TR, random_state, n_init, init_indices, centroid_update, keep_eigenvectors, _profile, scores)
        self._is_fitted = True
        self._run_stats(stats, paired_tests, n_perm, random_state)
        if save_results:                                                                   # _leida.py:306-310 and the stage writers
            from .results_io import save_clustering, save_dynamics, save_eigenvectors
            save_eigenvectors(self._results_path_, self.eigenvectors, binary=save_binary, csv=save_csv)
            save_clustering(self._results_path_, self.predictions, self._clustering_.performance, self._clustering_.models,
                            binary=save_binary, csv=save_csv)
            save_dynamics(self._results_path_, {"occupancies": self._dynamics_.occupancies,
                                                "dwell_times": self._dynamics_.dwell_times,
                                                "transitions": self._dynamics_.transitions})
            if self._dynamics_.get("stats") is not None:
                from .results_io import save_stats
                save_stats(self._results_path_, self._dynamics_["stats"])
        return self

    def _execute_all(self, TR, random_state, n_init, init_indices, centroid_update, keep_eigenvectors, profile=None,
                     scores=True):
        import pandas as pd
        torch = nv.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        comm = self._comm if self._comm is not None else Comm()
        ids = list(self.time_series.keys())
        N = self._N
        n_vols = [int(np.asarray(self.time_series[s]).shape[1]) - 2 for s in ids]
        if min(n_vols) < 1:
            raise Exception("every subject needs at least 3 volumes")
        self.h2d_bytes = self.d2h_bytes = 0
        import time
        t_mark = [time.perf_counter()]
        self.timings_ = {}

        def mark(name):
            t = time.perf_counter()
            self.timings_[name] = self.timings_.get(name, 0.0) + 1e3 * (t - t_mark[0])
            t_mark[0] = t
        # host -> device: consecutive subjects with the same T are staged as one pinned block
        blocks, i = [], 0
        while i < len(ids):
            T = n_vols[i] + 2
            j = i
            while j < len(ids) and n_vols[j] + 2 == T:
                j += 1
            host = torch.empty((j - i, N, T), dtype=torch.float64, pin_memory=True)
            hv = host.numpy()
            dev_block = torch.empty((j - i, N, T), dtype=torch.float64, device=dev)
            # stage in slices: worker threads fill pinned memory (numpy copies release the GIL); all fills are
            # queued at once and every slice goes over PCIe as soon as its own fills are done
            step = max(1, (j - i + 15) // 16)

            def fill(lo, hi, base=i):
                for b in range(lo, hi):
                    hv[b] = self.time_series[ids[base + b]]
            pool, pending = _staging_pool(), []
            for lo in range(0, j - i, step):
                hi = min(j - i, lo + step)
                cuts = np.linspace(lo, hi, min(4, hi - lo) + 1).astype(int)
                pending.append((lo, hi, [pool.submit(fill, int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]))
            for lo, hi, futs in pending:
                for f in futs:
                    f.result()
                dev_block[lo:hi].copy_(host[lo:hi], non_blocking=True)
            blocks.append(dev_block)
            self.h2d_bytes += host.numel() * 8
            i = j
        # per-volume metadata (_leida.py:290-297) and the metric groups
        sub_list = np.repeat(np.asarray(ids, dtype=object), n_vols)
        if any(len(self.classes[s]) > 1 for s in ids):
            cond_list = np.concatenate([
                np.asarray(self.classes[s], dtype=object)[1:nv_ + 1] if len(self.classes[s]) > 1
                else np.repeat(np.asarray(self.classes[s][:1], dtype=object), nv_)
                for s, nv_ in zip(ids, n_vols)])
            # ids keep their own dtype (np.unique of integer ids sorts 2 before 10, like the reference's
            # np.unique(metadata.subject_id)); only mixed-type ids fall back to their string form
            try:
                sub_native = np.repeat(np.asarray(ids), n_vols)
                np.unique(sub_native)
            except TypeError:
                sub_native = sub_list.astype(str)
            segments = Segments.from_arrays(sub_native, cond_list.astype(str))
        else:
            conds = [self.classes[s][0] for s in ids]
            cond_list = np.repeat(np.asarray(conds, dtype=object), n_vols)
            segments = Segments.contiguous(np.asarray(ids), np.asarray(conds), n_vols)
        self._meta = (sub_list, cond_list)
        mark("stage_inputs_h2d_ms")
        seg_off = torch.from_numpy(segments.off).to(dev)
        seg_rows = torch.from_numpy(segments.rows).to(dev)
        self.h2d_bytes += segments.off.nbytes + segments.rows.nbytes
        def to_host(t, stream=None):
            """device -> pinned host, asynchronous; synchronised once below"""
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            if stream is None:
                h.copy_(t, non_blocking=True)
            else:
                with torch.cuda.stream(stream):
                    h.copy_(t, non_blocking=True)
            return h
        pinned = {}
        side = torch.cuda.Stream(device=dev) if keep_eigenvectors else None

        def early_eigenvectors(lei64):
            # the fp64 eigenvectors (the largest result) are final after stage 2: copy them out on a side
            # stream while k-means runs
            side.wait_stream(torch.cuda.current_stream())
            pinned["lei"] = to_host(lei64, side)
        res = run_device(blocks, N, seg_off, seg_rows, K_min=self._K_min_, K_max=self._K_max_, n_init=n_init,
                         random_state=random_state, init_indices=init_indices, centroid_update=centroid_update,
                         keep_f64=keep_eigenvectors, comm=comm, profile=profile,
                         f64_sink=early_eigenvectors if keep_eigenvectors else None)
        self.device_result_ = res
        self.kmeans_batch_ = res.batch
        mark("device_pipeline_ms")
        # device -> host: labels (as the int64 the reference's frames hold, in both orientations: per-K rows for
        # the models, per-volume rows for the predictions frame), centroids, integer metric tables
        ks = list(range(self._K_min_, self._K_max_ + 1))
        labels64 = res.labels.to(torch.int64)
        pinned.update({"labels": to_host(labels64), "labels_t": to_host(labels64.t().contiguous()),
                       "occ": to_host(res.occ), "runs": to_host(res.runs), "trans": to_host(res.trans)})
        for k in ks:
            pinned[f"c{k}"] = to_host(res.models[k].centers_device_)
        torch.cuda.current_stream().synchronize()
        if side is not None:
            side.synchronize()
        mark("d2h_results_ms")
        labels, labels_t = pinned["labels"].numpy(), pinned["labels_t"].numpy()
        occ, runs, trans = pinned["occ"].numpy(), pinned["runs"].numpy(), pinned["trans"].numpy()
        self.d2h_bytes += labels.nbytes + labels_t.nbytes + occ.nbytes + runs.nbytes + trans.nbytes
        sc = None
        if scores:                                         # Dunn / silhouette / Davies-Bouldin (_clustering.py:331-339)
            from .clustering.kmeans import score_sample, sweep_runs
            from .clustering.scores import clustering_scores
            sc = clustering_scores(res.engine, res.labels, ks, [res.models[k]._centers_exact() for k in ks],
                                   score_sample(res.engine.m_global, sweep_runs.last_sample, comm))
        predictions = pd.DataFrame({"subject_id": pd.Series(sub_list, dtype=object),
                                    "condition": pd.Series(cond_list, dtype=object)})
        models, perf = {}, []
        for c, k in enumerate(ks):
            km = res.models[k]
            km.labels_ = labels[c]
            km.cluster_centers_ = pinned[f"c{k}"].numpy().copy()
            self.d2h_bytes += km.cluster_centers_.nbytes
            models[f"k_{k}"] = km
            perf.append({"k": k, "Dunn_score": sc["Dunn_score"][c] if sc else np.nan, "distortion": km.inertia_,
                         "silhouette_score": sc["silhouette_score"][c] if sc else np.nan,
                         "Davies-Bouldin_score": sc["Davies-Bouldin_score"][c] if sc else np.nan})
        predictions = pd.concat([predictions, pd.DataFrame(labels_t, columns=[f"k_{k}" for k in ks], copy=False)], axis=1)
        self._lei_host = None
        if keep_eigenvectors:
            self._lei_host = pinned["lei"].numpy()         # stays in the pinned buffer (no second copy)
            self.d2h_bytes += self._lei_host.nbytes
        from .dynamics_metrics.metrics import tables_from_counts
        dynamics = tables_from_counts([f"k_{k}" for k in ks], res.col_k, segments, occ, runs, trans, TR)
        self.predictions = predictions
        self._eigenvectors_df = None
        self._clustering_ = Bunch(predictions=predictions, performance=pd.DataFrame(perf), models=models)
        self._dynamics_ = Bunch(dwell_times=dynamics["dwell_times"], occupancies=dynamics["occupancies"],
                                transitions=dynamics["transitions"], stats=None)
        # np.unique(eigenvectors.condition) (_leida.py:190) without touching the per-volume column: the labels a
        # subject contributes are classes[s][1:n_volumes+1] (one per volume) or its single label
        used = [c for s_, nv_ in zip(ids, n_vols)
                for c in (self.classes[s_][1:nv_ + 1] if len(self.classes[s_]) > 1 else self.classes[s_][:1])]
        self._classes_lst_ = np.unique(np.asarray(used).astype(str)).tolist()
        self._N_classes_ = len(self._classes_lst_)
        mark("assemble_frames_ms")

    def _run_stats(self, stats, paired_tests, n_perm, random_state):
        """Stage 4b of the reference (_leida.py:336-342): permutation tests per state and K.  The tables are tested as
        they are (one row per subject and condition); results are kept in memory, `save_results` writes them with the
        other dynamics files."""
        if not stats:
            return
        comm = self._comm if self._comm is not None else Comm()
        occ, dwl = self._dynamics_.occupancies, self._dynamics_.dwell_times
        if comm.world > 1:
            # the tables are one row per (subject, condition): small.  Every rank receives every rank's rows, puts them
            # in the reference's row order (sorted condition, then sorted subject: _dynamics_metrics.py:180-182) and runs
            # the same tests, so Leida.stats() answers on every rank.
            import pandas as pd
            parts = comm.all_gather_object({m: {k: t[k] for k in t.keys()} for m, t in (("occupancies", occ), ("dwell_times", dwl))})

            def whole(metric, k):
                df = pd.concat([p[metric][k] for p in parts], ignore_index=True)
                return df.sort_values(["condition", "subject_id"], kind="stable").reset_index(drop=True)
            occ = {k: whole("occupancies", k) for k in occ.keys()}
            dwl = {k: whole("dwell_times", k) for k in dwl.keys()}
        first = occ[f"k_{self._K_min_}"]
        sizes = first.groupby("condition").size()
        if len(sizes) < 2 or int(sizes.min()) < 2:
            return                                   # nothing to compare (the reference would fail inside scipy here)
        if paired_tests and len(set(sizes.tolist())) != 1:
            raise ValueError("paired_tests=True needs the same number of rows in every condition")
        from .stats import _compute_stats
        tables = {"occupancies": occ, "dwell_times": dwl}
        self._dynamics_["stats"] = _compute_stats(tables, paired_tests=paired_tests, n_perm=n_perm, save_results=False,
                                                  random_state=random_state if isinstance(random_state, int) else None)

    def stats(self, k=2, metric="occupancies"):
        """Results of the statistical analysis for the K partition (_leida.py:475-500)."""
        if self._dynamics_.get("stats") is None:
            raise Exception("the statistics stage was not run (stats=False or a single condition)")
        return self._dynamics_["stats"][metric][f"k_{k}"]

    # ------------------------------------------------------------------------------------------
    @property
    def eigenvectors(self):
        """DataFrame(subject_id, condition, one column per ROI) like _leida.py:301-303; built on
        first access from the device copy."""
        import pandas as pd
        if getattr(self, "_eigenvectors_df", None) is None:
            if getattr(self, "_lei_host", None) is None:
                raise Exception("eigenvectors were not kept (keep_eigenvectors=False) or the model is not fitted")
            df = pd.DataFrame(self._lei_host, columns=self.rois_labels)
            df.insert(0, "subject_id", pd.Series(self._meta[0], dtype=object))
            df.insert(1, "condition", pd.Series(self._meta[1], dtype=object))
            self._eigenvectors_df = df
        return self._eigenvectors_df

    def _check_is_fitted(self):
        if not hasattr(self, "_is_fitted"):
            raise Exception("You have to fit the model first by using the 'fit_predict' method.")

    def load_model(self, k=2):
        self._check_is_fitted()
        return self._clustering_.models[f"k_{k}"]

    def load_centroids(self, k=2):
        import pandas as pd
        self._check_is_fitted()
        c = self._clustering_.models[f"k_{k}"].cluster_centers_
        return pd.DataFrame(c.T, columns=[f"PL_state_{i + 1}" for i in range(c.shape[0])], index=self.rois_labels)

    def centroids_distances(self, k=2):
        """Distance of every eigenvector to every centroid of partition k (_leida.py:447-473): DataFrame
        (subject_id, condition, centroid_1..k), fp64 cosine distances computed on the device rows."""
        import pandas as pd
        self._check_is_fitted()
        if not isinstance(k, int):
            raise TypeError("'k' must be a integer!")
        if k not in range(self._K_min_, self._K_max_ + 1):
            raise ValueError(f"'k' must be a fitted model: you selected 'k'= {k}, but K-Means models were fitted "
                             f"for the {self._K_min_} - {self._K_max_} range.")
        d = self.load_model(k).transform(self.device_result_.engine)
        out = pd.DataFrame(d, columns=[f"centroid_{c + 1}" for c in range(k)])
        out.insert(0, "subject_id", pd.Series(self._meta[0], dtype=object))
        out.insert(1, "condition", pd.Series(self._meta[1], dtype=object))
        return out

    def dwell_times(self, k=2):
        self._check_is_fitted()
        return self._dynamics_.dwell_times[f"k_{k}"]

    def occupancies(self, k=2):
        self._check_is_fitted()
        return self._dynamics_.occupancies[f"k_{k}"]

    def transitions(self, k=2):
        self._check_is_fitted()
        return self._dynamics_.transitions[f"k_{k}"]
