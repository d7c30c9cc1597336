"""DataLoader: explore a results folder written by `Leida.fit_predict(save_results=True)`.

Reference: pyleida/_data_loader.py:45-144 (constructor and checks) and :146-345 (accessors).  Same constructor,
same accessor names and return types; reads the reference's own folders as well as the ones this package writes
(results_io.py), preferring the binary eigenvector / prediction files when they exist.  The plotting methods of the
reference class are outside the accelerated path and are not provided.  `centroids_distances` runs on the GPU
(KMeansLeida.transform); everything else is file access.
"""
import os
import pickle

import numpy as np

from ._leida import _load_inputs
from .results_io import load_eigenvectors, load_predictions


def _check_k_input(k_min, k_max, k):
    if not isinstance(k, int):
        raise TypeError("'k' must be an integer!")
    if k < k_min or k > k_max:
        raise ValueError(f"'k' must be between {k_min} and {k_max}.")


class DataLoader:
    def __init__(self, data_path="data", results_path="LEiDA_results"):
        import pandas as pd
        for _path in (data_path, results_path):
            if not isinstance(_path, str):
                raise TypeError("'data_path' and 'results_path' must be strings!")
        if not os.path.exists(data_path):
            raise ValueError("The provided 'data_path' could't be founded.")
        if not os.path.exists(results_path):
            raise ValueError("The provided 'results_path' could't be founded.")
        self._data_path_, self._results_path_ = data_path, results_path
        self._clustering_ = results_path + "/clustering"
        self._dynamics_ = results_path + "/dynamics_metrics"
        self._models_ = self._clustering_ + "/models"
        if not os.path.exists(self._clustering_) or not os.path.exists(self._dynamics_):
            raise Exception("No results were found in the specified path.")
        try:
            self.eigenvectors = load_eigenvectors(results_path)
            nrois = self.eigenvectors.shape[1] - 2
        except Exception:
            raise Exception("The eigenvectors dataframe could't be loaded.")
        try:
            self.predictions = load_predictions(results_path)
        except Exception:
            raise Exception("The dataframe with the KMeans models predictions couldn't be loaded.")
        with open(os.path.join(data_path, "rois_labels.txt")) as f:
            self.rois_labels = [ln.strip() for ln in f if ln.strip()]
        self.rois_coordinates_ = None
        coords = os.path.join(data_path, "rois_coordinates.csv")
        if os.path.exists(coords):
            c = pd.read_csv(coords, sep=",").values
            self.rois_coordinates_ = c if c.shape[1] == 3 else None
        if nrois != len(self.rois_labels):
            raise Exception(f"The number of ROIs in 'eigenvectors.csv' ({nrois}) and 'rois_labels.txt' "
                            f"{len(self.rois_labels)} must coincide!")
        ks = sorted(int(c[2:]) for c in self.predictions.columns if str(c).startswith("k_"))
        self._K_min_, self._K_max_ = (ks[0], ks[-1]) if ks else (2, 20)
        self._classes_lst_ = np.unique(self.eigenvectors.condition).tolist()
        self._N_classes_ = len(self._classes_lst_)

    def time_series(self):
        return _load_inputs(self._data_path_)[0]

    def load_model(self, k=2):
        _check_k_input(self._K_min_, self._K_max_, k)
        try:
            with open(f"{self._models_}/model_k_{k}.pkl", "rb") as f:
                return pickle.load(f)
        except Exception:
            raise Exception("Can't find results for the selected k.")

    def load_centroids(self, k=2):
        import pandas as pd
        _check_k_input(self._K_min_, self._K_max_, k)
        return pd.DataFrame(self.load_model(k=k).cluster_centers_, columns=self.rois_labels)

    def centroids_distances(self, k=2):
        import pandas as pd
        model = self.load_model(k=k)
        d = pd.DataFrame(model.transform(np.ascontiguousarray(self.eigenvectors.iloc[:, 2:].values)))
        d.columns = [f"centroid_{c + 1}" for c in range(k)]
        return pd.concat((self.eigenvectors[["subject_id", "condition"]], d), axis=1)

    def _table(self, k, fname):
        import pandas as pd
        _check_k_input(self._K_min_, self._K_max_, k)
        return pd.read_csv(f"{self._dynamics_}/k_{k}/{fname}", sep="\t")

    def stats(self, k=2, metric="occupancies"):
        if metric not in ("occupancies", "dwell_times"):
            raise ValueError("'metric' must be either 'occupancies' or 'dwell_times'.")
        return self._table(k, f"{metric}_stats.csv")

    def dwell_times(self, k=2):
        return self._table(k, "dwell_times.csv")

    def occupancies(self, k=2):
        return self._table(k, "occupancies.csv")

    def transitions(self, k=2):
        return self._table(k, "transitions_probabilities.csv")
