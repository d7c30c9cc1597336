"""Read a results folder with the UNMODIFIED reference DataLoader (pyleida/_data_loader.py:45-345) in a process where
only the reference package is importable, and print what it saw as JSON.  TEST INFRASTRUCTURE ONLY (used by
tests/test_results_roundtrip.py, build container only: needs /root/reference).

    python oracle/ref_dataloader_check.py <data_path> <results_path> <k>
"""
import json
import os
import sys
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402  (only for its stub list and the reference location)

import numpy as np  # noqa: E402

if not hasattr(np, "float_"):
    np.float_ = np.float64
for name in ref_shim._STUBS:
    sys.modules.setdefault(name, MagicMock(name=name))
sys.path.insert(0, ref_shim.REFERENCE_ROOT)
from pyleida._data_loader import DataLoader  # noqa: E402  (the reference's)

data_path, results_path, k = sys.argv[1], sys.argv[2], int(sys.argv[3])
dl = DataLoader(data_path=data_path, results_path=results_path)
model = dl.load_model(k)
dist = dl.centroids_distances(k)
out = {
    "loader_module": DataLoader.__module__, "loader_file": sys.modules[DataLoader.__module__].__file__,
    "model_class": f"{type(model).__module__}.{type(model).__name__}",
    "model_file": sys.modules[type(model).__module__].__file__,
    "eig_shape": list(dl.eigenvectors.shape), "eig_sum": float(dl.eigenvectors.iloc[:, 2:].values.sum()),
    "pred_cols": list(dl.predictions.columns), "pred_k": dl.predictions[f"k_{k}"].tolist(),
    "centers": model.cluster_centers_.tolist(), "inertia": float(model.inertia_),
    "predict": model.predict(dl.eigenvectors.iloc[:, 2:].values.astype(np.float32)).tolist(),
    "dist_cols": list(dist.columns), "dist_sum": float(dist.iloc[:, 2:].values.sum()),
    "centroids_shape": list(dl.load_centroids(k).shape),
    "occ": dl.occupancies(k).iloc[:, 2:].values.tolist(), "dwell": dl.dwell_times(k).iloc[:, 2:].values.tolist(),
    "trans": dl.transitions(k).iloc[:, 2:].values.tolist(), "classes": dl._classes_lst_,
    "n_subjects": len(dl.time_series()),
}
print(json.dumps(out))
