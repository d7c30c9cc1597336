"""Results folder in the reference's on-disk layout, so that the reference's DataLoader
(pyleida/_data_loader.py:83-144) and users' scripts can read what the GPU path produced:

    <root>/eigenvectors.csv                                  _leida.py:308          (tab separated)
    <root>/clustering/models/model_k_{k}.pkl                 _clustering.py:345-350 (pickled KMeansLeida)
    <root>/clustering/clustering_performance.csv             _clustering.py:366
    <root>/clustering/predictions.csv                        _clustering.py:367
    <root>/dynamics_metrics/k_{k}/occupancies.csv            _dynamics_metrics.py:202
    <root>/dynamics_metrics/k_{k}/dwell_times.csv            _dynamics_metrics.py:538
    <root>/dynamics_metrics/k_{k}/transitions_probabilities.csv   _dynamics_metrics.py:338
Host-side pandas / pickle only; nothing here touches the GPU.

Fast path (not part of the reference's format): at BASELINE config 3 `eigenvectors.csv` is 1.2 M x 400 numbers of
text (~10 GB, minutes to write and to parse).  binary=True additionally writes `eigenvectors.npy` (float64 matrix) with
`eigenvectors_index.csv` (subject_id, condition) and `clustering/predictions.npy` (int16 labels, one column per K in
the csv's order); `pyleida.DataLoader` prefers them when present.  csv=False skips the two big text files (the folder is
then readable by this package only).
"""
import os
import pickle


def save_clustering(path, predictions, performance, models, binary=False, csv=True):
    import numpy as np
    root = os.path.join(path, "clustering")
    os.makedirs(os.path.join(root, "models"), exist_ok=True)
    for name, model in models.items():
        with open(os.path.join(root, "models", f"model_{name}.pkl"), "wb") as f:
            pickle.dump(model, f)
    performance.to_csv(os.path.join(root, "clustering_performance.csv"), sep="\t", index=False)
    if csv:
        predictions.to_csv(os.path.join(root, "predictions.csv"), sep="\t", index=False)
    if binary or not csv:
        cols = [c for c in predictions.columns if str(c).startswith("k_")]
        np.save(os.path.join(root, "predictions.npy"), predictions[cols].to_numpy(dtype=np.int16))
        with open(os.path.join(root, "predictions_columns.txt"), "w") as f:
            f.write("\n".join(cols) + "\n")
        predictions[["subject_id", "condition"]].to_csv(os.path.join(root, "predictions_index.csv"), sep="\t", index=False)


def save_dynamics(path, metrics):
    root = os.path.join(path, "dynamics_metrics")
    files = {"occupancies": "occupancies.csv", "dwell_times": "dwell_times.csv",
             "transitions": "transitions_probabilities.csv"}
    for kind, fname in files.items():
        for k, table in metrics[kind].items():
            os.makedirs(os.path.join(root, k), exist_ok=True)
            table.to_csv(os.path.join(root, k, fname), sep="\t", index=False)


def save_stats(path, stats):
    """{metric}_stats.csv next to the metric tables (stats/_stats.py:356-359)."""
    root = os.path.join(path, "dynamics_metrics")
    for metric, per_k in stats.items():
        for k, table in per_k.items():
            os.makedirs(os.path.join(root, k), exist_ok=True)
            table.to_csv(os.path.join(root, k, f"{metric}_stats.csv"), sep="\t", index=False)


def save_eigenvectors(path, eigenvectors, binary=False, csv=True):
    import numpy as np
    os.makedirs(path, exist_ok=True)
    if csv:
        eigenvectors.to_csv(os.path.join(path, "eigenvectors.csv"), sep="\t", index=False)
    if binary or not csv:
        np.save(os.path.join(path, "eigenvectors.npy"), eigenvectors.iloc[:, 2:].to_numpy(dtype=np.float64))
        eigenvectors.iloc[:, :2].to_csv(os.path.join(path, "eigenvectors_index.csv"), sep="\t", index=False)
        with open(os.path.join(path, "eigenvectors_columns.txt"), "w") as f:
            f.write("\n".join(str(c) for c in eigenvectors.columns[2:]) + "\n")


def load_eigenvectors(path):
    """eigenvectors DataFrame of a results folder: the binary pair when present, else the reference's csv."""
    import numpy as np
    import pandas as pd
    npy = os.path.join(path, "eigenvectors.npy")
    if os.path.exists(npy):
        idx = pd.read_csv(os.path.join(path, "eigenvectors_index.csv"), sep="\t")
        with open(os.path.join(path, "eigenvectors_columns.txt")) as f:
            cols = [ln.rstrip("\n") for ln in f if ln.strip()]
        df = pd.DataFrame(np.load(npy, mmap_mode="r"), columns=cols, copy=False)
        df.insert(0, "subject_id", idx["subject_id"].values)
        df.insert(1, "condition", idx["condition"].values)
        return df
    return pd.read_csv(os.path.join(path, "eigenvectors.csv"), sep="\t")


def load_predictions(path):
    """predictions DataFrame (subject_id, condition, k_2 ...) of a results folder, binary first."""
    import numpy as np
    import pandas as pd
    root = os.path.join(path, "clustering")
    npy = os.path.join(root, "predictions.npy")
    if os.path.exists(npy):
        idx = pd.read_csv(os.path.join(root, "predictions_index.csv"), sep="\t")
        with open(os.path.join(root, "predictions_columns.txt")) as f:
            cols = [ln.strip() for ln in f if ln.strip()]
        lab = pd.DataFrame(np.load(npy).astype(np.int64), columns=cols)
        return pd.concat([idx, lab], axis=1)
    return pd.read_csv(os.path.join(root, "predictions.csv"), sep="\t")
