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
"""
import os
import pickle


def save_clustering(path, predictions, performance, models):
    root = os.path.join(path, "clustering")
    os.makedirs(os.path.join(root, "models"), exist_ok=True)
    for name, model in models.items():
        with open(os.path.join(root, "models", f"model_{name}.pkl"), "wb") as f:
            pickle.dump(model, f)
    performance.to_csv(os.path.join(root, "clustering_performance.csv"), sep="\t", index=False)
    predictions.to_csv(os.path.join(root, "predictions.csv"), sep="\t", index=False)


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


def save_eigenvectors(path, eigenvectors):
    os.makedirs(path, exist_ok=True)
    eigenvectors.to_csv(os.path.join(path, "eigenvectors.csv"), sep="\t", index=False)
