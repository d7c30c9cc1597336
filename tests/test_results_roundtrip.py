"""CPU, build container only (needs /root/reference): a results folder written by this package's results_io is read
by the UNMODIFIED reference DataLoader (pyleida/_data_loader.py:83-144 and its accessors) in a separate process, the
pickled models included (they unpickle into the reference's own KMeansLeida and its predict/transform run).  The
tables here come from the CPU oracle -- what is under test is the on-disk format, which is host code."""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

import leida_oracle as O
import ref_shim
import synthetic_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree absent")


def test_reference_dataloader_reads_our_results_folder(tmp_path):
    import pandas as pd
    import pyleida                                                     # loads on CPU; nothing below launches a kernel
    from pyleida import results_io
    syn = synthetic_loader.load()
    ts, classes = syn.make_dataset(4, 9, 40, "states")
    res = O.execute_all(ts, classes, TR=0.72, K_min=2, K_max=4, n_init=2, random_state=0)
    rois = [f"roi{i}" for i in range(9)]
    data = tmp_path / "data"
    data.mkdir()
    with open(data / "time_series.pkl", "wb") as f:
        pickle.dump(ts, f)
    with open(data / "metadata.pkl", "wb") as f:
        pickle.dump(classes, f)
    (data / "rois_labels.txt").write_text("\n".join(rois) + "\n")
    out = str(tmp_path / "LEiDA_results")
    eig = pd.DataFrame(res["eigenvectors"], columns=rois)
    eig.insert(0, "subject_id", res["subject_id"])
    eig.insert(1, "condition", res["condition"])
    pred = pd.DataFrame({"subject_id": res["subject_id"], "condition": res["condition"]})
    models, perf, dyn = {}, [], {"occupancies": {}, "dwell_times": {}, "transitions": {}}
    for k in (2, 3, 4):
        om = res["models"][k]
        km = pyleida.clustering.KMeansLeida(k=k, n_init=2)
        km.cluster_centers_, km.labels_, km.inertia_, km._is_fitted = om.cluster_centers_, om.labels_, om.inertia_, True
        models[f"k_{k}"] = km
        pred[f"k_{k}"] = om.labels_
        perf.append({"k": k, "Dunn_score": 0.0, "distortion": om.inertia_, "silhouette_score": 0.0, "Davies-Bouldin_score": 0.0})
        met = res["metrics"][k]
        ids = pd.DataFrame(met["rows"], columns=["subject_id", "condition"])
        dyn["occupancies"][f"k_{k}"] = pd.concat([ids, pd.DataFrame(met["occupancies"], columns=[f"PL_state_{i + 1}" for i in range(k)])], axis=1)
        dyn["dwell_times"][f"k_{k}"] = pd.concat([ids, pd.DataFrame(met["dwell_times"], columns=[f"PL_state_{i + 1}" for i in range(k)])], axis=1)
        dyn["transitions"][f"k_{k}"] = pd.concat([ids, pd.DataFrame(met["transitions"], columns=[f"From_{i + 1}_to_{j + 1}" for i in range(k) for j in range(k)])], axis=1)
    results_io.save_eigenvectors(out, eig, binary=True)
    results_io.save_clustering(out, pred, pd.DataFrame(perf), models, binary=True)
    results_io.save_dynamics(out, dyn)
    # --- the reference reads it (separate process: only the reference package on its path)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    p = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_dataloader_check.py"), str(data), out, "3"],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr[-2000:]
    got = json.loads(p.stdout.strip().splitlines()[-1])
    assert got["loader_file"].startswith(ref_shim.REFERENCE_ROOT) and got["model_file"].startswith(ref_shim.REFERENCE_ROOT)
    assert got["model_class"] == "pyleida.clustering._clustering.KMeansLeida"
    assert got["eig_shape"] == [eig.shape[0], 11] and got["eig_sum"] == pytest.approx(res["eigenvectors"].sum(), rel=1e-9)
    assert got["pred_cols"] == ["subject_id", "condition", "k_2", "k_3", "k_4"]
    assert got["pred_k"] == res["models"][3].labels_.tolist()
    assert np.allclose(got["centers"], res["models"][3].cluster_centers_, rtol=0, atol=0)
    assert got["predict"] == res["models"][3].labels_.tolist()         # the reference's predict on our pickled centroids
    assert got["dist_cols"] == ["subject_id", "condition", "centroid_1", "centroid_2", "centroid_3"]
    assert got["centroids_shape"] == [3, 9] and got["n_subjects"] == 4 and got["classes"] == ["A", "B"]
    assert np.allclose(got["occ"], res["metrics"][3]["occupancies"], rtol=1e-12)
    assert np.allclose(got["dwell"], res["metrics"][3]["dwell_times"], rtol=1e-12)
    assert np.allclose(got["trans"], res["metrics"][3]["transitions"], rtol=1e-12)
    # --- and this package's DataLoader reads the same folder (binary files preferred), file access only
    dl = pyleida.DataLoader(str(data), out)
    assert dl.eigenvectors.shape == eig.shape and np.array_equal(dl.eigenvectors.iloc[:, 2:].values, res["eigenvectors"])
    assert np.array_equal(dl.predictions["k_4"].values, res["models"][4].labels_)
    assert (dl._K_min_, dl._K_max_) == (2, 4) and dl._classes_lst_ == ["A", "B"]
    assert np.array_equal(dl.load_model(2).cluster_centers_, res["models"][2].cluster_centers_)
    assert np.allclose(dl.occupancies(2).iloc[:, 2:].values, res["metrics"][2]["occupancies"], rtol=1e-12)
    assert dl.load_centroids(4).shape == (4, 9) and len(dl.time_series()) == 4
