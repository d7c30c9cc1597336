"""Clustering quality scores on the GPU (reference: pyleida/clustering/_clustering.py:331-339 and
dunn_fast :934-973; sklearn silhouette_score(metric='cosine') and davies_bouldin_score).

All K columns are scored together: the silhouette through per-cluster sums of the normalised rows
(no n^2 pass), Dunn through one tiled pairwise pass shared by every K, Davies-Bouldin from the exact
centroids that k-means already holds.  Only K x K sized arithmetic stays on the host.
"""
import numpy as np

from .. import _native as nv


def clustering_scores(engine, labels, ks, centroids64, sample_idx=None):
    """engine: KMeansEngine (rows on the device); labels: CUDA int32 (n_cols, M_local) final (remapped)
    labels; ks[c] = K of column c; centroids64[c]: (K, N) float64 exact member means of column c;
    sample_idx: GLOBAL indices of the rows scored by Dunn / silhouette (None = all).  Davies-Bouldin always uses
    every row, like the reference.  Returns dict of three float64 arrays (n_cols).

    Row-sharded engines (several GPUs): the scored rows -- at most the reference's 30 000-row sample -- and their
    labels are gathered on every rank (each row is owned by one rank, the others add zeros), so Dunn and the
    silhouette are computed from exactly the single-GPU inputs; Davies-Bouldin adds the ranks' int64 distance sums
    and member counts.  Every rank returns the same values."""
    torch = nv.require_cuda()
    from ..parallel import gather_seed_rows
    dev = engine.X.device
    comm = engine.comm
    n_cols = len(ks)
    col_k = np.asarray(ks, dtype=np.int32)
    col0 = np.zeros(n_cols, dtype=np.int32)
    col0[1:] = np.cumsum(col_k[:-1])
    total = int(col_k.sum())
    labels = labels.to(torch.int32).contiguous()
    M, N, XP = engine.M, engine.N, engine.XP
    if comm.world > 1:
        idx = np.arange(engine.m_global, dtype=np.int64) if sample_idx is None else np.asarray(sample_idx, dtype=np.int64)
        if len(idx) > 60_000:
            raise ValueError("scoring more than 60 000 rows on a sharded engine is not supported (the reference samples 30 000)")
        Xs = gather_seed_rows(engine.X, N, idx, engine.row_lo, comm)                 # (n, XP) on every rank
        gi = torch.from_numpy(idx).to(dev)
        mine = (gi >= engine.row_lo) & (gi < engine.row_lo + M)
        labs = torch.zeros((n_cols, len(idx)), dtype=torch.int32, device=dev)
        if bool(mine.any()):
            labs[:, mine] = labels[:, gi[mine] - engine.row_lo]
        comm.all_reduce_sum_(labs)
        n, sel, lab_s, pitch_s = len(idx), None, labs, len(idx)
    else:
        Xs = engine.X
        n = M if sample_idx is None else len(sample_idx)
        sel = None if sample_idx is None else torch.from_numpy(np.ascontiguousarray(sample_idx, dtype=np.int64)).to(dev)
        lab_s, pitch_s = labels, M
    d_col0, d_colk = torch.from_numpy(col0).to(dev), torch.from_numpy(col_k).to(dev)
    xs = torch.empty((n, XP), dtype=torch.float32, device=dev)
    G = torch.empty((total, XP), dtype=torch.int64, device=dev)
    cnt = torch.empty(total, dtype=torch.int32, device=dev)
    sil = torch.empty(n_cols, dtype=torch.int64, device=dev)
    mn = torch.empty(n_cols, dtype=torch.int32, device=dev)
    mx = torch.empty(n_cols, dtype=torch.int32, device=dev)
    nv.call("leida_scores_cosine", nv.ptr(Xs), XP, N, nv.ptr(sel), n, nv.ptr(lab_s), pitch_s, n_cols, nv.ptr(d_col0),
            nv.ptr(d_colk), total, nv.ptr(xs), nv.ptr(G), nv.ptr(cnt), nv.ptr(sil), nv.ptr(mn), nv.ptr(mx), nv.stream())
    # Davies-Bouldin is a ratio of Euclidean distances: evaluated in the engine's (power-of-two scaled) units
    cen_all = [np.asarray(c, dtype=np.float64) / engine.scale for c in centroids64]
    cents = torch.from_numpy(np.ascontiguousarray(np.vstack(cen_all), dtype=np.float64)).to(dev)
    dsum = torch.zeros(total, dtype=torch.int64, device=dev)
    if M > 0:
        nv.call("leida_scores_db", nv.ptr(engine.X), XP, N, M, nv.ptr(labels), M, n_cols, nv.ptr(d_col0), nv.ptr(cents), total,
                nv.ptr(dsum), nv.stream())
    counts_all = torch.zeros(total, dtype=torch.int64, device=dev)
    for c in range(n_cols):
        K = int(col_k[c])
        counts_all[col0[c]:col0[c] + K] = torch.bincount(labels[c].long(), minlength=K)[:K]
    if comm.world > 1:
        comm.all_reduce_sum_(dsum)
        comm.all_reduce_sum_(counts_all)
    sil_h = sil.cpu().numpy().astype(np.float64) * 2.0 ** -40 / n
    mn_h = mn.cpu().numpy().view(np.float32).astype(np.float64)
    mx_h = mx.cpu().numpy().view(np.float32).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        dunn = mn_h / mx_h                                       # np.min(deltas) / np.max(big_deltas)   (:958)
    dsum_h = dsum.cpu().numpy().astype(np.float64) * 2.0 ** -40
    counts_h = counts_all.cpu().numpy().astype(np.float64)
    db = np.empty(n_cols)
    for c in range(n_cols):
        K = int(col_k[c])
        with np.errstate(divide="ignore", invalid="ignore"):
            intra = dsum_h[col0[c]:col0[c] + K] / counts_h[col0[c]:col0[c] + K]   # mean distance of the members to their centroid
        cen = cen_all[c]
        diff = cen[:, None, :] - cen[None, :, :]
        cd = np.sqrt((diff * diff).sum(-1))                      # pairwise_distances(centroids)
        if np.allclose(intra, 0) or np.allclose(cd, 0):
            db[c] = 0.0
            continue
        cd[cd == 0] = np.inf
        db[c] = np.mean(np.max((intra[:, None] + intra[None, :]) / cd, axis=1))
    return {"Dunn_score": dunn, "silhouette_score": sil_h, "Davies-Bouldin_score": db}


def dunn_fast(points, labels):
    """Dunn index of a labelling on cosine distances (reference: clustering/_clustering.py:934-973): the smallest
    non-zero distance between rows of different clusters over the largest distance between rows of the same cluster,
    from one tiled pairwise pass on the GPU (the reference materialises the n x n matrix with sklearn).  `labels` may
    be any values; they are ranked like np.unique does."""
    torch = nv.require_cuda()
    from .kmeans import KMeansEngine
    pts = np.ascontiguousarray(np.asarray(points), dtype=np.float32)
    uniq, lab = np.unique(np.asarray(labels), return_inverse=True)
    if pts.ndim != 2 or lab.shape[0] != pts.shape[0]:
        raise ValueError("'points' must be (N_samples, N_features) and 'labels' (N_samples,)")
    eng = KMeansEngine(pts)
    K = int(len(uniq))
    cents = np.vstack([pts[lab == c].astype(np.float64).mean(0) for c in range(K)])
    dev_labels = torch.from_numpy(lab.astype(np.int32)[None, :].copy()).to(eng.X.device)
    return float(clustering_scores(eng, dev_labels, [K], [cents])["Dunn_score"][0])
