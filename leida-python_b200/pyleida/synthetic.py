"""Seeded synthetic BOLD-like inputs for the five BASELINE.json configurations.

Not part of the hot path: a host-side generator shared by the tests and ``bench.py`` so that both
arms of every comparison see the same bytes.  Pure numpy, deterministic given the arguments.
"""
import numpy as np

CONFIGS = {
    # name: (subjects, rois, volumes)
    "config1": (20, 90, 200),
    "config2": (100, 116, 1200),
    "config3": (1000, 400, 1200),
    "config4": (200, 1000, 1200),
}


def subject_series(subject_index, n_rois, n_volumes, kind="states", n_patterns=6, mean_dwell=8.0):
    """One subject's (n_rois, n_volumes) float64 series, rng = default_rng(1000 + subject_index).

    kind="states": P latent phase-locking patterns p_k in {-1,+1}^N (fixed, default_rng(7)), a
    Markov state sequence z_t with the given mean dwell, and
    x[j,t] = cos(2 pi 0.05 t + pi/2 (p_{z_t}[j] < 0) + 0.3 xi_j) + 0.5 eps[j,t].
    kind="iid": N(0,1) noise (worst case for k-means iteration counts).
    """
    rng = np.random.default_rng(1000 + int(subject_index))
    if kind == "iid":
        return rng.standard_normal((n_rois, n_volumes))
    if kind != "states":
        raise ValueError("kind must be 'states' or 'iid'")
    patterns = np.random.default_rng(7).choice([-1.0, 1.0], size=(n_patterns, n_rois))
    stay = 1.0 - 1.0 / mean_dwell
    z = np.empty(n_volumes, dtype=np.int64)
    z[0] = rng.integers(n_patterns)
    jumps = rng.random(n_volumes) >= stay
    targets = rng.integers(n_patterns, size=n_volumes)
    for t in range(1, n_volumes):
        z[t] = targets[t] if jumps[t] else z[t - 1]
    xi = rng.standard_normal(n_rois)
    eps = rng.standard_normal((n_rois, n_volumes))
    t = np.arange(n_volumes, dtype=np.float64)
    lag = (np.pi / 2) * (patterns[z].T < 0)                       # (N, T)
    return np.cos(2 * np.pi * 0.05 * t[None, :] + lag + 0.3 * xi[:, None]) + 0.5 * eps


def make_dataset(n_subjects, n_rois, n_volumes, kind="states", first_subject=0):
    """dict subject_id -> series, dict subject_id -> [condition]; first half "A", second half "B".

    ``first_subject`` lets a rank generate only its own shard with the same global seeds.
    The A/B split is decided by the GLOBAL subject index against ``n_subjects_total`` = the
    number passed by the caller for the whole job (see :func:`make_shard`).
    """
    return make_shard(n_subjects, n_rois, n_volumes, 0, n_subjects, kind)


def make_shard(n_subjects_total, n_rois, n_volumes, lo, hi, kind="states"):
    """Subjects lo..hi-1 of an ``n_subjects_total`` job."""
    ts, classes = {}, {}
    for s in range(lo, hi):
        sid = f"sub-{s:05d}"
        ts[sid] = subject_series(s, n_rois, n_volumes, kind)
        classes[sid] = ["A" if s < (n_subjects_total + 1) // 2 else "B"]
    return ts, classes


def init_indices(M, K_min, K_max, n_init, seed=0):
    """Explicit k-means initial rows {k: [idx_r ...]} from default_rng(seed) ('k-means
    seeds'): cheap to draw, and can be handed to both the CUDA path and the CPU oracle."""
    rng = np.random.default_rng(seed)
    return {k: [rng.choice(M, size=k, replace=False).astype(np.int64) for _ in range(n_init)]
            for k in range(K_min, K_max + 1)}
