"""The hot path as one device-resident function: BOLD on the GPU in, labels / centroids / integer
metric tables on the GPU out.  `Leida._execute_all` wraps it with the host<->device copies and the
DataFrame assembly; bench.py times it directly for the HBM-resident number.

Reference flow: pyleida/_leida.py:279-331 (per-subject phase -> eigenvector, identify_states,
compute_dynamics_metrics).
"""
import numpy as np

from . import _native as nv
from .clustering.kmeans import KMeansEngine, select_models, sweep_runs
from .parallel import Comm, row_offsets
from .signal_tools.phase import leading_eigenvectors, padded_pitch


class DeviceResult:
    """Device-resident outputs of run_device()."""
    __slots__ = ("lei64", "x32", "engine", "batch", "models", "labels", "col_k", "occ", "runs", "trans",
                 "n_rows", "seg_off", "seg_rows")


def run_device(blocks, n_rois, seg_off, seg_rows, K_min=2, K_max=20, n_init=15, random_state=None,
               init_indices=None, centroid_update="exact", keep_f64=True, comm=None, profile=None, f64_sink=None):
    """blocks: list of CUDA fp64 tensors (S_b, N, T_b), subjects in pipeline order.
    seg_off / seg_rows: CUDA int64 tensors describing the metric groups (see Segments).
    f64_sink: optional callable(lei64) invoked right after the stage-1/2 launches are queued; `Leida`
    uses it to start the device->host copy of the fp64 eigenvectors on a side stream while k-means runs.
    Returns a DeviceResult; nothing is copied to the host except the few control words k-means
    needs (active-run count, per-run distortion and cluster sizes for restart selection)."""
    torch = nv.require_cuda()
    comm = comm if comm is not None else Comm()
    dev = blocks[0].device
    N = int(n_rois)
    xp = padded_pitch(N)
    rows = [int(b.shape[0]) * (int(b.shape[2]) - 2) for b in blocks]
    M = int(sum(rows))
    x32 = torch.empty((M, xp), dtype=torch.float32, device=dev)
    lei64 = torch.empty((M, N), dtype=torch.float64, device=dev) if keep_f64 else None
    def timed(name, fn):
        """bench.py's live per-kernel timing (same convention as KMeansEngine.run)"""
        if profile is None or not (profile.get("*") or name in profile):
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        profile.setdefault(name, []).append((e0, e1))
        return out

    def stage12():
        r0 = 0
        for b, nr in zip(blocks, rows):                                    # stage 1+2 (_leida.py:287-289)
            leading_eigenvectors(b, want_f64=keep_f64, out_f64=None if lei64 is None else lei64[r0:r0 + nr],
                                 out_f32=x32[r0:r0 + nr])
            r0 += nr
    timed("stage12", stage12)
    if f64_sink is not None and lei64 is not None:
        f64_sink(lei64)
    off = row_offsets(comm.all_gather_int(M))
    engine = KMeansEngine(x32, n_features=N, comm=comm, row_lo=int(off[comm.rank]), m_global=int(off[-1]))
    runs, plan = sweep_runs(engine.m_global, K_min, K_max, n_init, random_state, init_indices, comm=comm)
    batch = engine.run(runs, n_iter=1_000, centroid_update=centroid_update, profile=profile)   # stage 3 (:314-321)
    models, labels = select_models(engine, batch, plan, n_init, to_host=False, stacked=True)   # labels: (n_cols, M) int32
    ks = list(range(K_min, K_max + 1))
    col_k = np.asarray(ks, dtype=np.int32)
    G = int(seg_off.numel()) - 1
    kmax = int(col_k.max())
    occ = torch.empty((len(ks), G, kmax), dtype=torch.int32, device=dev)
    runs_t = torch.empty((len(ks), G, kmax), dtype=torch.int32, device=dev)
    trans = torch.empty((len(ks), G, kmax, kmax), dtype=torch.int32, device=dev)
    d_k = torch.from_numpy(col_k).to(dev)
    timed("metrics", lambda: nv.call(
        "leida_dynamics_counts_i32", nv.ptr(labels), M, len(ks), nv.ptr(d_k), nv.ptr(seg_off), nv.ptr(seg_rows), G,
        kmax, nv.ptr(occ), nv.ptr(runs_t), nv.ptr(trans), nv.stream()))   # stage 4 (:326-331)
    out = DeviceResult()
    out.lei64, out.x32, out.engine, out.batch, out.models = lei64, x32, engine, batch, models
    out.labels, out.col_k, out.occ, out.runs, out.trans = labels, col_k, occ, runs_t, trans
    out.n_rows, out.seg_off, out.seg_rows = M, seg_off, seg_rows
    return out
