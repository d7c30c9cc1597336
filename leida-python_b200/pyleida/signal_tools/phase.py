"""Stage 1 + 2 host wrappers: Hilbert phase, phase coherence, leading eigenvectors.

Every function accepts numpy arrays (copied to the GPU, result returned as numpy -- the reference's
types at the API edge) or CUDA tensors (result stays on the device).  All arithmetic happens in
libleida_b200.so; there is no CPU path.
"""
import numpy as np

from .. import _native as nv


def padded_pitch(n):
    """Row pitch (floats) of the fp32 eigenvector matrix handed to k-means: next multiple of 32."""
    return (int(n) + 31) // 32 * 32


def _to_device(a, dtype):
    torch = nv.require_cuda()
    if isinstance(a, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda(), True
    if torch.is_tensor(a):
        if not a.is_cuda:
            a = a.cuda()
        tdt = {np.float64: torch.float64, np.float32: torch.float32}[dtype]
        return a.to(tdt).contiguous(), False
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=dtype)).cuda(), True


def _back(t, as_numpy):
    return t.cpu().numpy() if as_numpy else t


def hilbert_phase(signals):
    """Instantaneous phase of each ROI series (reference: signal_tools/_signal_tools.py:24-44).

    signals: (N_rois, N_time_points) -> phase (N_rois, N_time_points) float64 in (-pi, pi].
    """
    torch = nv.require_cuda()
    x, as_np = _to_device(signals, np.float64)
    if x.dim() != 2:
        raise Exception("'signals' must be a 2D array.")
    n, t = x.shape
    out = torch.empty((n, t), dtype=torch.float64, device=x.device)
    ws_bytes = nv.lib.leida_hilbert_workspace_bytes(t)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    nv.call("leida_hilbert_phase_f64", nv.ptr(x), n, t, t, nv.ptr(out), t, 0, t, nv.ptr(ws), ws_bytes, nv.stream())
    return _back(out, as_np)


def phase_coherence(signals_phases):
    """Phase-coherence matrices (reference: signal_tools/_signal_tools.py:137-183).

    signals_phases: (N, T) -> dFC (N, N, T-2), first and last volume dropped (:172).  The Leida
    pipeline never calls this (the N x N matrices are not materialised); it exists for API parity.
    """
    torch = nv.require_cuda()
    is_arr = isinstance(signals_phases, np.ndarray) or torch.is_tensor(signals_phases)
    if not is_arr or signals_phases.ndim != 2:
        raise Exception("'signals_phases' must be a 2D array.")
    ph, as_np = _to_device(signals_phases, np.float64)
    n, t = ph.shape
    if t < 3:
        raise Exception("'signals_phases' needs at least 3 time points.")
    out = torch.empty((n, n, t - 2), dtype=torch.float64, device=ph.device)
    nv.call("leida_phase_coherence_f64", nv.ptr(ph), n, t, t, nv.ptr(out), nv.stream())
    return _back(out, as_np)


def get_eigenvectors(dFC, n=1, tol=1e-12, max_iter=200, return_info=False):
    """Leading eigenvector of every dFC[:, :, t] (reference: signal_tools/_signal_tools.py:185-224).

    General symmetric input (blocked subspace iteration on the GPU).  Only n=1 is supported -- it
    is what Leida uses (_leida.py:288), and for n>1 the reference's block order is ARPACK's,
    which is not reproducible.
    """
    torch = nv.require_cuda()
    is_arr = isinstance(dFC, np.ndarray) or torch.is_tensor(dFC)
    if not is_arr or dFC.ndim != 3:
        raise Exception("'dFC' must be a 3D array!")
    if n != 1:
        raise NotImplementedError("only n=1 (the leading eigenvector) is available on the GPU path")
    a, as_np = _to_device(dFC, np.float64)
    N, N2, T = a.shape
    if N != N2:
        raise Exception("'dFC' must have shape (N_rois, N_rois, N_volumes).")
    out = torch.empty((T, N), dtype=torch.float64, device=a.device)
    info = torch.zeros(4, dtype=torch.int32, device=a.device)
    ws_bytes = nv.lib.leida_eigvec_dense_workspace_bytes(N, T)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=a.device)
    nv.call("leida_eigvec_dense_f64", nv.ptr(a), N, T, nv.ptr(out), N, float(tol), int(max_iter), nv.ptr(info),
            nv.ptr(ws), ws_bytes, nv.stream())
    if return_info:
        i = info.cpu().numpy()
        return _back(out, as_np), {"iterations": int(i[0]), "unconverged": int(i[1])}
    return _back(out, as_np)


def eigenvectors_from_phases(signals_phases, want_f32=False):
    """get_eigenvectors(phase_coherence(phases)) for n=1 without forming the N x N matrices
    (rank-2 closed form; signal_tools/_signal_tools.py:137-224).  phases: (N, T) or (S, N, T) ->
    (S*(T-2), N) fp64 rows (first and last volume dropped like :172)."""
    torch = nv.require_cuda()
    ph, as_np = _to_device(signals_phases, np.float64)
    if ph.dim() == 2:
        ph = ph.unsqueeze(0)
    if ph.dim() != 3:
        raise Exception("'signals_phases' must be a 2D array.")
    S, N, T = ph.shape
    if T < 3:
        raise Exception("'signals_phases' needs at least 3 time points.")
    M = S * (T - 2)
    lei = torch.empty((M, N), dtype=torch.float64, device=ph.device)
    xp = padded_pitch(N)
    xf = torch.empty((M, xp), dtype=torch.float32, device=ph.device) if want_f32 else None
    nv.call("leida_eigvec_rank2_f64", ph.data_ptr() + 8, S, N, T - 2, N * T, T, nv.ptr(lei), N, nv.ptr(xf), xp,
            nv.stream())
    lei = _back(lei, as_np)
    return (lei, xf) if want_f32 else lei


def leading_eigenvectors(time_series, want_f64=True, want_f32=True, out_f64=None, out_f32=None):
    """Fused stage 1+2 for a batch of subjects with the same shape: BOLD (S, N, T) ->
    eigenvector rows (S*(T-2), N) [fp64] and the zero-padded fp32 copy (S*(T-2), padded_pitch(N))
    that the k-means kernels read.  Equivalent to, per subject,
    get_eigenvectors(phase_coherence(hilbert_phase(x))) (_leida.py:287-289) via the rank-2 closed form.
    out_f64 / out_f32: optional preallocated CUDA row blocks ((S*(T-2), N) fp64 contiguous and
    (S*(T-2), padded_pitch(N)) fp32 contiguous) to write into (slices of a larger matrix).
    Returns (lei64 or None, x32 or None); x32 always stays on the device."""
    torch = nv.require_cuda()
    x, as_np = _to_device(time_series, np.float64)
    if x.dim() == 2:
        x = x.unsqueeze(0)
    if x.dim() != 3:
        raise Exception("'time_series' must have shape (S, N, T).")
    S, N, T = x.shape
    if T < 3:
        raise Exception("time series need at least 3 volumes.")
    M = S * (T - 2)
    xp = padded_pitch(N)
    lei = out_f64 if out_f64 is not None else (torch.empty((M, N), dtype=torch.float64, device=x.device) if want_f64 else None)
    xf = out_f32 if out_f32 is not None else (torch.empty((M, xp), dtype=torch.float32, device=x.device) if want_f32 else None)
    for o, shape, dt in ((lei, (M, N), torch.float64), (xf, (M, xp), torch.float32)):
        if o is not None and (tuple(o.shape) != shape or o.dtype != dt or not o.is_contiguous() or not o.is_cuda):
            raise ValueError(f"output block must be a contiguous CUDA {dt} tensor of shape {shape}")
    ws_bytes = nv.lib.leida_leading_eigvec_workspace_bytes(S, N, T)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    nv.call("leida_leading_eigvec_f64", nv.ptr(x), S, N, T, nv.ptr(lei), N, nv.ptr(xf), xp, nv.ptr(ws), ws_bytes,
            nv.stream())
    if as_np and lei is not None and out_f64 is None:
        lei = lei.cpu().numpy()
    return lei, xf
