"""clean_signals: detrending, Butterworth band-pass and standardisation of the BOLD series on the GPU.

Reference: pyleida/signal_tools/_signal_tools.py:66-133, a thin wrapper over nilearn.signal.clean (nilearn 0.9.1).
nilearn is not a dependency of this package: the options the reference exposes follow nilearn 0.9.1's published
algorithm (linear detrend -> scipy.signal.butter(5, ...) applied with filtfilt -> z-score / percent signal change);
the filter DESIGN (eleven scalar coefficients) is done on the host with scipy, every pass over the data runs in
libleida_b200.so (leida_clean_signals_f64).  filter_type='cosine' (nilearn's DCT drift regression) is not available.
"""
import numpy as np

from .. import _native as nv

_STANDARDIZE = {False: 0, None: 0, "zscore": 1, "psc": 2}


def butterworth_design(TR, low_pass=None, high_pass=None, order=5):
    """(b, a, zi) of nilearn.signal.butterworth's filter, or None when no cut-off is given."""
    from scipy import signal as sp_signal
    if low_pass is None and high_pass is None:
        return None
    if low_pass is not None and high_pass is not None and high_pass >= low_pass:
        raise ValueError(f"High pass cutoff frequency ({high_pass}) is greater than or equal to low pass filter frequency "
                         f"({low_pass}). This case is not handled by this function.")
    nyq = 0.5 / float(TR)
    crit, btype = [], None
    if high_pass is not None:
        btype = "high"
        crit.append(high_pass / nyq)
    if low_pass is not None:
        btype = "low"
        crit.append(min(low_pass / nyq, 1.0))
    if len(crit) == 2:
        btype = "band"
    else:
        crit = crit[0]
    b, a = sp_signal.butter(order, crit, btype=btype, output="ba")
    return np.ascontiguousarray(b, dtype=np.float64), np.ascontiguousarray(a, dtype=np.float64), \
        np.ascontiguousarray(sp_signal.lfilter_zi(b, a), dtype=np.float64)


def clean_array(x, detrend=True, standardize="zscore", filter_type=None, low_pass=None, high_pass=None, TR=None):
    """The same cleaning for one array of series, (..., T) with time last: numpy in -> numpy out, CUDA tensor in ->
    CUDA tensor out (the batched form `Leida` would call for all subjects of equal length at once)."""
    torch = nv.require_cuda()
    if standardize not in _STANDARDIZE:
        raise ValueError(f"{standardize} is no valid standardize strategy.")
    if filter_type not in (None, False, "butterworth"):
        if filter_type == "cosine":
            raise NotImplementedError("filter_type='cosine' (nilearn's DCT drift regression) is not available on the GPU path")
        raise ValueError(f"Filter method {filter_type} not implemented.")
    design = None
    if filter_type == "butterworth":
        if TR is None:
            raise ValueError("Repetition time (TR) must be specified for butterworth filtering.")
        design = butterworth_design(TR, low_pass, high_pass)
    as_np = not torch.is_tensor(x)
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda() if as_np else x.to(torch.float64).contiguous()
    if not xt.is_cuda:
        xt = xt.cuda()
    T = int(xt.shape[-1])
    n_series = int(xt.numel() // T) if T > 0 else 0
    n_coef = 0 if design is None else len(design[0])
    if n_coef and T <= 3 * n_coef:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {3 * n_coef}.")
    out = torch.empty_like(xt)
    ws_bytes = nv.lib.leida_clean_workspace_bytes(n_series, T, n_coef)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xt.device)
    b, a, zi = design if design is not None else (None, None, None)
    nv.call("leida_clean_signals_f64", nv.ptr(xt), n_series, T, T, nv.ptr(out), T, int(bool(detrend)),
            _STANDARDIZE[standardize], None if b is None else b.ctypes.data, None if a is None else a.ctypes.data,
            None if zi is None else zi.ctypes.data, n_coef, nv.ptr(ws), ws_bytes, nv.stream())
    return out.cpu().numpy() if as_np else out


def clean_signals(signals, detrend=True, standardize="zscore", filter_type=None, low_pass=None, high_pass=None, TR=None):
    """_signal_tools.py:66-133: `signals` is a dict subject -> (N_rois, N_volumes) array; returns the dict of cleaned
    arrays.  Subjects with the same shape are cleaned in one batch."""
    if not isinstance(signals, dict):
        raise ValueError("'signals' must be a dictionary!")
    groups = {}
    for sub, arr in signals.items():
        groups.setdefault(np.asarray(arr).shape, []).append(sub)
    cleaned = {}
    for shape, subs in groups.items():
        stack = np.stack([np.asarray(signals[s], dtype=np.float64) for s in subs])
        res = clean_array(stack, detrend, standardize, filter_type, low_pass, high_pass, TR)
        for i, s in enumerate(subs):
            cleaned[s] = res[i]
    return {s: cleaned[s] for s in signals}
