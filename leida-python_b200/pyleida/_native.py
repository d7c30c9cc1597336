"""ctypes binding of libleida_b200.so (the C-ABI declared in include/leida_b200.h).

There is NO CPU fallback: if the library is missing the import fails, and every wrapper raises
when handed anything but CUDA tensors.  PyTorch is used for device memory and streams only.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LEIDA_B200_LIB") or os.path.join(_HERE, "_lib", "libleida_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the LEiDA CUDA library has not been built. "
        "Run `python -c 'import __graft_entry__ as g; g.build()'` (or leida-python_b200/csrc/build.py). "
        "This package has no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

LEIDA_KM_FRAC_BITS = 40
STATE_WORDS = 16
ST_CHANGED, ST_DIST_HI, ST_DIST_LO, ST_RECHECKED, ST_ACTIVE, ST_PASSES, ST_EMPTY, ST_K = range(8)
ST_RECHECK_TOTAL, ST_CHANGED_TOTAL, ST_TAU_WATCH = 8, 9, 10
KM_MAX_K = 64

_P = c_void_p
# name -> (restype, argtypes); every symbol of include/leida_b200.h is listed (tests check this).
SIGNATURES = {
    "leida_version": (c_int, []),
    "leida_last_error": (c_char_p, []),
    "leida_device_sm_count": (c_int, []),
    "leida_launch_count": (c_int64, []),
    "leida_clean_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "leida_clean_signals_f64": (c_int, [_P, c_int64, c_int, c_int64, _P, c_int64, c_int, c_int, _P, _P, _P, c_int, _P, c_size_t, _P]),
    "leida_hilbert_workspace_bytes": (c_size_t, [c_int]),
    "leida_hilbert_phase_f64": (c_int, [_P, c_int64, c_int, c_int64, _P, c_int64, c_int, c_int, _P, c_size_t, _P]),
    "leida_phase_coherence_f64": (c_int, [_P, c_int, c_int, c_int64, _P, _P]),
    "leida_eigvec_rank2_f64": (c_int, [_P, c_int64, c_int, c_int, c_int64, c_int64, _P, c_int64, _P, c_int64, _P]),
    "leida_leading_eigvec_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "leida_leading_eigvec_f64": (c_int, [_P, c_int64, c_int, c_int, _P, c_int64, _P, c_int64, _P, c_size_t, _P]),
    "leida_eigvec_dense_workspace_bytes": (c_size_t, [c_int, c_int]),
    "leida_eigvec_dense_f64": (c_int, [_P, c_int, c_int, _P, c_int64, c_double, c_int, _P, _P, c_size_t, _P]),
    "leida_kmeans_row_norms": (c_int, [_P, c_int64, c_int, c_int64, _P, _P]),
    "leida_kmeans_pack_rows": (c_int, [_P, c_int64, c_int, c_int64, _P, c_int64, _P]),
    "leida_kmeans_init": (c_int, [_P, c_int64, _P, c_int, _P, _P, c_int, c_int64, _P, _P, _P, _P, c_int64, _P, _P]),
    "leida_kmeans_assign": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P,
                                    c_int64, c_int, _P]),
    "leida_kmeans_recheck": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int, _P]),
    "leida_kmeans_tc_pitch": (c_int64, [c_int]),
    "leida_kmeans_tc_tile_columns": (c_int, []),
    "leida_kmeans_tc_split_rows": (c_int, [_P, c_int64, c_int, c_int64, _P, _P, _P]),
    "leida_kmeans_tc_plan": (c_int, [c_int, _P, _P, _P, _P, _P, c_int, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P]),
    "leida_kmeans_tc_assign": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, c_double, _P]),
    "leida_kmeans_tc_tau": (c_double, [c_int]),
    "leida_kmeans_tc_scores": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, _P]),
    "leida_kmeans_tc_apply": (c_int, [_P, c_int64, c_int, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "leida_kmeans_update": (c_int, [c_int, _P, _P, c_int, c_int64, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "leida_kmeans_update_peers": (c_int, [c_int, _P, _P, c_int, c_int64, _P, c_int, _P, c_int64, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "leida_kmeans_update_fused": (c_int, [c_int, _P, _P, c_int, c_int64, _P, c_int64, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P,
                                          _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "leida_kmeans_exchange_publish": (c_int, [c_int, _P, _P, c_int64, _P, c_int64, _P, _P, c_int, c_int, _P, _P, _P]),
    "leida_kmeans_distortion": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, _P]),
    "leida_kmeans_means_f32_sequential": (c_int, [_P, c_int64, c_int, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "leida_kmeans_transform_f64": (c_int, [_P, c_int64, c_int, c_int64, _P, c_int, c_int64, _P, _P, c_int, _P]),
    "leida_remap_labels_i32": (c_int, [_P, c_int64, _P, c_int, _P]),
    "leida_scores_cosine": (c_int, [_P, c_int64, c_int, _P, c_int64, _P, c_int64, c_int, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "leida_scores_db": (c_int, [_P, c_int64, c_int, c_int64, _P, c_int64, c_int, _P, _P, c_int, _P, _P]),
    "leida_remap_labels_batch_i32": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, _P]),
    "leida_stats_moments_f64": (c_int, [_P, c_int64, c_int, _P, c_int, _P, c_int, _P, _P]),
    "leida_perm_ttest_ind_f64": (c_int, [_P, c_int64, c_int, _P, c_int, _P, c_int, _P, c_int, _P, ctypes.c_uint64, _P, _P, _P, _P]),
    "leida_perm_paired_f64": (c_int, [_P, c_int64, c_int, _P, _P, c_int, c_int, _P, ctypes.c_uint64, _P, _P, _P, _P]),
    "leida_dynamics_counts_i32": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int, c_int, _P, _P, _P, _P]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = the library is stale / incomplete
    _fn.restype = _res
    _fn.argtypes = _args


class LeidaCudaError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib.leida_last_error().decode("utf-8", "replace")
        raise LeidaCudaError(f"{what or 'libleida_b200'} failed (code {rc}): {msg}")


def _torch():
    import torch
    return torch


def require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise LeidaCudaError("no CUDA device: the LEiDA hot path runs on the GPU only (no CPU fallback)")
    return torch


def ptr(t):
    """Device pointer of a CUDA tensor (or None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise LeidaCudaError("expected a CUDA tensor")
    return t.data_ptr()


def stream():
    return _torch().cuda.current_stream().cuda_stream


def call(name, *args):
    check(getattr(lib, name)(*args), name)


def prepared(name, *args):
    """A zero-argument callable for a call whose arguments never change (the per-pass launches of the k-means
    loop): the symbol lookup and the argument tuple are built once, so issuing it costs one ctypes call."""
    fn = getattr(lib, name)

    def run():
        rc = fn(*args)
        if rc != 0:
            check(rc, name)
    return run


def launch_count():
    return int(lib.leida_launch_count())
