"""pyleida.signal_tools -- BOLD phase tools on the GPU (drop-in names of the reference module,
pyleida/signal_tools/__init__.py:4-18).  ``clean_signals`` and ``txt_matrix`` (nilearn band-pass,
plots) are outside the accelerated path and are not provided."""
from .phase import (
    hilbert_phase,
    phase_coherence,
    get_eigenvectors,
    leading_eigenvectors,
    eigenvectors_from_phases,
    padded_pitch,
)

__all__ = ["hilbert_phase", "phase_coherence", "get_eigenvectors", "leading_eigenvectors", "eigenvectors_from_phases",
           "padded_pitch"]
