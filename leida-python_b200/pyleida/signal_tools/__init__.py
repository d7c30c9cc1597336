"""pyleida.signal_tools -- BOLD signal tools on the GPU (drop-in names of the reference module,
pyleida/signal_tools/__init__.py:4-18).  ``clean_signals`` follows nilearn 0.9.1's published algorithm (detrend,
Butterworth filtfilt, standardise) without depending on nilearn; ``txt_matrix`` (a plot helper) is not provided."""
from .phase import (
    hilbert_phase,
    phase_coherence,
    get_eigenvectors,
    leading_eigenvectors,
    eigenvectors_from_phases,
    padded_pitch,
)
from .clean import clean_signals, clean_array

__all__ = ["hilbert_phase", "phase_coherence", "get_eigenvectors", "leading_eigenvectors", "eigenvectors_from_phases",
           "padded_pitch", "clean_signals", "clean_array"]
