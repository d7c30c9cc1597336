"""pyleida (B200-native build): the LEiDA hot path -- Hilbert phase, leading eigenvector of the
phase-coherence matrix, cosine k-means over a K sweep, dynamics metrics -- behind the reference's
call surface (pyleida/__init__.py:32-33 and the three stage modules), executed by hand-written
sm_100a CUDA kernels in libleida_b200.so.  No CPU fallback."""
from . import _native  # noqa: F401  (fails loudly when the CUDA library is missing)
from . import signal_tools, clustering, dynamics_metrics, stats  # noqa: F401
from ._leida import Leida
from ._data_loader import DataLoader

__all__ = ["Leida", "DataLoader", "signal_tools", "clustering", "dynamics_metrics", "stats"]
__version__ = "0.1.0"
