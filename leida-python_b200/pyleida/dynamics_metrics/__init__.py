"""pyleida.dynamics_metrics -- occupancy / dwell time / transition metrics from GPU integer counts
(drop-in names of pyleida/dynamics_metrics/__init__.py:4-26 for the compute part)."""
from .metrics import (
    compute_dynamics_metrics,
    fractional_occupancy,
    fractional_occupancy_group,
    dwell_times,
    dwell_times_group,
    transition_probabilities,
    transition_probabilities_group,
    dynamics_counts,
    Segments,
)

__all__ = [
    "compute_dynamics_metrics", "fractional_occupancy", "fractional_occupancy_group", "dwell_times",
    "dwell_times_group", "transition_probabilities", "transition_probabilities_group", "dynamics_counts",
    "Segments",
]
