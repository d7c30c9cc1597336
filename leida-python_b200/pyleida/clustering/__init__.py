"""pyleida.clustering -- cosine k-means over the leading eigenvectors on the GPU (drop-in names of
pyleida/clustering/__init__.py:6-32 for the accelerated part: KMeansLeida, identify_states, patterns_stability, dunn_fast)."""
from .kmeans import KMeansLeida, KMeansEngine, identify_states, patterns_stability
from .scores import clustering_scores, dunn_fast

__all__ = ["KMeansLeida", "KMeansEngine", "identify_states", "clustering_scores", "dunn_fast", "patterns_stability"]
