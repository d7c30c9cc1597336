"""Module path of the reference's clustering classes (pyleida/clustering/_clustering.py).

The fitted models of a results folder are pickled (`clustering/models/model_k_{k}.pkl`,
_clustering.py:345-350) and pickles name their class by module path.  KMeansLeida therefore reports this module
as its home: a folder written by the GPU package is read by the reference's DataLoader (which unpickles into its own
pyleida.clustering._clustering.KMeansLeida: same attribute names, host arrays only), and a folder written by the
reference is read by this package's DataLoader."""
from .kmeans import KMeansLeida, identify_states, patterns_stability  # noqa: F401
