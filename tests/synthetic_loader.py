"""Load pyleida/synthetic.py by path (it is pure numpy) so CPU-only tests that must not import the
CUDA-backed product package can still build the shared synthetic inputs."""
import importlib.util
import os


def load():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "leida_synthetic", os.path.join(root, "leida-python_b200", "pyleida", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
