"""CPU: the C-ABI library loads and exports every symbol that include/leida_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "leida_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(leida_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    path = os.path.join(ROOT, "leida-python_b200", "pyleida", "_lib", "libleida_b200.so")
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/leida_b200.h but not exported"
    lib.leida_version.restype = ctypes.c_int
    assert lib.leida_version() >= 100


def test_python_binding_lists_every_declared_symbol():
    from pyleida import _native
    assert sorted(_native.SIGNATURES) == declared_symbols()


def test_no_product_import_of_oracle():
    pkg = os.path.join(ROOT, "leida-python_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "leida_oracle" not in src and "ref_shim" not in src, f
