"""The C-ABI shared library loads on a CPU-only box and exports every function that
include/poccala_b200.h declares (no compute call is made without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "poccala_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_function():
    from poccala_b200 import _native, build

    lib_path = build.build(force=False)
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: %s" % n
    # the ctypes binding table covers the header (same names, nothing stale)
    assert set(_native.EXPORTS) <= set(names)
    missing = set(names) - set(_native.EXPORTS) - {"pc_debug_read", "pc_debug_read_fb"}
    assert not missing, missing
    assert lib.pc_abi_version() == 3


def test_product_path_has_no_cpu_fallback():
    """No module under poccala_b200/ imports the oracle, and creating an engine without a CUDA
    device fails loudly."""
    import torch

    pkg = os.path.join(ROOT, "poccala_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    if not torch.cuda.is_available():
        from poccala_b200.engine import Engine

        with pytest.raises(RuntimeError):
            Engine(0)
