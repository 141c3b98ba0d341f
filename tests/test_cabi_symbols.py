"""CPU: the C-ABI library loads and exports every symbol include/fmk.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

from finmlkit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fmk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = declared_symbols()
    assert len(names) >= 40
    L = ctypes.CDLL(_lib.SO_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in fmk.h but not exported: {missing}"


def test_python_binding_covers_header():
    names = set(declared_symbols())
    bound = set(_lib.SIGNATURES)
    assert names <= bound, f"not bound in _lib.py: {sorted(names - bound)}"
    assert bound <= names, f"bound but not declared in fmk.h: {sorted(bound - names)}"
    _lib.lib()   # resolves every symbol with its argtypes


def test_no_device_means_loud_failure():
    """Without a CUDA device the product path must fail, not fall back."""
    L = _lib.lib()
    if L.fmk_device_count() > 0:
        return
    import pytest
    from finmlkit_b200 import core
    with pytest.raises(core.FmkError):
        core.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "finmlkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "fmk_oracle" not in src, f
