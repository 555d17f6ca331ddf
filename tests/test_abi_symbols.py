"""The C-ABI library loads and exports exactly the symbols include/rbgtopo.h
declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rbgtopo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rbgtopo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from rbg_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rbgtopo.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, (set(names) ^ set(_lib.SIGNATURES))
    assert _lib.load().rbgtopo_abi_version() == 1


def test_no_device_fails_loudly_without_fallback():
    """Without a CUDA device create() must fail with ENODEVICE, never fall back."""
    import torch
    if torch.cuda.is_available():
        return
    from rbg_b200.engine import RbgTopoError, TopoPlacer
    try:
        TopoPlacer()
    except RbgTopoError as e:
        assert e.code == -2 and "no CPU path" in str(e)
    else:
        raise AssertionError("TopoPlacer() succeeded without a GPU")


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under rbg_b200/ may reference it."""
    pkg = os.path.join(ROOT, "rbg_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "liboracle" not in txt, f
