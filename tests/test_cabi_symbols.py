"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/navsim_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_gpu

HEADER = os.path.join(ROOT, "include", "navsim_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nvb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    return ctypes.CDLL(g.LIB)


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "%s declared in navsim_b200.h but not exported" % n


def test_binding_covers_header():
    from navsim import _cabi
    assert sorted(_cabi.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback(lib):
    """Without a B200 the engine refuses to start instead of computing on the CPU."""
    if has_gpu():
        pytest.skip("a GPU is present")
    from navsim import _cabi
    import numpy as np
    import navsim
    with pytest.raises(_cabi.NavsimB200Error):
        navsim.NavEngine(np.zeros((64, 64, 3), np.uint8), (8, 2), 1.0)
    lib.nvb_last_error.restype = ctypes.c_char_p
    assert b"no CPU fallback" in lib.nvb_last_error() or b"sm_100a" in lib.nvb_last_error()


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import,
    load or link it."""
    pkg = os.path.join(ROOT, "navigation-by-deja-vu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "navsim_oracle" not in src and "ref_loader" not in src, f
