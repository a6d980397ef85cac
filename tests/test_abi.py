"""CPU test: libsw4b200.so loads and exports every symbol include/sw4b200.h declares; compute
entry points refuse to run without a CUDA device (no CPU fallback)."""
import os
import re
import ctypes as C
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sw4b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sw4b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import sw4lite_b200 as S
    lib = S.load()
    names = declared_symbols()
    assert len(names) > 40
    for n in names:
        assert hasattr(lib, n), n
    # and the binding types every one of them
    from sw4lite_b200.lib import SIGNATURES
    assert sorted(SIGNATURES) == names


def test_no_cpu_fallback():
    import torch
    import sw4lite_b200 as S
    lib = S.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.sw4b200_init(0) != 0
    assert b"no CPU fallback" in lib.sw4b200_last_error()
    assert lib.sw4b200_rhs4sg(1, 0, 9, 0, 9, 0, 9, 6, (C.c_int * 6)(), None, None, None, None, 1.0, None, None, None, None) != 0
    assert lib.sw4b200_malloc(64) is None


def test_builtin_coefficients_match_oracle():
    import numpy as np
    import sw4lite_b200 as S
    from oracle import port
    lib = S.load()
    a = [np.zeros(384), np.zeros(6), np.zeros(48), np.zeros(5)]
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.sw4b200_get_stencil_coefficients(*[dp(x) for x in a]) == 0
    acof, ghcof, bope, sbop = port.get_stencil_coefficients()
    assert np.array_equal(a[0], acof) and np.array_equal(a[1], ghcof)
    assert np.array_equal(a[2], bope) and np.array_equal(a[3], sbop)
