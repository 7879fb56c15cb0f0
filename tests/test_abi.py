"""The C-ABI library loads and exports every symbol include/madicp_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from mad_icp_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("madicp_b200.h", "madicp_b200_debug.h"):  # the drop-in surface + the tuning/diagnostic header
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(mad(?:icp|tree)_[a-z_0-9]+)\s*\(", src))
    return sorted(names)


def test_debug_entry_points_are_not_in_the_drop_in_header():
    src = open(os.path.join(ROOT, "include", "madicp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert not re.findall(r"\bmadicp_(debug_[a-z_]+|set_gn_grid|set_walk_mode)\s*\(", src)


def test_header_symbols_all_exported(built):
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/madicp_b200.h but not exported"


def test_binding_table_matches_header(built):
    assert sorted(_capi.SYMBOLS) == _declared_symbols()


def test_no_cpu_fallback_without_gpu(built):
    """On a box without a GPU the device API must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mad_icp_b200 import MadIcpError, Registrar
    with pytest.raises(MadIcpError, match="no usable CUDA device|CUDA"):
        Registrar(device=0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mad_icp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle: product code must not touch it"


def test_host_entry_points_reject_bad_arguments(built):
    """madtree_build / madicp_deskew are host code: their argument checks run without a GPU."""
    import ctypes as C
    import numpy as np
    L = _capi.lib()
    pts = np.zeros((4, 3))
    dp = C.POINTER(C.c_double)
    T = np.eye(4)[:3].copy()
    ok = lambda a: a.ctypes.data_as(dp)
    assert L.madicp_deskew(None, 4, ok(T), ok(T), 10.0, 1) < 0
    assert L.madicp_deskew(ok(pts), 0, ok(T), ok(T), 10.0, 1) < 0
    assert L.madicp_deskew(ok(pts), 4, ok(T), ok(T), 0.0, 1) < 0
    assert L.madicp_deskew(ok(pts), 4, None, ok(T), 10.0, 1) < 0
    assert b"madicp_deskew" in L.madicp_last_error()
    assert L.madicp_deskew(ok(pts), 4, ok(T), ok(T), 10.0, 1) == 0  # identity motion, four points at the origin
    out = C.c_void_p()
    assert L.madtree_build(None, 4, 0.2, 0.1, 1, C.byref(out)) < 0
    assert L.madtree_build(ok(pts), 0, 0.2, 0.1, 1, C.byref(out)) < 0
    # b_max <= 0 / NaN would make `bbox(2) < b_max` unsatisfiable: one-point ranges would split for ever
    big = np.random.RandomState(0).normal(size=(100, 3))
    for bad in (0.0, -1.0, float("nan"), float("inf")):
        assert L.madtree_build(ok(big), 100, bad, 0.1, 1, C.byref(out)) < 0, bad
        assert b"b_max" in L.madicp_last_error()
    assert L.madtree_build(ok(big), 100, 0.2, float("nan"), 1, C.byref(out)) < 0
