"""The C-ABI library loads and exports every symbol include/madicp_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from mad_icp_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "madicp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mad(?:icp|tree)_[a-z_0-9]+)\s*\(", src)))


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
