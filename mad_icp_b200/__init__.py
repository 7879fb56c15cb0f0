"""mad_icp_b200 -- B200 (sm_100a) implementation of MAD-ICP's per-scan registration hot path.

Layout: csrc/ (CUDA kernels + C ABI + host flat-tree builder), engine.py (ctypes handles),
api.py (reference-named facade: MADtree / MADicp / VectorEigen3d / Pipeline), synth.py (inputs).
"""
from .engine import FlatTree, Registrar, MadIcpError  # noqa: F401

__version__ = "0.1.0"
