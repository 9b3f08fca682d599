"""B200-native Poisson-GPFA EM hot path behind the API of mackelab/poisson-gpfa.

Sub-modules mirror the reference (funs/util.py, funs/inference.py, funs/learning.py, funs/engine.py);
``kernels`` and ``_lib`` are the ctypes layer over libpgpfa_b200.so (include/pgpfa_b200.h).
Importing the package does not need a GPU; computing anything does (no CPU fallback).
"""
__version__ = "0.1.0"
