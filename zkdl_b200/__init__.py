"""zkdl_b200 — B200-native prover for zkDL's fully-connected layer proof (sm_100a CUDA behind a C ABI).

The product is zkdl_b200/libzkdl_b200.so (hand-written CUDA, include/zkdl_b200.h) plus the reference-named C++ host
classes in zkdl_b200/host/.  This package is the ctypes binding used by the tests and bench.py.
"""
from . import capi  # noqa: F401
