"""gten-b200: B200-native (sm_100a) forward hot path of tinyllama.cpp behind the gten API.

Layout: csrc/ holds the CUDA kernels and the C-ABI (include/gten_b200.h) built into
libgten_b200.so; capi.py is the ctypes binding used by tests and bench.py; weights.py is the
gten-format tooling (synthetic weights, converter quantisers, .gten reader/writer).
The C++ drop-in headers are in include/gten/.
"""
from . import weights  # noqa: F401
