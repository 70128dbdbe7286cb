"""B200-native LINS iterated-ESKF update path.

Layout (only what the hot path needs):
  csrc/cuda/   hand-written sm_100a kernels + the C-ABI shared library (include/lins_gpu.h)
  csrc/host/   host-side C++ mirror of the reference classes that stay on the CPU
  capi.py      ctypes binding of the C-ABI (the call a Python user makes)
  synth.py     synthetic scan-pair generator binding (inputs only)
"""
from .ctypes_defs import Batch, LinsParams, LinsReport, POINT_DTYPE, make_points  # noqa: F401
