"""struct3d_b200 -- B200-native (sm_100a) implementation of Struct3d's native structured-grid
linear-algebra path.

Layers:
  csrc/   hand-written CUDA kernels + the C ABI of include/s3d_cuda.h   -> lib/libs3d_cuda.so
  host/   C++ host classes mirroring the reference interface (SVec, SMat, SolverBase,
          createSolver, PDEBase ...) + a flat C wrapper (include/s3d_host.h) -> lib/libstruct3d_host.so
  capi.py / hostapi.py   ctypes bindings of the two libraries (tests, bench, smoke)

There is no CPU fallback: importing works anywhere, but creating a context without the built
CUDA library or without a GPU raises.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
