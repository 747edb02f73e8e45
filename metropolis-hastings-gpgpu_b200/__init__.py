"""B200-native parallel Metropolis-Hastings layout optimiser -- host-side Python mirror.

The product is libKernel.so (csrc/: hand-written sm_100a CUDA + a plain-C host wrapper behind
the C ABI of include/mh_kernel.h).  This package only (1) mirrors the wire structs as numpy
dtypes, (2) generates the synthetic rooms of BASELINE.json, and (3) binds the C ABI with
ctypes for the tests and bench.py.  It never computes a cost itself and has no CPU fallback.
"""
from . import dist, layout, synth  # noqa: F401
from .binding import Kernel, KernelError, lib_path  # noqa: F401
