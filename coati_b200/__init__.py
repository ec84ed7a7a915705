"""coati_b200 -- B200-native marginal Gotoh hot path of COATi behind a C ABI.

The product is ``libcoati_gpu.so`` (include/coati_gpu.h): hand-written sm_100a CUDA kernels plus
the C++ host layer.  This Python package is only the ctypes binding used by tests/ and bench.py;
it never computes alignments itself and it raises if the CUDA library is missing.
"""
from .capi import (Context, CoatiGpuError, load_library, library_path)  # noqa: F401

__all__ = ["Context", "CoatiGpuError", "load_library", "library_path"]
