"""Seeded synthetic codon-sequence pairs for the benchmark workloads.

Bench / test infrastructure, NOT part of the product: its own small library (synth/libsynth.so, built by
`make -C synth` / __graft_entry__.build()), so that the reference arm of bench.py generates the same
inputs without mapping libcoati_gpu.so.  Pair p depends only on (seed, first + p): any subset or sharding of
a batch reproduces the same sequences."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libsynth.so")
_u64p = C.POINTER(C.c_uint64)
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "synth.cc")
    if force or not os.path.exists(_LIBPATH) or os.path.getmtime(src) > os.path.getmtime(_LIBPATH):
        subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", _LIBPATH,
                        src, "-lpthread"], check=True)
    return _LIBPATH


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.coati_synth_offsets.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_double,
                                             C.c_int, _u64p, _u64p]
        _lib.coati_synth_offsets.restype = None
        _lib.coati_synth_fill.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_int,
                                          _u64p, _u64p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.coati_synth_fill.restype = None
    return _lib


def synth_offsets(n: int, workload: int = 5, seed: int = 42, first: int = 0, sub: float = 0.05, indel: float = 0.005,
                  threads: int = 0):
    """CSR offsets (n + 1 entries each) of the pairs [first, first + n) of the seeded stream."""
    lib = _load()
    a_off = np.zeros(n + 1, dtype=np.uint64)
    b_off = np.zeros(n + 1, dtype=np.uint64)
    lib.coati_synth_offsets(seed, first, n, workload, sub, indel, threads or (os.cpu_count() or 1),
                            a_off.ctypes.data_as(_u64p), b_off.ctypes.data_as(_u64p))
    return a_off, b_off


def synth_pairs(n: int, workload: int = 5, seed: int = 42, first: int = 0, sub: float = 0.05,
                indel: float = 0.005, threads: int = 0, alloc=None):
    """Returns dict(a_off, b_off, a_all, b_all, anc_all, des_all): encoded (utils.cc:496-528) and raw symbols;
    `alloc(nbytes)` may supply pinned / shared uint8 buffers."""
    lib = _load()
    threads = threads or (os.cpu_count() or 1)
    a_off, b_off = synth_offsets(n, workload, seed, first, sub, indel, threads)
    alloc = alloc or (lambda nbytes: np.zeros(nbytes, dtype=np.uint8))
    ta, tb = int(a_off[-1]), int(b_off[-1])
    out = dict(a_off=a_off, b_off=b_off, a_all=alloc(ta + 1), b_all=alloc(tb + 1), anc_all=alloc(ta + 1),
               des_all=alloc(tb + 1))
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    lib.coati_synth_fill(seed, first, n, workload, sub, indel, threads, a_off.ctypes.data_as(_u64p),
                         b_off.ctypes.data_as(_u64p), vp(out["anc_all"]), vp(out["des_all"]), vp(out["a_all"]),
                         vp(out["b_all"]))
    return out
