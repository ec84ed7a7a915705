"""CPU: the host restatement of glibc's expf/logf/log1pf (oracle/libm_ports.c -- the same algorithms
the device runs in coati_b200/csrc/devmath.cuh) against the running libm on strided sweeps of float
bit patterns.  (Exhaustive sweeps were run once, see the file header; these keep it pinned.)"""
import ctypes as C
import struct

import pytest

import oracle

oracle.lib.orc_libm_check.restype = C.c_uint64


def bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


@pytest.mark.parametrize("op,lo,hi", [
    (0, bits(0.0), bits(89.5)), (0, bits(-0.0), bits(-104.5)),          # expf
    (1, 1, 0x7f7fffff),                                                 # logf: every positive float
    (2, bits(0.0), bits(1e30)), (2, bits(-0.0), bits(-0.99999994)),     # log1pf
])
def test_port_matches_running_libm(op, lo, hi):
    n = C.c_uint64(0)
    bad = oracle.lib.orc_libm_check(op, C.c_uint32(lo), C.c_uint32(hi), C.c_uint32(509), C.byref(n))
    assert n.value > 1_000_000
    assert bad == 0
