"""CPU: the addressing of rows_to_host_kernel (coati_b200/csrc/traceback.cuh), restated lane by lane.  A row of
n_out bytes goes to a host slot of any alignment as aligned 16-byte words plus at most two shared edge units;
several GPUs write neighbouring slots of one arena at the same time, so the invariant is: every byte of
[mis, mis + n_out) is written exactly once, from the right source byte, and no byte outside is touched.  (The GPU
test tests/test_gpu_multi.py::test_rows_written_straight_into_page_locked_arenas checks the kernel itself.)"""
import numpy as np
import pytest


def _warp_pass(n_out, mis):
    hi = mis + n_out
    written = np.zeros(((hi + 511) // 512) * 512 + 16, np.int32)
    src_of = np.full(written.shape, -1, np.int64)
    words = bytes_alone = 0
    base = 0
    while base < hi:                                    # for(base = 0; base < hi; base += 512)
        part = []
        for lane in range(32):
            v0 = base + 16 * lane
            full = v0 >= mis and v0 + 16 <= hi
            if full:                                    # one aligned 16-byte store
                assert v0 % 16 == 0
                written[v0:v0 + 16] += 1
                src_of[v0:v0 + 16] = np.arange(v0 - mis, v0 - mis + 16)
                words += 1
            elif v0 < hi and v0 + 16 > mis:             # ballot: an edge unit shared with a neighbouring slot
                part.append(lane)
        for pl in part:                                 # sixteen lanes, a byte each, one instruction
            for lane in range(16):
                v = base + 16 * pl + lane
                if mis <= v < hi:
                    written[v] += 1
                    src_of[v] = v - mis
                    bytes_alone += 1
        base += 512
    return written, src_of, words, bytes_alone


@pytest.mark.parametrize("mis", range(16))
def test_every_byte_once_and_nothing_else(mis):
    rng = np.random.RandomState(mis)
    sizes = list(range(1, 40)) + [495, 496, 497, 511, 512, 513, 1023, 1024, 1025, 4803] + list(rng.randint(1, 5000, 40))
    for n_out in sizes:
        written, src_of, words, alone = _warp_pass(int(n_out), mis)
        hi = mis + n_out
        assert (written[mis:hi] == 1).all() and written[:mis].sum() == 0 and written[hi:].sum() == 0, (n_out, mis)
        assert np.array_equal(src_of[mis:hi], np.arange(n_out))
        # the bulk leaves as words: at most two edge units (< 16 bytes each) go byte-wise
        assert alone <= 30 and 16 * words + alone == n_out
