"""CPU: pin the C restatement bit-for-bit against the UNMODIFIED reference hot path
(oracle/_ref/libcoati_ref.so = reference align_pair.cc + contrib/random behind oracle/ref_shim.cc).
Skipped when the reference objects were not built (they are prebuilt in the build container and
travel with the snapshot; /root/reference itself is never read at test time)."""
import numpy as np
import pytest

import oracle
from tests import util

pytestmark = pytest.mark.skipif(oracle.ref is None, reason="oracle/_ref not built")


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("semiring", [0, 1])
def test_fill_bit_exact(k, semiring, tables):
    rng = np.random.RandomState(100 + 10 * k + semiring)
    for trial in range(12):
        anc, des = util.random_pair(rng, n_codons=int(rng.randint(1, 50)), k=k, ambiguous=trial % 3 == 0)
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        if len(des) % k or len(anc) % k or not des:
            continue
        a, b = oracle.encode_pair(anc, des)
        T = tables[["mg_golden", "ecm_default", "mg_c5"][trial % 3]]
        g = np.float32([0.001, 0.01, 0.2][trial % 3])
        e = np.float32([5.0 / 6.0, 0.5, 0.9][trial % 3])
        if semiring == 0:
            mo = oracle.fill(0, a, b, T, g, e, k)
            mr = oracle.fill(0, a, b, T, g, e, k, impl="ref")
        else:
            *mo3, to = oracle.fill(1, a, b, T, g, e, k, with_trans=True)
            *mr3, tr = oracle.fill(1, a, b, T, g, e, k, with_trans=True, impl="ref")
            mo, mr = mo3, mr3
            # ref member order: mch_mch mch_del mch_ins del_mch del_del ins_mch ins_del ins_ins
            assert np.array_equal(_bits(to), _bits(tr))
        for x, y in zip(mo, mr):
            assert np.array_equal(_bits(x), _bits(y))


@pytest.mark.parametrize("k", [1, 2, 3])
def test_viterbi_strings_and_scores(k, tables):
    rng = np.random.RandomState(7 + k)
    for trial in range(25):
        anc, des = util.random_pair(rng, n_codons=int(rng.randint(1, 120)), k=k, ambiguous=trial % 4 == 0)
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        if len(des) % k or len(anc) % k:
            continue
        T = tables[["mg_golden", "ecm_default"][trial % 2]]
        o = oracle.viterbi(anc, des, T, k=k)
        r = oracle.viterbi(anc, des, T, k=k, impl="ref")
        assert o[:2] == r[:2]
        assert util.f32_bits(o[2]) == util.f32_bits(r[2])


def test_sampling_stream_identical(tables):
    rng = np.random.RandomState(99)
    for trial in range(8):
        k = [1, 3][trial % 2]
        anc, des = util.random_pair(rng, n_codons=int(rng.randint(2, 60)), k=k)
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        st = oracle.seed_state([str(trial), "s"])
        assert np.array_equal(st, oracle.ref_seed_state([str(trial), "s"]))
        o = oracle.sample(anc, des, tables["mg_golden"], st, 40, k=k)
        r = oracle.sample(anc, des, tables["mg_golden"], st, 40, k=k, impl="ref")
        assert o[0] == r[0]
        assert np.array_equal(_bits(o[1]), _bits(r[1]))
        assert np.array_equal(o[2], r[2])


def test_seed_strings_match_reference():
    for seeds in (["42"], [""], ["random42"], ["-5"], ["2147483648"], ["0042"], ["+7"], ["a", "b", "c"],
                  ["1", "2", "3", "4", "5", "6", "7", "8", "9", "10"]):
        assert np.array_equal(oracle.seed_state(seeds), oracle.ref_seed_state(seeds)), seeds
