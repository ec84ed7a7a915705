"""GPU (-m gpu): Forward fill + seeded sampleback through the C ABI against the oracle and the
committed outputs of the unmodified reference.  Bars: libm twins bit-identical to the host libm;
forward matrices bit-exact (tolerance stated in BASELINE north_star: 1e-4 relative -- asserted too);
samples identical to the reference for every seed (strings, float32 score bits, final RNG state)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import oracle
from tests import util

pytestmark = pytest.mark.gpu

SMP = util.load_json("sample_golden.json")
_libm = C.CDLL("libm.so.6")
for _n in ("expf", "logf", "log1pf"):
    getattr(_libm, _n).restype = C.c_float
    getattr(_libm, _n).argtypes = [C.c_float]


def _host(fn, xs):
    f = getattr(_libm, fn)
    return np.array([f(float(x)) for x in xs], dtype=np.float32)


@pytest.mark.parametrize("op,fn,lo,hi", [(0, "expf", -104.0, 0.0), (0, "expf", -20.0, 20.0),
                                         (1, "logf", 1e-30, 3.0), (1, "logf", 0.5, 1e6),
                                         (2, "log1pf", 0.0, 1.0), (2, "log1pf", 1e-9, 1e-3)])
def test_libm_twins_bit_identical(op, fn, lo, hi, gpu_ctx, tables):
    gpu_ctx.set_model(tables["mg_golden"])
    rng = np.random.RandomState(op * 7 + 1)
    xs = rng.uniform(lo, hi, 200_000).astype(np.float32)
    xs[:8] = np.float32([lo, hi, (lo + hi) / 2, lo, hi, lo, hi, lo])
    got = gpu_ctx.libm_eval(op, xs)
    want = _host(fn, xs)
    bad = int((got.view(np.uint32) != want.view(np.uint32)).sum())
    assert bad == 0, f"{fn}: {bad} of {len(xs)} results differ from the host libm"


def test_log1p_exp_matches_oracle(gpu_ctx, tables):
    gpu_ctx.set_model(tables["mg_golden"])
    xs = np.concatenate([np.random.RandomState(3).uniform(-40, 20, 100_000),
                         [-16.0, 8.0, 14.5, -16.000002, 8.000001, 14.500001, 0.0, -0.0]]).astype(np.float32)
    got = gpu_ctx.libm_eval(3, xs)
    oracle.lib.orc_log1p_exp.restype = C.c_float
    want = np.array([oracle.lib.orc_log1p_exp(C.c_float(float(x))) for x in xs], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _lattice(m, k):
    """reference (La+k) x (Lb+k) matrix -> lattice (La+1) x (Lb+1) view (drop k-1 padding rows/cols)"""
    return m[k - 1:, k - 1:]


@pytest.mark.parametrize("k", [1, 2, 3])
def test_forward_matrices_vs_oracle(k, gpu_ctx, tables):
    rng = np.random.RandomState(300 + k)
    T = tables["mg_golden"]
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, k)
    for trial in range(6):
        anc, des = util.random_pair(rng, n_codons=int(rng.randint(1, 70)), k=k, ambiguous=trial == 0)
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        a, b = oracle.encode_pair(anc, des)
        Mo, Do, Io = oracle.fill(1, a, b, T, k=k)
        fw = gpu_ctx.forward(a, b)
        M, D, I = fw.matrices()
        term, _ = fw.terminal()
        fw.free()
        Mo, Do, Io = (_lattice(x, k).copy() for x in (Mo, Do, Io))
        # the reference adjusts the terminal cell in place (align_pair.cc:130-138)
        assert [util.f32_bits(x) for x in term] == [util.f32_bits(Mo[-1, -1]), util.f32_bits(Do[-1, -1]),
                                                    util.f32_bits(Io[-1, -1])]
        for X, Xo in ((M, Mo), (D, Do), (I, Io)):
            X, Xo = X.copy(), Xo.copy()
            X[-1, -1] = Xo[-1, -1] = 0
            assert np.array_equal(X.view(np.uint32), Xo.view(np.uint32))
            np.testing.assert_allclose(X, Xo, rtol=1e-4)   # the stated tolerance (met with margin 0)


@pytest.mark.parametrize("s", SMP, ids=lambda s: s["name"])
def test_samples_vs_reference_golden(s, gpu_ctx, tables):
    anc, _ = oracle.trim_end_stop(s["anc"])
    des, _ = oracle.trim_end_stop(s["des"])
    a, b = oracle.encode_pair(anc, des)
    T = tables[s["table"]]
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, s["k"])
    fw = gpu_ctx.forward(a, b)
    st = np.array(s["state0"], dtype=np.uint64)
    rows, sc, st2, _ = fw.sampleback(anc, des, st, s["n"])
    term, _ = fw.terminal()
    fw.free()
    # match rate vs the oracle, sample by sample (reported on failure)
    orows, osc, ost, oll = oracle.sample(anc, des, T, st, s["n"], k=s["k"])
    same = sum(1 for x, y, p, q in zip(rows, orows, sc, osc) if x == y and util.f32_bits(p) == util.f32_bits(q))
    first_bad = next((i for i, (x, y) in enumerate(zip(rows, orows)) if x != y), None)
    assert same == s["n"], f"match rate {same}/{s['n']}, first mismatch at sample {first_bad}"
    assert [int(st2[0]), int(st2[1])] == s["state1"]
    h = hashlib.sha256()
    for (ra, rb), x in zip(rows, sc):
        h.update((ra + "\0" + rb + "\0" + util.f32_bits(x)).encode())
    assert h.hexdigest() == s["sha256"]
    for (ra, rb), x, f in zip(rows, sc, s["first"]):
        assert [ra, rb, util.f32_bits(x)] == f
    # forward log-likelihood: log_sum_exp of the adjusted terminal scores, within 1e-4 relative
    ll = oracle.lib.orc_log_sum_exp(C.c_float(oracle.lib.orc_log_sum_exp(C.c_float(term[0]), C.c_float(term[1]))),
                                    C.c_float(term[2]))
    assert ll == pytest.approx(float(oll), rel=1e-4)


def test_reference_sample_known_answers(gpu_ctx, tables):
    """align_marginal.cc:653-671 (seed "42")."""
    T = tables["mg_golden"]
    gpu_ctx.set_model(T)
    a, b = oracle.encode_pair("CCCCCC", "CCCCCCCC")
    fw = gpu_ctx.forward(a, b)
    rows, sc, _, _ = fw.sampleback("CCCCCC", "CCCCCCCC", oracle.seed_state(["42"]), 3)
    fw.free()
    assert [r[0] for r in rows] == ["CC--CCCC", "CCCCCC--", "CCCC--CC"]
    np.testing.assert_allclose(sc, [-1.9466571807861328, -1.9466569423675537, -1.9466572999954224], rtol=1e-6)
    a, b = oracle.encode_pair("CCCCCC", "CCCC")
    fw = gpu_ctx.forward(a, b)
    rows, sc, _, _ = fw.sampleback("CCCCCC", "CCCC", oracle.seed_state(["42"]), 1)
    fw.free()
    assert rows[0] == ("CCCCCC", "--CCCC")


def test_forward_errors(gpu_ctx, tables):
    import coati_b200
    gpu_ctx.set_model(tables["mg_golden"], gap_len=3)
    with pytest.raises(coati_b200.CoatiGpuError) as e:
        gpu_ctx.forward(np.zeros(3, np.uint8), np.zeros(4, np.uint8))   # Lb % k != 0
    assert e.value.code == -5
    with pytest.raises(coati_b200.CoatiGpuError) as e:
        gpu_ctx.forward(np.array([0, 1, 250], np.uint8), np.zeros(3, np.uint8))
    assert e.value.code == -4
