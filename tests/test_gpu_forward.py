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


def _pair(rng, n_codons, k=1, ambiguous=False):
    anc, des = util.random_pair(rng, n_codons=n_codons, k=k, ambiguous=ambiguous)
    return oracle.trim_end_stop(anc)[0], oracle.trim_end_stop(des)[0]


@pytest.mark.parametrize("n_codons", [4, 40, 350, 1400])
def test_forward_band_kernel_matrices_bit_exact(n_codons, gpu_ctx, tables):
    """k = 1 runs the banded register pipeline (forward_band.cuh), a single pair as a wavefront over the SMs:
    every cell of M, D, I equals the oracle's bits, up to 4200 x 4200 (one, several and hundreds of bands,
    ragged last band), on two tables."""
    rng = np.random.RandomState(n_codons)
    for tname in ("mg_golden", "ecm_default"):
        T = tables[tname]
        gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
        anc, des = _pair(rng, n_codons, ambiguous=tname == "ecm_default")
        a, b = oracle.encode_pair(anc, des)
        Mo, Do, Io = oracle.fill(1, a, b, T)
        fw = gpu_ctx.forward(a, b)
        M, D, I = fw.matrices()
        term, _ = fw.terminal()
        fw.free()
        assert [util.f32_bits(x) for x in term] == [util.f32_bits(Mo[-1, -1]), util.f32_bits(Do[-1, -1]),
                                                    util.f32_bits(Io[-1, -1])]
        for X, Xo in ((M, Mo), (D, Do), (I, Io)):
            X, Xo = X.copy(), Xo.copy()
            X[-1, -1] = Xo[-1, -1] = 0     # the reference adjusts the terminal cell in place
            assert np.array_equal(X.view(np.uint32), Xo.view(np.uint32))


def test_forward_generic_kernel_equals_band_kernel(tables):
    """COATI_GPU_FORWARD_GENERIC=1 (the any-k kernel on k = 1) against the band kernel and the oracle."""
    import os
    import coati_b200
    rng = np.random.RandomState(9)
    anc, des = _pair(rng, 120)
    a, b = oracle.encode_pair(anc, des)
    T = tables["mg_golden"]
    Mo, Do, Io = oracle.fill(1, a, b, T)
    os.environ["COATI_GPU_FORWARD_GENERIC"] = "1"
    try:
        ctx = coati_b200.Context(0)
    finally:
        del os.environ["COATI_GPU_FORWARD_GENERIC"]
    ctx.set_model(T)
    fw = ctx.forward(a, b)
    M, D, I = fw.matrices()
    fw.free()
    ctx.close()
    for X, Xo in ((M, Mo), (D, Do), (I, Io)):
        X, Xo = X.copy(), Xo.copy()
        X[-1, -1] = Xo[-1, -1] = 0
        assert np.array_equal(X.view(np.uint32), Xo.view(np.uint32))


@pytest.mark.parametrize("k,npairs", [(1, 5), (1, 70), (3, 40)])
def test_forward_batch_and_per_pair_sampling(k, npairs, gpu_ctx, tables):
    """coati_gpu_forward_batch + coati_gpu_sampleback_batch: every pair's adjusted terminal scores, forward
    log-likelihood and its seeded samples (own RNG stream per pair) equal the oracle's, for a batch run as
    per-pair wavefronts (<= 16 pairs), as one warp per pair (k = 1) and by the any-k kernel (k = 3); empty and
    one-sided pairs included."""
    from coati_b200.capi import PackedPairs
    rng = np.random.RandomState(500 + npairs)
    T = tables["mg_golden"]
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, k)
    ancs, dess = [], []
    while len(ancs) < npairs - 3:
        anc, des = _pair(rng, int(rng.randint(1, 60)), k=k, ambiguous=len(ancs) % 5 == 0)
        if len(anc) % k or len(des) % k:
            continue
        ancs.append(anc), dess.append(des)
    ancs += ["", "AAA", ""]
    dess += ["", "", "ACG"]
    As, Bs = zip(*[oracle.encode_pair(x, y) for x, y in zip(ancs, dess)])
    fb = gpu_ctx.forward_batch(PackedPairs(list(As), list(Bs), ancs, dess))
    term, ll, _ = fb.terminal()
    nsamp = 6
    states = np.array([oracle.seed_state([str(1000 + p)]) for p in range(npairs)], dtype=np.uint64)
    rows, sc, st2, _ = fb.sampleback(states, nsamp)
    fb.free()
    for p in range(npairs):
        Mo, Do, Io = oracle.fill(1, As[p], Bs[p], T, k=k)
        assert [util.f32_bits(x) for x in term[p]] == [util.f32_bits(Mo[-1, -1]), util.f32_bits(Do[-1, -1]),
                                                       util.f32_bits(Io[-1, -1])], p
        orows, osc, ost, oll = oracle.sample(ancs[p], dess[p], T, states[p], nsamp, k=k)
        assert ll[p] == pytest.approx(float(oll), rel=1e-4), p
        assert rows[p] == orows, p
        assert np.array_equal(sc[p].view(np.uint32), osc.view(np.uint32)), p
        assert np.array_equal(st2[p], ost), p


def test_branch_free_log1p_exp_on_every_float_of_its_domain(gpu_ctx, tables):
    """The banded Forward kernel evaluates log1p_exp (utils.hpp:134-146) through straight-line twins of glibc's
    expf / log1pf that are only valid for y <= 0 (devmath.cuh: log1p_exp_neg).  Compared here with the host libm
    (oracle: orc_log1p_exp -> expf, log1pf of the box's glibc) on EVERY float in [-104.5, -0.0] (1.12e9 inputs:
    below -104 expf underflows to 0 on both sides), plus +0 and the most negative finite value."""
    from concurrent.futures import ThreadPoolExecutor
    import struct
    gpu_ctx.set_model(tables["mg_golden"])
    lo_bits = struct.unpack("<I", struct.pack("<f", -104.5))[0]
    fn = oracle.lib.orc_log1p_exp_array
    fn.restype = None
    chunk = 1 << 26
    nthreads = 16

    def host(xs):
        out = np.empty_like(xs)
        parts = np.array_split(np.arange(len(xs)), nthreads)
        def run(idx):
            if len(idx):
                a, b = xs[idx[0]:idx[-1] + 1], out[idx[0]:idx[-1] + 1]
                fn(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)))
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(run, parts))
        return out

    total = bad = 0
    for first in range(0x80000000, lo_bits + 1, chunk):
        n = min(chunk, lo_bits + 1 - first)
        xs = (np.arange(n, dtype=np.uint64) + first).astype(np.uint32).view(np.float32)
        got = gpu_ctx.libm_eval(4, xs)
        want = host(xs)
        bad += int((got.view(np.uint32) != want.view(np.uint32)).sum())
        total += n
    xs = np.float32([0.0, -3.4028234663852886e38, -1e30, -16.0, -15.999999, -16.000002])
    bad += int((gpu_ctx.libm_eval(4, xs).view(np.uint32) != host(xs).view(np.uint32)).sum())
    assert total > 1.1e9 and bad == 0, f"{bad} of {total} inputs differ from the host libm"
