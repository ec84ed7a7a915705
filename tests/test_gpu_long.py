"""GPU (-m gpu): long single pairs (BASELINE config 3: example-10k .. 160k, benchmark_32k) and every
environment-selected kernel path, bit-exact through the C ABI.

Validators for pairs the full-matrix oracle cannot hold (oracle/long_pair.c, pinned to the oracle in
tests/test_oracle_long.py):
  (a) the returned path re-scored through forward_impl's own terms and association
      (/root/reference/src/lib/align_pair.cc:81-138) must equal the GPU score bit for bit -- the path IS an
      arg-max chain of the recurrence;
  (b) a score-only rolling-row CPU Viterbi must give the same bits -- the GPU score IS the optimum;
and for 20k / 40k the unmodified reference itself (oracle/_ref, 4.8 / 19 GB of matrices)."""
import os

import numpy as np
import pytest

import oracle
from tests import util

pytestmark = pytest.mark.gpu


def _load(name):
    (_, anc), (_, des) = util.load_fasta(name)
    if name.startswith("example-"):
        anc = util.sanitise_ancestor(anc)     # SURVEY 8(d) C3 protocol
    anc, _ = oracle.trim_end_stop(anc)
    des, _ = oracle.trim_end_stop(des)
    return anc, des


def _check_exact(anc, des, a, b, ra, rb, sc, T):
    assert len(ra) == len(rb)
    assert ra.replace("-", "") == anc and rb.replace("-", "") == des
    assert util.f32_bits(oracle.path_score(ra, rb, a, b, T)) == util.f32_bits(sc), "path re-score != GPU score"
    assert util.f32_bits(oracle.viterbi_score(a, b, T)) == util.f32_bits(sc), "rolling-row optimum != GPU score"


@pytest.mark.parametrize("name", ["benchmark_32k", "example-20k", "example-40k", "example-80k", "example-160k"])
def test_long_pairs_bit_exact(name, gpu_ctx, tables):
    """Default configuration: wavefront fill (R = 4 up to 100k rows, R = 10 beyond), run-at-a-time traceback,
    segment-parallel expansion."""
    anc, des = _load(name)
    T = tables["mg_golden"]
    a, b = oracle.encode_pair(anc, des)
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
    ra, rb, sc = gpu_ctx.viterbi(a, b, anc, des)
    _check_exact(anc, des, a, b, ra, rb, sc, T)


def _avail_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2**30
    except Exception:
        return 0.0


@pytest.mark.parametrize("name,need_gb", [("example-20k", 8), ("example-40k", 30)])
def test_long_pairs_vs_unmodified_reference(name, need_gb, gpu_ctx, tables):
    """The reference's own viterbi_mem + traceback_viterbi (oracle/_ref) on the 20k and 40k pairs: identical
    rows and score bits."""
    if oracle.ref is None:
        pytest.skip("oracle/_ref not built")
    if _avail_gb() < need_gb:
        pytest.skip(f"needs {need_gb} GB of host memory for the reference's matrices")
    anc, des = _load(name)
    T = tables["mg_golden"]
    a, b = oracle.encode_pair(anc, des)
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
    ra, rb, sc = gpu_ctx.viterbi(a, b, anc, des)
    oa, ob, osc = oracle.viterbi(anc, des, T, impl="ref", enc=(a, b))
    assert (ra, rb) == (oa, ob)
    assert util.f32_bits(sc) == util.f32_bits(osc)


def _ctx_with(env):
    import coati_b200
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return coati_b200.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.mark.parametrize("wave_r", [2, 4, 8, 10])
@pytest.mark.parametrize("name", ["benchmark_4k", "example-10k"])
def test_every_wavefront_configuration(name, wave_r, tables):
    """COATI_GPU_WAVE_R in {2, 4, 8, 10}: every registered wavefront kernel (16- and 4-column variants: the
    raw-sequence entry point picks by the symbols it finds) against the oracle / the reference's golden."""
    anc, des = _load(name)
    T = tables["mg_golden"]
    a, b = oracle.encode_pair(anc, des)
    if name == "example-10k":
        c = next(c for c in util.load_json("viterbi_golden.json") if c["name"] == "example-10k:sanitised")
        want_bits = c["score_bits"]
        want_rows = None
    else:
        oa, ob, osc = oracle.viterbi(anc, des, T, enc=(a, b))
        want_bits, want_rows = util.f32_bits(osc), (oa, ob)
    ctx = _ctx_with({"COATI_GPU_WAVE_R": str(wave_r)})
    ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
    ra, rb, sc = ctx.viterbi(a, b, anc, des)                     # encoded entry: column variant from the host scan
    assert util.f32_bits(sc) == want_bits
    if want_rows:
        assert (ra, rb) == want_rows
    _check_exact(anc, des, a, b, ra, rb, sc, T)
    # raw entry: both column variants launched, the device flag picks (here: ACGT only -> 4 columns)
    rows_a, rows_b, score, status = ctx.alignpair_batch([anc], [des])
    assert status[0] == 0 and (rows_a[0], rows_b[0]) == (ra, rb) and util.f32_bits(score[0]) == want_bits
    # an ambiguity code in the descendant -> the 16-column variant
    des_n = des[:100] + "N" + des[101:]
    a2, b2 = oracle.encode_pair(anc, des_n)
    rows_a, rows_b, score, status = ctx.alignpair_batch([anc], [des_n])
    assert status[0] == 0
    _check_exact(anc, des_n, a2, b2, rows_a[0], rows_b[0], score[0], T)
    ctx.close()


@pytest.mark.parametrize("env", [{"COATI_GPU_FORCE_R": "4"}, {"COATI_GPU_FORCE_R": "8"}, {"COATI_GPU_FORCE_R": "10"},
                                 {"COATI_GPU_PIPE_SCALAR": "1"}, {"COATI_GPU_PIPE_SCALAR": "1", "COATI_GPU_FORCE_R": "8"},
                                 {"COATI_GPU_TB_SERIAL": "1"}, {"COATI_GPU_NO_WAVE": "1", "COATI_GPU_FORCE_R": "10"}],
                         ids=lambda e: ",".join(f"{k[10:]}={v}" for k, v in e.items()))
def test_every_inter_pair_configuration(env, tables):
    """Every inter-pair fill the planner can be forced into (rows per lane 4 / 8 / 10, the scalar template, the
    column-at-a-time traceback), k = 1 and k = 3, ACGT-only and ambiguous batches, encoded and raw entry points:
    rows and score bits equal the oracle's."""
    from coati_b200.capi import PackedPairs
    ctx = _ctx_with(env)
    rng = np.random.RandomState(4242)
    for k, tname in ((1, "mg_c5"), (3, "ecm_default")):
        if k == 3 and "COATI_GPU_FORCE_R" in env:
            continue  # FORCE_R names K = 1 tiles
        T = tables[tname]
        ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, k)
        for ambiguous in (False, True):
            ancs, dess, As, Bs = [], [], [], []
            while len(ancs) < 40:
                # up to 1300 nt: several 320-row bands per pair, ragged last bands
                anc, des = util.random_pair(rng, int(rng.randint(1, 430)), k=k, ambiguous=ambiguous and len(ancs) % 2 == 0)
                anc, _ = oracle.trim_end_stop(anc)
                des, _ = oracle.trim_end_stop(des)
                if len(anc) % k or len(des) % k:
                    continue
                ea, eb = oracle.encode_pair(anc, des)
                ancs.append(anc), dess.append(des), As.append(ea), Bs.append(eb)
            want = [oracle.viterbi(ancs[p], dess[p], T, k=k, enc=(As[p], Bs[p])) for p in range(40)]
            rows_a, rows_b, score, status = ctx.viterbi_batch(PackedPairs(As, Bs, ancs, dess))
            assert (status == 0).all()
            for p in range(40):
                assert (rows_a[p], rows_b[p]) == want[p][:2], p
                assert util.f32_bits(score[p]) == util.f32_bits(want[p][2]), p
            rows_a, rows_b, score, status = ctx.alignpair_batch(ancs, dess)
            assert (status == 0).all()
            for p in range(40):
                assert (rows_a[p], rows_b[p]) == want[p][:2], p
                assert util.f32_bits(score[p]) == util.f32_bits(want[p][2]), p
    ctx.close()


def test_serial_sampleback_equals_reference_golden(tables):
    """COATI_GPU_SAMPLE_SERIAL=1 (one thread draws every sample, the reference's own order of operations) against
    the committed samples of the unmodified reference, as the parallel path is in test_gpu_forward.py."""
    import hashlib
    ctx = _ctx_with({"COATI_GPU_SAMPLE_SERIAL": "1"})
    for s in util.load_json("sample_golden.json"):
        if s["n"] > 200:
            continue  # the serial walk is the slow A/B form: 1000-sample cases stay with the parallel path
        anc, _ = oracle.trim_end_stop(s["anc"])
        des, _ = oracle.trim_end_stop(s["des"])
        a, b = oracle.encode_pair(anc, des)
        ctx.set_model(tables[s["table"]], oracle.DEFAULT_G, oracle.DEFAULT_E, s["k"])
        fw = ctx.forward(a, b)
        rows, sc, st2, _ = fw.sampleback(anc, des, np.array(s["state0"], dtype=np.uint64), s["n"])
        fw.free()
        assert [int(st2[0]), int(st2[1])] == s["state1"], s["name"]
        h = hashlib.sha256()
        for (ra, rb), x in zip(rows, sc):
            h.update((ra + "\0" + rb + "\0" + util.f32_bits(x)).encode())
        assert h.hexdigest() == s["sha256"], s["name"]
    ctx.close()


def test_stale_forward_handle_is_rejected(gpu_ctx, tables):
    """A Forward handle belongs to the model it was filled under (include/coati_gpu.h)."""
    import coati_b200
    gpu_ctx.set_model(tables["mg_golden"])
    a, b = oracle.encode_pair("CCCCCC", "CCCCCCCC")
    fw = gpu_ctx.forward(a, b)
    gpu_ctx.set_model(tables["ecm_default"])
    with pytest.raises(coati_b200.CoatiGpuError) as e:
        fw.sampleback("CCCCCC", "CCCCCCCC", oracle.seed_state(["42"]), 4)
    assert e.value.code == -2
    fw.free()
