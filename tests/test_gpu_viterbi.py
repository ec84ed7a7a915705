"""GPU (-m gpu): the CUDA Viterbi path, called through the C ABI (include/coati_gpu.h), against
  * the committed outputs of the unmodified reference (tests/golden/viterbi_golden.json),
  * the CPU oracle on seeded random pairs (bit-exact rows, scores and direction bytes),
  * size-independent properties at sizes the oracle cannot reach quickly.
Bar: BIT-EXACT alignment rows and float32 scores."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from coati_b200.capi import PackedPairs
from tests import util

pytestmark = pytest.mark.gpu

VIT = util.load_json("viterbi_golden.json")


def _inputs(c):
    if "anc" in c:
        return c["anc"], c["des"]
    (_, anc), (_, des) = util.load_fasta(c["file"])
    if c.get("sanitised"):
        anc = util.sanitise_ancestor(anc)
    return anc, des


def _sha(a, b):
    h = hashlib.sha256()
    for part in (a, b):
        h.update(part.encode())
        h.update(b"\0")
    return h.hexdigest()


@pytest.mark.parametrize("c", VIT, ids=lambda c: c["name"])
def test_reference_goldens(c, gpu_ctx, tables):
    anc, des = _inputs(c)
    g, e = util.bits_to_f32(c["g_bits"]), util.bits_to_f32(c["e_bits"])
    anc_t, s0 = oracle.trim_end_stop(anc)
    des_t, s1 = oracle.trim_end_stop(des)
    a, b = oracle.encode_pair(anc_t, des_t)
    gpu_ctx.set_model(tables[c["table"]], g, e, c["k"])
    ra, rb, sc = gpu_ctx.viterbi(a, b, anc_t, des_t)
    assert util.f32_bits(sc) == c["score_bits"]
    ra, rb, sc = oracle.restore_end_stops(ra, rb, sc, (s0, s1), g, e)
    assert util.f32_bits(sc) == c["final_score_bits"]
    assert len(ra) == c["len"]
    if "aln_a" in c:
        assert (ra, rb) == (c["aln_a"], c["aln_b"])
    assert _sha(ra, rb) == c["sha256"]


def _random_batch(rng, n, k, max_codons, ambiguous_every=4):
    ancs, dess, As, Bs = [], [], [], []
    while len(ancs) < n:
        anc, des = util.random_pair(rng, n_codons=int(rng.randint(1, max_codons)), k=k,
                                    ambiguous=len(ancs) % ambiguous_every == 0)
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        if len(anc) % k or len(des) % k:
            continue
        a, b = oracle.encode_pair(anc, des)
        ancs.append(anc), dess.append(des), As.append(a), Bs.append(b)
    return ancs, dess, As, Bs


@pytest.mark.parametrize("k,g,e,tname", [(1, 0.001, 5.0 / 6.0, "mg_golden"), (3, 0.001, 5.0 / 6.0, "ecm_default"),
                                         (2, 0.01, 0.5, "mg_c5"), (1, 0.2, 0.9, "mg_c5"),
                                         (6, 0.001, 5.0 / 6.0, "mg_golden")])
def test_random_batch_vs_oracle(k, g, e, tname, gpu_ctx, tables):
    rng = np.random.RandomState(1000 + k)
    ancs, dess, As, Bs = _random_batch(rng, 96, k, 90)
    g, e = np.float32(g), np.float32(e)
    T = tables[tname]
    gpu_ctx.set_model(T, g, e, k)
    rows_a, rows_b, score, status = gpu_ctx.viterbi_batch(PackedPairs(As, Bs, ancs, dess))
    assert (status == 0).all()
    for p in range(len(ancs)):
        oa, ob, osc = oracle.viterbi(ancs[p], dess[p], T, g, e, k, enc=(As[p], Bs[p]))
        assert (rows_a[p], rows_b[p]) == (oa, ob), p
        assert util.f32_bits(score[p]) == util.f32_bits(osc), p


@pytest.mark.parametrize("k", [1, 2, 3])
def test_direction_bytes_vs_oracle(k, gpu_ctx, tables):
    """Cell-level parity: every decision byte the fill emits equals what the reference's traceback
    expressions give on the reference's matrices (align_pair.cc:275-296)."""
    rng = np.random.RandomState(50 + k)
    T = tables["mg_golden"]
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, k)
    ancs, dess, As, Bs = _random_batch(rng, 6, k, 70)
    for a, b in zip(As, Bs):
        M, D, I = oracle.fill(0, a, b, T, k=k)
        want = oracle.directions(M, D, I, len(a), len(b), k=k)[k:, k:]
        got, score = gpu_ctx.directions(a, b)
        want = want.copy()
        got = got.copy()
        want[-1, -1] = got[-1, -1] = 0   # terminal cell holds the adjusted scores upstream
        assert np.array_equal(got, want)
        assert util.f32_bits(score) == util.f32_bits(max(M[-1, -1], D[-1, -1], I[-1, -1]))


def test_edge_cases(gpu_ctx, tables):
    """Empty and ragged inputs, per-pair errors (the reference would be UB / throw on these)."""
    T = tables["mg_golden"]
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
    pairs = [("", ""), ("AAA", ""), ("", "ACG"), ("AAA", "A"), ("AAACCC", "AAACCCGGGTTTAAACCCGGGTTT"),
             ("CTCTGGATAGTG", "CTATAGTG")]
    As, Bs = zip(*[oracle.encode_pair(x, y) for x, y in pairs])
    ancs, dess = zip(*pairs)
    rows_a, rows_b, score, status = gpu_ctx.viterbi_batch(PackedPairs(list(As), list(Bs), list(ancs), list(dess)))
    assert (status == 0).all()
    for p, (x, y) in enumerate(pairs):
        oa, ob, osc = oracle.viterbi(x, y, T)
        assert (rows_a[p], rows_b[p]) == (oa, ob)
        assert util.f32_bits(score[p]) == util.f32_bits(osc)
    # bad symbols and bad lengths are reported per pair, the rest of the batch is unaffected
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 3)
    As = [np.array([0, 1, 2], np.uint8), np.array([0, 1, 200], np.uint8), np.array([0, 1, 2], np.uint8),
          np.array([0, 1, 2], np.uint8)]
    Bs = [np.array([0, 1, 2], np.uint8), np.array([0, 1, 2], np.uint8), np.array([0, 15, 2], np.uint8),
          np.array([0, 1], np.uint8)]
    ancs, dess = ["AAA"] * 4, ["ACG", "ACG", "A-G", "AC"]
    rows_a, rows_b, score, status = gpu_ctx.viterbi_batch(PackedPairs(As, Bs, ancs, dess))
    assert list(status) == [0, -4, -4, -5]
    assert (rows_a[0], rows_b[0]) == oracle.viterbi("AAA", "ACG", T, k=3)[:2]
    assert rows_a[1] == "" and rows_a[3] == ""
    # zero pairs
    r = gpu_ctx.viterbi_batch(PackedPairs([], [], [], []))
    assert r[0] == [] and len(r[2]) == 0


def test_multi_chunk_equals_single(tables):
    """A direction-buffer budget smaller than the batch forces several fill/traceback rounds."""
    import coati_b200
    rng = np.random.RandomState(5)
    ancs, dess, As, Bs = _random_batch(rng, 64, 1, 120)
    T = tables["mg_golden"]
    os.environ["COATI_GPU_DIR_BUDGET_MB"] = "1"
    try:
        small = coati_b200.Context(0)
    finally:
        del os.environ["COATI_GPU_DIR_BUDGET_MB"]
    big = coati_b200.Context(0)
    outs = []
    for ctx in (small, big):
        ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
        pack = PackedPairs(As, Bs, ancs, dess)
        bt = ctx.batch(pack.a_off, pack.b_off)
        outs.append((ctx.viterbi_batch(pack), bt.stats()))
        bt.destroy()
        ctx.close()
    assert outs[0][1]["chunks"] > 1 and outs[1][1]["chunks"] == 1
    assert outs[0][0][0] == outs[1][0][0] and outs[0][0][1] == outs[1][0][1]
    assert np.array_equal(outs[0][0][2].view(np.uint32), outs[1][0][2].view(np.uint32))


def _check_alignment_properties(anc, des, ra, rb, sc, T, k):
    assert len(ra) == len(rb)
    assert ra.replace("-", "") == anc and rb.replace("-", "") == des
    assert not any(x == "-" and y == "-" for x, y in zip(ra, rb))
    # independent re-scoring of the returned rows (alignment_score semantics); different
    # association of the same float32 terms, so tolerance not bit-equality
    rescored = oracle.alignment_score(ra, rb, T, k=k)
    assert rescored == pytest.approx(float(sc), rel=2e-4, abs=1e-3)


@pytest.mark.parametrize("name", ["benchmark_8k", "benchmark_16k"])
def test_large_real_pairs_properties(name, gpu_ctx, tables):
    """Sizes beyond the committed goldens: round-trip + re-scoring properties, and (16k skipped on
    the oracle for time) the oracle at 8k."""
    (_, anc), (_, des) = util.load_fasta(name)
    anc, s0 = oracle.trim_end_stop(anc)
    des, s1 = oracle.trim_end_stop(des)
    T = tables["mg_golden"]
    a, b = oracle.encode_pair(anc, des)
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
    ra, rb, sc = gpu_ctx.viterbi(a, b, anc, des)
    _check_alignment_properties(anc, des, ra, rb, sc, T, 1)
    if name == "benchmark_8k":
        oa, ob, osc = oracle.viterbi(anc, des, T, enc=(a, b))
        assert (ra, rb) == (oa, ob) and util.f32_bits(sc) == util.f32_bits(osc)


def test_raw_sample_files_are_rejected_like_the_reference():
    """sampledata/example-10k..160k carry in-frame ancestor stop codons: the reference throws
    "Early stop codon in ancestor/reference." (utils.cc:511-514).  Encoding is host-side here."""
    (_, anc), (_, des) = util.load_fasta("example-10k")
    with pytest.raises(ValueError, match="Early stop codon"):
        oracle.encode_pair(anc, des)


def test_wavefront_kernel_equals_inter_pair_kernel(tables):
    """Long pairs run as an intra-pair wavefront (bands of one lattice spread over all SMs); the
    result must be identical to the one-warp-per-pair kernel and to the oracle."""
    import coati_b200
    T = tables["mg_golden"]
    (_, anc), (_, des) = util.load_fasta("benchmark_4k")
    anc, _ = oracle.trim_end_stop(anc)
    des, _ = oracle.trim_end_stop(des)
    a, b = oracle.encode_pair(anc, des)
    want = oracle.viterbi(anc, des, T, enc=(a, b))
    outs = []
    # wavefront fill or not; run-at-a-time warp traceback (default) or the column-at-a-time walk
    for no_wave, tb_serial in (("0", "0"), ("1", "0"), ("0", "1"), ("1", "1")):
        os.environ["COATI_GPU_NO_WAVE"] = no_wave
        os.environ["COATI_GPU_TB_SERIAL"] = tb_serial
        try:
            ctx = coati_b200.Context(0)
        finally:
            del os.environ["COATI_GPU_NO_WAVE"]
            del os.environ["COATI_GPU_TB_SERIAL"]
        ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
        outs.append(ctx.viterbi(a, b, anc, des))
        # ragged mini-batch: a long pair next to short ones
        rng = np.random.RandomState(11)
        ancs, dess, As, Bs = _random_batch(rng, 5, 1, 60)
        ancs.append(anc), dess.append(des), As.append(a), Bs.append(b)
        rows_a, rows_b, score, status = ctx.viterbi_batch(PackedPairs(As, Bs, ancs, dess))
        assert (status == 0).all()
        assert (rows_a[-1], rows_b[-1]) == want[:2] and util.f32_bits(score[-1]) == util.f32_bits(want[2])
        for p in range(5):
            o = oracle.viterbi(ancs[p], dess[p], T, enc=(As[p], Bs[p]))
            assert (rows_a[p], rows_b[p]) == o[:2] and util.f32_bits(score[p]) == util.f32_bits(o[2])
        ctx.close()
    for got in outs:
        assert got[:2] == want[:2] and util.f32_bits(got[2]) == util.f32_bits(want[2])


def test_alignpair_batch_raw_sequences(gpu_ctx, tables):
    """coati_gpu_alignpair_batch = marg_alignment per pair (length checks before trimming, end-stop
    trim/restore with its score penalty, encoding on the device) against the oracle pipeline."""
    rng = np.random.RandomState(77)
    T = tables["mg_golden"]
    g, e = oracle.DEFAULT_G, oracle.DEFAULT_E
    for k in (1, 3):
        gpu_ctx.set_model(T, g, e, k)
        ancs, dess = [], []
        stops = ["TAA", "TAG", "TGA", "taa", "UAG", ""]
        for i in range(48):
            anc, des = util.random_pair(rng, int(rng.randint(1, 70)), k=k, ambiguous=i % 6 == 0)
            anc, _ = oracle.trim_end_stop(anc)
            des, _ = oracle.trim_end_stop(des)
            ancs.append(anc + stops[i % 6])
            dess.append(des + stops[(i // 2) % 6])
        # error cases, each reported per pair without disturbing the others
        ancs += ["AAACCNGGG", "AAATAAGGG", "AAATAGCCNGGG", "AAACCC", "AAAC", "AAACCC"]
        dess += ["AAACCC", "AAACCC", "AAACCC", "AA-CCC", "AAA", "AAACC" if k == 3 else "AAAC?C"]
        want_status = [0] * 48 + [-6, -7, -7, -4, -5, -5 if k == 3 else -4]
        rows_a, rows_b, score, status = gpu_ctx.alignpair_batch(ancs, dess)
        assert list(status) == want_status
        for p in range(48):
            at, s0 = oracle.trim_end_stop(ancs[p])
            dt, s1 = oracle.trim_end_stop(dess[p])
            oa, ob, osc = oracle.viterbi(at, dt, T, g, e, k)
            oa, ob, osc = oracle.restore_end_stops(oa, ob, osc, (s0, s1), g, e)
            assert (rows_a[p], rows_b[p]) == (oa, ob), p
            assert util.f32_bits(score[p]) == util.f32_bits(osc), p
        assert all(rows_a[p] == "" for p in range(48, 54))
    # the reference's own driver cases (align_marginal.cc:149-240)
    gpu_ctx.set_model(T, g, e, 1)
    ra, rb, _, st = gpu_ctx.alignpair_batch(["CTCTGGATAGTG", "GCGACTGTT", "ACGTTAAGGGGT"],
                                            ["CTATAGTG", "GCGATTGCTGTT", "ACGAAT"])
    assert list(st) == [0, 0, 0]
    assert list(zip(ra, rb)) == [("CTCTGGATAGTG", "CT----ATAGTG"), ("GCGA---CTGTT", "GCGATTGCTGTT"),
                                 ("ACGTTAAGGGGT", "ACG--AA----T")]


@pytest.mark.parametrize("k,force_generic", [(1, False), (3, False), (1, True)])
def test_per_pair_models_leaf_batch(k, force_generic, tables):
    """The msa leaf batch (align_msa.cc:285-318): each pair aligned with its own substitution table."""
    import coati_b200
    if force_generic:
        os.environ["COATI_GPU_FORCE_GENERIC"] = "1"
    try:
        ctx = coati_b200.Context(0)
    finally:
        os.environ.pop("COATI_GPU_FORCE_GENERIC", None)
    names = ["mg_golden", "ecm_default", "mg_c5"]
    T = np.stack([tables[n] for n in names])
    ctx.set_models(T, oracle.DEFAULT_G, oracle.DEFAULT_E, k)
    rng = np.random.RandomState(31 + k)
    ancs, dess, As, Bs = _random_batch(rng, 45, k, 80)
    model = np.arange(45) % 3
    rows_a, rows_b, score, status = ctx.viterbi_batch(PackedPairs(As, Bs, ancs, dess), model_idx=model)
    assert (status == 0).all()
    for p in range(45):
        oa, ob, osc = oracle.viterbi(ancs[p], dess[p], tables[names[model[p]]], k=k, enc=(As[p], Bs[p]))
        assert (rows_a[p], rows_b[p]) == (oa, ob), p
        assert util.f32_bits(score[p]) == util.f32_bits(osc), p
    with pytest.raises(coati_b200.CoatiGpuError):
        ctx.viterbi_batch(PackedPairs(As[:2], Bs[:2], ancs[:2], dess[:2]), model_idx=[0, 3])
    ctx.close()


@pytest.mark.parametrize("nsub", [2, 3, 7])
def test_pipelined_lanes_equal_single_call(nsub, tables):
    """The sub-batch pipeline (three lanes, fills on their own low-priority streams, event fences) gives the
    rows, scores and per-pair status of the plain call: forced on a small batch with COATI_GPU_NSUB, for the
    encoded, the per-pair-model and the raw-sequence entry points, error pairs included."""
    import coati_b200
    rng = np.random.RandomState(100 + nsub)
    names = ["mg_golden", "ecm_default"]
    T = np.stack([tables[n] for n in names])
    ancs, dess, As, Bs = _random_batch(rng, 151, 1, 90)
    model = rng.randint(0, 2, size=151)
    raw_a = list(ancs) + ["AAACCNGGG", "AAATAAGGG", "AAAC"]
    raw_d = list(dess) + ["AAACCC", "AAACCC", "AAA"]
    for i in range(0, 151, 5):  # a few end stops, so trimming / restoring happens inside sub-batches
        raw_a[i] += "TAA"
    perm = rng.permutation(len(raw_a))
    raw_a, raw_d = [raw_a[i] for i in perm], [raw_d[i] for i in perm]
    ctx = coati_b200.Context(0)
    ctx.set_models(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
    pack = PackedPairs(As, Bs, ancs, dess)
    outs = []
    for n in (1, nsub):
        os.environ["COATI_GPU_NSUB"] = str(n)
        try:
            outs.append((ctx.viterbi_batch(pack), ctx.viterbi_batch(pack, model_idx=model),
                         ctx.alignpair_batch(raw_a, raw_d)))
        finally:
            del os.environ["COATI_GPU_NSUB"]
    for one, piped in zip(outs[0], outs[1]):
        assert one[0] == piped[0] and one[1] == piped[1]
        assert np.array_equal(one[2].view(np.uint32), piped[2].view(np.uint32))
        assert np.array_equal(one[3], piped[3])
    # and the plain call is right (oracle, per-pair table)
    rows_a, rows_b, score, status = outs[1][1]
    for p in range(0, 151, 7):
        oa, ob, osc = oracle.viterbi(ancs[p], dess[p], tables[names[model[p]]], k=1, enc=(As[p], Bs[p]))
        assert (rows_a[p], rows_b[p]) == (oa, ob) and util.f32_bits(score[p]) == util.f32_bits(osc), p
    assert sorted(outs[1][2][3])[:3] == [-7, -6, -5]
    ctx.close()


@pytest.mark.parametrize("workload,k,tname,n", [(5, 1, "mg_c5", 200000), (4, 3, "ecm_default", 20000)])
def test_workload_scale_properties(workload, k, tname, n, tables):
    """BASELINE configs[4] / configs[3] shapes at bench scale (a fifth of the pair count by default,
    COATI_TEST_PAIRS overrides): size-independent properties of EVERY alignment (tests/util.check_batch_properties:
    rows strip to the inputs, no gap-gap column, lengths, terminators), oracle equality on a sample and on the
    largest lattices, and identical results from the three entry points (staged batch, pipelined CSR call,
    raw-sequence call)."""
    import coati_b200
    from coati_b200 import capi
    from synth import synth_pairs
    n = int(os.environ.get("COATI_TEST_PAIRS", n))
    g, e = np.float32(0.001), np.float32(1.0) - np.float32(1.0) / np.float32(6.0)
    T = tables[tname]
    w = synth_pairs(n, workload, 42)
    total = int(w["a_off"][-1] + w["b_off"][-1]) + n

    def buffers():
        return (np.zeros(total + 1, np.uint8), np.zeros(total + 1, np.uint8), np.zeros(n, np.uint64),
                np.zeros(n, np.float32), np.zeros(n, np.int32))

    ctx = coati_b200.Context(0)
    ctx.set_model(T, g, e, k)
    # (1) staged batch API (what bench.py times as `value`)
    oa1, ob1, ln1, sc1, st1 = buffers()
    bt = ctx.batch(w["a_off"], w["b_off"])
    bt.upload(w["a_all"], w["b_all"], w["anc_all"], w["des_all"])
    bt.run()
    bt.download(oa1, ob1, ln1, sc1, st1)
    bt.destroy()
    assert util.check_batch_properties(w, oa1, ob1, ln1, sc1, st1, T, k, g, e, oracle) >= 4
    # (2) the pipelined public call on the same encoded input
    oa2, ob2, ln2, sc2, st2 = buffers()
    ctx._check(ctx.lib.coati_gpu_viterbi_batch(
        ctx.h, n, w["a_all"].ctypes.data, w["a_off"].ctypes.data_as(capi._u64p), w["b_all"].ctypes.data,
        w["b_off"].ctypes.data_as(capi._u64p), w["anc_all"].ctypes.data, w["des_all"].ctypes.data,
        oa2.ctypes.data, ob2.ctypes.data, ln2.ctypes.data_as(capi._u64p), sc2.ctypes.data_as(capi._fp),
        st2.ctypes.data_as(capi._i32p)))
    util.check_batch_properties(w, oa2, ob2, ln2, sc2, st2, T, k, g, e, oracle, exact=False)
    # (3) the raw-sequence call (what bench.py times as `e2e`); the generator's ancestors are stop-free, a
    # mutated descendant may end in a stop codon, which this entry point trims and restores
    oa3, ob3, ln3, sc3, st3 = buffers()
    ctx._check(ctx.lib.coati_gpu_alignpair_batch(
        ctx.h, n, w["anc_all"].ctypes.data, w["a_off"].ctypes.data_as(capi._u64p), w["des_all"].ctypes.data,
        w["b_off"].ctypes.data_as(capi._u64p), oa3.ctypes.data, ob3.ctypes.data, ln3.ctypes.data_as(capi._u64p),
        sc3.ctypes.data_as(capi._fp), st3.ctypes.data_as(capi._i32p)))
    util.check_batch_properties(w, oa3, ob3, ln3, sc3, st3, T, k, g, e, oracle, exact=False)
    util.compare_entry_points(w, (oa1, ob1, ln1, sc1, st1), (oa2, ob2, ln2, sc2, st2), (oa3, ob3, ln3, sc3, st3))
    ctx.close()
