"""CPU: the C restatement (oracle/coati_oracle.c) against the reference's own known answers and
against outputs of the unmodified reference recorded in tests/golden/ (tools/gen_golden.py)."""
import numpy as np
import pytest

import oracle
from oracle import table as otable
from tests import util

VIT = util.load_json("viterbi_golden.json")
SMP = util.load_json("sample_golden.json")


def _inputs(c):
    if "anc" in c:
        return c["anc"], c["des"]
    (_, anc), (_, des) = util.load_fasta(c["file"])
    if c.get("sanitised"):
        anc = util.sanitise_ancestor(anc)
    return anc, des


@pytest.mark.parametrize("c", [c for c in VIT if c["len"] < 9000], ids=lambda c: c["name"])
def test_viterbi_golden(c, tables):
    anc, des = _inputs(c)
    g, e = util.bits_to_f32(c["g_bits"]), util.bits_to_f32(c["e_bits"])
    anc_t, s0 = oracle.trim_end_stop(anc)
    des_t, s1 = oracle.trim_end_stop(des)
    a, b, sc = oracle.viterbi(anc_t, des_t, tables[c["table"]], g=g, e=e, k=c["k"])
    assert util.f32_bits(sc) == c["score_bits"]
    a, b, sc = oracle.restore_end_stops(a, b, sc, (s0, s1), g, e)
    assert util.f32_bits(sc) == c["final_score_bits"]
    assert len(a) == c["len"]
    if "aln_a" in c:
        assert (a, b) == (c["aln_a"], c["aln_b"])
    import hashlib
    h = hashlib.sha256()
    for part in (a, b):
        h.update(part.encode())
        h.update(b"\0")
    assert h.hexdigest() == c["sha256"]


def test_reference_known_answers(tables):
    """align_marginal.cc:149-240 -- the strings the reference's own test-suite expects."""
    T = tables["mg_golden"]
    assert oracle.viterbi("CTCTGGATAGTG", "CTATAGTG", T)[:2] == ("CTCTGGATAGTG", "CT----ATAGTG")
    assert oracle.viterbi("GCGACTGTT", "GCGATTGCTGTT", T)[:2] == ("GCGA---CTGTT", "GCGATTGCTGTT")
    assert oracle.viterbi("GCGACTGTT", "GCGATTGCTGTT", tables["ecm_default"])[:2] == \
        ("GCGA---CTGTT", "GCGATTGCTGTT")
    assert oracle.viterbi("ACGTTAAGGGGT", "ACGAAT", T)[:2] == ("ACGTTAAGGGGT", "ACG--AA----T")
    assert oracle.viterbi("ACGTTAAGGGGT", "ACGAAT", T, k=3)[:2] == ("ACGTTAAGGGGT", "AC------GAAT")
    assert oracle.viterbi("CTCTGGATAGTG", "CTATAGTR", T)[:2] == ("CTCTGGATAGTG", "CT----ATAGTR")
    assert oracle.viterbi("CTCTGGATAGTG", "CTATAGTR", tables["mg_golden_best"])[:2] == \
        ("CTCTGGATAGTG", "CT----ATAGTR")


SCORES = [  # align_marginal.cc:490-509
    ("CTCTGGATAGTG", "CT----ATAGTG", 1.50914), ("CTCT--AT", "CTCTGGAT", -0.83906),
    ("ACTCT-A", "ACTCTG-", -10.52864), ("ATGCTTTAC", "ATGCT-TAC", 2.13593),
    ("ATGCTT---", "ATGCTTTGA", 0.70607), ("A-CTAAC", "ACCTAAG", -8.2786),
    ("ACT---", "ACTCTG", -5.04197), ("ACTCTA", "ACT---", -5.04197), ("ACT----", "ACT-CTG", -5.04197),
    ("AAAAAA---AAA", "AAA---AAAAAA", -11.09557), ("AAA---AAAAAA", "AAAAAA---AAA", -11.09557),
    ("AAA-A-A-AAAA", "AAAA-A-A-AAA", -11.09557), ("---AAAAAA", "AAAAAAAAA", -2.03242),
    ("AAAAAA---", "AAAAAAAAA", -2.03242), ("AAAAAAAAA", "---AAAAAA", -2.03242),
    ("AAAAAAAAA", "AAAAAA---", -2.03242), ("ACTCTA", "ACTC--", -3.18537),
    ("ACTCTA-", "ACTCTAG", -10.45777), ("ACTCTA--", "ACTCT-AG", -10.45777)]


@pytest.mark.parametrize("a,b,exp", SCORES)
def test_alignment_score_goldens(a, b, exp, tables):
    got = oracle.alignment_score(a, b, tables["mg_golden"])
    assert got == pytest.approx(exp, rel=1e-5, abs=1e-5)   # doctest::Approx default epsilon


def test_alignment_score_failures(tables):
    with pytest.raises(ValueError):
        oracle.alignment_score("CTCTGGATAGTG", "CTATAGTG", tables["mg_golden"])   # :293 unequal length
    with pytest.raises(ValueError):
        oracle.alignment_score("ATAC", "ATA-", tables["mg_golden"])               # :523 La % 3


@pytest.mark.parametrize("s", SMP, ids=lambda s: s["name"])
def test_sample_golden(s, tables):
    import hashlib
    anc, _ = oracle.trim_end_stop(s["anc"])
    des, _ = oracle.trim_end_stop(s["des"])
    st = oracle.seed_state(s["seeds"])
    assert [int(st[0]), int(st[1])] == s["state0"]
    out, sc, st2, _ = oracle.sample(anc, des, tables[s["table"]], st, s["n"], k=s["k"])
    assert [int(st2[0]), int(st2[1])] == s["state1"]
    h = hashlib.sha256()
    for (a, b), x in zip(out, sc):
        h.update((a + "\0" + b + "\0" + util.f32_bits(x)).encode())
    assert h.hexdigest() == s["sha256"]
    for (a, b), x, f in zip(out, sc, s["first"]):
        assert [a, b, util.f32_bits(x)] == f


def test_sample_known_answers(tables):
    """align_marginal.cc:653-671: seed "42" sample strings (scores are pinned to Eigen's table)."""
    T = tables["mg_golden"]
    st = oracle.seed_state(["42"])
    out, sc, _, _ = oracle.sample("CCCCCC", "CCCCCCCC", T, st, 3)
    assert [o[0] for o in out] == ["CC--CCCC", "CCCCCC--", "CCCC--CC"]
    assert all(o[1] == "CCCCCCCC" for o in out)
    np.testing.assert_allclose(sc, [-1.9466571807861328, -1.9466569423675537, -1.9466572999954224],
                               rtol=1e-6)
    out, sc, _, _ = oracle.sample("CCCCCC", "CCCC", T, oracle.seed_state(["42"]), 1)
    assert out[0] == ("CCCCCC", "--CCCC")
    np.testing.assert_allclose(sc, [-1.6172490119934082], rtol=1e-6)


def test_rng_golden():
    import ctypes as C
    for r in util.load_json("rng_golden.json"):
        st = oracle.seed_state(r["seeds"])
        assert [int(st[0]), int(st[1])] == r["state"]
        s = (C.c_uint64 * 2)(int(st[0]), int(st[1]))
        assert [int(oracle.lib.orc_rng_bits(s)) for _ in range(8)] == r["bits"]
        assert [util.f32_bits(oracle.lib.orc_rng_f24(s)) for _ in range(4)] == r["f24"]


def test_encoding_goldens():
    """utils.cc:532-586 marginal_seq_encoding test; :971-1029 trim_end_stops; :1168-1227 codon maps."""
    a, b = oracle.encode_pair("AAAGGGTTTCCCACTAGA", "ACGTRYMKSWBDHVN-")
    assert list(a) == [0, 1, 2, 126, 127, 128, 180, 181, 182, 63, 64, 65, 21, 22, 23, 24, 25, 26]
    assert list(b) == list(range(16))
    a2, _ = oracle.encode_pair("aaagggtttcccacuaga", "acgu")
    assert list(a2) == list(a)
    with pytest.raises(ValueError, match="Ambiguous"):
        oracle.encode_pair("AAAGGGTTTCCCACTAGR", "A")
    with pytest.raises(ValueError, match="stop"):
        oracle.encode_pair("AAATAA", "A")
    for c64, c61 in [(0, 0), (20, 20), (47, 47), (49, 48), (51, 49), (52, 50), (53, 51), (57, 54),
                     (60, 57), (63, 60)]:
        assert oracle.lib.orc_cod64_to_61(c64) == c61
        assert oracle.lib.orc_cod61_to_64(c61) == c64
    assert oracle.lib.orc_cod64_to_61(48) == -2 and oracle.lib.orc_cod64_to_61(64) == -1
    assert oracle.trim_end_stop("AGATTTTGA") == ("AGATTT", "TGA")
    assert oracle.trim_end_stop("AGATTTtag") == ("AGATTT", "tag")
    assert oracle.trim_end_stop("AGATTTUAA") == ("AGATTT", "UAA")
    assert oracle.trim_end_stop("AGATTT") == ("AGATTT", "")
    assert oracle.trim_end_stop("TA") == ("TA", "")


def test_restore_end_stops():
    """utils.cc:1067-1094."""
    g, e = oracle.DEFAULT_G, oracle.DEFAULT_E
    assert oracle.restore_end_stops("AAA", "AAA", 1.0, ("", "")) == ("AAA", "AAA", 1.0)
    assert oracle.restore_end_stops("AAA", "AAA", 1.0, ("TAA", "TGA")) == ("AAATAA", "AAATGA", 1.0)
    a, b, s = oracle.restore_end_stops("AAA", "AAA", 1.0, ("", "TAG"))
    assert (a, b) == ("AAA---", "AAATAG")
    assert s == pytest.approx(1.0 + np.log(np.float32(g) * e * e), rel=1e-6)


def test_mg94_p_golden():
    """mutation_coati.cc:129-145: builder vs the golden mg94P, doctest::Approx tolerance."""
    P = np.load(util.GOLDEN + "/mg94_p_default.npy")
    np.testing.assert_allclose(otable.mg94_p(), P, rtol=1e-5, atol=1e-9)
    with pytest.raises(ValueError):
        otable.mg94_p(br_len=0)


def test_marginal_p_normalised(tables):
    """mutation_coati.cc:206-222: sum_nuc exp(p) * pi == 1 per (codon, phase)."""
    pi = np.array(otable.DEFAULT_PI, dtype=np.float64)
    for name in ("mg_golden", "mg_default", "ecm_default"):
        s = (np.exp(tables[name][:, :4].astype(np.float64)) * pi).sum(axis=1)
        np.testing.assert_allclose(s, 1.0, rtol=2e-5)


def test_gtr_q_golden():
    """mutation_coati.cc:358-386."""
    q = otable.gtr_q((0.308, 0.185, 0.199, 0.308), (0.009489730, 0.039164824, 0.004318182,
                                                    0.015438693, 0.038734091, 0.008550000))
    exp = [[-0.010879400, 0.001755600, 0.007793800, 0.00133],
           [0.002922837, -0.017925237, 0.003072300, 0.0119301],
           [0.012062766, 0.002856158, -0.017552324, 0.0026334],
           [0.001330000, 0.007165807, 0.001701450, -0.010197257]]
    np.testing.assert_allclose(q, exp, rtol=2e-5, atol=1e-8)
    with pytest.raises(ValueError):
        otable.gtr_q((0.25,) * 4, (-0.1, 0, 0, 0, 0, 0))


def test_batch_property_checker_on_oracle_output(tables):
    """The vectorised property checker the GPU tests run at workload scale (tests/util.check_batch_properties),
    exercised here on arenas filled by the oracle -- and on deliberately corrupted arenas, which it must reject."""
    from synth import synth_pairs
    g, e = np.float32(0.001), np.float32(1.0) - np.float32(1.0) / np.float32(6.0)
    for workload, k, tname in ((5, 1, "mg_c5"), (4, 3, "ecm_default")):
        w = synth_pairs(24 if workload == 5 else 6, workload, 42, threads=1)
        T = tables[tname]
        a_off, b_off = w["a_off"].astype(np.int64), w["b_off"].astype(np.int64)
        n = len(a_off) - 1
        total = int(a_off[-1] + b_off[-1]) + n
        out_a, out_b = np.zeros(total + 1, np.uint8), np.zeros(total + 1, np.uint8)
        out_len, score, status = np.zeros(n, np.uint64), np.zeros(n, np.float32), np.zeros(n, np.int32)
        for p in range(n):
            anc = w["anc_all"][a_off[p]:a_off[p + 1]].tobytes().decode()
            des = w["des_all"][b_off[p]:b_off[p + 1]].tobytes().decode()
            ra, rb, sc = oracle.viterbi(anc, des, T, g, e, k,
                                        enc=(w["a_all"][a_off[p]:a_off[p + 1]], w["b_all"][b_off[p]:b_off[p + 1]]))
            o = int(a_off[p] + b_off[p]) + p
            out_a[o:o + len(ra)] = np.frombuffer(ra.encode(), np.uint8)
            out_b[o:o + len(rb)] = np.frombuffer(rb.encode(), np.uint8)
            out_len[p], score[p] = len(ra), sc
        assert util.check_batch_properties(w, out_a, out_b, out_len, score, status, T, k, g, e, oracle, sample=8) >= 4
        # corruptions: a flipped symbol, a wrong score, a failed pair, a missing terminator
        o3 = int(a_off[3] + b_off[3]) + 3
        for mutate in ("symbol", "score", "status", "nul"):
            ca, cs, st = out_a.copy(), score.copy(), status.copy()
            if mutate == "symbol":
                j = o3 + int(np.flatnonzero(ca[o3:o3 + int(out_len[3])] != ord("-"))[0])
                ca[j] = ord("A") if ca[j] != ord("A") else ord("C")
            elif mutate == "score":
                cs[:] = cs + np.float32(1.0)
            elif mutate == "status":
                st[5] = -5
            else:
                ca[o3 + int(out_len[3])] = ord("A")
            with pytest.raises(AssertionError):
                util.check_batch_properties(w, ca, out_b, out_len, cs, st, T, k, g, e, oracle, sample=n * 4)


def test_entry_point_comparison_on_oracle_output(tables):
    """tests/util.compare_entry_points (used by the GPU workload-scale test) on arenas made by the oracle the
    way the three entry points fill them: scratch bytes after the terminators differ, and the raw-sequence
    call trims and restores end stops (two descendants are made to end in a stop codon)."""
    from synth import synth_pairs
    g, e = np.float32(0.001), np.float32(1.0) - np.float32(1.0) / np.float32(6.0)
    T = tables["mg_c5"]
    w = synth_pairs(20, 5, 42, threads=1)
    a_off, b_off = w["a_off"].astype(np.int64), w["b_off"].astype(np.int64)
    for p in (2, 7):  # ...TAA at the end of two descendants (IUPAC codes T = 3, A = 0)
        w["des_all"][b_off[p + 1] - 3:b_off[p + 1]] = np.frombuffer(b"TAA", np.uint8)
        w["b_all"][b_off[p + 1] - 3:b_off[p + 1]] = (3, 0, 0)
    n = 20
    total = int(a_off[-1] + b_off[-1]) + n
    rng = np.random.RandomState(3)

    def fill(raw):
        oa = rng.randint(1, 255, total + 1).astype(np.uint8)  # scratch everywhere, rows written over it
        ob = rng.randint(1, 255, total + 1).astype(np.uint8)
        ln, sc, st = np.zeros(n, np.uint64), np.zeros(n, np.float32), np.zeros(n, np.int32)
        for p in range(n):
            anc = w["anc_all"][a_off[p]:a_off[p + 1]].tobytes().decode()
            des = w["des_all"][b_off[p]:b_off[p + 1]].tobytes().decode()
            if raw:
                at, s0 = oracle.trim_end_stop(anc)
                dt, s1 = oracle.trim_end_stop(des)
                ra, rb, x = oracle.viterbi(at, dt, T, g, e, 1)
                ra, rb, x = oracle.restore_end_stops(ra, rb, x, (s0, s1), g, e)
            else:
                ra, rb, x = oracle.viterbi(anc, des, T, g, e, 1)
            o = int(a_off[p] + b_off[p]) + p
            oa[o:o + len(ra)] = np.frombuffer(ra.encode(), np.uint8)
            ob[o:o + len(rb)] = np.frombuffer(rb.encode(), np.uint8)
            oa[o + len(ra)] = ob[o + len(rb)] = 0
            ln[p], sc[p] = len(ra), x
        return oa, ob, ln, sc, st

    staged, piped, raw = fill(False), fill(False), fill(True)
    for r, exact in ((staged, True), (piped, False), (raw, False)):
        util.check_batch_properties(w, *r, T, 1, g, e, oracle, sample=6, exact=exact)
    assert util.compare_entry_points(w, staged, piped, raw) == 2
    bad = tuple(x.copy() for x in piped)
    bad[3][4] += np.float32(0.5)
    with pytest.raises(AssertionError):
        util.compare_entry_points(w, staged, bad, raw)
    bad = tuple(x.copy() for x in raw)
    o = int(a_off[9] + b_off[9]) + 9
    bad[1][o] = ord("-") if bad[1][o] != ord("-") else ord("A")
    with pytest.raises(AssertionError):
        util.compare_entry_points(w, staged, piped, bad)
