"""CPU: the C++ host layer above the C ABI (coati_b200/csrc/host) -- table builder, sequence prep,
seeding, re-scoring, output formatting -- against the oracle's independent numpy/C restatements and the
reference's known answers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from coati_b200 import capi
from coati_b200 import build as cbuild
from oracle import table as otable
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def lib():
    cbuild.build()
    return capi.load_library()


def test_mg94_p_vs_reference_golden(lib):
    """mutation_coati.cc:129-145: P against the reference's golden mg94P at its own tolerance."""
    P = np.zeros((61, 61), np.float32)
    pi = np.float32([0.308, 0.185, 0.199, 0.308])
    assert lib.coati_host_mg94_p(0.0133, 0.2, pi.ctypes.data_as(_fp), None, P.ctypes.data_as(_fp)) == 0
    np.testing.assert_allclose(P, np.load(util.GOLDEN + "/mg94_p_default.npy"), rtol=1e-5, atol=1e-9)
    assert lib.coati_host_mg94_p(0.0, 0.2, pi.ctypes.data_as(_fp), None, P.ctypes.data_as(_fp)) != 0  # br_len <= 0


@pytest.mark.parametrize("kw", [dict(), dict(model="mar-ecm"), dict(br_len=0.05, omega=0.5, pi=(0.25,) * 4),
                                dict(amb="BEST"), dict(msub="MAX"), dict(model="mar-ecm", br_len=0.4),
                                dict(br_len=2.5, omega=1.3)])
def test_marginal_table_vs_oracle_builder(kw, lib):
    """C++ float Pade expm + marginalisation vs the oracle's float64 scipy expm (table parity is pinned to
    the reference's 1e-5 tolerance only)."""
    got = capi.host_marginal_table(**kw)
    want = otable.build_table(**kw)
    np.testing.assert_allclose(got, want, rtol=3e-5, atol=3e-6)


def test_marginal_p_normalised(lib):
    """mutation_coati.cc:206-222."""
    T = capi.host_marginal_table()
    s = (np.exp(T[:, :4].astype(np.float64)) * np.array(otable.DEFAULT_PI)).sum(axis=1)
    np.testing.assert_allclose(s, 1.0, rtol=2e-5)


def test_gtr_q_golden(lib):
    """mutation_coati.cc:358-386."""
    pi = np.float32([0.308, 0.185, 0.199, 0.308])
    sg = np.float32([0.009489730, 0.039164824, 0.004318182, 0.015438693, 0.038734091, 0.008550000])
    q = np.zeros(16, np.float32)
    assert lib.coati_host_gtr_q(pi.ctypes.data_as(_fp), sg.ctypes.data_as(_fp), q.ctypes.data_as(_fp)) == 0
    exp = [[-0.010879400, 0.001755600, 0.007793800, 0.00133], [0.002922837, -0.017925237, 0.003072300, 0.0119301],
           [0.012062766, 0.002856158, -0.017552324, 0.0026334], [0.001330000, 0.007165807, 0.001701450, -0.010197257]]
    np.testing.assert_allclose(q.reshape(4, 4), exp, rtol=2e-5, atol=1e-8)
    sg[0] = -0.1
    assert lib.coati_host_gtr_q(pi.ctypes.data_as(_fp), sg.ctypes.data_as(_fp), q.ctypes.data_as(_fp)) != 0


def test_encoding_matches_oracle_and_reference_goldens(lib):
    """utils.cc:532-586."""
    a, b = capi.host_encode("AAAGGGTTTCCCACTAGA", "ACGTRYMKSWBDHVN-")
    assert list(a) == [0, 1, 2, 126, 127, 128, 180, 181, 182, 63, 64, 65, 21, 22, 23, 24, 25, 26]
    assert list(b) == list(range(16))
    rng = np.random.RandomState(4)
    for _ in range(30):
        anc, des = util.random_pair(rng, int(rng.randint(1, 40)), ambiguous=True)
        oa, ob = oracle.encode_pair(anc, des)
        ha, hb = capi.host_encode(anc, des)
        assert np.array_equal(oa, ha) and np.array_equal(ob, hb)
    with pytest.raises(capi.CoatiGpuError) as e:
        capi.host_encode("AAATAA", "A")
    assert e.value.code == -7 and "Early stop codon" in str(e.value)
    with pytest.raises(capi.CoatiGpuError) as e:
        capi.host_encode("AANAAA", "A")
    assert e.value.code == -6 and "Ambiguous" in str(e.value)


def test_seeding_matches_oracle(lib):
    for seeds in (["42"], [""], ["random42"], ["-5", "x"], ["2147483648"], ["1", "2", "3", "4", "5", "6", "7", "8", "9"]):
        st = (C.c_uint64 * 2)()
        arr = (C.c_char_p * len(seeds))(*[s.encode() for s in seeds])
        lib.coati_host_seed(arr, len(seeds), st)
        o = oracle.seed_state(seeds)
        assert [st[0], st[1]] == [int(o[0]), int(o[1])], seeds


def test_alignment_score_goldens(lib, tables):
    """align_marginal.cc:490-509 through the C++ host implementation."""
    from tests.test_oracle_golden import SCORES
    T = tables["mg_golden"]
    for a, b, exp in SCORES:
        sc = C.c_float(0)
        rc = lib.coati_host_alignment_score(a.encode(), b.encode(), T.ctypes.data_as(_fp), oracle.DEFAULT_G,
                                            oracle.DEFAULT_E, 1, C.byref(sc))
        assert rc == 0
        assert sc.value == pytest.approx(exp, rel=1e-5, abs=1e-5)
        assert util.f32_bits(sc.value) == util.f32_bits(oracle.alignment_score(a, b, T))
    sc = C.c_float(0)
    assert lib.coati_host_alignment_score(b"CTCTGGATAGTG", b"CTATAGTG", T.ctypes.data_as(_fp), oracle.DEFAULT_G,
                                          oracle.DEFAULT_E, 1, C.byref(sc)) != 0


def test_json_number_is_shortest_roundtrip_double(lib):
    """json.cc: nlohmann dumps the float score widened to double (align_marginal.cc:655-670 literals)."""
    buf = C.create_string_buffer(64)
    for v, s in ((np.float32(-1.9466571807861328), "-1.9466571807861328"), (0.0, "0.0"), (0.1, "0.10000000149011612"),
                 (np.float32(-1.6172490119934082), "-1.6172490119934082"), (3.0, "3.0")):
        lib.coati_host_json_number(C.c_float(float(v)), buf, 64)
        assert buf.value.decode() == s


def test_cli_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    fa = tmp_path / "x.fasta"
    fa.write_text(">1\nCTCTGGATAGTG\n>2\nCTATAGTG\n")
    r = subprocess.run([os.path.join(ROOT, "coati_b200", "bin", "coati-gpu"), "alignpair", str(fa)],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def _convert(lib, src, dst):
    lib.coati_host_convert.argtypes = [C.c_char_p, C.c_char_p]
    return lib.coati_host_convert(str(src).encode(), str(dst).encode())


def test_phylip_reader_and_writers_known_answers(lib, tmp_path):
    """read_phylip / write_phylip / write_fasta / write_json against the reference's own cases
    (phylip.cc:101-147 interleaved input with a 10-character name field; phylip.cc:230-275 output;
    json.cc:37-42 object layout) and a FASTA -> PHYLIP -> JSON -> FASTA round trip."""
    src = tmp_path / "in.phy"
    s0 = "CTCTGGATAG" * 10
    s1 = "CTATA" * 20
    src.write_text(" 2 100\nVeryLongNa" + s0[:60] + "\n2         " + s1[:60] + "\n\n" + s0[60:] + "\n" + s1[60:] + "\n\n")
    assert _convert(lib, src, tmp_path / "a.fasta") == 0
    assert (tmp_path / "a.fasta").read_text() == (">VeryLongNa\n" + s0[:60] + "\n" + s0[60:] + "\n>2\n" + s1[:60] + "\n"
                                                   + s1[60:] + "\n")
    assert _convert(lib, tmp_path / "a.fasta", tmp_path / "b.phy") == 0
    lines = (tmp_path / "b.phy").read_text().split("\n")
    assert lines[0] == "2 100"
    assert lines[1] == "VeryLongNa" + s0[:50] and lines[2] == "2         " + s1[:50] and lines[3] == ""
    assert lines[4] == s0[50:] and lines[5] == s1[50:]          # 60 per line after the first 50
    assert _convert(lib, tmp_path / "b.phy", tmp_path / "c.json") == 0
    assert (tmp_path / "c.json").read_text() == ('{\n  "alignment": {\n    "VeryLongNa": "%s",\n    "2": "%s"\n  },\n'
                                                 '  "score": 0.0\n}\n' % (s0, s1))
    assert _convert(lib, tmp_path / "c.json", tmp_path / "d.fa") == 0
    assert (tmp_path / "d.fa").read_text() == (tmp_path / "a.fasta").read_text()
    # short rows: the reference's write_phylip case
    (tmp_path / "s.fa").write_text(">tx_1\nCTCTGGATAGTG\n>taxa_2\nCT----ATAGTG\n")
    assert _convert(lib, tmp_path / "s.fa", tmp_path / "s.phy") == 0
    assert (tmp_path / "s.phy").read_text().split("\n")[:3] == ["2 12", "tx_1      CTCTGGATAGTG", "taxa_2    CT----ATAGTG"]
    # failures are exceptions upstream (io.cc:196-221): unknown extension, missing file
    assert _convert(lib, tmp_path / "s.fa", tmp_path / "s.xyz") == -2
    assert _convert(lib, tmp_path / "missing.fa", tmp_path / "o.fa") == -2


def test_parse_matrix_csv_self_consistency(lib, tmp_path):
    """--sub CSV (io.cc:48-88): expm(Q * t) of a "codon,codon,rate" file equals mg94_p built from the same
    rates -- the reference's own test (io.cc:91-131, Approx) with Q from the oracle's builder."""
    pi = np.array([0.308, 0.185, 0.199, 0.308], dtype=np.float32)
    Q, d = otable.mg94_q(0.2, pi)
    codons = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT" if a + b + c not in ("TAA", "TAG", "TGA")]
    t = np.float32(0.0133)
    path = tmp_path / "q.csv"
    with open(path, "w") as fh:
        fh.write("%.9g\n" % float(t / d))  # the file carries the normalised branch length
        for i in range(61):
            for j in range(61):
                fh.write("%s,%s,%.9g\n" % (codons[i], codons[j], float(Q[i, j])))
    lib.coati_host_parse_matrix_csv.argtypes = [C.c_char_p, _fp]
    P = np.zeros(61 * 61, dtype=np.float32)
    assert lib.coati_host_parse_matrix_csv(str(path).encode(), P.ctypes.data_as(_fp)) == 0
    want = np.zeros(61 * 61, dtype=np.float32)
    sigma = np.zeros(6, dtype=np.float32)
    assert lib.coati_host_mg94_p(C.c_float(0.0133), C.c_float(0.2), pi.ctypes.data_as(_fp), sigma.ctypes.data_as(_fp),
                                 want.ctypes.data_as(_fp)) == 0
    assert np.allclose(P, want, rtol=2e-5, atol=1e-7)
    # wrong line count / unreadable file -> invalid_argument upstream
    with open(path, "a") as fh:
        fh.write("AAA,AAA,0\n")
    assert lib.coati_host_parse_matrix_csv(str(path).encode(), P.ctypes.data_as(_fp)) == -2
    assert lib.coati_host_parse_matrix_csv(b"", P.ctypes.data_as(_fp)) == -2


def test_gtr_sigma_reaches_the_table_only_when_asked():
    """upstream parses -x/--sigma and then drops it (utils.cc:606).  Default here: the same table as
    without sigma (drop-in).  alignment_t::use_sigma (--gtr): the rates go through gtr_q into mg94_p
    (mutation_coati.cc:317-354); checked against the numpy restatement with scipy's expm (table parity is pinned to
    the reference's own 1e-5 class of tolerance, not to bits: Eigen is absent)."""
    from coati_b200 import capi
    from oracle import table as otable
    pi = (0.308, 0.185, 0.199, 0.308)
    sigma = (0.009489730, 0.039164824, 0.004318182, 0.015438693, 0.038734091, 0.008550000)   # mutation_coati.cc:360
    plain = capi.host_marginal_table("mar-mg", br_len=0.0133, omega=0.2, pi=pi)
    wired = capi.host_marginal_table_gtr(0.0133, 0.2, pi, sigma)
    want = otable.marginal_p(otable.mg94_p(0.0133, 0.2, pi, sigma), pi)
    assert np.allclose(wired, want, rtol=3e-5, atol=3e-6)
    assert not np.allclose(wired, plain, rtol=1e-3)           # the rates matter
    zero = capi.host_marginal_table_gtr(0.0133, 0.2, pi, (0,) * 6)
    assert np.array_equal(zero, plain)                         # all-zero sigma = Yang-94 rates, as mg94_p decides


def test_phylip_round_trip_and_layout(tmp_path):
    """write_phylip / read_phylip (phylip.cc:37-97, 194-217): names cut or padded to 10 characters, 50 columns in
    the first block, 60 in the following ones, a blank line after every block; the reader gives the rows back."""
    import ctypes as C
    import coati_b200
    lib = coati_b200.load_library()
    if not hasattr(lib, "coati_host_phylip_roundtrip"):
        import pytest
        pytest.skip("hook missing")
    lib.coati_host_phylip_roundtrip.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
    a = "ACGT" * 43 + "A-"          # 174 columns: 50 + 60 + 60 + 4
    b = "TG-A" * 43 + "CC"
    out = C.create_string_buffer(4096)
    assert lib.coati_host_phylip_roundtrip(b"a_very_long_name", a.encode(), b"s2", b.encode(), out, 4096) == 0
    text = out.value.decode()
    lines = text.split("\n")
    assert lines[0] == "2 174"
    assert lines[1] == "a_very_lon" + a[:50] and lines[2] == "s2        " + b[:50] and lines[3] == ""
    assert lines[4] == a[50:110] and lines[5] == b[50:110] and lines[6] == ""
    assert lines[10] == a[170:] and lines[11] == b[170:] and lines[12] == ""
