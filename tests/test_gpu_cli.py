"""GPU (-m gpu): the coati-gpu front end (C++ host layer + C ABI) on the reference's own driver tests
(align_marginal.cc:149-361 marg_alignment, :598-723 marg_sample): same inputs, same output files."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "coati_b200", "bin", "coati-gpu")


def run(*args, cwd=None):
    return subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, cwd=cwd)


def fasta(tmp_path, text, name="test-marg.fasta"):
    p = tmp_path / name
    p.write_text(text)
    return p


def test_alignpair_fasta_output(tmp_path):
    f = fasta(tmp_path, ">1\nCTCTGGATAGTG\n>2\nCTATAGTG\n")
    out = tmp_path / "out.fasta"
    r = run("alignpair", f, "-m", "mar-mg", "-o", out)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == ">1\nCTCTGGATAGTG\n>2\nCT----ATAGTG\n"


def test_alignpair_ref_swap_and_phylip(tmp_path):
    f = fasta(tmp_path, ">A\nGCGATTGCTGTT\n>B\nGCGACTGTT\n")
    out = tmp_path / "out.phy"
    r = run("alignpair", f, "-m", "mar-ecm", "-v", "-o", out)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == "2 12\nB         GCGA---CTGTT\nA         GCGATTGCTGTT\n\n"
    f = fasta(tmp_path, ">1\nCTATAGTG\n>2\nCTCTGGATAGTG\n")
    out = tmp_path / "o2.fasta"
    assert run("alignpair", f, "-r", "2", "-o", out).returncode == 0
    assert out.read_text() == ">2\nCTCTGGATAGTG\n>1\nCT----ATAGTG\n"


def test_alignpair_json_stdout_and_gap_len(tmp_path):
    f = fasta(tmp_path, ">1\nACGTTAAGGGGT\n>2\nACGAAT\n")
    r = run("alignpair", f)
    assert r.returncode == 0, r.stderr
    j = json.loads(r.stdout)
    assert j["alignment"] == {"1": "ACGTTAAGGGGT", "2": "ACG--AA----T"}
    r = run("alignpair", f, "-k", "3")
    assert json.loads(r.stdout)["alignment"] == {"1": "ACGTTAAGGGGT", "2": "AC------GAAT"}
    # ambiguity handling (hidden -a option) and end-stop restoration with its score penalty
    f = fasta(tmp_path, ">1\nCTCTGGATAGTG\n>2\nCTATAGTR\n")
    assert json.loads(run("alignpair", f, "-a", "BEST").stdout)["alignment"]["2"] == "CT----ATAGTR"
    f = fasta(tmp_path, ">1\nCTCTGGATAGTGTAA\n>2\nCTATAGTG\n")
    j = json.loads(run("alignpair", f).stdout)
    assert j["alignment"] == {"1": "CTCTGGATAGTGTAA", "2": "CT----ATAGTG---"}


def test_alignpair_failures(tmp_path):
    for text, extra in ((">1\nGCGATTGCTGT\n>2\nGCGACTGTT\n", ["-k", "3"]), (">A\nCTCGGA\n>B\nCTCGG\n", ["-k", "3"]),
                        (">1\nCTCTGGATAGTG\n", []), (">1\nCTCTGGATAGTG\n>2\nCTATAGTG\n", ["-r", "seq_name"]),
                        (">1\nCTCTGGATAGTG\n>2\nCTATAGTG\n", ["-s"]), (">1\nCTCTAAATAGTG\n>2\nCTATAGTG\n", [])):
        r = run("alignpair", fasta(tmp_path, text), *extra)
        assert r.returncode != 0 and r.stderr.startswith("ERROR: ")


def test_score_flag(tmp_path):
    f = fasta(tmp_path, ">1\nCTCTGGATAGTG\n>2\nCT----ATAGTG\n")
    r = run("alignpair", f, "-s")
    assert r.returncode == 0 and float(r.stdout) == pytest.approx(1.50914, rel=1e-5)


def test_sample_matches_reference_strings_and_oracle_scores(tmp_path, tables):
    f = fasta(tmp_path, ">A\nCCCCCC\n>B\nCCCCCCCC\n")
    out = tmp_path / "s.json"
    r = run("sample", f, "-n", "3", "-s", "42", "-o", out)
    assert r.returncode == 0, r.stderr
    j = json.loads(out.read_text())
    assert [x["alignment"]["A"] for x in j] == ["CC--CCCC", "CCCCCC--", "CCCC--CC"]      # align_marginal.cc:664-667
    assert all(x["alignment"]["B"] == "CCCCCCCC" for x in j)
    np.testing.assert_allclose([x["score"] for x in j],
                               [-1.9466571807861328, -1.9466569423675537, -1.9466572999954224], rtol=1e-6)
    # byte-level layout of the JSON array (json.cc:211-227)
    assert out.read_text().startswith('[\n{\n  "alignment": {\n    "A": "CC--CCCC",\n    "B": "CCCCCCCC"\n  },\n  "score": ')
    assert out.read_text().endswith("\n}\n]\n")
    # example-003 with the string seed of BASELINE config 2, first samples against the reference golden
    g = next(s for s in util.load_json("sample_golden.json") if s["name"] == "example-003:default-seed")
    f = fasta(tmp_path, ">a\n%s\n>b\n%s\n" % (g["anc"], g["des"]), "e3.fasta")
    r = run("sample", f, "-n", g["n"], "-t", "0.0133")
    assert r.returncode == 0, r.stderr
    j = json.loads(r.stdout)
    # the CLI builds its own table (C++ expm); strings must still match the golden made with the
    # reference's golden P unless a near-tie flips, scores agree to table tolerance
    same = sum(1 for x, fgold in zip(j, g["first"]) if [x["alignment"]["a"], x["alignment"]["b"]] == fgold[:2])
    assert same == len(g["first"])


def test_sample_failures(tmp_path):
    assert run("sample", fasta(tmp_path, ">seq1\nAC\n>seq2\nACG\n")).returncode != 0
    assert run("sample", fasta(tmp_path, ">A\nCCC\n>B\nCCCC\n"), "-k", "3").returncode != 0
    assert run("sample", fasta(tmp_path, ">A\nCCC\n")).returncode != 0
    assert run("sample", fasta(tmp_path, ">A\nCCC\n>B\nCCC\n"), "-o", "/nonexistent-dir/x.json").returncode != 0


def test_phylip_and_json_input_and_sub_matrix(tmp_path, tables):
    """io.cc:184-222 readers (read_input test cases, io.cc:226-306) and --sub (io.cc:48-88;
    align_marginal.cc:304-344: a CSV of the MG94 Q gives the mar-mg alignment)."""
    phy = tmp_path / "in.phy"
    phy.write_text("2 12\n1         CTCTGGATAGTG\n2         CTATAGTG\n")
    j = json.loads(run("alignpair", phy).stdout)
    assert j["alignment"] == {"1": "CTCTGGATAGTG", "2": "CT----ATAGTG"}
    js = tmp_path / "in.json"
    js.write_text('{\n  "alignment": {\n    "a": "CTCTGGATAGTG",\n    "b": "CTATAGTG"\n  },\n  "score": 0.1\n}\n')
    out = tmp_path / "o.fa"
    assert run("alignpair", js, "-o", out).returncode == 0
    assert out.read_text() == ">a\nCTCTGGATAGTG\n>b\nCT----ATAGTG\n"
    # --sub: normalised MG94 Q (so that expm(Q * t) == mg94_p) written as codon,codon,value lines
    from oracle import table as otable
    Q, d = otable.mg94_q(0.2, otable.DEFAULT_PI)
    Q = Q / d
    codons = util.SENSE_CODONS
    csv = tmp_path / "q.csv"
    with open(csv, "w") as f:
        f.write("0.0133\n")
        for i in range(61):
            for k in range(61):
                f.write("%s,%s,%r\n" % (codons[i], codons[k], float(Q[i, k])))
    f = fasta(tmp_path, ">1\nCTCTGGATAGTG\n>2\nCTATAGTG\n")
    out = tmp_path / "sub.fasta"
    r = run("alignpair", f, "--sub", csv, "-o", out)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == ">1\nCTCTGGATAGTG\n>2\nCT----ATAGTG\n"
    bad = tmp_path / "bad.csv"
    bad.write_text("0.0133\nAAA,AAA,0.1\n")
    assert run("alignpair", f, "--sub", bad).returncode != 0
