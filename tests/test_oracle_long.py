"""CPU: the two O(La + Lb)-memory validators for long pairs (oracle/long_pair.c) are pinned bit for bit to the
full-matrix oracle (itself pinned to the reference, tests/test_oracle_vs_ref.py) before the GPU tests trust them."""
import numpy as np
import pytest

import oracle
from tests import util


def _pairs(seed, n, k, max_codons):
    rng = np.random.RandomState(seed)
    out = []
    while len(out) < n:
        anc, des = util.random_pair(rng, int(rng.randint(1, max_codons)), k=k, ambiguous=len(out) % 3 == 0)
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        if len(anc) % k or len(des) % k:
            continue
        out.append((anc, des))
    return out


@pytest.mark.parametrize("tname,g,e", [("mg_golden", 0.001, 5.0 / 6.0), ("ecm_default", 0.001, 5.0 / 6.0),
                                       ("mg_c5", 0.2, 0.9)])
def test_path_score_equals_viterbi_score_k1(tname, g, e, tables):
    """k = 1: the returned path re-scored through the fill's own terms IS the optimum, bit for bit."""
    T = tables[tname]
    g, e = np.float32(g), np.float32(e)
    for anc, des in _pairs(7, 60, 1, 120) + [("", "ACG"), ("AAA", ""), ("AAACCC", "A")]:
        a, b = oracle.encode_pair(anc, des)
        ra, rb, sc = oracle.viterbi(anc, des, T, g, e, 1, enc=(a, b))
        assert util.f32_bits(oracle.path_score(ra, rb, a, b, T, g, e, 1)) == util.f32_bits(sc)


@pytest.mark.parametrize("k", [2, 3])
def test_path_score_follows_the_fill_for_k_gt_1(k, tables):
    """k > 1: the path value is read off the full matrices: the terminal state's adjusted value when the walk is the
    fill's own arg-max chain (traceback compares with +ge where the fill used +ge*k, align_pair.cc:285-296, so the
    traced path need not be the optimum; the re-scored value is still a lower bound of the score)."""
    T = tables["mg_golden"]
    for anc, des in _pairs(11 + k, 40, k, 90):
        a, b = oracle.encode_pair(anc, des)
        ra, rb, sc = oracle.viterbi(anc, des, T, k=k, enc=(a, b))
        ps = oracle.path_score(ra, rb, a, b, T, k=k)
        assert ps <= sc
        assert float(ps) == pytest.approx(float(sc), rel=1e-4, abs=1.0)


def test_path_score_rejects_non_alignments(tables):
    T = tables["mg_golden"]
    a, b = oracle.encode_pair("AAACCC", "AAACCC")
    with pytest.raises(ValueError):
        oracle.path_score("AAACC", "AAACC", a, b, T)          # too short
    with pytest.raises(ValueError):
        oracle.path_score("AAACCC-", "AAACCC-", a, b, T)      # gap-gap column
    with pytest.raises(ValueError):
        oracle.path_score("A-AACCC", "-AAACCC", a, b, T)      # insertion right after a deletion (no D->I)
    a, b = oracle.encode_pair("AAACCC", "AAAC")
    with pytest.raises(ValueError):
        oracle.path_score("AAACCC--", "AA--AC", a, b, T)      # rows of different content


@pytest.mark.parametrize("threads", [1, 2, 3, 8])
def test_rolling_row_score_equals_full_matrix(threads, tables):
    rng = np.random.RandomState(threads)
    cases = _pairs(21 + threads, 25, 1, 200) + [("", "ACG"), ("AAA", ""), ("", ""), ("AAACCC", "A")]
    # a pair wider than one column block (2048) and taller than the strip count
    anc, des = util.random_pair(rng, 900, k=1)
    cases.append((oracle.trim_end_stop(anc)[0], oracle.trim_end_stop(des)[0]))
    for tname in ("mg_golden", "mg_c5"):
        T = tables[tname]
        for anc, des in cases:
            a, b = oracle.encode_pair(anc, des)
            M, D, I = oracle.fill(0, a, b, T)
            want = max(M[-1, -1], D[-1, -1], I[-1, -1])
            assert util.f32_bits(oracle.viterbi_score(a, b, T, threads=threads)) == util.f32_bits(want)


def test_rolling_row_score_on_a_golden_long_pair(tables):
    """example-10k (sanitised): the committed reference score, reproduced with O(Lb) memory."""
    c = next(c for c in util.load_json("viterbi_golden.json") if c["name"] == "example-10k:sanitised")
    (_, anc), (_, des) = util.load_fasta("example-10k")
    anc = util.sanitise_ancestor(anc)
    a, b = oracle.encode_pair(anc, des)
    assert util.f32_bits(oracle.viterbi_score(a, b, tables["mg_golden"])) == c["score_bits"]
