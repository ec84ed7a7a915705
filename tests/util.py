"""Shared helpers for the test-suite and tools/gen_golden.py (no reference access at run time)."""
from __future__ import annotations

import gzip
import json
import os
import struct

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SENSE_CODONS = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT"
                if a + b + c not in ("TAA", "TAG", "TGA")]


def read_fasta_text(text: str):
    """FASTA semantics of the reference reader (src/lib/fasta.cc:39-76): ';' comment lines and
    empty lines skipped, whitespace stripped from sequence lines, multi-line records joined."""
    recs, name, content = [], None, []
    for line in text.split("\n"):
        if not line or line[0] == ";":
            continue
        if line[0] == ">":
            if name is not None:
                recs.append((name, "".join(content)))
            name = line[1:]
            if not name:
                raise ValueError("Input fasta file contains a sequence without a name.")
            content = []
        elif name is not None:
            content.append("".join(line.split()))
    if name is not None:
        recs.append((name, "".join(content)))
    return recs


def load_fasta(name: str):
    with gzip.open(os.path.join(GOLDEN, "data", name + ".fasta.gz"), "rt") as f:
        return read_fasta_text(f.read())


def load_tables():
    z = np.load(os.path.join(GOLDEN, "tables.npz"))
    return {k: np.ascontiguousarray(z[k], dtype=np.float32) for k in z.files}


def load_json(name: str):
    return json.load(open(os.path.join(GOLDEN, name)))


def bits_to_f32(h: str) -> np.float32:
    return np.float32(struct.unpack("<f", struct.pack("<I", int(h, 16)))[0])


def f32_bits(x) -> str:
    return "%08x" % struct.unpack("<I", struct.pack("<f", float(x)))[0]


def sanitise_ancestor(anc: str) -> str:
    """SURVEY 8(d) C3 protocol: every in-frame ancestor stop codon gets its 3rd base set to C
    (TAA/TAG -> TAC, TGA -> TGC); case preserved for the other symbols."""
    out = list(anc)
    for i in range(0, len(anc) - 2, 3):
        cod = anc[i:i + 3].upper().replace("U", "T")
        if cod in ("TAA", "TAG", "TGA"):
            out[i + 2] = "C"
    return "".join(out)


def random_pair(rng: np.random.RandomState, n_codons: int, k: int = 1, ambiguous: bool = False,
                sub=0.08, indel=0.04):
    """Small seeded ancestor/descendant pair: ancestor of sense codons, descendant = ancestor with
    substitutions and gap-unit (k) indels; lengths are multiples of lcm(3, k) as the reference
    requires (utils.cc:819-837)."""
    unit = 3 if k in (1, 3) else 3 * k
    n_codons = max(1, n_codons)
    if (n_codons * 3) % unit:
        n_codons += (unit - (n_codons * 3) % unit) // 3 + 0
        while (n_codons * 3) % unit:
            n_codons += 1
    anc = "".join(SENSE_CODONS[rng.randint(61)] for _ in range(n_codons))
    des = []
    i = 0
    while i < len(anc):
        r = rng.rand()
        if r < indel:                      # deletion of one gap unit
            i += k
            continue
        if r < 2 * indel:                  # insertion of one gap unit
            des.extend("ACGT"[rng.randint(4)] for _ in range(k))
        ch = anc[i]
        if rng.rand() < sub:
            ch = "ACGT"[rng.randint(4)]
        des.append(ch)
        i += 1
    des = des[:len(des) - len(des) % k] if len(des) >= k else list("ACGT"[rng.randint(4)] * k)
    if not des:
        des = list("A" * k)
    if ambiguous:
        codes = "RYMKSWBDHVN"
        for _ in range(max(1, len(des) // 10)):
            des[rng.randint(len(des))] = codes[rng.randint(len(codes))]
    if k not in (1, 3) and "".join(des[-3:]).upper().replace("U", "T") in ("TAA", "TAG", "TGA"):
        # the reference checks Lb % k before stripping a terminal stop codon (utils.cc:819-837), so
        # for k not dividing 3 a stripped stop leaves an unreachable terminal cell (UB upstream)
        des[-1] = "C"
    if rng.rand() < 0.3:                   # exercise case-insensitivity + U
        des = [c.lower() if rng.rand() < 0.5 else c for c in des]
        anc = "".join(c.lower() if rng.rand() < 0.2 else c for c in anc).replace("T", "U", 1)
    return anc, "".join(des)


def check_batch_properties(w, out_a, out_b, out_len, score, status, table, k, g, e, oracle_mod, sample=48, seed=0,
                           exact=True):
    """Size-independent checks of a whole batch of Viterbi alignments, vectorised over the output arenas
    (rows of pair p at byte a_off[p] + b_off[p] + p; layout of include/coati_gpu.h):
      * every pair succeeded, rows are NUL-terminated, max(La, Lb) <= length <= La + Lb,
      * the rows with the gaps removed ARE the input sequences (content, order and per-pair counts),
      * no column pairs a gap with a gap,
      * exact: a random sample and the largest lattices equal the oracle bit for bit (rows and float32 score).
    (Re-scoring the rows with alignment_score is NOT a general property: it charges adjacent indels differently
    from the three-state model, by several score units on some pairs.)
    Raises AssertionError naming the first failed property."""
    a_off, b_off = w["a_off"].astype(np.int64), w["b_off"].astype(np.int64)
    n = len(a_off) - 1
    la, lb = np.diff(a_off), np.diff(b_off)
    ln = out_len.astype(np.int64)
    assert int((status != 0).sum()) == 0, "pairs failed"
    assert np.all(ln >= np.maximum(la, lb)) and np.all(ln <= la + lb), "alignment length out of range"
    off = a_off[:-1] + b_off[:-1] + np.arange(n, dtype=np.int64)
    assert not out_a[off + ln].any() and not out_b[off + ln].any(), "rows are not NUL-terminated"
    start = np.concatenate(([0], np.cumsum(ln)[:-1]))
    idx = np.repeat(off - start, ln) + np.arange(int(ln.sum()), dtype=np.int64)
    ra, rb = out_a[idx], out_b[idx]
    gap = np.uint8(ord("-"))
    na, nb = ra != gap, rb != gap
    assert not np.any(~na & ~nb), "a column pairs two gaps"
    nz = ln > 0
    assert np.array_equal(np.add.reduceat(na.astype(np.int64), start[nz]), la[nz]), "ancestor symbols per row"
    assert np.array_equal(np.add.reduceat(nb.astype(np.int64), start[nz]), lb[nz]), "descendant symbols per row"
    assert np.array_equal(ra[na], w["anc_all"][: int(a_off[-1])]), "row a without gaps is not the ancestor"
    assert np.array_equal(rb[nb], w["des_all"][: int(b_off[-1])]), "row b without gaps is not the descendant"
    if not exact:
        return 0
    rng = np.random.RandomState(seed)
    picks = set(rng.randint(0, n, size=sample).tolist()) | set(np.argsort(la * lb)[-4:].tolist())
    for p in sorted(picks):
        anc = w["anc_all"][a_off[p]:a_off[p + 1]].tobytes().decode("latin-1")
        des = w["des_all"][b_off[p]:b_off[p + 1]].tobytes().decode("latin-1")
        enc = (w["a_all"][a_off[p]:a_off[p + 1]], w["b_all"][b_off[p]:b_off[p + 1]])
        oa, ob, osc = oracle_mod.viterbi(anc, des, table, g, e, k, enc=enc)
        got_a = out_a[off[p]:off[p] + ln[p]].tobytes().decode("latin-1")
        got_b = out_b[off[p]:off[p] + ln[p]].tobytes().decode("latin-1")
        assert (got_a, got_b) == (oa, ob), f"pair {p}: rows differ from the oracle"
        assert f32_bits(score[p]) == f32_bits(osc), f"pair {p}: score differs from the oracle"
    return len(picks)


def rows_equal(w, out_len, mask, a1, b1, a2, b2):
    """The rows of the pairs selected by `mask` are byte-identical in two pairs of output arenas (bytes after a
    row's terminator are scratch and are not compared)."""
    a_off, b_off = w["a_off"].astype(np.int64), w["b_off"].astype(np.int64)
    n = len(a_off) - 1
    off = (a_off[:-1] + b_off[:-1] + np.arange(n, dtype=np.int64))[mask]
    ln = out_len.astype(np.int64)[mask]
    start = np.concatenate(([0], np.cumsum(ln)[:-1]))
    idx = np.repeat(off - start, ln) + np.arange(int(ln.sum()), dtype=np.int64)
    return bool(np.array_equal(a1[idx], a2[idx]) and np.array_equal(b1[idx], b2[idx]))


def ends_with_stop(arena, off):
    """per pair: the last three symbols are TAA / TAG / TGA (upper case, as the generator writes them)"""
    end = off[1:].astype(np.int64)
    ok = np.diff(off.astype(np.int64)) >= 3
    e3 = np.where(ok, end, 3)
    x, y, z = arena[e3 - 3], arena[e3 - 2], arena[e3 - 1]
    T, A, G = ord("T"), ord("A"), ord("G")
    return ok & (x == T) & (((y == A) & ((z == A) | (z == G))) | ((y == G) & (z == A)))


def compare_entry_points(w, staged, piped, raw):
    """Results (out_a, out_b, out_len, score, status) of the same batch through the staged batch API, the
    pipelined CSR call and the raw-sequence call: the first two are identical pair by pair; the raw call is
    identical wherever no end stop is trimmed and restored."""
    oa1, ob1, ln1, sc1, st1 = staged
    oa2, ob2, ln2, sc2, st2 = piped
    oa3, ob3, ln3, sc3, st3 = raw
    n = len(ln1)
    assert np.array_equal(st1, st2) and np.array_equal(ln1, ln2), "pipelined call: status / length"
    assert np.array_equal(sc1.view(np.uint32), sc2.view(np.uint32)), "pipelined call: scores"
    assert rows_equal(w, ln1, np.ones(n, dtype=bool), oa1, ob1, oa2, ob2), "pipelined call: rows"
    assert int((st3 != 0).sum()) == 0, "raw call: pairs failed"
    plain = ~(ends_with_stop(w["anc_all"], w["a_off"]) | ends_with_stop(w["des_all"], w["b_off"]))
    assert plain.mean() > 0.8
    assert np.array_equal(ln1[plain], ln3[plain]), "raw call: length"
    assert np.array_equal(sc1[plain].view(np.uint32), sc3[plain].view(np.uint32)), "raw call: scores"
    assert rows_equal(w, ln1, plain, oa1, ob1, oa3, ob3), "raw call: rows"
    return int((~plain).sum())
