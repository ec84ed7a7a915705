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
