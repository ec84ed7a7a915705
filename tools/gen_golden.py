#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ from the reference at /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    make -C oracle && python tools/gen_golden.py

What it writes
  mg94_p_default.npy      the reference's golden 61x61 P matrix (src/include/coati/mg94p.tcc:26)
  ecm_unrest.npz          ECM exchangeabilities + codon frequencies (Kosiol et al. 2007 supplement,
                          as tabulated in src/include/coati/ecm_unrest.tcc:28,581)
  tables.npz              named 183x15 float32 marginal tables used by every DP parity test
  data/*.fasta.gz         the reference's sample/benchmark inputs (sampledata/, benchmark/data/)
  viterbi_golden.json     alignments + float32 score bits produced HERE by the unmodified reference
                          (oracle/_ref/libcoati_ref.so) on those inputs and on seeded random pairs
  sample_golden.json      forward+sampleback outputs of the unmodified reference for fixed seeds
Nothing here copies reference source code: only numeric data tables and reference *outputs*.
"""
import gzip
import hashlib
import json
import os
import re
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

import oracle  # noqa: E402
from oracle import table as otable  # noqa: E402
from tests import util  # noqa: E402


def parse_floats(path, name, count):
    src = open(path).read()
    m = re.search(r"%s\s*(\[[^\]]*\])+\s*=?\s*\{" % re.escape(name), src)
    body = src[m.end():]
    depth, end = 1, 0
    for end, ch in enumerate(body):
        depth += ch == "{"
        depth -= ch == "}"
        if depth == 0:
            break
    vals = re.findall(r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)", body[:end])
    assert len(vals) == count, (name, len(vals))
    return np.array([float(v) for v in vals], dtype=np.float64).astype(np.float32)


def fbits(x):
    return "%08x" % struct.unpack("<I", struct.pack("<f", float(x)))[0]


def sha(*parts):
    h = hashlib.sha256()
    for p in parts:
        h.update(p.encode() if isinstance(p, str) else p)
        h.update(b"\0")
    return h.hexdigest()


def main():
    assert oracle.ref is not None, "build oracle/_ref first (make -C oracle)"
    os.makedirs(os.path.join(OUT, "data"), exist_ok=True)

    # ---- numeric data tables -----------------------------------------------------------
    P = parse_floats(f"{REF}/src/include/coati/mg94p.tcc", "mg94P", 61 * 61).reshape(61, 61)
    np.save(os.path.join(OUT, "mg94_p_default.npy"), P)
    ex = parse_floats(f"{REF}/src/include/coati/ecm_unrest.tcc", "exchang", 61 * 61).reshape(61, 61)
    epi = parse_floats(f"{REF}/src/include/coati/ecm_unrest.tcc", "ecm_pi", 61)
    np.savez(os.path.join(OUT, "ecm_unrest.npz"), exchang=ex, ecm_pi=epi)

    tables = {
        # the reference's own golden P, marginalised: closest thing to "the reference's table"
        "mg_golden": otable.marginal_p(P),
        "mg_golden_best": otable.marginal_p(P, amb="BEST"),
        "mg_default": otable.build_table("mar-mg"),
        "ecm_default": otable.build_table("mar-ecm"),
        # BASELINE config 5: -w 0.5 -p 0.25 0.25 0.25 0.25 -t 0.05
        "mg_c5": otable.build_table("mar-mg", br_len=0.05, omega=0.5, pi=(0.25,) * 4),
    }
    np.savez(os.path.join(OUT, "tables.npz"), **tables)

    # ---- input data --------------------------------------------------------------------
    files = {}
    for d, names in (("sampledata", ["example-001", "example-002", "example-003", "example-10k",
                                     "example-20k", "example-40k", "example-80k", "example-160k"]),
                     ("benchmark/data", ["benchmark_156", "benchmark_1k", "benchmark_2k",
                                         "benchmark_4k", "benchmark_8k", "benchmark_16k",
                                         "benchmark_32k"])):
        for n in names:
            raw = open(f"{REF}/{d}/{n}.fasta", "rb").read()
            with gzip.GzipFile(os.path.join(OUT, "data", n + ".fasta.gz"), "wb", mtime=0) as f:
                f.write(raw)
            files[n] = util.read_fasta_text(raw.decode())

    # ---- Viterbi goldens from the unmodified reference ---------------------------------
    cases = []

    def add(name, anc, des, tname, k=1, inline=True, g=oracle.DEFAULT_G, e=oracle.DEFAULT_E):
        anc_t, s0 = oracle.trim_end_stop(anc)
        des_t, s1 = oracle.trim_end_stop(des)
        a, b, sc = oracle.viterbi(anc_t, des_t, tables[tname], g=g, e=e, k=k, impl="ref")
        a2, b2, sc2 = oracle.restore_end_stops(a, b, sc, (s0, s1), g, e)
        c = {"name": name, "table": tname, "k": k, "g_bits": fbits(g), "e_bits": fbits(e),
             "score_bits": fbits(sc), "final_score_bits": fbits(sc2), "len": len(a2),
             "sha256": sha(a2, b2)}
        if inline:
            c.update(anc=anc, des=des, aln_a=a2, aln_b=b2)
        else:
            c["file"] = name
        cases.append(c)
        return a2, b2

    # the reference's own known answers (align_marginal.cc:149-240): assert them while recording
    known = [
        ("ref-test-fasta", "CTCTGGATAGTG", "CTATAGTG", "mg_golden", 1, "CTCTGGATAGTG", "CT----ATAGTG"),
        ("ref-test-phylip", "GCGACTGTT", "GCGATTGCTGTT", "mg_golden", 1, "GCGA---CTGTT", "GCGATTGCTGTT"),
        ("ref-test-ecm", "GCGACTGTT", "GCGATTGCTGTT", "ecm_default", 1, "GCGA---CTGTT", "GCGATTGCTGTT"),
        ("ref-test-2dels", "ACGTTAAGGGGT", "ACGAAT", "mg_golden", 1, "ACGTTAAGGGGT", "ACG--AA----T"),
        ("ref-test-k3", "ACGTTAAGGGGT", "ACGAAT", "mg_golden", 3, "ACGTTAAGGGGT", "AC------GAAT"),
        ("ref-test-amb-sum", "CTCTGGATAGTG", "CTATAGTR", "mg_golden", 1, "CTCTGGATAGTG", "CT----ATAGTR"),
        ("ref-test-amb-best", "CTCTGGATAGTG", "CTATAGTR", "mg_golden_best", 1, "CTCTGGATAGTG", "CT----ATAGTR"),
    ]
    for name, anc, des, t, k, ea, eb in known:
        a, b = add(name, anc, des, t, k)
        assert (a, b) == (ea, eb), (name, a, b)

    for n in ["example-001", "example-002", "example-003", "benchmark_156", "benchmark_1k",
              "benchmark_2k", "benchmark_4k"]:
        (_, anc), (_, des) = files[n]
        for t in (["mg_golden", "ecm_default", "mg_c5"] if "bench" not in n or n.endswith("156")
                  else ["mg_golden"]):
            add(f"{n}:{t}", anc, des, t, 1, inline=len(anc) <= 500)
            cases[-1]["file"] = n
    # k = 3 on real data whose lengths allow it
    for n in ["example-003", "benchmark_156"]:
        (_, anc), (_, des) = files[n]
        if len(anc) % 3 == 0 and len(des) % 3 == 0:
            add(f"{n}:mg_golden:k3", anc, des, "mg_golden", 3, inline=True)
            cases[-1]["file"] = n
    # sanitised 10k (reference rejects the raw file: in-frame ancestor stops, utils.cc:511-514)
    (_, anc), (_, des) = files["example-10k"]
    add("example-10k:sanitised", util.sanitise_ancestor(anc), des, "mg_golden", 1, inline=False)
    cases[-1]["file"] = "example-10k"
    cases[-1]["sanitised"] = True

    # seeded random pairs (generator in tests/util.py), incl. ambiguity codes and k = 2, 3
    rng = np.random.RandomState(20240603)
    for idx in range(40):
        k = [1, 1, 3, 2][idx % 4]
        amb = idx % 5 == 0
        anc, des = util.random_pair(rng, n_codons=int(rng.randint(1, 60)), k=k, ambiguous=amb)
        t = ["mg_golden", "ecm_default", "mg_c5"][idx % 3]
        add(f"rand-{idx}", anc, des, t, k)
    json.dump(cases, open(os.path.join(OUT, "viterbi_golden.json"), "w"), indent=0)

    # ---- sampling goldens ----------------------------------------------------------------
    samples = []

    def add_sample(name, anc, des, tname, seeds, n, k=1, keep=10, expect=None):
        st = oracle.ref_seed_state(seeds)
        anc_t, s0 = oracle.trim_end_stop(anc)
        des_t, s1 = oracle.trim_end_stop(des)
        out, sc, st2, _ = oracle.sample(anc_t, des_t, tables[tname], st, n, k=k, impl="ref")
        if expect:
            assert [o[0] for o in out] == expect, (name, out)
        h = hashlib.sha256()
        for (a, b), s in zip(out, sc):
            h.update((a + "\0" + b + "\0" + fbits(s)).encode())
        samples.append({
            "name": name, "anc": anc, "des": des, "table": tname, "seeds": seeds, "n": n, "k": k,
            "state0": [int(st[0]), int(st[1])], "state1": [int(st2[0]), int(st2[1])],
            "first": [[a, b, fbits(s)] for (a, b), s in list(zip(out, sc))[:keep]],
            "sha256": h.hexdigest(), "lens_sum": int(sum(len(a) for a, _ in out))})

    # align_marginal.cc:653-671 (strings are the reference's goldens; scores are table-dependent)
    add_sample("ref-test-size1", "CCCCCC", "CCCCCCCC", "mg_golden", ["42"], 1, expect=["CC--CCCC"])
    add_sample("ref-test-del", "CCCCCC", "CCCC", "mg_golden", ["42"], 1, expect=["CCCCCC"])
    add_sample("ref-test-size3", "CCCCCC", "CCCCCCCC", "mg_golden", ["42"], 3,
               expect=["CC--CCCC", "CCCCCC--", "CCCC--CC"])
    (_, anc), (_, des) = files["example-003"]
    add_sample("example-003:random42", anc, des, "mg_golden", ["random42"], 1000, keep=5)
    add_sample("example-003:default-seed", anc, des, "mg_golden", [""], 20, keep=3)
    (_, anc), (_, des) = files["benchmark_156"]
    add_sample("benchmark_156:s7", anc, des, "mg_c5", ["7", "x"], 50, keep=3)
    rng = np.random.RandomState(7)
    for idx in range(6):
        k = [1, 3, 2][idx % 3]
        anc, des = util.random_pair(rng, n_codons=int(rng.randint(2, 40)), k=k, ambiguous=idx == 4)
        add_sample(f"rand-{idx}", anc, des, ["mg_golden", "ecm_default"][idx % 2],
                   [str(idx), "seed"], 25, k=k, keep=3)
    json.dump(samples, open(os.path.join(OUT, "sample_golden.json"), "w"), indent=0)

    # ---- RNG goldens ---------------------------------------------------------------------
    rngs = []
    for seeds in (["42"], ["random42"], [""], ["-17", "abc", "4294967296"], ["2147483647"],
                  ["café"], ["1", "2", "3", "4", "5", "6", "7", "8", "9"]):
        st = oracle.ref_seed_state(seeds)
        import ctypes as C
        s = (C.c_uint64 * 2)(int(st[0]), int(st[1]))
        bits = [int(oracle.ref.coati_ref_rng_bits(s)) for _ in range(8)]
        f24 = [fbits(oracle.ref.coati_ref_rng_f24(s)) for _ in range(4)]
        rngs.append({"seeds": seeds, "state": [int(st[0]), int(st[1])], "bits": bits, "f24": f24})
    json.dump(rngs, open(os.path.join(OUT, "rng_golden.json"), "w"), indent=0)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
