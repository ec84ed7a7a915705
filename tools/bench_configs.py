#!/usr/bin/env python
"""Secondary BASELINE configs (C1 latency, C2 sampling, C3 long single pairs, C4 k=3 batch) timed on
the GPU next to the reference CPU path on the same box.  Prints one JSON line per case; the headline
metric (C5) is bench.py.  usage: python tools/bench_configs.py [c1] [c2] [c3] [c4] [--max-cpu-len N]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import coati_b200  # noqa: E402
import oracle  # noqa: E402
from tests import util  # noqa: E402


def wall(fn, reps=3):
    best = 1e30
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c2", "c3"]
    max_cpu = 20000
    if "--max-cpu-len" in sys.argv:
        max_cpu = int(sys.argv[sys.argv.index("--max-cpu-len") + 1])
    tables = util.load_tables()
    ctx = coati_b200.Context(0)
    g, e = oracle.DEFAULT_G, oracle.DEFAULT_E

    def viterbi_case(name, anc, des, T, k, cpu=True):
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        a, b = oracle.encode_pair(anc, des)
        ctx.set_model(T, g, e, k)
        ctx.viterbi(a, b, anc, des)  # warm-up
        t_gpu, (ra, rb, sc) = wall(lambda: ctx.viterbi(a, b, anc, des))
        from coati_b200.capi import PackedPairs
        pk = PackedPairs([a], [b], [anc], [des])
        bt = ctx.batch(pk.a_off, pk.b_off)
        bt.upload(pk.a_all, pk.b_all, pk.anc_all, pk.des_all)
        bt.run()
        bt.run()
        tm = bt.timing()
        bt.destroy()
        rec = {"case": name, "la": len(a), "lb": len(b), "k": k, "gpu_e2e_ms": 1e3 * t_gpu,
               "gpu_e2e_gcups": len(a) * len(b) / t_gpu / 1e9, "fill_ms": tm["fill_ms"],
               "fill_gcups": len(a) * len(b) / (tm["fill_ms"] / 1e3) / 1e9, "traceback_ms": tm["traceback_ms"],
               "score": float(sc), "aln_len": len(ra)}
        if cpu and oracle.ref is not None:
            t_cpu, (oa, ob, osc) = wall(lambda: oracle.viterbi(anc, des, T, g, e, k, impl="ref", enc=(a, b)), reps=1)
            rec.update(cpu_ms=1e3 * t_cpu, cpu_gcups=len(a) * len(b) / t_cpu / 1e9, speedup=t_cpu / t_gpu,
                       identical=(ra, rb) == (oa, ob) and np.float32(sc).tobytes() == np.float32(osc).tobytes())
        else:
            rescored = oracle.alignment_score(ra, rb, T, g, e, k)
            rec.update(rescored=float(rescored), roundtrip=ra.replace("-", "") == anc and rb.replace("-", "") == des)
        print(json.dumps(rec), flush=True)

    if "c1" in which:
        (_, anc), (_, des) = util.load_fasta("example-001")
        viterbi_case("C1 example-001 mar-mg", anc, des, tables["mg_golden"], 1)
    if "c2" in which:
        (_, anc), (_, des) = util.load_fasta("example-003")
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        a, b = oracle.encode_pair(anc, des)
        T = tables["mg_golden"]
        ctx.set_model(T, g, e, 1)
        st = oracle.seed_state(["random42"])
        fw = ctx.forward(a, b)
        fw.sampleback(anc, des, st, 10)
        fw.free()
        t0 = time.perf_counter()
        fw = ctx.forward(a, b)
        t1 = time.perf_counter()
        rows, sc, st2, smp_ms = fw.sampleback(anc, des, st, 1000)
        t2 = time.perf_counter()
        _, fill_ms = fw.terminal()
        fw.free()
        tm = {}
        orows, osc, ost, _ = oracle.sample(anc, des, T, st, 1000, impl="ref" if oracle.ref is not None else "oracle",
                                           timings=tm)
        match = sum(1 for x, y, p, q in zip(rows, orows, sc, osc) if x == y and p.tobytes() == q.tobytes())
        first_bad = next((i for i, (x, y) in enumerate(zip(rows, orows)) if x != y), None)
        print(json.dumps({"case": "C2 example-003 sample -n 1000 -s random42", "la": len(a), "lb": len(b),
                          "gpu_forward_kernel_ms": fill_ms, "gpu_forward_e2e_ms": 1e3 * (t1 - t0),
                          "gpu_sampleback_kernel_ms": smp_ms, "gpu_sampleback_cabi_ms": 1e3 * fw.last_call_s,
                          "gpu_sampleback_python_e2e_ms": 1e3 * (t2 - t1),
                          "cpu_forward_ms": 1e3 * tm.get("fill_s", 0), "cpu_sampleback_ms": 1e3 * tm.get("sample_s", 0),
                          "sample_match_rate": match / 1000.0, "first_mismatch": first_bad,
                          "rng_state_identical": bool(np.array_equal(st2, ost))}), flush=True)
    if "c3" in which:
        T = tables["mg_golden"]
        for name in ("benchmark_1k", "benchmark_4k", "benchmark_16k", "benchmark_32k"):
            (_, anc), (_, des) = util.load_fasta(name)
            viterbi_case("C3 " + name, anc, des, T, 1, cpu=len(anc) <= max_cpu)
        for name in ("example-10k", "example-20k", "example-40k", "example-80k", "example-160k"):
            (_, anc), (_, des) = util.load_fasta(name)
            try:
                viterbi_case("C3 " + name + " (sanitised)", util.sanitise_ancestor(anc), des, T, 1,
                             cpu=len(anc) <= max_cpu)
            except coati_b200.CoatiGpuError as ex:
                print(json.dumps({"case": "C3 " + name, "error": str(ex)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
