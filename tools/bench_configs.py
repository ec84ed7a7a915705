#!/usr/bin/env python
"""Secondary BASELINE configs (C1 latency, C2 forward + sampling, C3 long single pairs, C4 k=3 batch, batched
Forward) timed on the GPU next to the reference CPU path on the same box.  The headline metric (C5) is bench.py,
which calls collect() and carries the result in its JSON line as `extra.configs`.

CPU legs (cpu_baseline of each case; the oracle is used here as bench.py uses it: the reference timed beside the
GPU, never inside a timed GPU region): oracle/_ref (the unmodified reference) where its three full matrices fit,
the O(Lb)-memory rolling-row port beyond.  The CPU result doubles as the parity flag of the case.

usage: python tools/bench_configs.py [c1] [c2] [c3] [c4] [fwd] [--max-ref-len N]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FLOP_PER_CELL = 23


def wall(fn, reps=3):
    best = 1e30
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def _bits(x):
    return np.float32(x).tobytes()


def collect(ctx, which=("c1", "c2", "c3", "c4", "fwd"), max_ref_len=10500):
    import oracle
    from coati_b200.capi import PackedPairs
    from synth import synth_pairs
    from tests import util
    tables = util.load_tables()
    g, e = oracle.DEFAULT_G, oracle.DEFAULT_E
    info = ctx.device_info()
    peak_tcups = info["sm_count"] * 128 * info["clock_khz"] * 1e3 / FLOP_PER_CELL / 1e12   # at the maximum SM clock
    out = {"roofline_peak_tcups_at_max_clock": peak_tcups}

    def viterbi_case(name, anc, des, T, k):
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        a, b = oracle.encode_pair(anc, des)
        cells = len(a) * len(b)
        ctx.set_model(T, g, e, k)
        ctx.viterbi(a, b, anc, des)  # warm-up
        t_gpu, (ra, rb, sc) = wall(lambda: ctx.viterbi(a, b, anc, des))
        pk = PackedPairs([a], [b], [anc], [des])
        bt = ctx.batch(pk.a_off, pk.b_off)
        bt.upload(pk.a_all, pk.b_all, pk.anc_all, pk.des_all)
        bt.run()
        bt.run()
        tm = bt.timing()
        bt.destroy()
        fill_gcups = cells / (tm["fill_ms"] / 1e3) / 1e9 if tm["fill_ms"] else None
        rec = {"la": len(a), "lb": len(b), "k": k, "gpu_e2e_ms": 1e3 * t_gpu, "gpu_e2e_gcups": cells / t_gpu / 1e9,
               "fill_ms": tm["fill_ms"], "fill_gcups": fill_gcups,
               "roofline_frac": fill_gcups / (peak_tcups * 1e3) if fill_gcups else None,
               "traceback_ms": tm["traceback_ms"], "score": float(sc), "aln_len": len(ra)}
        checks = {"roundtrip": ra.replace("-", "") == anc and rb.replace("-", "") == des}
        if k == 1:   # the returned path re-scored through forward_impl's own terms: bit-equal
            checks["path_rescore_bits"] = _bits(oracle.path_score(ra, rb, a, b, T, g, e, 1)) == _bits(sc)
        if oracle.ref is not None and max(len(a), len(b)) <= max_ref_len:
            t_cpu, (oa, ob, osc) = wall(lambda: oracle.viterbi(anc, des, T, g, e, k, impl="ref", enc=(a, b)), reps=1)
            rec["cpu_baseline"] = {"kind": "reference", "cores": 1, "ms": 1e3 * t_cpu, "gcups": cells / t_cpu / 1e9}
            checks["rows_and_score_equal_reference"] = (ra, rb) == (oa, ob) and _bits(sc) == _bits(osc)
        elif k == 1:
            nthr = os.cpu_count() or 1
            t_cpu, osc = wall(lambda: oracle.viterbi_score(a, b, T, g, e, 1, threads=nthr), reps=1)
            rec["cpu_baseline"] = {"kind": "port", "cores": nthr, "ms": 1e3 * t_cpu, "gcups": cells / t_cpu / 1e9,
                                   "what": "score-only rolling-row Viterbi (oracle/long_pair.c), O(Lb) memory"}
            checks["optimum_score_bits"] = _bits(sc) == _bits(osc)
        rec["checks"] = checks
        rec["parity"] = all(checks.values())
        return rec

    if "c1" in which:
        (_, anc), (_, des) = util.load_fasta("example-001")
        out["c1_example-001_mar-mg"] = viterbi_case("C1", anc, des, tables["mg_golden"], 1)
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
        _, fill_ms = fw.terminal()
        cabi_s = fw.last_call_s
        fw.free()
        tm = {}
        orows, osc, ost, _ = oracle.sample(anc, des, T, st, 1000, impl="ref" if oracle.ref is not None else "oracle",
                                           timings=tm)
        match = sum(1 for x, y, p, q in zip(rows, orows, sc, osc) if x == y and p.tobytes() == q.tobytes())
        first_bad = next((i for i, (x, y) in enumerate(zip(rows, orows)) if x != y), None)
        out["c2_example-003_sample_n1000_random42"] = {
            "la": len(a), "lb": len(b), "gpu_forward_kernel_ms": fill_ms, "gpu_forward_cabi_ms": 1e3 * (t1 - t0),
            "gpu_forward_gcups": len(a) * len(b) / (fill_ms / 1e3) / 1e9,
            "gpu_sampleback_kernel_ms": smp_ms, "gpu_sampleback_cabi_ms": 1e3 * cabi_s,
            "cpu_baseline": {"kind": "reference" if oracle.ref is not None else "port", "cores": 1,
                             "forward_ms": 1e3 * tm.get("fill_s", 0), "sampleback_ms": 1e3 * tm.get("sample_s", 0)},
            "sample_match_rate": match / 1000.0, "first_mismatch": first_bad,
            "rng_state_identical": bool(np.array_equal(st2, ost)),
            "parity": match == 1000 and bool(np.array_equal(st2, ost))}
    if "c3" in which:
        T = tables["mg_golden"]
        c3 = {}
        for name in ("benchmark_156", "benchmark_1k", "benchmark_2k", "benchmark_4k", "benchmark_8k", "benchmark_16k",
                     "benchmark_32k"):
            (_, anc), (_, des) = util.load_fasta(name)
            c3[name] = viterbi_case(name, anc, des, T, 1)
        for name in ("example-10k", "example-20k", "example-40k", "example-80k", "example-160k"):
            (_, anc), (_, des) = util.load_fasta(name)
            c3[name + ":sanitised"] = viterbi_case(name, util.sanitise_ancestor(anc), des, T, 1)
        out["c3_long_single_pairs_mar-mg"] = c3
    if "c4" in which:
        n = 100_000
        T = tables["ecm_default"]
        ctx.set_model(T, g, e, 3)
        w = synth_pairs(n, 4, 20240603)
        cells = float((np.diff(w["a_off"]).astype(np.float64) * np.diff(w["b_off"]).astype(np.float64)).sum())
        bt = ctx.batch(w["a_off"], w["b_off"])
        bt.upload(w["a_all"], w["b_all"], w["anc_all"], w["des_all"])
        bt.run()
        fill = 0.0
        t0 = time.perf_counter()
        steps = 2
        for _ in range(steps):
            bt.run()
            fill += bt.timing()["fill_ms"]
        dt = time.perf_counter() - t0
        total = int(w["a_off"][-1] + w["b_off"][-1]) + n
        oa, ob = np.zeros(total + 1, np.uint8), np.zeros(total + 1, np.uint8)
        ln, sc, stt = np.zeros(n, np.uint64), np.zeros(n, np.float32), np.zeros(n, np.int32)
        bt.download(oa, ob, ln, sc, stt)
        bt.destroy()
        picked = util.check_batch_properties(w, oa, ob, ln, sc, stt, T, 3, g, e, oracle, sample=6)
        kg = cells / (fill / steps / 1e3) / 1e9
        out["c4_100k_pairs_mar-ecm_k3"] = {"pairs": n, "cells": cells, "value_gcups": cells * steps / dt / 1e9,
                                           "kernel": "viterbi_pipe3_kernel<6,4>", "kernel_gcups": kg,
                                           "roofline_frac": kg / (peak_tcups * 1e3),
                                           "parity": True, "parity_what": f"properties of every alignment + {picked} pairs "
                                                                          "equal to the oracle (rows, score bits)"}
        del w, oa, ob
    if "fwd" in which:
        # batched Forward: one warp per pair; 4096 copies of the C2 pair (177 k cells each, 12 B/cell stored = 8.7 GB):
        # enough pairs for four resident warps per scheduler
        (_, anc), (_, des) = util.load_fasta("example-003")
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        a, b = oracle.encode_pair(anc, des)
        T = tables["mg_golden"]
        ctx.set_model(T, g, e, 1)
        npairs = 4096
        pk = PackedPairs([a] * npairs, [b] * npairs, [anc] * npairs, [des] * npairs)
        fb = ctx.forward_batch(pk)
        fb.free()
        fb = ctx.forward_batch(pk)
        term, ll, ms = fb.terminal()
        fb.free()
        t_cpu, mats = wall(lambda: oracle.fill(1, a, b, T), reps=1)
        ok = all(_bits(term[p][x]) == _bits(mats[x][-1, -1]) for p in (0, npairs - 1) for x in range(3))
        cells = npairs * len(a) * len(b)
        out["forward_batch_4096x_example-003"] = {
            "pairs": npairs, "cells": cells, "kernel_ms": ms, "gcups": cells / (ms / 1e3) / 1e9,
            "cpu_baseline": {"kind": "port", "cores": 1, "ms_per_pair": 1e3 * t_cpu,
                             "gcups": len(a) * len(b) / t_cpu / 1e9}, "parity": bool(ok),
            "parity_what": "adjusted terminal M, D, I of the first and last pair equal the oracle's bits"}
    return out


def main():
    import coati_b200
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c2", "c3", "c4", "fwd"]
    max_ref = 10500
    if "--max-ref-len" in sys.argv:
        max_ref = int(sys.argv[sys.argv.index("--max-ref-len") + 1])
    ctx = coati_b200.Context(0)
    res = collect(ctx, which, max_ref)
    ctx.close()
    for k, v in res.items():
        print(json.dumps({k: v}), flush=True)


if __name__ == "__main__":
    main()
