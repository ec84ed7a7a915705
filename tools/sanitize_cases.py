#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
inter-pair fills (k = 1 with 4- and 16-column tables, k = 3, generic k = 2), the intra-pair wavefront with its
sentinel hand-off, traceback (thread, burst list, burst long) and expansion, Forward (banded wavefront, batch,
any-k kernel) and both samplebacks.  Results are checked against the oracle so that a sanitizer-clean run is also
a correct one.  usage: compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import coati_b200  # noqa: E402
import oracle  # noqa: E402
from coati_b200.capi import PackedPairs  # noqa: E402
from tests import util  # noqa: E402


def main():
    tables = util.load_tables()
    ctx = coati_b200.Context(0)
    rng = np.random.RandomState(1)
    checked = 0
    for k, tname in ((1, "mg_c5"), (3, "ecm_default"), (2, "mg_golden")):
        T = tables[tname]
        ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, k)
        ancs, dess, As, Bs = [], [], [], []
        while len(ancs) < 12:
            anc, des = util.random_pair(rng, int(rng.randint(1, 260)), k=k, ambiguous=len(ancs) % 3 == 0)
            anc, des = oracle.trim_end_stop(anc)[0], oracle.trim_end_stop(des)[0]
            if len(anc) % k or len(des) % k:
                continue
            a, b = oracle.encode_pair(anc, des)
            ancs.append(anc), dess.append(des), As.append(a), Bs.append(b)
        for res in (ctx.viterbi_batch(PackedPairs(As, Bs, ancs, dess)), ctx.alignpair_batch(ancs, dess)):
            for p in range(12):
                oa, ob, osc = oracle.viterbi(ancs[p], dess[p], T, k=k, enc=(As[p], Bs[p]))
                assert (res[0][p], res[1][p]) == (oa, ob) and util.f32_bits(res[2][p]) == util.f32_bits(osc)
                checked += 1
    # intra-pair wavefront (small batch, >= 2^21 cells) + burst traceback + segment-parallel expansion
    T = tables["mg_golden"]
    ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
    (_, anc), (_, des) = util.load_fasta("benchmark_2k")
    anc, des = oracle.trim_end_stop(anc)[0], oracle.trim_end_stop(des)[0]
    a, b = oracle.encode_pair(anc, des)
    ra, rb, sc = ctx.viterbi(a, b, anc, des)
    oa, ob, osc = oracle.viterbi(anc, des, T, enc=(a, b))
    assert (ra, rb) == (oa, ob) and util.f32_bits(sc) == util.f32_bits(osc)
    checked += 1
    # Forward: banded wavefront (one pair), batch (warp per pair), any-k kernel; both samplebacks
    for k in (1, 3):
        ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, k)
        anc, des = util.random_pair(rng, 60, k=k)
        anc, des = oracle.trim_end_stop(anc)[0], oracle.trim_end_stop(des)[0]
        a, b = oracle.encode_pair(anc, des)
        fw = ctx.forward(a, b)
        st = oracle.seed_state(["7"])
        rows, scs, st2, _ = fw.sampleback(anc, des, st, 12)
        rows1, scs1, _, _ = fw.sampleback(anc, des, st, 2)     # serial kernel (n < 4)
        fw.free()
        orows, oscs, ost, _ = oracle.sample(anc, des, T, st, 12, k=k)
        assert rows == orows and np.array_equal(scs.view(np.uint32), oscs.view(np.uint32)) and np.array_equal(st2, ost)
        assert rows1 == orows[:2]
        checked += 1
    ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, 1)
    ancs, dess, As, Bs = [], [], [], []
    for _ in range(20):
        anc, des = util.random_pair(rng, int(rng.randint(1, 40)), k=1)
        anc, des = oracle.trim_end_stop(anc)[0], oracle.trim_end_stop(des)[0]
        a, b = oracle.encode_pair(anc, des)
        ancs.append(anc), dess.append(des), As.append(a), Bs.append(b)
    fb = ctx.forward_batch(PackedPairs(As, Bs, ancs, dess))
    term, ll, _ = fb.terminal()
    states = np.array([oracle.seed_state([str(p)]) for p in range(20)], dtype=np.uint64)
    rows, sc, _, _ = fb.sampleback(states, 3)
    fb.free()
    for p in range(20):
        orows, osc, _, _ = oracle.sample(ancs[p], dess[p], T, states[p], 3)
        assert rows[p] == orows and np.array_equal(sc[p].view(np.uint32), osc.view(np.uint32))
        checked += 1
    ctx.close()
    print("sanitize_cases OK:", checked, "checks")


if __name__ == "__main__":
    main()
