"""GPU (-m gpu): one batch over several contexts (coati_gpu_multi_alignpair_batch: shared queue of weight-balanced
chunks, one host thread per context) and over the shards of coati_gpu_plan_shards (coati_gpu_alignpair_batch_ranges),
against the single-device call and the oracle.  With one GPU the contexts share device 0 (the queue, the threads
and the arena addressing are the same); with more, every context gets its own device."""
import numpy as np
import pytest

import oracle
from tests import util

pytestmark = pytest.mark.gpu


def _buffers(w, n):
    total = int(w["a_off"][-1] + w["b_off"][-1]) + n
    return (np.zeros(total + 1, np.uint8), np.zeros(total + 1, np.uint8), np.zeros(n, np.uint64),
            np.zeros(n, np.float32), np.zeros(n, np.int32))


@pytest.mark.parametrize("n_ctx", [2, 3])
def test_multi_context_batch_equals_single_call(n_ctx, tables):
    import torch
    import coati_b200
    from coati_b200 import capi
    from synth import synth_pairs
    n = 60_000
    g, e = np.float32(0.001), np.float32(1.0) - np.float32(1.0) / np.float32(6.0)
    T = tables["mg_c5"]
    w = synth_pairs(n, 5, 42)
    ndev = torch.cuda.device_count()
    ctxs = [coati_b200.Context(i % ndev) for i in range(n_ctx)]
    for c in ctxs:
        c.set_model(T, g, e, 1)
    one = _buffers(w, n)
    ctxs[0]._check(ctxs[0].lib.coati_gpu_alignpair_batch(
        ctxs[0].h, n, w["anc_all"].ctypes.data, w["a_off"].ctypes.data_as(capi._u64p), w["des_all"].ctypes.data,
        w["b_off"].ctypes.data_as(capi._u64p), one[0].ctypes.data, one[1].ctypes.data,
        one[2].ctypes.data_as(capi._u64p), one[3].ctypes.data_as(capi._fp), one[4].ctypes.data_as(capi._i32p)))
    many = _buffers(w, n)
    capi.multi_alignpair_batch(ctxs, w, many)
    assert np.array_equal(one[4], many[4]) and int((many[4] != 0).sum()) == 0
    assert np.array_equal(one[2], many[2])
    assert np.array_equal(one[3].view(np.uint32), many[3].view(np.uint32))
    assert util.rows_equal(w, one[2], np.ones(n, dtype=bool), one[0], one[1], many[0], many[1])
    # every context took part
    assert all(c.launches > 0 for c in ctxs)
    # and the answer is right: properties of every alignment + a sample against the oracle (raw entry point:
    # pairs whose descendant ends in a stop codon are trimmed and restored, so compare those without)
    plain = ~(util.ends_with_stop(w["anc_all"], w["a_off"]) | util.ends_with_stop(w["des_all"], w["b_off"]))
    rng = np.random.RandomState(3)
    for p in rng.choice(np.flatnonzero(plain), size=24, replace=False):
        sa = slice(int(w["a_off"][p]), int(w["a_off"][p + 1]))
        sb = slice(int(w["b_off"][p]), int(w["b_off"][p + 1]))
        oa, ob, osc = oracle.viterbi(w["anc_all"][sa].tobytes().decode(), w["des_all"][sb].tobytes().decode(), T, g, e, 1,
                                     enc=(w["a_all"][sa], w["b_all"][sb]))
        off = int(w["a_off"][p] + w["b_off"][p]) + int(p)
        ln = int(many[2][p])
        assert many[0][off:off + ln].tobytes().decode() == oa and many[1][off:off + ln].tobytes().decode() == ob
        assert util.f32_bits(many[3][p]) == util.f32_bits(osc)
    # mismatched models are refused
    ctxs[1].set_model(T, g, e, 3)
    with pytest.raises(coati_b200.CoatiGpuError):
        capi.multi_alignpair_batch(ctxs, w, many)
    for c in ctxs:
        c.close()


def test_shard_ranges_deliver_the_whole_batch(gpu_ctx, tables):
    """Three shards of coati_gpu_plan_shards, run one after another through coati_gpu_alignpair_batch_ranges into
    the same arenas (what the ranks of bench.py do concurrently), equal the single call pair for pair; pairs
    outside a call's ranges are left untouched."""
    from coati_b200 import capi
    from synth import synth_pairs
    n = 50_000
    g, e = np.float32(0.001), np.float32(1.0) - np.float32(1.0) / np.float32(6.0)
    w = synth_pairs(n, 5, 7)
    gpu_ctx.set_model(tables["mg_c5"], g, e, 1)
    first, last, shard = capi.plan_shards(w["a_off"], w["b_off"], 3)
    one = _buffers(w, n)
    capi.alignpair_batch_ranges(gpu_ctx, w, one, np.array([0], np.uint64), np.array([n], np.uint64))
    assert int((one[4] != 0).sum()) == 0
    parts = _buffers(w, n)
    parts[4][:] = 99      # sentinel: untouched pairs keep it
    for s in range(3):
        sel = shard == s
        capi.alignpair_batch_ranges(gpu_ctx, w, parts, first[sel], last[sel])
        done = np.zeros(n, dtype=bool)
        for s2 in range(s + 1):
            for f, l in zip(first[shard == s2], last[shard == s2]):
                done[int(f):int(l)] = True
        assert (parts[4][done] == 0).all() and (parts[4][~done] == 99).all()
    assert np.array_equal(one[2], parts[2]) and np.array_equal(one[3].view(np.uint32), parts[3].view(np.uint32))
    assert util.rows_equal(w, one[2], np.ones(n, dtype=bool), one[0], one[1], parts[0], parts[1])
    # empty range list and an invalid range
    import coati_b200
    capi.alignpair_batch_ranges(gpu_ctx, w, parts, np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    with pytest.raises(coati_b200.CoatiGpuError):
        capi.alignpair_batch_ranges(gpu_ctx, w, parts, np.array([5], np.uint64), np.array([n + 1], np.uint64))


def test_pinned_arena_round_trip(gpu_ctx, tables):
    """coati_gpu_host_alloc / _free and _register / _unregister: arenas a C++ caller would use."""
    import ctypes as C
    from coati_b200 import capi
    arena = capi.PinnedArena(1 << 20)
    arena.array[:] = 7
    assert int(arena.array.sum()) == 7 << 20
    arena.free()
    buf = np.zeros(1 << 20, np.uint8)
    lib = gpu_ctx.lib
    assert lib.coati_gpu_host_register(C.c_void_p(buf.ctypes.data), buf.nbytes) == 0
    assert lib.coati_gpu_host_unregister(C.c_void_p(buf.ctypes.data)) == 0
    assert lib.coati_gpu_host_register(None, 16) == -2


@pytest.mark.parametrize("model,kw,k", [("mar-mg", dict(br_len=0.05, omega=0.5, pi=(0.25, 0.25, 0.25, 0.25)), 1),
                                        ("mar-ecm", {}, 3), ("mar-mg", {}, 1)])
def test_parity_on_the_products_own_table(model, kw, k, gpu_ctx):
    """The tables of every other DP test come from oracle/table.py; here the 183 x 15 table is the one the product's
    own builder makes (coati::set_subst -> mg94_p / ecm_p -> marginal_p, csrc/host/coati_host.cc): the same bytes go
    to the GPU and to the oracle, and rows and score bits must agree -- the DP parity contract does not depend on
    who built the table."""
    from coati_b200 import capi
    from coati_b200.capi import PackedPairs
    T = capi.host_marginal_table(model, **kw)
    rng = np.random.RandomState(77 + k)
    ancs, dess, As, Bs = [], [], [], []
    while len(ancs) < 48:
        anc, des = util.random_pair(rng, int(rng.randint(1, 250)), k=k, ambiguous=len(ancs) % 4 == 0)
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        if len(anc) % k or len(des) % k:
            continue
        ea, eb = oracle.encode_pair(anc, des)
        ancs.append(anc), dess.append(des), As.append(ea), Bs.append(eb)
    gpu_ctx.set_model(T, oracle.DEFAULT_G, oracle.DEFAULT_E, k)
    rows_a, rows_b, score, status = gpu_ctx.viterbi_batch(PackedPairs(As, Bs, ancs, dess))
    assert (status == 0).all()
    for p in range(48):
        oa, ob, osc = oracle.viterbi(ancs[p], dess[p], T, k=k, enc=(As[p], Bs[p]))
        assert (rows_a[p], rows_b[p]) == (oa, ob), p
        assert util.f32_bits(score[p]) == util.f32_bits(osc), p
