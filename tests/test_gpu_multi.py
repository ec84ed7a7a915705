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


def _csr(ancs, dess):
    n = len(ancs)
    a_off, b_off = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.uint64)
    np.cumsum([len(x) for x in ancs], out=a_off[1:])
    np.cumsum([len(x) for x in dess], out=b_off[1:])
    return {"a_off": a_off, "b_off": b_off,
            "anc_all": np.frombuffer(("".join(ancs) + "\0").encode("latin-1"), np.uint8).copy(),
            "des_all": np.frombuffer(("".join(dess) + "\0").encode("latin-1"), np.uint8).copy()}


def _only_used_bytes_written(w, out_len, arena, fill):
    """every slot of the arena holds its row, one NUL, and the fill byte it was given before the call"""
    a_off, b_off = w["a_off"].astype(np.int64), w["b_off"].astype(np.int64)
    n = len(a_off) - 1
    off = a_off[:-1] + b_off[:-1] + np.arange(n, dtype=np.int64)
    end = a_off[1:] + b_off[1:] + np.arange(1, n + 1, dtype=np.int64)
    ln = out_len.astype(np.int64)
    used = np.zeros(int(end[-1]) + 1, np.int32)
    np.add.at(used, off, 1)
    np.add.at(used, off + ln + 1, -1)
    used = np.cumsum(used[:-1]) > 0
    body = arena[:int(end[-1])]
    return bool((body[off + ln] == 0).all() and (body[~used] == fill).all() and (body[used] != fill).all())


def _ctx_with_env(name, value, device=0):
    import os
    import coati_b200
    os.environ[name] = value
    try:
        return coati_b200.Context(device)
    finally:
        os.environ.pop(name, None)


def test_rows_written_straight_into_page_locked_arenas(gpu_ctx, tables):
    """Row arenas that are page-locked (coati_gpu_host_alloc) can be written by the GPU itself, used bytes only
    (rows_to_host_kernel), instead of a D2H copy of the padded slots.  Same rows, lengths, score bits and status
    either way -- on a pipelined batch (several sub-batches, end stops restored), over two contexts, as shards,
    on a small batch with rejected pairs and on one holding a wavefront pair (which takes the copy).  The default
    picks the kernel for a share of a multi-device batch and the copy for a call that has the link to itself;
    COATI_GPU_ROWS_DIRECT=1 / 0 forces either."""
    import coati_b200
    from coati_b200 import capi
    from synth import synth_pairs
    g, e = np.float32(0.001), np.float32(1.0) - np.float32(1.0) / np.float32(6.0)
    T = tables["mg_c5"]
    gpu_ctx.set_model(T, g, e, 1)
    forced = _ctx_with_env("COATI_GPU_ROWS_DIRECT", "1")
    forced.set_model(T, g, e, 1)
    FILL = 0xAA

    def whole(n):
        return np.array([0], np.uint64), np.array([n], np.uint64)

    def both(w, run):
        """the batch into pageable arenas (copy, default context) and through `run` into page-locked ones"""
        n = len(w["a_off"]) - 1
        total = int(w["a_off"][-1] + w["b_off"][-1]) + n
        copied = _buffers(w, n)
        capi.alignpair_batch_ranges(gpu_ctx, w, copied, *whole(n))
        pa, pb = capi.PinnedArena(total + 1), capi.PinnedArena(total + 1)
        pa.array[:] = FILL
        pb.array[:] = FILL
        direct = (pa.array, pb.array) + _buffers(w, n)[2:]
        run(direct)
        assert np.array_equal(copied[4], direct[4]) and np.array_equal(copied[2], direct[2])
        assert np.array_equal(copied[3].view(np.uint32), direct[3].view(np.uint32))
        assert util.rows_equal(w, copied[2], np.ones(n, dtype=bool), copied[0], copied[1], direct[0], direct[1])
        tidy = (_only_used_bytes_written(w, direct[2], direct[0], FILL) and
                _only_used_bytes_written(w, direct[2], direct[1], FILL))
        pa.free(), pb.free()
        return copied, direct[2:], tidy

    # 1. pipelined (>= 65536 pairs): C5-shaped pairs, some with an end stop to trim and restore
    n = 70_000
    w = synth_pairs(n, 5, 11)
    one_stop = util.ends_with_stop(w["anc_all"], w["a_off"]) ^ util.ends_with_stop(w["des_all"], w["b_off"])
    assert int(one_stop.sum()) > 100       # the restored-stop branch (score penalty, "---" column) is exercised
    copied, res, tidy = both(w, lambda outs: capi.alignpair_batch_ranges(forced, w, outs, *whole(n)))
    assert tidy and int((res[2] != 0).sum()) == 0
    util.check_batch_properties(w, copied[0], copied[1], res[0], res[1], res[2], T, 1, g, e, oracle, exact=False)
    # ... the default on the same call: the device has the link to itself, the padded slots are copied
    _, _, tidy = both(w, lambda outs: capi.alignpair_batch_ranges(gpu_ctx, w, outs, *whole(n)))
    assert not tidy

    # 2. default, over two contexts (chunks of a shared queue, arenas addressed by global offsets) ...
    ndev = __import__("torch").cuda.device_count()
    ctxs = [coati_b200.Context(i % ndev) for i in range(2)]
    for c in ctxs:
        c.set_model(T, g, e, 1)
    _, _, tidy = both(w, lambda outs: capi.multi_alignpair_batch(ctxs, w, outs))
    assert tidy
    for c in ctxs:
        c.close()
    # ... and as three shards run one after another into the same arenas
    first, last, shard = capi.plan_shards(w["a_off"], w["b_off"], 3)

    def shards(outs):
        for s in range(3):
            capi.alignpair_batch_ranges(gpu_ctx, w, outs, first[shard == s], last[shard == s])
    _, _, tidy = both(w, shards)
    assert tidy

    # 3. small batch, every row alignment (slot offsets mod 16 vary), rows around the 512-byte strides of the
    #    kernel, every end-stop case, rejected pairs (status != 0: an empty row)
    rng = np.random.RandomState(5)
    ancs, dess = [], []
    stops = ["TAA", "TAG", "", "TGA"]
    for i, nc in enumerate([1, 2, 5, 80, 84, 85, 86, 170, 171, 172, 300, 301, 302, 303, 330] + list(rng.randint(1, 300, 60))):
        anc, des = util.random_pair(rng, int(nc), k=1, ambiguous=i % 5 == 0)
        anc, _ = oracle.trim_end_stop(anc)
        des, _ = oracle.trim_end_stop(des)
        ancs.append(anc + stops[i % 4])
        dess.append(des + stops[(i // 4) % 4])
    ancs += ["AAACCNGGG", "AAATAAGGG", "AAAC", "AAACCC"]
    dess += ["AAACCC", "AAACCC", "AAA", "AAAC?C"]
    w3 = _csr(ancs, dess)
    n3 = len(ancs)
    total = int(w3["a_off"][-1] + w3["b_off"][-1]) + n3
    pa, pb = capi.PinnedArena(total + 1), capi.PinnedArena(total + 1)
    pa.array[:] = FILL
    pb.array[:] = FILL
    direct = (pa.array, pb.array) + _buffers(w3, n3)[2:]
    capi.alignpair_batch_ranges(forced, w3, direct, *whole(n3))
    assert list(direct[4][-4:]) == [-6, -7, -5, -4] and int((direct[4][:-4] != 0).sum()) == 0
    assert _only_used_bytes_written(w3, direct[2], direct[0], FILL)
    assert _only_used_bytes_written(w3, direct[2], direct[1], FILL)
    for p in range(n3 - 4):
        at, s0 = oracle.trim_end_stop(ancs[p])
        dt, s1 = oracle.trim_end_stop(dess[p])
        oa, ob, osc = oracle.viterbi(at, dt, T, g, e, 1)
        oa, ob, osc = oracle.restore_end_stops(oa, ob, osc, (s0, s1), g, e)
        o, ln = int(w3["a_off"][p] + w3["b_off"][p]) + p, int(direct[2][p])
        assert direct[0][o:o + ln].tobytes().decode() == oa and direct[1][o:o + ln].tobytes().decode() == ob, p
        assert util.f32_bits(direct[3][p]) == util.f32_bits(osc), p
    pa.free(), pb.free()

    # 4. a batch with a wavefront pair takes the copy (one warp per pair would send a long row 512 bytes at a time)
    anc, des = util.random_pair(rng, 1100, k=1)
    w4 = _csr(ancs[:20] + [oracle.trim_end_stop(anc)[0]], dess[:20] + [oracle.trim_end_stop(des)[0]])
    _, res, tidy = both(w4, lambda outs: capi.alignpair_batch_ranges(forced, w4, outs, *whole(21)))
    assert not tidy and int((res[2] != 0).sum()) == 0
    forced.close()

    # 5. switched off: shards take the copy as well
    off = _ctx_with_env("COATI_GPU_ROWS_DIRECT", "0")
    off.set_model(T, g, e, 1)
    _, _, tidy = both(w3, lambda outs: [capi.alignpair_batch_ranges(off, w3, outs, np.array([f], np.uint64),
                                                                    np.array([l], np.uint64))
                                        for f, l in ((0, 40), (40, n3))])
    assert not tidy
    off.close()

    # the library's own count of what crossed the link
    h2d, d2h = gpu_ctx.transfer_bytes
    assert h2d > 0 and d2h > 0
