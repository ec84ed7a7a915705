"""CPU: the multi-GPU shard plan (coati_gpu_plan_shards: contiguous chunks, heaviest first, greedy
longest-processing-time over the shards; north_star (4)) -- a host function of the product library,
no GPU needed -- and, over gloo with world_size 2, the way bench.py uses it: every rank computes the same plan and
delivers its ranges into ONE shared host arena."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from coati_b200 import capi
from synth import synth_offsets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("npairs,shards", [(1_000_000, 8), (1_000_000, 1), (100_000, 4), (40_000, 3), (5_000, 8), (1, 2)])
def test_plan_covers_balances_and_is_deterministic(npairs, shards):
    a_off, b_off = synth_offsets(npairs, 5, 42, 0)
    first, last, shard = capi.plan_shards(a_off, b_off, shards)
    f2, l2, s2 = capi.plan_shards(a_off, b_off, shards)
    assert np.array_equal(first, f2) and np.array_equal(last, l2) and np.array_equal(shard, s2)
    # the ranges tile [0, npairs) exactly once
    order = np.argsort(first)
    assert first[order][0] == 0 and last[order][-1] == npairs
    assert np.array_equal(first[order][1:], last[order][:-1])
    assert (last > first).all() and shard.max() < shards
    # heaviest first
    cells = np.diff(a_off).astype(np.float64) * np.diff(b_off).astype(np.float64)
    csum = np.concatenate(([0.0], np.cumsum(cells)))
    cost = csum[last.astype(np.int64)] - csum[first.astype(np.int64)]
    assert np.all(np.diff(cost) <= 1e-6 * cost.max())
    assert (last - first).max() <= 1 << 17
    # chunks are sized by weight: at least a fill's worth of work each unless the batch is too small for that
    if len(first) > shards:
        assert cost.min() > 0.5 * 1.3e10
    # greedy LPT: no shard exceeds the mean by more than one chunk; on the bench workload within 4 %
    load = np.bincount(shard, weights=cost, minlength=shards)
    used = min(shards, len(first))
    assert load.max() <= load.sum() / used + cost.max() + 1e-6
    if npairs >= 100_000:
        assert load.max() / (load.sum() / shards) < 1.04
        assert shards <= len(first) <= npairs // 8192 + shards + 1


def test_plan_orders_a_sorted_batch_by_weight():
    """A batch sorted by length (worst case for contiguous sharding by count): LPT still balances it."""
    a_off, b_off = synth_offsets(200_000, 5, 42, 0)
    la, lb = np.diff(a_off), np.diff(b_off)
    idx = np.argsort(la, kind="stable")
    a2 = np.concatenate(([0], np.cumsum(la[idx]))).astype(np.uint64)
    b2 = np.concatenate(([0], np.cumsum(lb[idx]))).astype(np.uint64)
    first, last, shard = capi.plan_shards(a2, b2, 8)
    cells = (la[idx].astype(np.float64) * lb[idx].astype(np.float64))
    csum = np.concatenate(([0.0], np.cumsum(cells)))
    load = np.bincount(shard, weights=csum[last.astype(np.int64)] - csum[first.astype(np.int64)], minlength=8)
    assert load.max() / load.mean() < 1.25      # equal-count contiguous shards would be > 3x off here
    naive = np.array([cells[i * 25_000:(i + 1) * 25_000].sum() for i in range(8)])
    assert naive.max() / naive.mean() > 2.0


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from coati_b200 import capi
    from synth import synth_offsets

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N = 50_000
    a_off, b_off = synth_offsets(N, 5, 42, 0)
    first, last, shard = capi.plan_shards(a_off, b_off, world)
    name = [None]
    if rank == 0:
        name[0] = "/dev/shm/coati_test_%%d" %% os.getpid()
        with open(name[0], "wb") as f:
            f.truncate(8 * N)
    dist.broadcast_object_list(name, src=0)
    arena = np.memmap(name[0], dtype=np.uint64, mode="r+", shape=(N,))
    cells = np.diff(a_off) * np.diff(b_off)
    mine = np.flatnonzero(shard == rank)
    done = 0
    for j in mine:                       # "deliver" every pair of my ranges into the one arena
        f, l = int(first[j]), int(last[j])
        arena[f:l] = cells[f:l] + 1
        done += l - f
    arena.flush()
    t = torch.tensor([done], dtype=torch.int64)
    dist.all_reduce(t)
    dist.barrier()
    if rank == 0:
        assert int(t[0]) == N
        assert np.array_equal(np.asarray(arena), cells + 1), "a pair was not delivered (or delivered twice)"
        os.unlink(name[0])
        print("SHARDS_OK", len(first), done)
    dist.destroy_process_group()
""") % ROOT


def test_two_ranks_deliver_one_batch_into_one_arena(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SHARDS_OK" in r.stdout
