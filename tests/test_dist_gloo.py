"""CPU, world_size 2 over gloo: the N > 1 host path -- sharding of the seeded pair stream by rank and
the gather of per-rank results to rank 0 -- reproduces the single-process result.  The per-pair
"results" are produced by the oracle here (no GPU); on GPUs the same code moves the device buffers."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from coati_b200 import dist as cdist
    from synth import synth_pairs
    import oracle
    from tests import util

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N = 21
    T = util.load_tables()["mg_c5"]
    first, last = cdist.shard_range(N, rank, world)
    w = synth_pairs(last - first, 5, 42, first=first, threads=1)
    rows, scores = [], []
    for p in range(last - first):
        sa = slice(int(w["a_off"][p]), int(w["a_off"][p + 1])); sb = slice(int(w["b_off"][p]), int(w["b_off"][p + 1]))
        anc = w["anc_all"][sa].tobytes().decode(); des = w["des_all"][sb].tobytes().decode()
        a, b, sc = oracle.viterbi(anc, des, T, enc=(w["a_all"][sa], w["b_all"][sb]))
        rows.append(a + "\\0" + b + "\\0"); scores.append(sc)
    payload = {"rows": torch.frombuffer(bytearray("".join(rows).encode()), dtype=torch.uint8),
               "scores": torch.from_numpy(np.asarray(scores, np.float32).view(np.uint8).copy())}
    got = cdist.gather_to_root(payload)
    if rank == 0:
        all_rows = b"".join(bytes(g["rows"].numpy()) for g in got).decode().split("\\0")[:-1]
        all_scores = np.concatenate([g["scores"].numpy().view(np.float32) for g in got])
        # single-process answer over the whole stream
        w1 = synth_pairs(N, 5, 42, first=0, threads=1)
        for p in range(N):
            sa = slice(int(w1["a_off"][p]), int(w1["a_off"][p + 1])); sb = slice(int(w1["b_off"][p]), int(w1["b_off"][p + 1]))
            a, b, sc = oracle.viterbi(w1["anc_all"][sa].tobytes().decode(), w1["des_all"][sb].tobytes().decode(), T,
                                      enc=(w1["a_all"][sa], w1["b_all"][sb]))
            assert (all_rows[2 * p], all_rows[2 * p + 1]) == (a, b), p
            assert np.float32(all_scores[p]).tobytes() == np.float32(sc).tobytes(), p
        print("GATHER_OK", len(all_scores))
    dist.destroy_process_group()
""") % ROOT


def test_shard_ranges_cover_the_stream():
    from coati_b200.dist import shard_range
    for n, world in ((21, 2), (1_000_000, 8), (5, 8), (0, 4)):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_equals_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29573", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GATHER_OK 21" in r.stdout
