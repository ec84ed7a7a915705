#!/usr/bin/env python
"""bench.py -- GCUPS / pairs-per-second of the marginal Gotoh Viterbi hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--pairs P] [--workload c5|c4]
                    [--scaling strong|weak] [--no-cpu] [--no-extra]

A "step" is one pass of the hot path (fill + traceback + row expansion) over ONE batch of synthetic
codon-sequence pairs.  Default workload = BASELINE.json configs[4] (the configuration the metric is quoted on):
1 000 000 length-binned pairs {150,300,600,1200,2400} nt, mar-mg with omega=0.5 pi=0.25 t=0.05, k=1, seed 42.

Multi-GPU (one process per GPU under torchrun), default `--scaling strong` = the config as written: the ONE
batch lives in one shared, page-locked host arena; coati_gpu_plan_shards cuts it into contiguous chunks and
gives them to the ranks by greedy longest-processing-time on the sum of La * Lb; no data-path collective.
`--scaling weak` gives every rank its own P pairs (the round-1 definition).

value  : whole-job GCUPS with inputs resident in HBM (CUDA events on the context's stream, max over ranks);
         at N > 1 every step ends with the NCCL gather of all rows and result records to rank 0's device,
         issued from a double-buffered staging copy on a side stream so it overlaps the next step's fill
e2e    : same metric through the C ABI call a user makes (coati_gpu_alignpair_batch_ranges: raw sequences
         in, aligned rows out) with pinned HOST arenas: validation + plan + H2D + encode + kernels + D2H inside
         the timed region; at N > 1 every rank writes its pairs' rows into the one shared output arena
roofline / cpu_baseline / extra: see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import uuid

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_CELL = 23          # 18 FADD + 5 FMAX of forward_impl's body (align_pair.cc:97-124)
LANES_PER_SM = 128          # FP32 lanes per SM per clock
WORKLOADS = {
    "c5": dict(id=5, k=1, table="mg_c5", pairs=1_000_000, seed=42,
               model=dict(model="mar-mg", br_len=0.05, omega=0.5, pi=(0.25, 0.25, 0.25, 0.25)),
               kernel="viterbi_pipe1_kernel<10,4>",
               desc="BASELINE configs[4]: length-binned pairs {150,300,600,1200,2400} nt "
                    "(40/30/20/8/2 %), mar-mg w=0.5 pi=0.25 t=0.05, k=1"),
    "c4": dict(id=4, k=3, table="ecm_default", pairs=100_000, seed=20240603,
               model=dict(model="mar-ecm"), kernel="viterbi_pipe3_kernel<6,4>",
               desc="BASELINE configs[3]: 300-3000 nt pairs, mar-ecm, gap unit k=3"),
}
G_OPEN, G_EXT = np.float32(0.001), np.float32(1.0) - np.float32(1.0) / np.float32(6.0)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs of the batch (strong) / per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1-C4 / pageable records")
    return ap.parse_args()


def load_table(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", "tables.npz"))
    return np.ascontiguousarray(z[name], dtype=np.float32)


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, reasons, mx, power = [], set(), None, []
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                clocks.append(float(p[1]))
                mx = float(p[2])
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if clocks:
            out.update(sm_mhz=float(np.median(clocks)), sm_max_mhz=mx, reasons=sorted(reasons),
                       samples=len(clocks), power_w_max=max(power) if power else None)
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def emit(line):
    """The JSON line is the LAST line of stdout: flush whatever C libraries (NCCL's version banner) still
    hold in their stdio buffers first."""
    sys.stdout.flush()
    try:
        import ctypes
        ctypes.CDLL(None).fflush(None)
    except Exception:
        pass
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# CPU legs: the reference's own viterbi_mem + traceback_viterbi (oracle/_ref, else the C port), ONE PROCESS PER
# CORE, each on its own contiguous block of the same seeded pair stream.
def cpu_reference_run(wl, npairs_total, seconds, procs, table, first=0, lib=None, flags="-O3 -DNDEBUG"):
    from tools import cpu_worker
    est_gcups_core = 0.05
    cells_per_pair = 338_000 if wl["id"] == 5 else 2_900_000
    n = int(max(procs, min(npairs_total, seconds * est_gcups_core * 1e9 * procs / cells_per_pair)))
    res = cpu_worker.run_pool(wl["id"], wl["seed"], first, n, table, float(G_OPEN), float(G_EXT), wl["k"], procs, lib)
    cells = sum(r["cells"] for r in res)
    secs = max(r["seconds"] for r in res)         # all workers start together: the slowest one is the wall time
    kind = res[0]["kind"]
    return dict(value=cells / secs / 1e9, unit="GCUPS", cores=procs, kind=kind, seconds=secs, pairs=n,
                pairs_per_s=n / secs, cells=cells,
                sample=f"first {n} pairs of the seeded {wl['desc'].split(':')[0]} stream (i.i.d. length bins, "
                       f"{cells:.3g} cells), viterbi_mem+traceback_viterbi, {procs} processes x 1 thread, {flags}")


def workload_config(wl, npairs):
    """`config` of the JSON line: what was measured, identical in both arms (the driver compares them); what is
    particular to a run -- shards, arenas, chunk counts -- goes into `run`."""
    return {"workload": wl["desc"], "pairs": npairs, "k": wl["k"], "seed": wl["seed"],
            "l2": "GPU arm: inputs + decision stream of a step >> 126 MB L2 (no flush needed)"}


def run_reference(args, wl, table):
    rank, world, local = dist_env()
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    npairs = args.pairs or wl["pairs"]
    vals, last = [], None
    for _ in range(args.warmup + args.steps):
        last = cpu_reference_run(wl, npairs, max(2.0, min(args.cpu_seconds, 150.0 / (args.warmup + args.steps))),
                                 procs, table)
        vals.append(last)
    timed = vals[args.warmup:]
    secs = sum(v["seconds"] for v in timed)
    cells = sum(v["cells"] for v in timed)
    value = cells / secs / 1e9
    line = {
        "impl": "reference", "metric": "mar-mg Viterbi GCUPS (fill + traceback)", "value": value,
        "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "pairs_per_s": sum(v["pairs"] for v in timed) / secs,
        "config": workload_config(wl, npairs),
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU path on host cores; each step = bounded sample of the workload",
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
class Arena:
    """One block of page-locked host memory carved into named uint8 arrays.  world == 1: cudaHostAlloc through
    the library; world > 1: one file in /dev/shm mapped by every rank and registered with CUDA, so that all ranks
    read the one input batch from, and deliver their rows to, the same physical memory."""

    def __init__(self, sizes, world, rank, dist, lib):
        self.lib, self.world, self.rank, self.path, self.pinned = lib, world, rank, None, True
        self.off, total = {}, 0
        for name, nbytes in sizes:
            self.off[name] = (total, nbytes)
            total += (nbytes + 4095) & ~4095
        self.total = max(total, 4096)
        self.shared = world > 1 or os.environ.get("COATI_BENCH_SHM") == "1"   # (the env: A/B of the arena kind at N = 1)
        if not self.shared:
            from coati_b200.capi import PinnedArena
            self.block = PinnedArena(self.total)
            self.buf = self.block.array
        else:
            import ctypes as C
            name = [None]
            if rank == 0:
                name[0] = "/dev/shm/coati_bench_%s" % uuid.uuid4().hex
                with open(name[0], "wb") as f:
                    f.truncate(self.total)
            if dist is not None:
                dist.broadcast_object_list(name, src=0)
            self.path = name[0]
            self.buf = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(self.total,))
            self.pinned = lib.coati_gpu_host_register(C.c_void_p(self.buf.ctypes.data), self.total) == 0

    def view(self, name, dtype=np.uint8):
        o, n = self.off[name]
        return self.buf[o:o + n].view(dtype)

    def close(self, dist):
        if not self.shared:
            self.buf = None
            self.block.free()
            return
        import ctypes as C
        if self.pinned:
            self.lib.coati_gpu_host_unregister(C.c_void_p(self.buf.ctypes.data))
        if dist is not None:
            dist.barrier()
        if self.rank == 0:
            os.unlink(self.path)


def _fixed_alloc(bufs):
    """alloc callback for synth_pairs that hands out the given arrays in order (a_all, b_all, anc_all, des_all)."""
    it = iter(bufs)

    def alloc(nbytes):
        b = next(it)
        assert len(b) >= nbytes
        return b
    return alloc


def _compact_local(w, firsts, lasts):
    """The pairs of the given ranges as one local CSR batch (for the device-resident metric)."""
    a_off, b_off = w["a_off"], w["b_off"]
    z = [np.zeros(0, np.uint64)]
    la = np.concatenate([np.diff(a_off[int(f):int(l) + 1]) for f, l in zip(firsts, lasts)] or z)
    lb = np.concatenate([np.diff(b_off[int(f):int(l) + 1]) for f, l in zip(firsts, lasts)] or z)
    lo_a = np.zeros(len(la) + 1, np.uint64)
    lo_b = np.zeros(len(lb) + 1, np.uint64)
    np.cumsum(la, out=lo_a[1:])
    np.cumsum(lb, out=lo_b[1:])

    def cat(name, off):
        parts = [w[name][int(off[int(f)]):int(off[int(l)])] for f, l in zip(firsts, lasts)]
        return np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0, np.uint8)
    return dict(a_off=lo_a, b_off=lo_b, a_all=cat("a_all", a_off), b_all=cat("b_all", b_off),
                anc_all=cat("anc_all", a_off), des_all=cat("des_all", b_off))


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, load_table(wl["table"]))
        return

    import torch
    import coati_b200
    from coati_b200 import capi
    from synth import synth_offsets, synth_pairs

    rank, world, local = dist_env()
    if os.environ.get("COATI_TRACE_DIR"):   # diagnostics: every rank's stderr (COATI_GPU_TRACE timeline) to its own file
        os.makedirs(os.environ["COATI_TRACE_DIR"], exist_ok=True)
        fd = os.open(os.path.join(os.environ["COATI_TRACE_DIR"], "rank%d.log" % rank), os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
        os.dup2(fd, 2)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    strong = args.scaling == "strong"
    npairs_arg = args.pairs or wl["pairs"]
    npairs = npairs_arg if strong else npairs_arg * world      # pairs of the whole job
    # the model: the product's own table builder (coati::set_subst), checked against the committed table
    table = capi.host_marginal_table(**wl["model"])
    want = load_table(wl["table"])
    assert np.allclose(table, want, rtol=3e-5, atol=3e-6), "host table builder disagrees with tests/golden/tables.npz"

    ctx = coati_b200.Context(local)           # raises if the CUDA library/device is missing
    ctx.set_model(table, G_OPEN, G_EXT, wl["k"])
    info = ctx.device_info()
    lib = ctx.lib
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    dev = torch.device("cuda", local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the one batch, in one page-locked arena ---------------------------------------------------------
    t0 = time.perf_counter()
    a_off, b_off = synth_offsets(npairs, wl["id"], wl["seed"], 0)
    ta, tb = int(a_off[-1]), int(b_off[-1])
    out_total = ta + tb + npairs
    arena = Arena([("anc_all", ta + 1), ("des_all", tb + 1), ("a_all", ta + 1), ("b_all", tb + 1),
                   ("out_a", out_total + 1), ("out_b", out_total + 1), ("out_len", 8 * npairs),
                   ("score", 4 * npairs), ("status", 4 * npairs)], world, rank, dist, lib)
    w = dict(a_off=a_off, b_off=b_off, anc_all=arena.view("anc_all"), des_all=arena.view("des_all"),
             a_all=arena.view("a_all"), b_all=arena.view("b_all"))
    if rank == 0:
        synth_pairs(npairs, wl["id"], wl["seed"], 0,
                    alloc=_fixed_alloc([w["a_all"], w["b_all"], w["anc_all"], w["des_all"]]))
    out_a, out_b = arena.view("out_a"), arena.view("out_b")
    out_len, score = arena.view("out_len", np.uint64), arena.view("score", np.float32)
    status = arena.view("status", np.int32)
    barrier()
    gen_s = time.perf_counter() - t0

    # ---- shards: contiguous chunks, heaviest first, greedy LPT over the ranks ------------------------------
    if strong:
        r_first, r_last, r_shard = capi.plan_shards(a_off, b_off, world)
    else:  # weak: rank r owns pairs [r * P, (r + 1) * P)
        r_first = np.arange(world, dtype=np.uint64) * np.uint64(npairs_arg)
        r_last = r_first + np.uint64(npairs_arg)
        r_shard = np.arange(world, dtype=np.uint32)
    mine = np.flatnonzero(r_shard == rank)
    my_first, my_last = r_first[mine], r_last[mine]
    order = np.argsort(my_first)
    cells_pair = np.diff(a_off).astype(np.float64) * np.diff(b_off).astype(np.float64)
    cells_total = float(cells_pair.sum())
    my_cells = float(sum(cells_pair[int(f):int(l)].sum() for f, l in zip(my_first, my_last)))

    # ---- device-resident metric: the rank's pairs as one local CSR batch ----------------------------------
    loc = _compact_local(w, my_first[order], my_last[order])
    batch = ctx.batch(loc["a_off"], loc["b_off"])
    batch.upload(loc["a_all"], loc["b_all"], loc["anc_all"], loc["des_all"])
    gather_bytes = 0
    gather = None
    if dist is not None:
        from coati_b200 import dist as cdist
        bufs = batch.device_buffers()
        payload = {"out_a": cdist.DeviceBytes(bufs["out_a"], bufs["out_bytes"]).tensor(dev),
                   "out_b": cdist.DeviceBytes(bufs["out_b"], bufs["out_bytes"]).tensor(dev),
                   "results": cdist.DeviceBytes(bufs["results"], bufs["result_bytes"]).tensor(dev)}
        gather = cdist.OverlappedGather(payload, stream, root=0)
        gather_bytes = gather.bytes_to_root

    fill_ms = trace_ms = compact_ms = 0.0

    def run_step(timed):
        nonlocal fill_ms, trace_ms, compact_ms
        batch.run()
        if gather is not None:
            gather.step()
        if timed:  # per-family device time of THIS step (synchronises the batch's stream, not the gather's)
            tm = batch.timing()
            fill_ms += tm["fill_ms"]
            trace_ms += tm["traceback_ms"]
            compact_ms += tm["compact_ms"]

    for _ in range(args.warmup):
        run_step(False)
    if gather is not None:
        gather.finish()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        run_step(True)
    if gather is not None:
        gather.finish()                        # the last step's gather is inside the timed region
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches - launches0
    fill_ms, trace_ms, compact_ms = fill_ms / args.steps, trace_ms / args.steps, compact_ms / args.steps
    stats = batch.stats()
    # check the batch ran clean before reporting anything
    n_loc = len(loc["a_off"]) - 1
    tot_loc = int(loc["a_off"][-1] + loc["b_off"][-1]) + n_loc + 1
    chk = (np.zeros(tot_loc, np.uint8), np.zeros(tot_loc, np.uint8), np.zeros(n_loc, np.uint64),
           np.zeros(n_loc, np.float32), np.zeros(n_loc, np.int32))
    batch.download(*chk)
    assert int((chk[4] != 0).sum()) == 0, "pairs failed"
    assert n_loc == 0 or int(chk[2].min()) > 0
    if gather is not None and rank == 0:
        # the root holds every rank's result records (32 B: term[3], score, len, start, status, pad): all clean
        torch.cuda.synchronize()
        got_pairs = 0
        for r, bufsr in enumerate(gather.root_buffers()):
            rec = bufsr["results"].cpu().numpy().view(np.int32).reshape(-1, 8)
            assert int((rec[:, 6] != 0).sum()) == 0 and (len(rec) == 0 or int(rec[:, 4].min()) > 0), f"rank {r} records"
            got_pairs += len(rec)
        assert got_pairs == npairs, "gather incomplete"
    batch.destroy()
    del loc, chk

    # ---- end-to-end through the public C ABI with host buffers ---------------------------------------------
    outs = (out_a, out_b, out_len, score, status)

    def e2e_once():
        # the call a user makes: raw sequences in, aligned rows + scores out (marg_alignment semantics:
        # length checks, end-stop trim/restore, encoding on the device, Viterbi, traceback)
        capi.alignpair_batch_ranges(ctx, w, outs, my_first, my_last)

    for _ in range(max(1, args.warmup)):         # warm-up (pool allocations, page faults)
        e2e_once()
    barrier()
    moved0 = ctx.transfer_bytes
    t0 = time.perf_counter()
    e2e_steps = max(1, args.steps)
    for _ in range(e2e_steps):
        e2e_once()
    torch.cuda.synchronize()
    e2e_own_s = time.perf_counter() - t0       # this rank alone; the reported time also waits for the slowest rank
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    if rank == 0:  # every rank's rows are in the one arena
        assert int((status != 0).sum()) == 0 and int(out_len.min()) > 0, "e2e run failed"
    # bytes per step as the library counted them: symbols + pair descriptors in; result records and rows out (the
    # rows are written by the expansion kernel straight into the page-locked arena, length + terminator each;
    # with COATI_GPU_ROWS_DIRECT=0 or a pageable arena it is a copy of the padded slots, 2 (La + Lb + 1) per pair)
    moved1 = ctx.transfer_bytes
    h2d, d2h = ((moved1[i] - moved0[i]) // e2e_steps for i in (0, 1))

    # ---- reduce over ranks ------------------------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()
        c = torch.tensor([float(launches), float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        launches_all, h2d_all, d2h_all = c.tolist()
        per_rank = torch.zeros(2, world, dtype=torch.float64, device="cuda")
        per_rank[0, rank] = my_cells
        per_rank[1, rank] = 1e3 * e2e_own_s / e2e_steps
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        shard_cells, e2e_rank_ms = per_rank[0].tolist(), per_rank[1].tolist()
    else:
        launches_all, h2d_all, d2h_all, shard_cells = float(launches), float(h2d), float(d2h), [my_cells]
        e2e_rank_ms = [1e3 * e2e_own_s / e2e_steps]

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        extra = run_extras(ctx, capi, w, npairs, cells_total, outs, e2e_steps)
    arena_pinned, arena_shared = arena.pinned, arena.shared
    del w, outs, out_a, out_b, out_len, score, status
    arena.close(dist)
    if rank != 0:
        dist.destroy_process_group()
        return

    secs = ms / 1e3
    value = cells_total * args.steps / secs / 1e9
    e2e_val = cells_total * e2e_steps / e2e_s / 1e9
    # roofline of the dominant kernel: FP32 issue; rank 0's shard, its own fill time
    sm_mhz = clocks.get("sm_mhz") or info["clock_khz"] / 1e3
    peak_tflops = info["sm_count"] * LANES_PER_SM * sm_mhz * 1e6 / 1e12
    peak_tflops_max = info["sm_count"] * LANES_PER_SM * (clocks.get("sm_max_mhz") or info["clock_khz"] / 1e3) * 1e6 / 1e12
    ach_tflops = my_cells * FLOP_PER_CELL / (fill_ms / 1e3) / 1e12 if fill_ms > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh).get(args.workload, {})
        if "bytes_per_cell" in tj:
            traffic = tj["bytes_per_cell"] * my_cells  # bytes per step, all fill launches of one GPU
            traffic_note = ("DRAM bytes of the fill launches per step on rank 0: %.4f B/cell measured by ncu --set full "
                            "(%s) x cells of the shard; algorithmic decision stream 0.625 B/cell"
                            % (tj["bytes_per_cell"], tj.get("source", "profiles/")))
    except (OSError, ValueError):
        pass
    line = {
        "metric": "mar-mg Viterbi GCUPS (fill + traceback)", "value": value, "unit": "GCUPS",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "pairs_per_s": npairs * args.steps / secs,
        "config": workload_config(wl, npairs),
        "run": {"cells": cells_total,
                "sharding": ("one batch; contiguous chunks, heaviest first, greedy LPT on sum(La*Lb) over ranks "
                             "(coati_gpu_plan_shards)" if strong else "rank r owns pairs [r*P, (r+1)*P)"),
                "chunks": int(len(r_first)),
                "shard_cells_max_over_mean": max(shard_cells) / (sum(shard_cells) / world),
                "decision_stream_bytes_rank0": stats["dir_bytes"], "dir_chunks_rank0": stats["chunks"],
                "gen_seconds": gen_s,
                "host_arena": "/dev/shm + cudaHostRegister" if arena_shared else "cudaHostAlloc",
                "host_arena_pinned": bool(arena_pinned),
                "collective": None if dist is None else ("NCCL send/recv gather of rows + result records to rank 0, "
                                                         "double-buffered staging copy, side stream"),
                "nccl_gather_bytes_to_root_per_step": gather_bytes},
        "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                "ms_per_step": 1e3 * e2e_s / e2e_steps, "pairs_per_s": npairs * e2e_steps / e2e_s,
                "steps": e2e_steps, "ms_per_step_by_rank": [round(x, 2) for x in e2e_rank_ms],
                "api": "coati_gpu_alignpair_batch_ranges, rows delivered to one host arena",
                # the library's default: the kernel for a share of a multi-device batch, the copy otherwise
                "row_delivery": ("rows_to_host_kernel writes the used bytes of every row into the page-locked arena"
                                 if arena_pinned and (os.environ.get("COATI_GPU_ROWS_DIRECT") == "1" or
                                                      (os.environ.get("COATI_GPU_ROWS_DIRECT") is None and world > 1))
                                 else "D2H copy of the padded slots")},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": {"bound": "fp32_issue", "kernel": wl["kernel"], "achieved": ach_tflops,
                     "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach_tflops / peak_tflops if peak_tflops else None,
                     "peak_at_max_clock": peak_tflops_max, "flop_per_cell": FLOP_PER_CELL,
                     "kernel_ms_per_step": fill_ms, "kernel_ms_note": "mean over the timed steps, rank 0's shard",
                     "kernel_gcups": my_cells / (fill_ms / 1e3) / 1e9 if fill_ms else None,
                     "traffic": traffic, "traffic_note": traffic_note,
                     "peak_source": f"{info['sm_count']} SMs x 128 FP32 lanes x median SM clock under load",
                     "hbm_stream": {"achieved_gbs": stats["dir_bytes"] / (fill_ms / 1e3) / 1e9 if fill_ms else None,
                                    "peak_gbs": hbm_peak, "of": "measured" if peaks else "fallback"},
                     "traceback_ms_per_step": trace_ms, "compact_ms_per_step": compact_ms},
    }
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_run(
                wl, npairs, args.cpu_seconds, os.cpu_count() or 1, want).items()
                if k in ("value", "unit", "cores", "kind", "sample", "pairs_per_s")}
        except Exception as ex:  # the checker failing must not hide the measurement
            line["cpu_baseline"] = {"error": repr(ex)}
        o2g = os.path.join(ROOT, "oracle", "_ref", "libcoati_ref_o2g.so")
        if os.path.exists(o2g) and not args.no_extra:
            try:  # once: the reference as its DEFAULT Meson build compiles it
                r = cpu_reference_run(wl, npairs, min(4.0, args.cpu_seconds), os.cpu_count() or 1, want, lib=o2g,
                                      flags="-O2 -g, assertions on (Meson default buildtype=debugoptimized)")
                line.setdefault("extra", {})["cpu_baseline_meson_default_build"] = {
                    k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:
                line.setdefault("extra", {})["cpu_baseline_meson_default_build"] = {"error": repr(ex)}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def run_extras(ctx, capi, w, npairs, cells_total, outs, e2e_steps):
    """Driver-visible records beside the headline (N = 1 only): e2e from pageable host memory, and BASELINE
    configs 1-4 (tools/bench_configs.py)."""
    extra = {}
    # (1) the same e2e call from PAGEABLE caller memory (what a C++ caller holding std::string / std::vector gets
    #     unless it allocates with coati_gpu_host_alloc or registers its buffers)
    try:
        wp = dict(a_off=w["a_off"], b_off=w["b_off"], anc_all=np.array(w["anc_all"]), des_all=np.array(w["des_all"]))
        po = (np.zeros(len(outs[0]), np.uint8), np.zeros(len(outs[1]), np.uint8), np.zeros(npairs, np.uint64),
              np.zeros(npairs, np.float32), np.zeros(npairs, np.int32))
        f = np.array([0], np.uint64)
        l = np.array([npairs], np.uint64)
        capi.alignpair_batch_ranges(ctx, wp, po, f, l)
        t0 = time.perf_counter()
        steps = min(2, e2e_steps)
        for _ in range(steps):
            capi.alignpair_batch_ranges(ctx, wp, po, f, l)
        dt = time.perf_counter() - t0
        assert int((po[4] != 0).sum()) == 0
        extra["e2e_pageable"] = {"value": cells_total * steps / dt / 1e9, "unit": "GCUPS", "ms_per_step": 1e3 * dt / steps,
                                 "note": "same call, caller arenas in pageable memory (numpy): copies are staged by the driver"}
        del wp, po
    except Exception as ex:
        extra["e2e_pageable"] = {"error": repr(ex)}
    # (2) BASELINE configs 1-4
    try:
        from tools import bench_configs
        extra["configs"] = bench_configs.collect(ctx)
    except Exception as ex:
        extra["configs"] = {"error": repr(ex)}
    return extra


if __name__ == "__main__":
    main()
