#!/usr/bin/env python
"""bench.py -- GCUPS / pairs-per-second of the marginal Gotoh Viterbi hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--pairs P] [--workload c5|c4]

A "step" is one pass of the hot path (fill + traceback + row compaction) over one batch of
synthetic codon-sequence pairs.  Default workload = BASELINE.json configs[4] (the configuration the
metric is quoted on): 1 000 000 length-binned pairs {150,300,600,1200,2400} nt, mar-mg with
omega=0.5 pi=0.25 t=0.05, k=1, seed 42 -- PER GPU (weak scaling: rank r takes pairs
[r*P, (r+1)*P) of the same seeded stream; no data-path collective, results gathered to rank 0).

value  : whole-job GCUPS with inputs resident in HBM (CUDA events on the context's stream,
         max over ranks)
e2e    : same metric through the C ABI call a user makes (coati_gpu_alignpair_batch: raw sequences
         in, aligned rows out) with pinned HOST buffers: validation + plan + H2D + encode + kernels +
         D2H inside the timed region
roofline / cpu_baseline: see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_CELL = 23          # 18 FADD + 5 FMAX of forward_impl's body (align_pair.cc:97-124)
LANES_PER_SM = 128          # FP32 lanes per SM per clock
WORKLOADS = {
    "c5": dict(id=5, k=1, table="mg_c5", pairs=1_000_000, seed=42,
               desc="BASELINE configs[4]: length-binned pairs {150,300,600,1200,2400} nt "
                    "(40/30/20/8/2 %), mar-mg w=0.5 pi=0.25 t=0.05, k=1"),
    "c4": dict(id=4, k=3, table="ecm_default", pairs=100_000, seed=20240603,
               desc="BASELINE configs[3]: 300-3000 nt pairs, mar-ecm, gap unit k=3"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the config's)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu", action="store_true")
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


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(wl, npairs_total, seconds, threads, table, first=0):
    """Time the reference's own viterbi_mem + traceback_viterbi (oracle/_ref, else the C port) on a
    bounded, length-stratified sample of the same seeded workload.  Returns dict."""
    import ctypes as C
    import oracle
    from coati_b200.capi import synth_pairs

    kind = "reference" if oracle.ref is not None else "port"
    # stratified sample: every (npairs_total // n)-th pair of the stream keeps the bin weights
    est_gcups_core = 0.045
    cells_per_pair = 338_000 if wl["id"] == 5 else 2_900_000
    n = int(max(threads, min(npairs_total, seconds * est_gcups_core * 1e9 * threads / cells_per_pair)))
    # pairs are i.i.d. draws of the seeded stream, so a contiguous block keeps the bin weights
    w = synth_pairs(n, wl["id"], wl["seed"], first)
    a_off, b_off = w["a_off"], w["b_off"]
    la, lb = np.diff(a_off), np.diff(b_off)
    a_all, anc_all, b_all, des_all = w["a_all"], w["anc_all"], w["b_all"], w["des_all"]
    total = int(a_off[-1] + b_off[-1]) + n
    out_a = np.zeros(total + 1, np.uint8)
    out_b = np.zeros(total + 1, np.uint8)
    out_len = np.zeros(n, np.uint64)
    score = np.zeros(n, np.float32)
    cells = float((la.astype(np.float64) * lb.astype(np.float64)).sum())
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    g, e = np.float32(0.001), np.float32(1.0) - np.float32(1.0) / np.float32(6.0)
    if kind == "reference":
        fn = oracle.ref.coati_ref_viterbi_batch
        fn.restype = C.c_double
        secs = fn(C.c_size_t(n), vp(a_all), vp(a_off), vp(b_all), vp(b_off), vp(anc_all), vp(des_all),
                  vp(table), C.c_float(g), C.c_float(e), C.c_size_t(wl["k"]), C.c_int(threads),
                  vp(out_a), vp(out_b), vp(out_len), vp(score))
    else:
        t0 = time.perf_counter()
        for p in range(n):
            sl_a = slice(int(a_off[p]), int(a_off[p + 1]))
            sl_b = slice(int(b_off[p]), int(b_off[p + 1]))
            oracle.viterbi(anc_all[sl_a].tobytes().decode(), des_all[sl_b].tobytes().decode(), table,
                           g, e, wl["k"], enc=(a_all[sl_a], b_all[sl_b]))
        secs = time.perf_counter() - t0
        threads = 1
    if secs <= 0:
        raise RuntimeError("CPU reference run failed")
    return dict(value=cells / secs / 1e9, unit="GCUPS", cores=threads, kind=kind, seconds=secs,
                pairs=n, pairs_per_s=n / secs, cells=cells,
                sample=f"first {n} pairs of the seeded {wl['desc'].split(':')[0]} stream (i.i.d. length bins, "
                       f"{cells:.3g} cells), viterbi_mem+traceback_viterbi, {threads} threads")


def run_reference(args, wl, table):
    rank, world, local = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    npairs = args.pairs or wl["pairs"]
    vals, last = [], None
    for _ in range(args.warmup + args.steps):
        last = cpu_reference_run(wl, npairs, max(2.0, min(args.cpu_seconds, 150.0 / (args.warmup + args.steps))),
                                 threads, table)
        vals.append(last)
    timed = vals[args.warmup:]
    secs = sum(v["seconds"] for v in timed)
    cells = sum(v["cells"] for v in timed)
    value = cells / secs / 1e9
    line = {
        "impl": "reference", "metric": "mar-mg Viterbi GCUPS (fill + traceback)", "value": value,
        "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "pairs_per_s": sum(v["pairs"] for v in timed) / secs,
        "config": {"workload": wl["desc"], "pairs_per_gpu": npairs, "k": wl["k"], "seed": wl["seed"],
                   "note": "reference CPU path on host cores; each step = bounded stratified sample"},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
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


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    table = load_table(wl["table"])
    if args.impl == "reference":
        run_reference(args, wl, table)
        return

    import torch
    import coati_b200
    from coati_b200.capi import synth_pairs

    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
        torch.cuda.set_device(local)
    npairs = args.pairs or wl["pairs"]
    g, e = np.float32(0.001), np.float32(1.0) - np.float32(1.0) / np.float32(6.0)

    # ---- synthetic batch in pinned host memory --------------------------------------------------
    pinned = []

    def alloc(nbytes):
        t = torch.empty(max(1, nbytes), dtype=torch.uint8, pin_memory=True)
        pinned.append(t)
        return t.numpy()

    t0 = time.perf_counter()
    w = synth_pairs(npairs, wl["id"], wl["seed"], first=rank * npairs, alloc=alloc)
    gen_s = time.perf_counter() - t0
    la = np.diff(w["a_off"]).astype(np.float64)
    lb = np.diff(w["b_off"]).astype(np.float64)
    cells = float((la * lb).sum())
    out_total = int(w["a_off"][-1] + w["b_off"][-1]) + npairs
    out_a, out_b = alloc(out_total + 1), alloc(out_total + 1)
    out_len = np.zeros(npairs, np.uint64)
    score = np.zeros(npairs, np.float32)
    status = np.zeros(npairs, np.int32)

    ctx = coati_b200.Context(local)           # raises if the CUDA library/device is missing
    ctx.set_model(table, g, e, wl["k"])
    info = ctx.device_info()
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident metric -----------------------------------------------------------------
    batch = ctx.batch(w["a_off"], w["b_off"])
    batch.upload(w["a_all"], w["b_all"], w["anc_all"], w["des_all"])
    gather_bytes = 0
    if dist is not None:
        # result gather to rank 0 (the path's only collective): device-resident rows + per-pair records
        from coati_b200 import dist as cdist
        bufs = batch.device_buffers()
        dev = torch.device("cuda", local)
        payload = {"out_a": cdist.DeviceBytes(bufs["out_a"], bufs["out_bytes"]).tensor(dev),
                   "out_b": cdist.DeviceBytes(bufs["out_b"], bufs["out_bytes"]).tensor(dev),
                   "results": cdist.DeviceBytes(bufs["results"], bufs["result_bytes"]).tensor(dev)}
        recv_cache = {}
        gather_bytes = 2 * bufs["out_bytes"] + bufs["result_bytes"]

    def run_step():
        batch.run()
        if dist is not None:
            with torch.cuda.stream(stream):
                cdist.gather_to_root(payload, 0, recv_cache)

    for _ in range(args.warmup):
        run_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    fill_ms = trace_ms = compact_ms = 0.0
    for _ in range(args.steps):
        run_step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches - launches0
    tm = batch.timing()                        # per-family device time of the LAST step
    fill_ms, trace_ms, compact_ms = tm["fill_ms"], tm["traceback_ms"], tm["compact_ms"]
    stats = batch.stats()
    # check the batch ran clean before reporting anything
    batch.download(out_a, out_b, out_len, score, status)
    assert int((status != 0).sum()) == 0, "pairs failed"
    assert int(out_len.min()) > 0

    # ---- end-to-end through the public C ABI with host buffers -----------------------------------
    def e2e_once():
        # the call a user makes: raw sequences in, aligned rows + scores out (marg_alignment semantics:
        # length checks, end-stop trim/restore, encoding on the device, Viterbi, traceback)
        ctx._check(ctx.lib.coati_gpu_alignpair_batch(
            ctx.h, npairs, w["anc_all"].ctypes.data, w["a_off"].ctypes.data_as(coati_b200.capi._u64p),
            w["des_all"].ctypes.data, w["b_off"].ctypes.data_as(coati_b200.capi._u64p),
            out_a.ctypes.data, out_b.ctypes.data, out_len.ctypes.data_as(coati_b200.capi._u64p),
            score.ctypes.data_as(coati_b200.capi._fp), status.ctypes.data_as(coati_b200.capi._i32p)))

    batch.destroy()
    for _ in range(max(1, args.warmup)):         # warm-up (pool allocations, page faults)
        e2e_once()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, args.steps)
    for _ in range(e2e_steps):
        e2e_once()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    assert int((status != 0).sum()) == 0 and int(out_len.min()) > 0, "e2e run failed"
    h2d = int(w["a_off"][-1] + w["b_off"][-1])
    d2h = int(2 * out_total + npairs * 32)

    # ---- reduce over ranks ------------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([ms, e2e_s, fill_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, fill_ms_max = t.tolist()
        c = torch.tensor([cells, float(npairs), float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cells_all, pairs_all, launches_all = c.tolist()
    else:
        cells_all, pairs_all, launches_all, fill_ms_max = cells, float(npairs), float(launches), fill_ms
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    secs = ms / 1e3
    value = cells_all * args.steps / secs / 1e9
    e2e_val = cells_all * e2e_steps / e2e_s / 1e9
    # roofline of the dominant kernel (viterbi_pipe_kernel): FP32 issue, SURVEY 8(d)
    sm_mhz = clocks.get("sm_mhz") or info["clock_khz"] / 1e3
    peak_tflops = info["sm_count"] * LANES_PER_SM * sm_mhz * 1e6 / 1e12
    peak_tflops_max = info["sm_count"] * LANES_PER_SM * (clocks.get("sm_max_mhz") or info["clock_khz"] / 1e3) * 1e6 / 1e12
    ach_tflops = cells * FLOP_PER_CELL / (fill_ms / 1e3) / 1e12 if fill_ms > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # DRAM bytes of the fill launches per step, from the committed ncu capture (bytes per cell there x cells here)
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_traffic.json")) as fh:
            tj = json.load(fh).get(args.workload, {})
        if "bytes_per_cell" in tj:
            traffic = tj["bytes_per_cell"] * cells  # bytes per step, all fill launches of one GPU
            traffic_note = ("DRAM bytes of the fill launches per step and GPU: %.4f B/cell measured by ncu --set full "
                            "(profiles/r01_traffic.json) x cells; algorithmic decision stream 0.625 B/cell"
                            % tj["bytes_per_cell"])
    except (OSError, ValueError):
        pass
    line = {
        "metric": "mar-mg Viterbi GCUPS (fill + traceback)", "value": value, "unit": "GCUPS",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "pairs_per_s": pairs_all * args.steps / secs,
        "config": {"workload": wl["desc"], "pairs_per_gpu": npairs, "k": wl["k"], "seed": wl["seed"],
                   "cells_per_gpu": cells, "l2": "inputs + decision stream >> 126 MB L2 (no flush needed)",
                   "decision_stream_bytes_per_gpu": stats["dir_bytes"], "chunks": stats["chunks"],
                   "gen_seconds": gen_s,
                   "nccl_gather_bytes_per_rank_per_step": gather_bytes},
        "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / e2e_steps, "pairs_per_s": pairs_all * e2e_steps / e2e_s,
                "steps": e2e_steps},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": {"bound": "fp32_issue", "kernel": "viterbi_pipe_kernel", "achieved": ach_tflops,
                     "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach_tflops / peak_tflops if peak_tflops else None,
                     "peak_at_max_clock": peak_tflops_max, "flop_per_cell": FLOP_PER_CELL,
                     "kernel_ms_per_step": fill_ms, "kernel_gcups": cells / (fill_ms / 1e3) / 1e9 if fill_ms else None,
                     "traffic": traffic, "traffic_note": traffic_note,
                     "peak_source": f"{info['sm_count']} SMs x 128 FP32 lanes x median SM clock under load",
                     "hbm_stream": {"achieved_gbs": stats["dir_bytes"] / (fill_ms / 1e3) / 1e9 if fill_ms else None,
                                    "peak_gbs": hbm_peak, "of": "measured" if peaks else "fallback"},
                     "traceback_ms_per_step": trace_ms, "compact_ms_per_step": compact_ms},
    }
    if world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_run(
                wl, npairs, args.cpu_seconds, os.cpu_count() or 1, table).items()
                if k in ("value", "unit", "cores", "kind", "sample", "pairs_per_s")}
        except Exception as ex:  # the checker failing must not hide the measurement
            line["cpu_baseline"] = {"error": repr(ex)}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
