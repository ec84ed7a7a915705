#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the text summary committed under profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep "note" > profiles/rNN_x.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("Duration", "SM Frequency", "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread",
        "Theoretical Occupancy", "Achieved Occupancy", "DRAM Throughput", "Memory Throughput",
        "Compute (SM) Throughput", "Block Limit Shared Mem", "Block Limit Registers", "Grid Size",
        "Eligible Warps Per Scheduler", "Issued Warp Per Scheduler", "No Eligible", "L2 Cache Throughput",
        "Dynamic Shared Memory Per Block", "Waves Per SM", "Executed Instructions")
RAW = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
       "sm__inst_executed.avg.per_cycle_elapsed")


def table(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    print("#", sys.argv[2] if len(sys.argv) > 2 else "")
    print("# source:", rep, "(binary report kept out of git); per-launch numbers are cold-cache and serialised")
    rows = table(rep, "details")
    ix = {h: i for i, h in enumerate(rows[0])}
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"]:
            continue
        name = r[ix["Metric Name"]]
        if name in KEEP:
            print(f'{r[ix["Kernel Name"]][:28]:28s} | {r[ix["Section Name"]][:26]:26s} | {name:34s} | {r[ix["Metric Value"]]:>16s} {r[ix["Metric Unit"]]}')
    rows = table(rep, "raw")
    hdr, units = rows[0], rows[1]
    print("# raw metrics")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name", "")[:60])
        for k in RAW:
            if k in d:
                print(f"  {k:70s} {d[k]:>16s} {u[k]}")
        for k in hdr:
            if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued") and d[k] not in ("0", ""):
                print(f"  {k:70s} {d[k]:>16s}")


if __name__ == "__main__":
    main()
