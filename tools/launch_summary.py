#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python tools/launch_summary.py gpurun_out/launches.csv "note" > profiles/rNN_bench_launches.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
ix = {h: i for i, h in enumerate(rows[0])}
tot = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    name = re.sub(r"^void |coati_gpu::", "", name)
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += float(r[ix["Metric Value"]]) / 1e6
total = sum(v[1] for v in tot.values())
print("#", sys.argv[2] if len(sys.argv) > 2 else "")
print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.")
print(f'{"kernel":70s} {"launches":>8s} {"total_ms":>12s} {"share":>7s}')
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {n:8d} {ms:12.3f} {100 * ms / total:6.1f}%")
print(f'{"TOTAL":70s} {sum(v[0] for v in tot.values()):8d} {total:12.3f}')
