"""Host-side gaps (> 12 ms between two marks) in a COATI_GPU_TRACE=1 log: where the batch pipeline's host
thread was blocked.  usage: python tools/trace_gaps.py gpurun_out/trace.log"""
import re,sys
prev=None
for l in open(sys.argv[1]):
    m=re.match(r"\[coati_gpu trace\]\s+([\d.]+) ms\s+(.*)",l)
    if not m: 
        if l.startswith("nsub") or l.startswith("e2e"): print(l.strip())
        continue
    t=float(m.group(1)); what=m.group(2).strip()
    if prev and t-prev[0]>12 and "wait end" not in what and "encode done" not in what: print("GAP %.1f ms  %s  ->  %s"%(t-prev[0],prev[1],what))
    prev=(t,what)
