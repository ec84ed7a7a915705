mkdir -p gpurun_out
COATI_GPU_NSUB=31 python tools/e2e_trace.py 2> gpurun_out/r2_trace31.log | tail -2
grep -A400 "traced call" gpurun_out/r2_trace31.log | grep -E "plan begin|plan end|wait begin|wait end|sub@|download enq" | head -120
