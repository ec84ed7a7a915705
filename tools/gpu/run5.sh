mkdir -p gpurun_out
TRACE=0 REPS=3 python tools/e2e_trace.py 8 16 31 62 2>/dev/null | tee gpurun_out/r2_e2e_nsub.log
python tools/e2e_trace.py 2> gpurun_out/r2_trace.log | tail -2
grep -A400 "traced call" gpurun_out/r2_trace.log | head -150
