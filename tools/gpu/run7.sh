mkdir -p gpurun_out
echo two fill streams; TRACE=0 REPS=3 python tools/e2e_trace.py 8 16 31 62 2>/dev/null
echo one fill stream; COATI_GPU_ONE_FILL=1 TRACE=0 REPS=3 python tools/e2e_trace.py 16 31 2>/dev/null
