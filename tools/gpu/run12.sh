mkdir -p gpurun_out
COATI_GPU_TRACE=1 python bench.py --pairs 125000 --steps 2 --warmup 3 --no-extra --no-cpu 2> gpurun_out/r2_trace_b125k.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('pairs',d['config']['pairs'],'chunks',d['config']['chunks'],'value ms',round(d['ms_per_step'],1),'e2e ms',round(d['e2e']['ms_per_step'],1))"
grep -E "plan begin|plan end|wait begin|wait end|sub@|download enq|run enq" gpurun_out/r2_trace_b125k.log | tail -75
