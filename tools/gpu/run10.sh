mkdir -p gpurun_out
for P in 500000 125000; do
python bench.py --steps 4 --warmup 2 --no-extra --no-cpu --pairs $P 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('pairs',d['config']['pairs'],'chunks',d['config']['chunks'],'value',round(d['value']),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),'ms',round(d['e2e']['ms_per_step'],1))"
done
N=125000 COATI_GPU_TRACE=1 python tools/e2e_trace.py 2> gpurun_out/r2_trace125k.log | tail -1
grep -A200 "traced call" gpurun_out/r2_trace125k.log | grep -E "plan begin|plan end|wait begin|wait end|sub@|download enq|upload|run enq" | head -60
