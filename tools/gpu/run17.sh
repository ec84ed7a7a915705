mkdir -p gpurun_out
for SHM in 0 1; do
COATI_BENCH_SHM=$SHM COATI_GPU_TRACE=1 python bench.py --pairs 125000 --steps 3 --warmup 3 --no-extra --no-cpu 2> gpurun_out/r2_trace_shm$SHM.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('shm',d['config']['host_arena'],'value ms',round(d['ms_per_step'],1),'e2e ms',round(d['e2e']['ms_per_step'],1))"
grep -E "plan begin|plan end|wait begin|wait end|sub@" gpurun_out/r2_trace_shm$SHM.log | tail -22
done
