mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_viterbi.py -x -q 2>&1 | tail -3
for P in 1000000 125000; do
python bench.py --steps 4 --warmup 2 --no-extra --no-cpu --pairs $P 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('pairs',d['config']['pairs'],'chunks',d['config']['chunks'],'value',round(d['value']),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),'ms',round(d['e2e']['ms_per_step'],1), 'tb',round(d['roofline']['traceback_ms_per_step'],2),'exp',round(d['roofline']['compact_ms_per_step'],2))"
done
