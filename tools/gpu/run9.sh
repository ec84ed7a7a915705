mkdir -p gpurun_out
COATI_BENCH_SHM=1 python bench.py --steps 4 --warmup 2 --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('shm arena: value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['config']['host_arena'])"
python bench.py --steps 4 --warmup 2 --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('cudaHostAlloc: value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['config']['host_arena'])"
