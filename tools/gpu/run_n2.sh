mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/r02b_bench_n2.err | tail -1 > gpurun_out/r02b_bench_n2.json
python -c "
import json
d=json.loads(open('gpurun_out/r02b_bench_n2.json').read().strip().splitlines()[-1]); print('N2 value %.1f (%.1f ms) e2e %.1f (%.1f ms) by rank %s gather %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_by_rank'], d.get('run', d['config'])['nccl_gather_bytes_to_root_per_step']))"
