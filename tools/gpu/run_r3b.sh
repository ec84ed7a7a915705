# round 2, third session: N-GPU bench with the rows written by rows_to_host_kernel (the default for shards)
mkdir -p gpurun_out
N=${1:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/r3_bench_n$N.err | tail -1 > gpurun_out/r3_bench_n$N.json
python -c "
import json
d=json.loads(open('gpurun_out/r3_bench_n$N.json').read().strip().splitlines()[-1]); print('N$N value %.1f (%.1f ms) e2e %.1f (%.1f ms) by rank %s d2h %s %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_by_rank'], d['e2e']['d2h_bytes_per_step'], d['e2e']['row_delivery']))"
