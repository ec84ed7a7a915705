mkdir -p gpurun_out
N=8
COATI_TRACE_DIR=gpurun_out/trace8 COATI_GPU_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),'ms',round(d['e2e']['ms_per_step'],1), d['e2e']['ms_per_step_by_rank'])"
for r in 0 5; do echo "== rank $r"; grep -E "plan begin|plan end|wait begin|wait end|sub@" gpurun_out/trace8/rank$r.log | tail -20; done
