N=8
for NR in 0 1; do
COATI_DIAG_NO_ROWS=$NR timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('norows=$NR N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),'ms',round(d['e2e']['ms_per_step'],1), d['e2e']['ms_per_step_by_rank'])"
done
