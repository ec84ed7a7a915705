python -m pytest tests -x -q -m gpu 2>&1 | tail -2
COATI_GPU_PIPE_SCALAR=1 python -m pytest tests/test_gpu_viterbi.py -x -q -k "random_batch or direction or edge" 2>&1 | tail -1
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'kernel',round(d['roofline']['kernel_gcups'],1),'frac',round(d['roofline']['frac'],3))"
