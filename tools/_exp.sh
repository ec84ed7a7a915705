python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for w in c5; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1),'kernel',round(d['roofline']['kernel_gcups'],1),'frac',round(d['roofline']['frac'],3),'tb',round(d['roofline']['traceback_ms_per_step'],2))"; done
for n in example-10k example-160k; do python tools/long_pair.py $n 2>&1 | tail -1; done
