python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for m in 1 0; do echo "TB_SERIAL=$m"; COATI_GPU_TB_SERIAL=$m python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'kernel',round(d['roofline']['kernel_gcups'],1),'tb',round(d['roofline']['traceback_ms_per_step'],2),'exp',round(d['roofline']['compact_ms_per_step'],2))"; done
