timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
one() { timeout 300 python bench.py $2 --steps 4 --warmup 2 --no-cpu --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('$1 value %.1f e2e %.1f kernel_gcups %.1f frac %.4f fill_ms %.2f tb %.2f' % (d['value'], d['e2e']['value'], r['kernel_gcups'], r['frac'], r['kernel_ms_per_step'], r['traceback_ms_per_step']))"; }
one c5 "--pairs 250000"
one c4 "--workload c4 --pairs 30000"
one c4full "--workload c4"
