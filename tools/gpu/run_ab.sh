one() { timeout 300 python bench.py $2 --steps 5 --warmup 2 --no-cpu --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('$1 value %.1f (%.1f ms) e2e %.1f (%.1f ms) kernel_gcups %.1f frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], r['kernel_gcups'], r['frac']))"; }
one r120_a ""; one r120_b ""
export COATI_GPU_LIB=$PWD/tools/gpu/ab_r128.so
one r128_a ""; one r128_b ""
