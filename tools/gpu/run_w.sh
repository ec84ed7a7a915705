timeout 1200 python -m pytest tests/test_gpu_long.py -x -q -m gpu 2>&1 | tail -2
timeout 400 python tools/bench_configs.py c3 > gpurun_out/r02b_c3.json 2> gpurun_out/r02b_c3.err
python - <<'P'
import json
for line in open('gpurun_out/r02b_c3.json'):
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    for k,v in d.get('c3_long_single_pairs_mar-mg',{}).items():
        print(k, v['la'], 'e2e %.2f fill %.2f gcups %.1f frac %.3f tb %.2f parity %s' % (v['gpu_e2e_ms'], v['fill_ms'], v['fill_gcups'], v['roofline_frac'], v['traceback_ms'], v.get('parity')))
P
