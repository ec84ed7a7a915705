mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02b_gpu_tests.log
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r02b_bench_n1.json') if l.startswith('{')][-1]; r=d['roofline']
print('C5 value %.1f (%.1f ms) e2e %.1f (%.1f ms) kernel %.1f frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], r['kernel_gcups'], r['frac']))"
