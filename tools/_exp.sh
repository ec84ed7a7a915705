python bench.py --steps 4 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'kernel',round(d['roofline']['kernel_gcups'],1),'kernel ms',round(d['roofline']['kernel_ms_per_step'],1))"
for r in 4 8 10; do for n in example-10k example-40k example-160k; do echo "wave R=$r $n"; COATI_GPU_WAVE_R=$r python tools/long_pair.py $n 2>&1 | tail -1; done; done
