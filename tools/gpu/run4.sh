set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.json
