set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q 2>&1 | tail -15 > gpurun_out/r2_fwd_tests.log
cat gpurun_out/r2_fwd_tests.log
timeout 300 python tools/bench_configs.py c2 > gpurun_out/r2_c2.log 2>&1
cat gpurun_out/r2_c2.log
