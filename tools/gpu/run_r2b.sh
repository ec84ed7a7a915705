mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2b_gpu_tests.log
timeout 600 python tools/bench_configs.py c3 > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err; tail -c 600 gpurun_out/r2b_c3.err
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/r2b_racecheck.log 2>&1; tail -3 gpurun_out/r2b_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r2b_memcheck.log 2>&1; tail -3 gpurun_out/r2b_memcheck.log
