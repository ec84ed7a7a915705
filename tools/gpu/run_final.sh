mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02b_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/r02b_sanitizer_racecheck.log 2>&1; tail -n 2 gpurun_out/r02b_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r02b_sanitizer_memcheck.log 2>&1; tail -n 2 gpurun_out/r02b_sanitizer_memcheck.log
