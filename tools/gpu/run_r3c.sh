# round 2, third session: last verification inside what is left of the GPU budget (files in order of relevance to
# the change; -v so that a cut-off run still says what passed)
mkdir -p gpurun_out
timeout 112 python -m pytest tests/test_gpu_multi.py tests/test_gpu_viterbi.py tests/test_gpu_cli.py tests/test_gpu_forward.py tests/test_gpu_long.py -x -v -m gpu -p no:cacheprovider > gpurun_out/r3_gpu_tests.log 2>&1
echo "pytest rc $?" >> gpurun_out/r3_gpu_tests.log
grep -c PASSED gpurun_out/r3_gpu_tests.log; tail -2 gpurun_out/r3_gpu_tests.log
timeout 25 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
