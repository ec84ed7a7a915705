set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2
timeout 1500 python -m pytest tests/test_gpu_long.py -x -q 2>&1 | tail -15 > gpurun_out/r2_long_tests.log
timeout 300 python tools/wave_exp.py > gpurun_out/r2_wave_exp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:viterbi_pipe1 -c 2 -o gpurun_out/r2_wave40k python tools/long_pair.py example-40k > gpurun_out/r2_ncu_wave.log 2>&1
cat gpurun_out/r2_long_tests.log gpurun_out/r2_wave_exp.log
