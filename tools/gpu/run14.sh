mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_long.py -x -q -k "wavefront or long_pairs_bit" 2>&1 | tail -3
timeout 300 python tools/wave_exp.py 2>&1 | tee gpurun_out/r2_wave_exp2.log
