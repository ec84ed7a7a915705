mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:viterbi_pipe1 -c 1 -o gpurun_out/r2_wave40k_new python tools/long_pair.py example-40k > gpurun_out/r2_ncu_wave_new.log 2>&1
tail -2 gpurun_out/r2_ncu_wave_new.log
