mkdir -p gpurun_out
(echo "# tools/wave_lag.py 40000 2 4 8 10: fill time of lattices with lb = 39 999 and 1920 / 3840 / 7680 / 15360 rows; slope over the band count = growth per band, intercept = lone-warp step (round 2 final kernel: eight-step groups at R <= 8)"; timeout 250 python tools/wave_lag.py 40000 2 4 8 10 2>&1 | grep -E "^R ") > gpurun_out/r02b_wave_lag.txt
(echo "# tools/wave_exp.py: fill / traceback ms per forced R (round 2 final kernel)"; timeout 300 python tools/wave_exp.py example-10k example-20k example-40k example-80k example-160k 2>&1 | grep "fill_ms") > gpurun_out/r02b_wave_exp.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:viterbi_wave1 -c 1 -f -o gpurun_out/r02b_wave40k python tools/long_pair.py example-40k > gpurun_out/r02b_ncu_wave40k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:viterbi_wave1 -c 1 -f -o gpurun_out/r02b_wave10k python tools/long_pair.py example-10k > gpurun_out/r02b_ncu_wave10k.log 2>&1
grep ": t_step" gpurun_out/r02b_wave_lag.txt
