set -x
mkdir -p gpurun_out
# (1) launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02_bench_under_ncu.log 2>&1
# (2) K = 1 fill, C5 at 200 k pairs: the R = 10 launch and the R = 8 launch of the device-resident pass
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_pipe1_kernel -c 2 -o gpurun_out/r02_pipe1 python bench.py --pairs 200000 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02_ncu_pipe1.log 2>&1
# (3) K = 3 fill, C4 at 30 k pairs
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_pipe3_kernel -c 1 -o gpurun_out/r02_pipe3 python bench.py --workload c4 --pairs 30000 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02_ncu_pipe3.log 2>&1
# (4) Forward: banded wavefront on C2 and the batch kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forward_band -c 3 -o gpurun_out/r02_forward python tools/bench_configs.py c2 fwd > gpurun_out/r02_ncu_forward.log 2>&1
# (5) wavefront at 40k (R = 4) and 160k (R = 10)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:viterbi_pipe1 -c 1 -o gpurun_out/r02_wave40k python tools/long_pair.py example-40k > gpurun_out/r02_ncu_wave40k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_pipe1 -c 1 -o gpurun_out/r02_wave160k python tools/long_pair.py example-160k > gpurun_out/r02_ncu_wave160k.log 2>&1
# (6) compute-sanitizer
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_cases.py > gpurun_out/r02_sanitizer_synccheck.log 2>&1
tail -4 gpurun_out/r02_sanitizer_*.log
ls -la gpurun_out/r02_*
