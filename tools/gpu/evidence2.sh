# final evidence of round 2 (names r02b_*): tests, bench arms, launch list, ncu captures, sanitizer
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02b_gpu_tests.log
python bench.py --impl reference > gpurun_out/r02b_bench_ref_n1.json 2> gpurun_out/r02b_bench_ref_n1.err
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
python bench.py --workload c4 --no-extra > gpurun_out/r02b_bench_c4_n1.json 2> gpurun_out/r02b_bench_c4_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02b_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_pipe1_kernel -c 2 -f -o gpurun_out/r02b_pipe1 python bench.py --pairs 200000 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02b_ncu_pipe1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_pipe3_kernel -c 1 -f -o gpurun_out/r02b_pipe3 python bench.py --workload c4 --pairs 30000 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02b_ncu_pipe3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:viterbi_wave1 -c 1 -f -o gpurun_out/r02b_wave40k python tools/long_pair.py example-40k > gpurun_out/r02b_ncu_wave40k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_wave1 -c 1 -f -o gpurun_out/r02b_wave160k python tools/long_pair.py example-160k > gpurun_out/r02b_ncu_wave160k.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r02b_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/r02b_sanitizer_racecheck.log 2>&1
tail -3 gpurun_out/r02b_sanitizer_*.log
