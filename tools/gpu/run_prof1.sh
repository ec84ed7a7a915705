mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbi_pipe1_kernel -c 2 -f -o gpurun_out/r02b_pipe1 python bench.py --pairs 200000 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02b_ncu_pipe1.log 2>&1
tail -2 gpurun_out/r02b_ncu_pipe1.log
