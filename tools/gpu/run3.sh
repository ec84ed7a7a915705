set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forward_band -c 1 -o gpurun_out/r2_fwdband python tools/bench_configs.py c2 > gpurun_out/r2_ncu_fwd.log 2>&1
tail -3 gpurun_out/r2_ncu_fwd.log
