mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "multi_context" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -5 gpurun_out/r2_bench_n$N.err; tail -1 gpurun_out/r2_bench_n$N.json
