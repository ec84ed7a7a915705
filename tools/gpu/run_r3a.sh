# round 2, third session: rows written straight into page-locked arenas (rows_to_host_kernel) -- parity + A/B
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r3_direct_test.log; tail -3 gpurun_out/r3_direct_test.log
timeout 150 python bench.py --no-cpu --no-extra --steps 5 --warmup 3 > gpurun_out/r3_bench_direct.json 2> gpurun_out/r3_bench_direct.err
COATI_GPU_ROWS_CTAS=64 timeout 150 python bench.py --no-cpu --no-extra --steps 5 --warmup 3 > gpurun_out/r3_bench_direct64.json 2> gpurun_out/r3_bench_direct64.err
COATI_GPU_ROWS_DIRECT=0 timeout 150 python bench.py --no-cpu --no-extra --steps 5 --warmup 3 > gpurun_out/r3_bench_copy.json 2> gpurun_out/r3_bench_copy.err
python - <<'PY'
import json
for n in ("direct", "direct64", "copy"):
    try:
        r = json.loads(open(f"gpurun_out/r3_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(r["value"], 1), "e2e", round(r["e2e"]["value"], 1), "ms", round(r["e2e"]["ms_per_step"], 2),
              "d2h", r["e2e"]["d2h_bytes_per_step"], "h2d", r["e2e"]["h2d_bytes_per_step"])
    except Exception as ex:
        print(n, "failed", ex)
PY
