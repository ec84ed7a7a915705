mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/r02b_bench_ref_n1.json 2> gpurun_out/r02b_bench_ref_n1.err
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
tail -c 400 gpurun_out/r02b_bench_n1.err
python bench.py --workload c4 --no-extra > gpurun_out/r02b_bench_c4_n1.json 2> gpurun_out/r02b_bench_c4_n1.err
