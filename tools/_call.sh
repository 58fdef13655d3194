set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; tail -3 gpurun_out/r2_bench_c3.err; cat gpurun_out/r2_bench_c3.json | cut -c1-1500
python bench.py --steps 10 --warmup 3 --scaling strong --batch 256 --no-cpu-baseline > gpurun_out/r2_bench_c3_b256.json 2> gpurun_out/r2_bench_c3_b256.err; tail -3 gpurun_out/r2_bench_c3_b256.err
python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu-baseline --no-graph > gpurun_out/r2_bench_c3_b256_nograph.json 2>/dev/null
python bench.py --steps 5 --warmup 3 --config c5 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err; tail -3 gpurun_out/r2_bench_c5.err
python bench.py --steps 5 --warmup 3 --config c1 --no-cpu-baseline > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; tail -3 gpurun_out/r2_bench_c1.err
