set -x
python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -2 gpurun_out/r02_bench_c3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_ncu.log 2>&1
python tools/time_configs.py > gpurun_out/r02_time_configs.txt 2>&1; cat gpurun_out/r02_time_configs.txt
for c in c1 c2 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; done
