set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/r02_bench_c4.json 2>gpurun_out/c4.err; tail -c 600 gpurun_out/r02_bench_c4.json
timeout 300 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/r02_bench_c5.json 2>gpurun_out/c5.err; tail -c 600 gpurun_out/r02_bench_c5.json
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3.json 2>gpurun_out/c3.err; tail -c 900 gpurun_out/r02_bench_c3.json
(timeout 300 python tools/time_ewald.py n29_rs1 256 2048 8192 32768; timeout 300 python tools/time_ewald.py n57_rs0.25 64 512 2048; timeout 300 python tools/time_ewald.py n13_rs10 256 4096) > gpurun_out/r02_time_ewald.txt 2>&1
timeout 600 python tools/time_configs.py > gpurun_out/r02_time_configs.txt 2>&1; tail -13 gpurun_out/r02_time_configs.txt
