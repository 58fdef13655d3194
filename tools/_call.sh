set -x
timeout 300 python tools/time_sweep.py n13_rs10 512 50 auto 5 2>&1 | tail -1
timeout 300 python tools/time_sweep.py n13_rs10 64 50 auto 5 2>&1 | tail -1
timeout 300 python tools/time_sweep.py n29_rs1 64 50 auto 5 2>&1 | tail -1
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --config c1 --steps 10 --warmup 3 > gpurun_out/r02_bench_c1.json 2>/dev/null
