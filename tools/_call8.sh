set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; grep -v "^\[W" gpurun_out/r02_bench_8gpu.err | tail -3; cut -c1-200 gpurun_out/r02_bench_8gpu.json
