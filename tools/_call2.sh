set -x
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/r02_bench_2gpu_strong.json 2> gpurun_out/r02_bench_2gpu_strong.err; grep -v "^\[W" gpurun_out/r02_bench_2gpu_strong.err | tail -5; cut -c1-300 gpurun_out/r02_bench_2gpu_strong.json
