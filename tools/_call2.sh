set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -3 gpurun_out/r02_bench_2gpu.err; cat gpurun_out/r02_bench_2gpu.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/r02_bench_2gpu_strong.json 2> gpurun_out/r02_bench_2gpu_strong.err; tail -3 gpurun_out/r02_bench_2gpu_strong.err
