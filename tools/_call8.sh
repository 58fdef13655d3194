set -x
P=29620
for c in c4 c5; do
  P=$((P+1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 10 --warmup 3 --config $c > gpurun_out/r02_bench_8gpu_$c.json 2> gpurun_out/r02_bench_8gpu_$c.err
  cut -c1-200 gpurun_out/r02_bench_8gpu_$c.json
done
