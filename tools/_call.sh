set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python tools/time_sweep.py n29_rs1 2048 50 warp 2>&1 | tail -1
python tools/time_sweep.py n13_rs10 4096 50 warp 5 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:wsweep_kernel -s 1 -c 1 -o gpurun_out/r2_w592 python tools/profile_sweep.py n29_rs1 592 50 warp > gpurun_out/r2_ncu592.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wsweep_kernel -s 1 -c 1 -o gpurun_out/r2_w2048 python tools/profile_sweep.py n29_rs1 2048 50 warp > gpurun_out/r2_ncu2048.log 2>&1
tail -3 gpurun_out/r2_ncu2048.log
ls -la gpurun_out/*.ncu-rep
