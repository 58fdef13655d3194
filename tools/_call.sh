set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/time_sweep.py n29_rs1 2048 50 warp 2>&1 | tail -1
python tools/time_sweep.py n29_rs1 8192 50 warp 5 2>&1 | tail -1
python tools/time_sweep.py n13_rs10 4096 50 warp 5 2>&1 | tail -1
python tools/time_sweep.py n29_rs1 256 50 auto 2>&1 | tail -1
CGB_LIB=build/variants/libcgb200_prof.so python tools/phase_clocks.py n29_rs1 2048 50 warp
