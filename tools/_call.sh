set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 120 python tools/time_sweep.py n29_rs1 2048 50 warp 2>&1 | tail -1
timeout 120 python tools/time_sweep.py n29_rs1 8192 50 warp 5 2>&1 | tail -1
timeout 120 python tools/time_sweep.py n57_rs0.25 512 50 auto 5 2>&1 | tail -1
timeout 120 python tools/time_sweep.py n29_rs1 256 50 auto 2>&1 | tail -1
