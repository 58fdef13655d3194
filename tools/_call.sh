python tools/_dbg2.py 2>&1 | tail -4
python -m pytest tests/test_gpu_rng.py -m gpu -q -x 2>&1 | tail -3
