set -x
timeout 600 python tools/check_slater32_cta.py 2>&1 | tail -8
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
