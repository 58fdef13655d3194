set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6
