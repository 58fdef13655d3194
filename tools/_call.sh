set -x
timeout 900 python tools/stress_spec.py 60 2>&1 | tail -3
for i in 1 2 3; do timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1; done
