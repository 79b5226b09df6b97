python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k degenerate 2>&1 | grep -v "^$" | tail -30
