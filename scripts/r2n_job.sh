python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ties or more_than_32" 2>&1 | tail -15
