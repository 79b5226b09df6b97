python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python scripts/time_train_pieces.py 2>&1 | tail -4
python scripts/time_train_repeat.py 2>&1 | grep -v "norm quant" | grep "grad"
