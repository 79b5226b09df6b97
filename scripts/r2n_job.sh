VARIANTS="cur:-" bash scripts/run_variants.sh
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python scripts/time_train_fitted.py 2>&1 | head -2
