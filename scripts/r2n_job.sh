python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python scripts/time_kmeans_repeat.py 2>&1 | tail -6
