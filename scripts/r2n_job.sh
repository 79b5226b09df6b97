python scripts/time_decode.py 2>&1 | tail -2
git stash -q 2>/dev/null
