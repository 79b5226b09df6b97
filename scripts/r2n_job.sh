python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python scripts/time_bound_modes.py 2>&1 | tail -8
python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 1500 gpurun_out/r2n_bench.json
