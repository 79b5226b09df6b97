# profiles of the final round-2 kernel: ncu full capture with source, raw metrics, launch list of the bench, tests
TAG=${TAG:-r2i}
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python scripts/prof_encode.py > gpurun_out/${TAG}_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tc_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_tc_encode python scripts/prof_encode.py > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_tc_encode.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_source_sass.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_tc_encode.ncu-rep --page raw --csv > gpurun_out/${TAG}_tc_encode_ncu_raw.csv 2>/dev/null
python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_ncu_bench.log 2>&1
ls -la gpurun_out | grep ${TAG}
