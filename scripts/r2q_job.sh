# final measurement set of round 2 on one B200: tests, smoke, bench (+ reference arm), sweep, ncu full capture of the search,
# launch lists of the bench and of the training forward
TAG=${TAG:-r2q}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python scripts/prof_encode.py > gpurun_out/${TAG}_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tc_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_tc_encode python scripts/prof_encode.py > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_tc_encode.ncu-rep --page raw --csv > gpurun_out/${TAG}_tc_encode_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_tc_encode.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_source_sass.csv 2>/dev/null
python scripts/prof_trained.py > gpurun_out/${TAG}_prof_trained_plain.log 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:tc_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_tc_encode_trained python scripts/prof_trained.py > gpurun_out/${TAG}_ncu_trained.log 2>&1
ncu -i gpurun_out/${TAG}_tc_encode_trained.ncu-rep --page raw --csv > gpurun_out/${TAG}_tc_encode_trained_ncu_raw.csv 2>/dev/null
python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_train.csv python scripts/prof_train_fitted.py > gpurun_out/${TAG}_ncu_train.log 2>&1
python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_ref.log 2>&1; echo "ref rc=$?"
python scripts/sweep.py --out gpurun_out/${TAG}_sweep.json > gpurun_out/${TAG}_sweep.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/${TAG}_sweep.log
rm -f gpurun_out/${TAG}_tc_encode_trained.ncu-rep
ls -la gpurun_out | grep ${TAG}
