# round-2 job 1: baseline timing of the r1h kernel + one ncu --set full capture with source (score-loop stall analysis)
python scripts/diag_tc.py > gpurun_out/r2a_diag.log 2>&1; tail -8 gpurun_out/r2a_diag.log
TAG=r2a bash scripts/run_ncu.sh
ncu -i gpurun_out/r2a_tc_encode.ncu-rep --page source --csv --print-source sass > gpurun_out/r2a_source_sass.csv 2> gpurun_out/r2a_source_err.log
ncu -i gpurun_out/r2a_tc_encode.ncu-rep --page raw --csv > gpurun_out/r2a_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
