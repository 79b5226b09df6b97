# GPU check of a kernel change: parity tests, timing vs the exact path, then the single-tile-phase trace of CTA 0
python -m pytest tests -m gpu -x -q > gpurun_out/chk_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/chk_tests.log
python scripts/diag_tc.py > gpurun_out/chk_diag.log 2>&1; cat gpurun_out/chk_diag.log
if [ -n "$TRACE" ]; then
RVQ_NVCC_DEFS="RVQ_TC_TRACE RVQ_TRACE_N0=36" python -m encodec_pytorch_b200.build --force >/dev/null 2>&1 && TRACE_N0=36 python scripts/trace_tc.py > gpurun_out/chk_single.log 2>&1
RVQ_NVCC_DEFS="RVQ_TC_TRACE" python -m encodec_pytorch_b200.build --force >/dev/null 2>&1 && python scripts/trace_tc.py > gpurun_out/chk_pair.log 2>&1
fi
