set -x
RVQ_NVCC_DEFS="RVQ_TC_TRACE" python -m encodec_pytorch_b200.build --force && python scripts/trace_tc.py > gpurun_out/tr1_pair.log 2>&1
RVQ_NVCC_DEFS="RVQ_TC_TRACE RVQ_TRACE_N0=36" python -m encodec_pytorch_b200.build --force && TRACE_N0=36 python scripts/trace_tc.py > gpurun_out/tr1_single.log 2>&1
RVQ_NVCC_DEFS="RVQ_TC_TIMERS" python -m encodec_pytorch_b200.build --force && python scripts/diag_tc.py > gpurun_out/tr1_diag.log 2>&1
tail -3 gpurun_out/tr1_diag.log
