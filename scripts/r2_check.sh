# build a kernel variant (SRC), run the search parity tests + diag timing under a hang guard
SRC=${SRC:-rvq_tc.cu}; TAG=${TAG:-chk}
RVQ_TC_SRC=$SRC python -m encodec_pytorch_b200.build --force > gpurun_out/${TAG}_build.log 2>&1 || { tail -5 gpurun_out/${TAG}_build.log; exit 1; }
timeout 120 python scripts/diag_tc.py > gpurun_out/${TAG}_diag.log 2>&1; echo "diag rc=$?"; grep "n_q=\|train variant" gpurun_out/${TAG}_diag.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
python -m encodec_pytorch_b200.build --force > /dev/null 2>&1
