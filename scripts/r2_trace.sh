# event traces of CTA 0 for a given kernel source (SRC), pair regime (steps 2..7) and the lone third tile (steps 36..41)
SRC=${SRC:-rvq_tc.cu}; TAG=${TAG:-tr}
RVQ_TC_SRC=$SRC RVQ_NVCC_DEFS="RVQ_TC_TRACE" python -m encodec_pytorch_b200.build --force && python scripts/trace_tc.py > gpurun_out/${TAG}_pair.log 2>&1
RVQ_TC_SRC=$SRC RVQ_NVCC_DEFS="RVQ_TC_TRACE RVQ_TRACE_N0=36" python -m encodec_pytorch_b200.build --force && TRACE_N0=36 python scripts/trace_tc.py > gpurun_out/${TAG}_single.log 2>&1
python -m encodec_pytorch_b200.build --force > /dev/null
