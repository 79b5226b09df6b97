# profile artefacts of the round on one B200 (TAG names the files under gpurun_out/): launch list of a short bench run, one
# ncu --set full capture of the fused encode at cfg2, launch list of a training forward, event trace of CTA 0
TAG=${TAG:-rX}
python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
TAG=$TAG bash scripts/run_ncu.sh
python scripts/prof_train.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_train.csv python scripts/prof_train.py > gpurun_out/${TAG}_ncu_train.log 2>&1; echo "train launch list rc=$?"
RVQ_NVCC_DEFS="RVQ_TC_TRACE" python -m encodec_pytorch_b200.build --force >/dev/null 2>&1 && python scripts/trace_tc.py > gpurun_out/${TAG}_trace_pair.log 2>&1
RVQ_NVCC_DEFS="RVQ_TC_TRACE RVQ_TRACE_N0=36" python -m encodec_pytorch_b200.build --force >/dev/null 2>&1 && TRACE_N0=36 python scripts/trace_tc.py > gpurun_out/${TAG}_trace_single.log 2>&1
python -m encodec_pytorch_b200.build --force >/dev/null 2>&1
echo done
