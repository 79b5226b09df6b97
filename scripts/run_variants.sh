# time compile-time variants of the tcgen05 search on the GPU box: VARIANTS="name:DEF1,DEF2 name2:-@other_tc_file.cu" (defs comma-separated, '-' = none, @file = another revision of rvq_tc.cu in csrc/)
for v in $VARIANTS; do
  name=${v%%:*}; defs=${v#*:}; src=""; case "$defs" in *@*) src=${defs#*@}; defs=${defs%%@*};; esac
  defs=${defs//,/ }; [ "$defs" = "-" ] && defs=""
  RVQ_TC_SRC="$src" RVQ_NVCC_DEFS="$defs" python -m encodec_pytorch_b200.build --force > gpurun_out/var_${name}_build.log 2>&1 || { echo "$name: build failed"; tail -5 gpurun_out/var_${name}_build.log; continue; }
  python scripts/diag_tc.py > gpurun_out/var_${name}.log 2>&1
  echo "== $name [$defs]"; grep "n_q=32\|train variant" gpurun_out/var_${name}.log
done
python -m encodec_pytorch_b200.build --force > /dev/null 2>&1
