# like run_variants.sh, plus the fitted-codebook timing (scripts/time_trained.py)
for v in $VARIANTS; do
  name=${v%%:*}; defs=${v#*:}; src=""; case "$defs" in *@*) src=${defs#*@}; defs=${defs%%@*};; esac
  defs=${defs//,/ }; [ "$defs" = "-" ] && defs=""
  RVQ_TC_SRC="$src" RVQ_NVCC_DEFS="$defs" python -m encodec_pytorch_b200.build --force > gpurun_out/var_${name}_build.log 2>&1 || { echo "$name: build failed"; tail -5 gpurun_out/var_${name}_build.log; continue; }
  timeout 120 python scripts/diag_tc.py > gpurun_out/var_${name}.log 2>&1
  echo "== $name [$defs]"; grep "n_q=32" gpurun_out/var_${name}.log
  timeout 120 python scripts/time_trained.py 2>/dev/null | tail -2
done
python -m encodec_pytorch_b200.build --force > /dev/null 2>&1
