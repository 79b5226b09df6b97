python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kmeans or train" 2>&1 | tail -3
python scripts/diag_tc.py 2>&1 | grep "n_q=32"
python scripts/prof_train.py > gpurun_out/r2e_prof_train_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_train.csv python scripts/prof_train.py > gpurun_out/r2e_ncu_train.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2e_launches_train.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
ks=[(r[ki].split('(')[0][-60:], float(r[vi].replace(',',''))*(1e-3 if r[ui]=='ns' else 1)) for r in rows[1:]]
# last training step: find the last tc_encode and print from there
idx=[i for i,k in enumerate(ks) if 'tc_encode' in k[0]]
for k,v in ks[idx[-2]:idx[-1]]: print(f"{v:9.1f} us  {k}")
PY
