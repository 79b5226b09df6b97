# multi-GPU measurement set of round 2: N = number of GPUs of this box (gpurun --gpus N)
N=${N:-2}; TAG=${TAG:-r2f}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/${TAG}_topo_${N}gpu.log 2>&1; lscpu | grep -i "numa\|model name\|socket\|^CPU(s)" >> gpurun_out/${TAG}_topo_${N}gpu.log 2>&1; nproc >> gpurun_out/${TAG}_topo_${N}gpu.log
torchrun_probe() { $TR --master-port 29551 scripts/pcie_probe.py; }
torchrun_probe > gpurun_out/${TAG}_pcie_${N}gpu.log 2>&1
$TR --master-port 29521 scripts/ddp_check.py > gpurun_out/${TAG}_ddp_check_${N}gpu.json 2> gpurun_out/${TAG}_ddp_check_${N}gpu.err; echo "ddp_check rc=$?"; cat gpurun_out/${TAG}_ddp_check_${N}gpu.json
$TR --master-port 29533 scripts/train_scaling.py --out gpurun_out/${TAG}_train_${N}gpu.json > /dev/null 2> gpurun_out/${TAG}_train_${N}gpu.err; echo "train rc=$?"; cat gpurun_out/${TAG}_train_${N}gpu.json
$TR --master-port 29541 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'binding', j.get('host_binding'))
PY
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_model_dropin.py -x -q -m gpu -k ddp 2>&1 | tail -2; fi
