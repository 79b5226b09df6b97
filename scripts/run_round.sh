# round-end measurement set on one B200 (TAG names the artefacts under gpurun_out/)
TAG=${TAG:-rX}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_ref.log 2>&1; echo "ref rc=$?"
python scripts/pcie_probe.py > gpurun_out/${TAG}_pcie.log 2>&1; cat gpurun_out/${TAG}_pcie.log
python scripts/sweep.py --out gpurun_out/${TAG}_sweep.json > gpurun_out/${TAG}_sweep.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/${TAG}_sweep.log
