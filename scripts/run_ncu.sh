# one ncu --set full capture of tc_encode_kernel at cfg2 (after a plain run of the same program exited 0)
TAG=${TAG:-x}
python scripts/prof_encode.py > gpurun_out/${TAG}_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tc_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_tc_encode python scripts/prof_encode.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
