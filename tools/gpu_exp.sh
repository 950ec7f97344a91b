#!/bin/bash
# one gpurun call: parity tests, then timings of whatever is being worked on
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "batch or group or config4 or golden" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
echo "== config 4 shape, batch 512 / 2368"
for b in 512 2368; do
timeout 600 python tools/profile_target.py --K 512 --T 1024 --batch $b --segments 32 --iters 1
done
P="python tools/profile_target.py --K 512 --T 96 --batch 2368 --segments 1 --iters 1"
$P && ncu --set full --clock-control none --import-source on -k regex:k_flash_group_cols -c 1 -f -o gpurun_out/prof_group $P > gpurun_out/ncu_group.log 2>&1
tail -3 gpurun_out/ncu_group.log
