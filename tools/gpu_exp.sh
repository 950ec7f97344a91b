#!/bin/bash
# one gpurun call: parity tests, then timings of whatever is being worked on
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
echo "== headline N=127 / N=64 / N=8"
for n in 127 64 8; do timeout 120 python tools/profile_target.py --engine persistent --iters 4 --segments $n; done
echo "== config 4 shape, batch 512 / 2048 (two 8-warp CTAs per SM, then one 16-warp CTA)"
for b in 512 2048; do
timeout 600 python tools/profile_target.py --K 512 --T 1024 --batch $b --segments 32 --iters 1
FLASHV_GROUP_WARPS=16 timeout 600 python tools/profile_target.py --K 512 --T 1024 --batch $b --segments 32 --iters 1
done
