#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "golden or random_models or headline_flash_vs or error" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for n in 127 64; do
timeout 120 python tools/profile_target.py --engine sparse --iters 6 --segments $n
done
FLASHV_SPARSE_RESIDENT=0 timeout 120 python tools/profile_target.py --engine sparse --iters 6 --segments 127
