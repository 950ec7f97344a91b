#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for n in 127 127 64 8; do
timeout 120 python tools/profile_target.py --engine persistent --iters 6 --segments $n
done
timeout 120 python tools/profile_target.py --engine sparse --iters 6 --segments 127
