#!/bin/bash
# one gpurun call: parity tests, then timings of whatever is being worked on
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
echo "== headline: dense persistent vs sparse engine, N=127 / 64 / 8"
for n in 127 64 8; do
timeout 120 python tools/profile_target.py --engine persistent --iters 4 --segments $n
timeout 120 python tools/profile_target.py --engine sparse --iters 4 --segments $n
done
FLASHV_SPARSE_RESIDENT=0 timeout 120 python tools/profile_target.py --engine sparse --iters 4 --segments 127
