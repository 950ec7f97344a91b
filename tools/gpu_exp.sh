#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "wide_model or trellis_step" > gpurun_out/pytest_res.log 2>&1
tail -2 gpurun_out/pytest_res.log
for K in 8192 16384; do
echo "== K=$K chain table"; python tools/profile_target.py --engine persistent --segments 63 --iters 3 --K $K --T 64
echo "== K=$K slow path"; FLASHV_LACL_MAX_GB=0 python tools/profile_target.py --engine persistent --segments 63 --iters 3 --K $K --T 64
done
