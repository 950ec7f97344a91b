#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "bs or golden_vectors" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
FLASHV_BS_TRACE=1 timeout 300 python tools/profile_target.py --beam 128 --segments 8 --iters 3 2>&1 | tail -3
timeout 300 python tools/profile_target.py --beam 128 --segments 127 --iters 3
timeout 300 python tools/profile_target.py --beam 32 --segments 8 --iters 3
