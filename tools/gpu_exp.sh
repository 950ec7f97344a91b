#!/bin/bash
# one gpurun call: parity tests, then timings of whatever is being worked on
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "bs or golden or batch" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
echo "== FLASH-BS K=3965 T=256 B=128, N=8 / N=127, then B=32 N=8"
FLASHV_BS_TRACE=1 timeout 300 python tools/profile_target.py --beam 128 --segments 8 --iters 2 2>&1 | tail -3
timeout 300 python tools/profile_target.py --beam 128 --segments 127 --iters 2
timeout 300 python tools/profile_target.py --beam 32 --segments 8 --iters 2
