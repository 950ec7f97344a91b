#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "bs or golden_vectors or batch_equals" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for d in 0 1; do
echo "== FLASHV_BS_DENSE=$d"
FLASHV_BS_DENSE=$d timeout 300 python tools/profile_target.py --beam 128 --segments 8 --iters 3
FLASHV_BS_DENSE=$d timeout 300 python tools/profile_target.py --beam 128 --segments 127 --iters 3
FLASHV_BS_DENSE=$d timeout 300 python tools/profile_target.py --beam 32 --segments 8 --iters 3
done
