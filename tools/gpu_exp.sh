#!/bin/bash
# scratch script for the experiment at hand: level kernel with its own grid barrier; FLASH-BS with one cluster barrier per step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for N in 8 1 64 127; do python tools/profile_target.py --engine persistent --segments $N --iters 4; FLASHV_LEVEL_STEPS=1 python tools/profile_target.py --engine persistent --segments $N --iters 4; done
python tools/profile_target.py --beam 128 --segments 8 --iters 3
python tools/profile_target.py --beam 128 --segments 127 --iters 3
python tools/profile_target.py --beam 32 --segments 1 --iters 3
