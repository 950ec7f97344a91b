#!/bin/bash
# scratch script for the experiment at hand: cp.async staging in the level / step kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "goldens or golden_vectors or bench_instance_flash or random_models or wide_model or headline_flash_vs or batch or level_kernel" > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -3 gpurun_out/pytest_res.log
for N in 8 1 64 127; do python tools/profile_target.py --engine persistent --segments $N --iters 4; FLASHV_LEVEL_STEPS=1 python tools/profile_target.py --engine persistent --segments $N --iters 4; done
