#!/bin/bash
# scratch script for the experiment at hand: static triple-buffered prefetch in the batched step kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "goldens or golden_vectors or bench_instance_flash or random_models or wide_model or headline_flash_vs or batch or level_kernel or grid_y" > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -3 gpurun_out/pytest_res.log
for N in 8 1 64 16 32; do python tools/profile_target.py --engine persistent --segments $N --iters 4; done
