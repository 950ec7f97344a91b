#!/bin/bash
# scratch script for the experiment at hand (one gpurun call): parity subset, then timings of the default decodes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "goldens or golden_vectors or bench_instance_flash or random_models or level_kernel or batch_equals or headline_flash_vs" > gpurun_out/pytest_res.log 2>&1
tail -2 gpurun_out/pytest_res.log
for N in 1 8 64 127; do python tools/profile_target.py --engine persistent --segments $N --iters 4; done
