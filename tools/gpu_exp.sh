#!/bin/bash
# scratch script for the experiment at hand
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "golden or random_models or headline_flash_vs or streaming" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for i in 1 2; do timeout 120 python tools/profile_target.py --engine sparse --iters 8 --segments 127; done
