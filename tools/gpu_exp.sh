#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -4 gpurun_out/pytest_res.log
FLASHV_PREP_TRACE=1 python tools/profile_target.py --engine persistent --segments 127 --iters 2 2>&1 | tail -8
