#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -3 gpurun_out/pytest_res.log
python tools/profile_target.py --engine persistent --segments 127 --iters 6
python tools/profile_target.py --engine persistent --segments 8 --iters 4
