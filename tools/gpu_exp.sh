#!/bin/bash
# scratch script for the experiment at hand (one gpurun call): the GPU suite, smoke, and the default decodes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -3 gpurun_out/pytest_res.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for N in 1 8 64 127; do python tools/profile_target.py --engine persistent --segments $N --iters 4; done
