#!/bin/bash
# scratch script: full GPU suite + bench on the half-precision filter kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -5 gpurun_out/pytest_res.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err
python tools/summarize_bench.py gpurun_out/bench_a.log | head -30
