#!/bin/bash
# scratch script for the experiment at hand: the whole GPU suite, smoke() and a short bench against HEAD
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1; nproc >> gpurun_out/gpus.txt; free -g >> gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"
tail -c 600 gpurun_out/bench.err
python tools/summarize_bench.py gpurun_out/bench.log 2>/dev/null | head -60
