#!/bin/bash
# One gpurun call: GPU parity tests, then short bench runs of both engines.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_persistent_n64.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --segments 8 --no-cpu > gpurun_out/bench_persistent_n8.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --engine step --no-cpu > gpurun_out/bench_step_n64.log 2>&1
FLASHV_L2_HINT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_persistent_nohint.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench_persistent_n64.log
