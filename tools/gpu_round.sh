#!/bin/bash
# One gpurun call: GPU parity tests, then short bench runs.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then echo "tests failed: skipping benches"; exit 0; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_persistent_n64.log 2>&1
timeout 120 python bench.py --steps 5 --warmup 3 --segments 8 --no-cpu > gpurun_out/bench_persistent_n8.log 2>&1
timeout 120 python bench.py --steps 5 --warmup 3 --segments 127 --no-cpu > gpurun_out/bench_persistent_n127.log 2>&1
timeout 120 python bench.py --steps 5 --warmup 3 --engine step --no-cpu > gpurun_out/bench_step_n64.log 2>&1
for c in; do
FLASHV_CHUNK=$c timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_persistent_chunk$c.log 2>&1
done
tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench_persistent_n64.log
python tools/profile_target.py --beam 128 --segments 8 --iters 2 > gpurun_out/bs_n8_b128.log 2>&1
python tools/profile_target.py --beam 128 --segments 1 --iters 2 >> gpurun_out/bs_n8_b128.log 2>&1
python tools/profile_target.py --beam 32 --segments 8 --iters 2 >> gpurun_out/bs_n8_b128.log 2>&1
cat gpurun_out/bs_n8_b128.log
bash tools/gpu_trace.sh
