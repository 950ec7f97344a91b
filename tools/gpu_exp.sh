#!/bin/bash
# scratch script for the experiment at hand: half-precision filter kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "trellis_step or goldens or golden_vectors or bench_instance_flash or random_models or wide_model or headline_flash_vs" > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -5 gpurun_out/pytest_res.log
timeout 120 python tools/profile_target.py --engine persistent --segments 127 --iters 6
echo "== float sweep"; FLASHV_F16=0 timeout 120 python tools/profile_target.py --engine persistent --segments 127 --iters 6
for it in 4 12 16; do echo "== tm iters $it"; FLASHV_F16_TM_ITERS=$it timeout 120 python tools/profile_target.py --engine persistent --segments 127 --iters 6; done
FLASHV_TRACE_FILE=gpurun_out/trace.bin timeout 120 python tools/profile_target.py --engine persistent --segments 127 --iters 3 > gpurun_out/trace_run.log 2>&1
python tools/trace_report.py gpurun_out/trace.bin 2>&1
