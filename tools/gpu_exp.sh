#!/bin/bash
# scratch script for the experiment at hand: L2 prefetch of the double table in the pinned persistent pass; new tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "grid_y or level_kernel or bench_instance_flash" > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -3 gpurun_out/pytest_res.log
for P in 0 1 0 1; do FLASHV_LAD_PREFETCH=$P python tools/profile_target.py --engine persistent --segments 127 --iters 6; done
FLASHV_LAD_PREFETCH=1 FLASHV_TRACE_FILE=gpurun_out/trace.bin python tools/profile_target.py --engine persistent --segments 127 --iters 3 > gpurun_out/trace_run.log 2>&1
python tools/trace_report.py gpurun_out/trace.bin 2>&1 | grep -v "resident part\|ring part" | head -16
