#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "goldens or bench_instance_flash or half_filter or trellis_step" > gpurun_out/pytest_res.log 2>&1
tail -2 gpurun_out/pytest_res.log
python tools/profile_target.py --engine persistent --segments 127 --iters 6
FLASHV_TRACE_FILE=gpurun_out/trace.bin python tools/profile_target.py --engine persistent --segments 127 --iters 3 > gpurun_out/trace_run.log 2>&1
python tools/trace_report.py gpurun_out/trace.bin 2>&1 | head -10
