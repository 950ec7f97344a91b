#!/bin/bash
# scratch script for the experiment at hand: pipelined heap replay in FLASH-BS
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py tests/test_host_programs.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "bs or golden or one_process or host_program or dag" > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -3 gpurun_out/pytest_res.log
FLASHV_BS_TRACE=1 python tools/profile_target.py --beam 128 --segments 8 --iters 3 2>&1 | grep "steps=255\|engine" | tail -2
python tools/profile_target.py --beam 128 --segments 127 --iters 3
FLASHV_BS_TRACE=1 python tools/profile_target.py --beam 32 --segments 1 --iters 3 2>&1 | grep "steps=255\|engine" | tail -2
