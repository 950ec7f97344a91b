#!/bin/bash
mkdir -p gpurun_out
FLASHV_BS_TRACE=1 python tools/profile_target.py --beam 128 --segments 8 --iters 1 2>&1 | tail -8
FLASHV_BS_TRACE=1 python tools/profile_target.py --beam 32 --segments 1 --iters 1 2>&1 | tail -11
