#!/bin/bash
# scratch script for the experiment at hand: ncu capture of the batched level kernel (N=8 decode)
mkdir -p gpurun_out
P="python tools/profile_target.py --engine persistent --segments 8 --iters 2"
ncu --set full --clock-control none --import-source on -k regex:k_flash_step -s 10 -c 1 -f -o gpurun_out/prof_level $P > gpurun_out/ncu_level.log 2>&1
tail -3 gpurun_out/ncu_level.log
