#!/bin/bash
mkdir -p gpurun_out
P="python tools/profile_target.py --beam 128 --segments 8 --T 64 --iters 2"
$P > gpurun_out/plain_bs.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_bs_pass -s 0 -c 1 -f -o gpurun_out/prof_bs \
    $P > gpurun_out/ncu_bs.log 2>&1
cat gpurun_out/plain_bs.log; tail -2 gpurun_out/ncu_bs.log
