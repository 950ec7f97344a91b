#!/bin/bash
# ncu captures (one gpurun call; each ncu run follows a plain run of the same command that exited 0).
mkdir -p gpurun_out
P="python tools/profile_target.py"
$P --engine persistent > gpurun_out/plain_persist.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_flash_persist -s 2 -c 1 -f -o gpurun_out/prof_persist \
    $P --engine persistent > gpurun_out/ncu_persist.log 2>&1
$P --engine persistent --iters 2 > gpurun_out/plain_level.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_flash_step -s 4 -c 2 -f -o gpurun_out/prof_level \
    $P --engine persistent --iters 2 > gpurun_out/ncu_level.log 2>&1
$P --engine persistent --iters 2 > gpurun_out/plain_persist2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_persist.csv \
    $P --engine persistent --iters 2 > gpurun_out/ncu_launch_persist.log 2>&1
cat gpurun_out/plain_persist.log gpurun_out/plain_level.log
