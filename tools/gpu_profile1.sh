#!/bin/bash
mkdir -p gpurun_out
P="python tools/profile_target.py"
$P --engine persistent --segments 127 > gpurun_out/plain_persist.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_flash_persist -s 2 -c 1 -f -o gpurun_out/prof_persist \
    $P --engine persistent --segments 127 > gpurun_out/ncu_persist.log 2>&1
cat gpurun_out/plain_persist.log
