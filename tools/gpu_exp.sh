#!/bin/bash
mkdir -p gpurun_out
P="python tools/profile_target.py --engine persistent --segments 63 --iters 2 --K 16384 --T 64"
$P > gpurun_out/plain_k16384.log 2>&1 && cat gpurun_out/plain_k16384.log &&
ncu --set full --clock-control none --import-source on -k regex:k_flash_persist -s 1 -c 1 -f -o gpurun_out/prof_persist_k16384 $P > gpurun_out/ncu_k16384.log 2>&1
tail -2 gpurun_out/ncu_k16384.log
