#!/bin/bash
mkdir -p gpurun_out
P="python tools/profile_target.py --engine sparse --segments 127 --iters 2"
$P > gpurun_out/plain_sparse.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_flash_sparse_pass -s 1 -c 1 -f -o gpurun_out/prof_sparse $P > gpurun_out/ncu_sparse.log 2>&1
cat gpurun_out/plain_sparse.log; tail -2 gpurun_out/ncu_sparse.log
