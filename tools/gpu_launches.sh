#!/bin/bash
mkdir -p gpurun_out
P="python tools/profile_target.py --engine persistent --segments 127 --iters 2"
$P > gpurun_out/plain_l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_n127.csv $P > gpurun_out/ncu_l.log 2>&1
cat gpurun_out/plain_l.log
