#!/bin/bash
mkdir -p gpurun_out
for n in 32 511; do
timeout 600 python tools/profile_target.py --K 512 --T 1024 --batch 8192 --segments $n --iters 2
done
