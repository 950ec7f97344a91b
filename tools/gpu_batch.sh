#!/bin/bash
# config 4 timing: batched FLASH, K=512, T=1024 (reduced batch first)
mkdir -p gpurun_out
for b in 512 2048; do
timeout 600 python tools/profile_target.py --K 512 --T 1024 --batch $b --segments 32 --iters 1 >> gpurun_out/batch_k512.log 2>&1
done
cat gpurun_out/batch_k512.log
