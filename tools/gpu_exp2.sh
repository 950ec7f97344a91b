#!/bin/bash
# multi-GPU tests (+ optionally a torchrun bench: BENCH=1) on the N GPUs of the box
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_mgpu_${N}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_mgpu_${N}.log
tail -n 4 gpurun_out/pytest_mgpu_${N}.log
if [ -n "$BENCH" ]; then
timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
echo "bench exit $?"
python tools/summarize_bench.py gpurun_out/bench_${N}gpu.log | cut -c1-600
fi
