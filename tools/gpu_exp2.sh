#!/bin/bash
# multi-GPU experiment: the multi-GPU tests, bench.py under torchrun on all visible GPUs, the NCCL baseline
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; nproc >> gpurun_out/topo.txt; free -g >> gpurun_out/topo.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_mgpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_mgpu.log
tail -5 gpurun_out/pytest_mgpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench exit $?"
tail -c 800 gpurun_out/bench_n$N.err
python tools/summarize_bench.py gpurun_out/bench_n$N.log 2>/dev/null | cut -c1-1800
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/nccl_baseline.py --K 32768 --T 64 > gpurun_out/nccl_baseline_n$N.log 2>&1
grep "^{" gpurun_out/nccl_baseline_n$N.log || tail -5 gpurun_out/nccl_baseline_n$N.log
