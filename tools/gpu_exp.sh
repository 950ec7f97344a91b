#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
echo "bench exit $?"
python tools/summarize_bench.py gpurun_out/bench_default.log | cut -c1-1200 | tail -4
