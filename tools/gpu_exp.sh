#!/bin/bash
# scratch script for the experiment at hand: persistent kernel variants — parity, timing, phase trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "trellis_step or goldens or golden_vectors or bench_instance_flash or random_models or wide_model or headline_flash_vs" > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -4 gpurun_out/pytest_res.log
for P in 0 1; do
FLASHV_PIN=$P timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_pin$P.log 2> gpurun_out/bench_pin$P.err
echo "bench PIN=$P exit $?"; tail -c 300 gpurun_out/bench_pin$P.err
done
python tools/summarize_bench.py gpurun_out/bench_pin0.log gpurun_out/bench_pin1.log
FLASHV_TRACE_FILE=gpurun_out/trace.bin python tools/profile_target.py --engine persistent --segments 127 --iters 3 > gpurun_out/trace_run.log 2>&1
python tools/trace_report.py gpurun_out/trace.bin > gpurun_out/trace_report.txt 2>&1
cat gpurun_out/trace_run.log gpurun_out/trace_report.txt
