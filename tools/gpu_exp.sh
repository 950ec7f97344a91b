#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "goldens or golden_vectors or bench_instance_flash or random_models or half_filter" > gpurun_out/pytest_res.log 2>&1
tail -2 gpurun_out/pytest_res.log
python tools/profile_target.py --engine persistent --segments 127 --iters 6
python tools/profile_target.py --engine persistent --segments 8 --iters 4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err
python tools/summarize_bench.py gpurun_out/bench_a.log | head -2
