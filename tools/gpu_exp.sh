#!/bin/bash
# scratch script: half-precision level kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "goldens or golden_vectors or bench_instance_flash or random_models or level_kernel or batch_equals or headline_flash_vs" > gpurun_out/pytest_res.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res.log
tail -12 gpurun_out/pytest_res.log
for N in 1 8 64 127; do
python tools/profile_target.py --engine persistent --segments $N --iters 4
FLASHV_LEVEL16=0 python tools/profile_target.py --engine persistent --segments $N --iters 4
done
echo "== VG 2"; for N in 1 8; do FLASHV_LEVEL16_VG=2 python tools/profile_target.py --engine persistent --segments $N --iters 4; done
