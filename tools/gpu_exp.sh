#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "trellis_step or golden or headline_flash_vs or random_models" 2>&1 | tail -3
for g in 0 1; do
echo "== FLASHV_TMEM=$g"
FLASHV_TMEM=$g timeout 120 python tools/profile_target.py --engine persistent --iters 4 --segments 127
FLASHV_TMEM=$g FLASHV_TRACE_FILE=gpurun_out/trace_g$g.bin timeout 120 python tools/profile_target.py --engine persistent --iters 3 > /dev/null 2>&1
python tools/trace_report.py gpurun_out/trace_g$g.bin > gpurun_out/trace_g$g.txt 2>&1; head -7 gpurun_out/trace_g$g.txt
done
