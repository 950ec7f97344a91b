#!/bin/bash
mkdir -p gpurun_out
FLASHV_TRACE_FILE=gpurun_out/trace.bin python tools/profile_target.py --engine persistent --iters 3 > gpurun_out/trace_run.log 2>&1
python tools/trace_report.py gpurun_out/trace.bin > gpurun_out/trace_report.txt 2>&1
cat gpurun_out/trace_run.log gpurun_out/trace_report.txt
