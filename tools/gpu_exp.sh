#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "corner or large_batch" 2>&1 | tail -4
