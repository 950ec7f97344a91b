#!/bin/bash
mkdir -p gpurun_out
FLASHV_PREP_TRACE=1 python - <<'PY'
import sys, time
sys.path.insert(0, "flash-viterbi_b200/host")
import numpy as np
from __graft_entry__ import load_pkg
import gen_hmm
fv = load_pkg()
A, B, Pi = gen_hmm.make_hmm(3965, 50, 0.112, 1)
ctx = fv.Context(0)
for it in range(3):
    t0 = time.perf_counter()
    m = fv.Model(ctx, A, B, Pi)
    t1 = time.perf_counter()
    print(f"== model {it}: wall {1e3*(t1-t0):.1f} ms, prep_ms {m.prep_ms:.1f}", flush=True)
    m.close()
PY
