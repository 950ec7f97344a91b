#!/usr/bin/env python3
"""Print the bench JSON lines of the given log files (or gpurun_out/bench*.log) in a readable form."""
import glob
import json
import sys

files = sys.argv[1:] or sorted(glob.glob("gpurun_out/bench*.log"))
for f in files:
    found = False
    for line in open(f):
        if not line.startswith("{"):
            continue
        found = True
        d = json.loads(line)
        r = d.get("roofline", {}) or {}
        print(f"== {f}: n_gpus {d.get('n_gpus')} ms/step {d.get('ms_per_step', 0):.3f} value {d.get('value', 0):.1f} "
              f"e2e_ms {d.get('e2e', {}).get('ms_per_step', 0):.3f} cold_ms {(d.get('e2e_cold') or {}).get('ms')} "
              f"launches {d.get('gpu_launches')} pass_ms {r.get('ms_per_launch') or 0:.3f} frac {r.get('frac') or 0:.3f} "
              f"parity {d.get('parity')} prep_ms {d.get('model_prep_ms')} sm_mhz {d.get('clocks', {}).get('sm_mhz')}")
        if r.get("physical"):
            print("   physical:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r["physical"].items() if not k.endswith(("note", "basis"))})
        if d.get("cpu_baseline"):
            print("   cpu_baseline:", json.dumps(d["cpu_baseline"])[:300])
        for k, v in (d.get("other_segment_counts") or {}).items():
            print("   ", k, json.dumps(v)[:260])
        for k, v in (d.get("other_configs") or {}).items():
            print("   ", k, json.dumps(v)[:1500])
    if not found:
        print(f, "NO JSON:", open(f).read()[-600:].replace("\n", " | "))
