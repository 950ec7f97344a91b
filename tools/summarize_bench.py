#!/usr/bin/env python3
"""Print one line per bench log in gpurun_out/."""
import glob
import json
import sys

for f in sorted(glob.glob((sys.argv[1] if len(sys.argv) > 1 else "gpurun_out") + "/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            r = d.get("roofline", {})
            print(f"{f.split('/')[-1]:36s} ms/step {d['ms_per_step']:.3f} value {d['value']:.1f} e2e_ms {d['e2e'].get('ms_per_step', 0):.3f} "
                  f"launches {d['gpu_launches']} pass_ms {r.get('ms_per_launch', 0):.3f} frac {r.get('frac', 0):.3f} parity {d.get('parity')} "
                  f"sm_mhz {d.get('clocks', {}).get('sm_mhz')}")
            break
    else:
        print(f.split("/")[-1], "NO JSON:", open(f).read()[-300:].replace("\n", " | "))
