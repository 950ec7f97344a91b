#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --set full` capture of k_flash_persist: DRAM bytes read + written per launch
(what bench.py reports as roofline.traffic), plus the issue-slot and shared-memory figures of roofline.physical.

    python tools/traffic_from_ncu.py gpurun_out/prof_persist.ncu-rep profiles/r02_persist_ncu_details.txt
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

rep, name = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def nbytes(key):
    u, v = d[key]
    return float(v) * scale[u]


out = {
    "k_flash_persist_dram_bytes_per_launch": int(nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")),
    "dram_bytes_read": int(nbytes("dram__bytes_read.sum")), "dram_bytes_written": int(nbytes("dram__bytes_write.sum")),
    "duration_ms_under_ncu": float(d["gpu__time_duration.sum"][1]),
    "issue_slots_busy_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][1]),
    "warp_instructions": float(d["smsp__inst_executed.sum"][1]),
    "source": f"profiles/{Path(name).name} (ncu --set full --clock-control none, K=3965 T=256, one launch = 255 steps)",
}
Path(__file__).resolve().parents[1].joinpath("profiles", "traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out))
