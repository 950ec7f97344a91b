#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump (SASS view) of ONE kernel: stall reasons over all samples, the
instructions with the most samples, the instruction mix, and time / instructions by execution frequency
(which separates inner loops from code that runs once per tile or step).

    ncu -i report.ncu-rep --page source --csv > src.csv ; python tools/ncu_stalls.py src.csv [warps]
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
total = sum(int(r[col["# Samples"]] or 0) for r in data)
instr = sum(int(r[col["Instructions Executed"]]) for r in data)
print(f"{rows[0][1] if rows and len(rows[0]) > 1 else ''}")
print(f"samples {total}, warp-instructions executed {instr}, static instructions {len(data)}")
agg = collections.Counter()
for r in data:
    for n in stalls:
        agg[n] += int(r[col[n]] or 0)
print("\nstall reasons (all samples):")
for n, v in agg.most_common(10):
    print(f"  {n:26s} {100 * v / max(total, 1):5.1f} %")
print("\ninstruction mix:")
mix = collections.Counter()
for r in data:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]].strip())
    mix[m.group(2).split(".")[0] if m else "?"] += int(r[col["Instructions Executed"]])
for op, v in mix.most_common(14):
    print(f"  {op:10s} {100 * v / max(instr, 1):5.1f} %")
print("\ninstructions with the most samples:")
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:16]:
    st = sorted(((n, int(r[col[n]] or 0)) for n in stalls), key=lambda x: -x[1])[:2]
    print(f"  {r[col['# Samples']]:>6s} samples  x{r[col['Instructions Executed']]:>9s}  {r[col['Source']].strip()[:52]:52s} {st}")
if len(sys.argv) > 2:
    warps = float(sys.argv[2])
    print(f"\nby executions per warp ({warps:.0f} warps launched):")
    b = collections.defaultdict(lambda: [0, 0, 0])
    for r in data:
        n = int(r[col["Instructions Executed"]])
        k = round(n / warps, 1) if n / warps < 20 else round(n / warps)
        e = b[k]
        e[0] += 1
        e[1] += n
        e[2] += int(r[col["# Samples"]] or 0)
    for k, e in sorted(b.items(), key=lambda x: -x[1][2])[:8]:
        print(f"  x{k:<7} {e[0]:5d} static, {e[1] / warps:8.1f} executed per warp ({100 * e[1] / instr:4.1f} %), {100 * e[2] / max(total, 1):4.1f} % of the samples")
