#!/usr/bin/env python3
"""Bench driver for the GPU host programs, compatible with the reference's src/run.py.

Same contract as /root/reference/src/run.py ("R:" below): a list of parameter dicts
(K_STATE, T_STATE, obserRouteLEN, prob, MAX_THREADS, BeamSearchWidth), the knobs are substituted
textually into the program source with the same regular expressions (R:29-47), the program is
compiled, run, its stdout parsed with R:75-76's two regexes, and one CSV row per run is appended
to result/<program>_result.csv with R:105's header — so rows from the reference and from the GPU
build can sit in one table.  The only differences: the compile line links libflashv.so instead of
-pthread -lm, and extra report lines (score, device time, roofline inputs) are kept in extra
CSV columns after the reference's nine.

    python3 run.py --data ./data/ [--result ./result/] [--programs FLASH_Viterbi_multithread ...]
"""
from __future__ import annotations

import argparse
import csv
import re
import subprocess
import sys
from datetime import datetime
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
LIB = HERE.parent / "lib"
PROGRAMS = ["FLASH_Viterbi_multithread", "FLASH_BS_Viterbi_multithread"]

# R:8-25 — the reference driver's own two configurations
parameters = [
    {"K_STATE": 3965, "T_STATE": 50, "obserRouteLEN": 256, "prob": 0.112, "MAX_THREADS": 1, "BeamSearchWidth": 32},
    {"K_STATE": 3965, "T_STATE": 50, "obserRouteLEN": 256, "prob": 0.169, "MAX_THREADS": 1, "BeamSearchWidth": 32},
]

HEADER = ["timestamp", "K_STATE", "T_STATE", "obserRouteLEN", "prob", "MAX_THREADS", "BeamSearchWidth", "time", "memory"]
EXTRA = ["score", "device_decode_ms", "model_prep_ms", "time_including_prep", "executed_steps", "device_bytes"]


def substitute(source: str, p: dict, data_path: str, program: str) -> str:
    rules = [
        (r"#define K_STATE \d+", f"#define K_STATE {p['K_STATE']}"),
        (r"#define T_STATE \d+", f"#define T_STATE {p['T_STATE']}"),
        (r"#define obserRouteLEN \d+", f"#define obserRouteLEN {p['obserRouteLEN']}"),
        (r"const float prob = \d+\.\d+;", f"const float prob = {p['prob']};"),
        (r'const char data_path\[\] = "[^"]*";', f'const char data_path[] = "{data_path}";'),
        (r"#define MAX_THREADS \d+", f"#define MAX_THREADS {p['MAX_THREADS']}"),
    ]
    if "BS" in program:
        rules.append((r"const int BeamSearchWidth = \d+;", f"const int BeamSearchWidth = {p['BeamSearchWidth']};"))
    text = str(p["prob"])
    places = len(text.split(".")[1]) if "." in text else 0
    rules.append((r"prob%\.\d+f", f"prob%.{places}f"))
    for pattern, repl in rules:
        source, hits = re.subn(pattern, repl, source)
        if hits == 0:
            raise RuntimeError(f"{program}: pattern {pattern!r} did not match")
    return source


def compile_program(program: str, p: dict, data_path: str, out_dir: Path) -> Path:
    out_dir.mkdir(parents=True, exist_ok=True)
    src = substitute((HERE / f"{program}.c").read_text(), p, data_path, program)
    modified = out_dir / f"{program}_modified.c"
    modified.write_text(src)
    binary = out_dir / f"{program}_modified"
    cmd = ["gcc", "-g", str(modified), "-o", str(binary), f"-I{ROOT / 'include'}", f"-L{LIB}", "-lflashv",
           f"-Wl,-rpath,{LIB}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"compile ERROR: {res.stderr}")
    return binary


def run_program(binary: Path) -> dict:
    res = subprocess.run([str(binary)], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"run ERROR: {res.stderr}")
    out = res.stdout
    row = {"time": re.search(r"time: ([\d.]+)", out).group(1), "memory": re.search(r"memory: (\d+)", out).group(1)}
    for key in EXTRA:
        m = re.search(rf"{key}: ([-\d.e+]+)", out)
        row[key] = m.group(1) if m else ""
    row["path"] = re.search(r"path: \[([^\]]*)\]", out).group(1).split()
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="./data/")
    ap.add_argument("--result", default="./result/")
    ap.add_argument("--build-dir", default="./build_host/")
    ap.add_argument("--programs", nargs="*", default=PROGRAMS)
    a = ap.parse_args()
    result = Path(a.result)
    result.mkdir(parents=True, exist_ok=True)
    for program in a.programs:
        csv_path = result / f"{program}_result.csv"
        fresh = not csv_path.exists()
        with open(csv_path, "a", encoding="utf-8", newline="") as fh:
            w = csv.writer(fh)
            if fresh:
                w.writerow(HEADER + EXTRA)
            for p in parameters:
                binary = compile_program(program, p, a.data, Path(a.build_dir))
                row = run_program(binary)
                print(f"{program} Time: {row['time']}, Memory: {row['memory']}")
                w.writerow([datetime.now().strftime("%Y-%m-%d %H:%M:%S"), p["K_STATE"], p["T_STATE"], p["obserRouteLEN"],
                            p["prob"], p.get("MAX_THREADS", "N/A"), p.get("BeamSearchWidth", "N/A"), row["time"],
                            row["memory"]] + [row[k] for k in EXTRA])
                fh.flush()


if __name__ == "__main__":
    sys.exit(main())
