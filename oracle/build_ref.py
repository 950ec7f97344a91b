#!/usr/bin/env python3
"""Build and run the UNMODIFIED reference programs (test infrastructure only).

Does exactly what /root/reference/src/run.py:29-54 does: regex-substitute the compile-time
knobs (K_STATE, T_STATE, obserRouteLEN, prob, MAX_THREADS, BeamSearchWidth, data_path and
the `prob%.Nf` format) and compile with
    gcc -g -pthread <src> -o <bin> -lm -Wl,-z,stack-size=268435456
The substituted source is piped to gcc on stdin, so no reference source text is ever
written into this repository; only binaries land in oracle/_ref/ (git-ignored, but they
travel to the GPU box with the repo snapshot — /root/reference does not exist there).

Usage:
    python3 oracle/build_ref.py --standard            # the binaries bench.py / smoke use
    python3 oracle/build_ref.py --prog FLASH --K 64 --M 50 --T 256 --prob 0.253 --N 8
"""
from __future__ import annotations

import argparse
import os
import re
import subprocess
from pathlib import Path

REF_SRC = Path("/root/reference/src")
REF_BASELINES = Path("/root/reference/Base_line/C implementations")
OUT_DIR = Path(__file__).resolve().parent / "_ref"
PROGS = {"FLASH": "FLASH_Viterbi_multithread", "FLASH_BS": "FLASH_BS_Viterbi_multithread"}
# the vanilla baseline: same knob shapes (no MAX_THREADS), built with the same recipe; its K x T tables are static
# arrays, fine at the small sizes the sanity fixtures use
BASELINE_PROGS = {"VANILLA": "vanilla Viterbi"}


def _source_path(prog) -> Path:
    return REF_BASELINES / f"{BASELINE_PROGS[prog]}.c" if prog in BASELINE_PROGS else REF_SRC / f"{PROGS[prog]}.c"


def reference_available() -> bool:
    return all((REF_SRC / f"{v}.c").exists() for v in PROGS.values())


def binary_name(prog, K, M, T, prob, N, B=None) -> str:
    n = f"{prog}_K{K}_M{M}_T{T}_p{prob}_N{N}"
    if prog == "FLASH_BS":
        n += f"_B{B}"
    return n


def _substituted_source(prog, K, M, T, prob, N, B, data_path) -> str:
    text = _source_path(prog).read_text()
    # run.py:29-37
    text = re.sub(r"#define K_STATE \d+", f"#define K_STATE {K}", text)
    text = re.sub(r"#define T_STATE \d+", f"#define T_STATE {M}", text)
    text = re.sub(r"#define obserRouteLEN \d+", f"#define obserRouteLEN {T}", text)
    text = re.sub(r"const float prob = \d+\.\d+;", f"const float prob = {prob};", text)
    text = re.sub(r'const char data_path\[\] = "[^"]*";', f'const char data_path[] = "{data_path}";', text)
    text = re.sub(r"#define MAX_THREADS \d+", f"#define MAX_THREADS {N}", text)
    if prog == "FLASH_BS":
        text = re.sub(r"const int BeamSearchWidth = \d+;", f"const int BeamSearchWidth = {B};", text)
    # run.py:39-47
    s = str(prob)
    places = len(s.split(".")[1]) if "." in s else 0
    text = re.sub(r"prob%\.\d+f", f"prob%.{places}f", text)
    return text


def build(prog, K, M, T, prob, N, B=None, data_path="./data/", out_dir: Path = OUT_DIR, force=False) -> Path:
    """Compile one configuration; returns the binary path (cached by name)."""
    out_dir.mkdir(parents=True, exist_ok=True)
    out = out_dir / binary_name(prog, K, M, T, prob, N, B)
    if out.exists() and not force:
        return out
    if not _source_path(prog).exists():
        raise FileNotFoundError(f"{_source_path(prog)} not present and {out.name} was not prebuilt")
    src = _substituted_source(prog, K, M, T, prob, N, B, data_path)
    cmd = ["gcc", "-g", "-pthread", "-x", "c", "-", "-o", str(out), "-lm", "-Wl,-z,stack-size=268435456"]  # run.py:54
    subprocess.run(cmd, input=src.encode(), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return out


def run(binary: Path, cwd: Path, timeout=None):
    """Run a reference binary with `cwd` holding ./data/; parse the three report lines
    (run.py:75-76 regexes plus the path line)."""
    res = subprocess.run([str(binary)], cwd=str(cwd), capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError(f"{binary.name}: exit {res.returncode}: {res.stderr[-400:]}")
    out = res.stdout
    t = float(re.search(r"time: ([\d.]+)", out).group(1))
    mem = int(re.search(r"memory: (\d+)", out).group(1))
    path = [int(x) for x in re.search(r"path: \[([^\]]*)\]", out).group(1).split()]
    return {"time": t, "memory": mem, "path": path, "stdout": out}


def data_file(data_dir: Path, kind, K, T, prob) -> Path:
    """File name of F:51 / data_script.py:98-101."""
    return data_dir / f"{kind}_K{K}_T{T}_prob{prob}.txt"


# What bench.py (cpu_baseline / --impl reference) and smoke() look for on the GPU box.
# K=3965 at T=34 is the bounded sample of the headline workload: the full T=256 run takes
# ~173 s on 8 cores.
STANDARD = [
    ("FLASH", 64, 50, 256, 0.253, 8, None),
    ("FLASH_BS", 64, 50, 256, 0.253, 8, 8),
    ("FLASH", 3965, 50, 34, 0.112, 8, None),
    ("FLASH", 3965, 50, 34, 0.112, 16, None),
    ("FLASH_BS", 3965, 50, 34, 0.112, 8, 128),
    ("FLASH_BS", 3965, 50, 256, 0.112, 8, 128),
    # bench.py --impl reference: the whole headline config at our segment count, and at the
    # reference-comparable ones (a run is ~45 s on the GPU box's host: the N-way pass is single-threaded)
    ("FLASH", 3965, 50, 256, 0.112, 127, None),
    ("FLASH", 3965, 50, 256, 0.112, 8, None),
    ("FLASH", 3965, 50, 256, 0.112, 64, None),
    ("FLASH", 3965, 50, 256, 0.112, 1, None),
    # tests/test_host_programs.py::test_host_programs_read_dag_named_files
    ("FLASH", 96, 50, 64, 0.9, 9, None),
    ("FLASH_BS", 96, 50, 64, 0.9, 4, 16),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--standard", action="store_true")
    ap.add_argument("--prog", choices=list(PROGS))
    ap.add_argument("--K", type=int)
    ap.add_argument("--M", type=int, default=50)
    ap.add_argument("--T", type=int)
    ap.add_argument("--prob", type=float)
    ap.add_argument("--N", type=int)
    ap.add_argument("--B", type=int)
    a = ap.parse_args()
    if a.standard:
        for cfg in STANDARD:
            print(build(*cfg))
    else:
        print(build(a.prog, a.K, a.M, a.T, a.prob, a.N, a.B))


if __name__ == "__main__":
    main()
