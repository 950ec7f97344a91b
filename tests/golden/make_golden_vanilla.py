#!/usr/bin/env python3
"""Golden vectors of the reference's vanilla Viterbi baseline ("Base_line/C implementations/vanilla Viterbi.c"),
produced by RUNNING it (oracle/build_ref.py, the src/run.py:29-54 recipe) on the models already frozen under
tests/golden/ — SURVEY §8f-4: vanilla as a sanity path beside FLASH.

Only runnable where /root/reference exists.  The text files are written from the frozen float32 values with 9
significant digits, which fscanf("%f") reads back to exactly those floats, so the baseline decodes the same
model the FLASH goldens were recorded on.  Output: tests/golden/vanilla.npz with one path (and the program's
`memory:` line) per (model, sequence)."""
from __future__ import annotations

import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import build_ref  # noqa: E402

MODELS = ["hmm_k64", "hmm_k37", "hmm_k128", "hmm_k257", "dag_k96"]


def main():
    out_dir = Path(__file__).resolve().parent
    work = Path(tempfile.mkdtemp(prefix="flashv_golden_vanilla_"))
    out = {}
    try:
        for name in MODELS:
            g = np.load(out_dir / f"{name}.npz")
            A, B, Pi, obs = g["A"], g["B"], g["Pi"], g["obs"]
            K, M = B.shape
            T = obs.shape[1]
            prob = float(g["prob"])
            for si, ob in enumerate(obs):
                run_dir = work / f"{name}_{si}"
                data = run_dir / "data"
                data.mkdir(parents=True)
                np.savetxt(build_ref.data_file(data, "A", K, T, prob), A, fmt="%.9e")
                np.savetxt(build_ref.data_file(data, "B", K, T, prob), B, fmt="%.9e")
                np.savetxt(build_ref.data_file(data, "Pi", K, T, prob), Pi, fmt="%.9e", newline=" ")
                np.savetxt(build_ref.data_file(data, "ob", K, T, prob), ob, fmt="%d", newline=" ")
                r = build_ref.run(build_ref.build("VANILLA", K, M, T, prob, 1, out_dir=work / "bin"), run_dir)
                out[f"{name}__{si}__path"] = np.array(r["path"], np.int32)
                out[f"{name}__{si}__memory"] = np.int64(r["memory"])
                print(name, si, r["path"][:10], r["memory"])
        np.savez_compressed(out_dir / "vanilla.npz", **out)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
