#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Only runnable where /root/reference exists (the build container).  For every HMM below it
  1. draws A and B with the reference's own generator functions
     (/root/reference/generate_data/data_script.py:5-49, imported, not copied) and writes the
     four text files with the same np.savetxt formats (data_script.py:98-101),
  2. builds the unmodified FLASH / FLASH-BS programs with oracle/build_ref.py (the
     src/run.py:29-54 recipe) for each (N, B) and runs them on those files,
  3. stores A/B/Pi as the float32 values fscanf("%f") produces (F:85-91), the observation
     sequences, and the reference's path / memory outputs in <name>.npz.
The observation draw of the reference is unseeded (data_script.py:86), so sequences are
frozen here; sequence 0 of hmm_k64 is the one recorded in SURVEY.md Appendix A.
"""
from __future__ import annotations

import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/generate_data")

from oracle import build_ref, oracle  # noqa: E402

APPENDIX_A_OB = """17 43 29 34 7 15 4 41 18 17 23 4 40 10 6 27 28 41 1 41 14 9 1 39 39 33 41 26 45 47 40 43 45 1 20 15 23 12 12 0 24 48 27 22 25 30 18 46 2 26 23 13 11 19 45 7 45 29 41 2 21 24 0 27 15 17 48 45 27 20 42 21 47 16 28 45 40 10 0 4 17 18 43 39 44 7 41 19 16 15 18 12 39 41 33 27 22 46 27 36 5 2 18 26 41 37 21 27 16 39 38 11 33 24 22 42 17 15 10 7 49 44 39 17 8 33 6 12 28 30 7 45 8 41 40 33 15 27 37 0 24 42 27 31 45 17 24 2 24 17 20 7 23 24 34 7 16 47 36 21 37 8 20 43 2 36 35 11 31 46 37 8 3 15 0 48 17 39 4 15 39 30 39 29 9 24 4 24 38 43 32 9 36 9 28 14 16 45 42 19 4 34 38 14 12 1 30 6 18 36 47 18 19 39 16 10 9 20 35 39 38 30 23 48 40 28 32 20 38 8 34 25 10 5 5 26 49 23 36 48 6 11 26 35 37 47 23 23 23 17 10 9 42 17 15 7"""

# name, K, M, T, prob, seed, n_extra_sequences, FLASH Ns, BS (N, B) pairs
HMMS = [
    ("hmm_k64", 64, 50, 256, 0.253, 1, 1,
     [1, 2, 3, 4, 8, 30, 64, 127],
     [(1, 8), (4, 8), (8, 8), (8, 1), (8, 2), (8, 3), (8, 16), (8, 32), (8, 64), (3, 5), (30, 4), (2, 16)]),
    ("hmm_k37", 37, 7, 61, 0.4, 3, 1,
     [1, 2, 3, 5, 7, 30, 31],
     [(1, 37), (3, 9), (7, 4), (30, 6), (2, 36)]),
    ("hmm_k128", 128, 50, 100, 0.1, 2, 1,
     [1, 4, 16, 49],
     [(4, 16), (16, 32), (8, 128), (1, 12)]),
    ("hmm_k257", 257, 11, 40, 0.9, 5, 0,
     [1, 3, 8, 19],
     [(3, 64), (8, 200)]),
]


def generate_text(data_dir: Path, K, M, T, prob, seed, ob):
    import data_script  # the reference's generator

    A = data_script.create_A_b(n_nodes=K, sd=seed, prob=prob)
    B = data_script.create_B(n_observables=M, n_states=K, sd=seed)
    pi = np.full(K, 1 / K)
    np.savetxt(build_ref.data_file(data_dir, "A", K, T, prob), A, fmt="%.16f")
    np.savetxt(build_ref.data_file(data_dir, "B", K, T, prob), B, fmt="%.16f")
    np.savetxt(build_ref.data_file(data_dir, "Pi", K, T, prob), pi, fmt="%.16f", newline=" ")
    np.savetxt(build_ref.data_file(data_dir, "ob", K, T, prob), ob, fmt="%d", newline=" ")


def main():
    out_dir = Path(__file__).resolve().parent
    work = Path(tempfile.mkdtemp(prefix="flashv_golden_"))
    bin_dir = work / "bin"
    try:
        for name, K, M, T, prob, seed, n_extra, flash_ns, bs_cfgs in HMMS:
            rng = np.random.RandomState(1000 + seed)
            seqs = []
            if name == "hmm_k64":
                seqs.append(np.array(APPENDIX_A_OB.split(), dtype=np.int32))
            while len(seqs) < 1 + n_extra:
                seqs.append(rng.randint(0, M, size=T).astype(np.int32))
            cases = []
            A32 = B32 = Pi32 = None
            for si, ob in enumerate(seqs):
                run_dir = work / f"{name}_s{si}"
                data_dir = run_dir / "data"
                data_dir.mkdir(parents=True)
                generate_text(data_dir, K, M, T, prob, seed, ob)
                if A32 is None:
                    A32 = oracle.read_floats(build_ref.data_file(data_dir, "A", K, T, prob), K * K).reshape(K, K)
                    B32 = oracle.read_floats(build_ref.data_file(data_dir, "B", K, T, prob), K * M).reshape(K, M)
                    Pi32 = oracle.read_floats(build_ref.data_file(data_dir, "Pi", K, T, prob), K)
                for N in flash_ns:
                    b = build_ref.build("FLASH", K, M, T, prob, N, out_dir=bin_dir)
                    r = build_ref.run(b, run_dir)
                    cases.append((0, si, N, 0, r["memory"], r["path"]))
                for N, Bw in bs_cfgs:
                    b = build_ref.build("FLASH_BS", K, M, T, prob, N, Bw, out_dir=bin_dir)
                    r = build_ref.run(b, run_dir)
                    cases.append((1, si, N, Bw, r["memory"], r["path"]))
                print(f"{name} seq {si}: {len(cases)} cases so far")
            np.savez_compressed(
                out_dir / f"{name}.npz",
                A=A32, B=B32, Pi=Pi32, obs=np.stack(seqs),
                prob=np.float64(prob), seed=np.int64(seed),
                case_prog=np.array([c[0] for c in cases], np.int32),
                case_seq=np.array([c[1] for c in cases], np.int32),
                case_N=np.array([c[2] for c in cases], np.int32),
                case_B=np.array([c[3] for c in cases], np.int32),
                case_memory=np.array([c[4] for c in cases], np.int64),
                case_path=np.array([c[5] for c in cases], np.int32),
            )
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
