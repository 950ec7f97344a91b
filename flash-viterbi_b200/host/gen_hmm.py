#!/usr/bin/env python3
"""Synthetic HMMs with the distribution (and, for a given seed, the very numbers) of the reference
generator generate_data/data_script.py: row s of A gets Binomial(K, prob) distinct out-edges with
U(0.01, 1) weights and is normalised (data_script.py:5-35); B ~ U(0.1, 1) row-normalised after
re-seeding with the same seed (data_script.py:38-49); pi uniform (data_script.py:94).  The
reference draws observations from an UNSEEDED random.randint (data_script.py:86); here they come
from random.Random(ob_seed) so that runs are reproducible.

    python3 gen_hmm.py -s 1 -n 50 -K 64 -T 256 -p 0.253 [-o DIR] [--ob-seed 1000]

writes A_/B_/Pi_/ob_K{K}_T{T}_prob{p}.txt exactly as data_script.py:98-101 names and formats them.
"""
from __future__ import annotations

import argparse
import random
from pathlib import Path

import numpy as np


def transition_matrix(K: int, prob: float, seed: int) -> np.ndarray:
    np.random.seed(seed)
    states = list(range(K))
    A = np.zeros((K, K))
    for s in range(K):
        fanout = np.random.binomial(K, p=prob, size=None)
        targets = np.random.choice(states, size=fanout, replace=False)
        A[s, targets] = np.random.uniform(0.01, 1, size=fanout)
    for s in range(K):
        A[s, ] = A[s, ] / np.sum(A[s, ])
    return A


def transition_matrix_f32_into(out: np.ndarray, K: int, prob: float, seed: int) -> None:
    """transition_matrix() followed by as_reference_floats(), written row by row into `out` ([K][K]
    float32, e.g. a shared-memory mapping): the same numpy calls in the same order, hence the same
    numbers, without ever holding the K x K float64 table (8.6 GB at K=32768).  Rows are independent
    after the draws (normalisation is per row), so producing them one at a time changes nothing."""
    np.random.seed(seed)
    states = np.arange(K)  # choice() draws the same permutation for an array as for the reference's list
    row = np.zeros(K)
    for s in range(K):
        fanout = np.random.binomial(K, p=prob, size=None)
        targets = np.random.choice(states, size=fanout, replace=False)
        row[:] = 0.0
        row[targets] = np.random.uniform(0.01, 1, size=fanout)
        out[s, :] = np.round(row / np.sum(row), 16).astype(np.float32)


def emission_matrix(K: int, M: int, seed: int) -> np.ndarray:
    np.random.seed(seed)
    B = np.random.uniform(0.1, 1, (K, M))
    return B / B.sum(axis=1)[:, None]


def observation_batch(count: int, T: int, M: int, first_seed: int, stride: int = 1) -> np.ndarray:
    """`count` sequences of uniform symbols, sequence q seeded with first_seed + q*stride (numpy streams:
    the per-call Python generator of observations() takes seconds for the 8192 x 1024 batches)."""
    return np.stack([np.random.RandomState(first_seed + q * stride).randint(0, M, T) for q in range(count)]).astype(np.int32)


def observations(T: int, M: int, ob_seed: int) -> np.ndarray:
    rng = random.Random(ob_seed)
    return np.array([rng.randint(0, M - 1) for _ in range(T)], dtype=np.int32)


def as_reference_floats(x: np.ndarray) -> np.ndarray:
    """float32 values the reference's fscanf("%f") would hold after the '%.16f' text round trip
    (decimal rounding first, then float; the double step in between is exact for < 1e-16 cases
    only up to double rounding, which is why file-based parity goes through the real text)."""
    return np.round(x, 16).astype(np.float32)


def make_hmm(K: int, M: int, prob: float, seed: int):
    A = transition_matrix(K, prob, seed)
    B = emission_matrix(K, M, seed)
    Pi = np.full(K, 1 / K)
    return A, B, Pi


def file_name(directory: Path, kind: str, K: int, T: int, prob: float) -> Path:
    return Path(directory) / f"{kind}_K{K}_T{T}_prob{prob}.txt"


def write_text(directory: Path, K: int, T: int, prob: float, A, B, Pi, ob) -> None:
    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    np.savetxt(file_name(directory, "A", K, T, prob), A, fmt="%.16f")
    np.savetxt(file_name(directory, "B", K, T, prob), B, fmt="%.16f")
    np.savetxt(file_name(directory, "Pi", K, T, prob), Pi, fmt="%.16f", newline=" ")
    np.savetxt(file_name(directory, "ob", K, T, prob), ob, fmt="%d", newline=" ")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-s", type=int, required=True, help="seed")
    ap.add_argument("-n", type=int, required=True, help="number of observation symbols (T_STATE)")
    ap.add_argument("-K", type=int, required=True)
    ap.add_argument("-T", type=int, required=True)
    ap.add_argument("-p", type=float, required=True)
    ap.add_argument("-b", type=int, default=0, help="accepted for data_script.py compatibility; unused")
    ap.add_argument("-o", default=".", help="output directory")
    ap.add_argument("--ob-seed", type=int, default=1000)
    a = ap.parse_args()
    A, B, Pi = make_hmm(a.K, a.n, a.p, a.s)
    write_text(Path(a.o), a.K, a.T, a.p, A, B, Pi, observations(a.T, a.n, a.ob_seed))


if __name__ == "__main__":
    main()
