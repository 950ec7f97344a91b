#!/usr/bin/env python3
"""Golden vectors for DAG-structured HMMs, produced by RUNNING THE REFERENCE.

Only runnable where /root/reference exists.  The transition graph is built exactly as
/root/reference/generate_data/data_script_dag.py:45-52 builds it (random.seed(sd); the observation draw;
nx.gnp_random_graph(K, 0.9, directed=True); edges u < v with random.uniform(0, 1) weights;
nx.to_numpy_array) and B comes from that script's own create_B (imported, DD:7-17).

One deliberate difference, stated because it matters: the script normalises with `A / A.sum(axis=1)`
(DD:54), which divides COLUMN j by the sum of ROW j and turns the sink's 0/0 and x/0 into 0 and
1.797e308 (DD:55 nan_to_num).  Its files therefore hold "probabilities" above 1 and an infinite column
after fscanf("%f"), and the unmodified reference programs crash on them [measured here: K=24, T=40,
exit -11 at MAX_THREADS=1].  The fixtures below use the row normalisation the script evidently intends,
`A / A.sum(axis=1)[:, None]` with the sink's all-zero row kept at 0 — a proper DAG HMM (upper-triangular
A, one absorbing sink with no way out), K > T so that live paths exist for the whole sequence.  Files are
named as DD:63-66 names them (`*_K{K}_T{T}_DAG.txt`); the reference programs cannot open those names
(F:51), so for the golden run they are copied to the `prob` names.
"""
from __future__ import annotations

import random
import shutil
import sys
import tempfile
from pathlib import Path

import networkx as nx
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/generate_data")

from oracle import build_ref, oracle  # noqa: E402

K, M, T, SEED, PROB_NAME = 96, 50, 64, 1, 0.9
FLASH_NS = [1, 2, 4, 9, 31]
BS_CFGS = [(1, 8), (4, 16), (9, 96), (31, 5)]


def dag_hmm():
    import data_script_dag  # the reference's script: only create_B is importable, the rest lives in main()

    random.seed(SEED)  # DD:45
    y = [random.randint(0, M - 1) for _ in range(T)]  # DD:46
    G = nx.gnp_random_graph(K, 0.9, directed=True)  # DD:50
    dag = nx.DiGraph([(u, v, {"weight": random.uniform(0, 1)}) for (u, v) in G.edges() if u < v])  # DD:51
    A = nx.to_numpy_array(dag, nodelist=sorted(dag.nodes()))  # DD:52 (node order made explicit)
    with np.errstate(invalid="ignore", divide="ignore"):
        A = np.nan_to_num(A / A.sum(axis=1)[:, None])  # see the module docstring
    B = data_script_dag.create_B(n_observables=M, n_states=K, sd=SEED)  # DD:58
    pi = np.full(K, 1 / K)  # DD:61
    return A, B, pi, np.array(y, np.int32)


def write_dag_text(directory: Path, A, B, pi, ob, suffix="DAG"):
    directory.mkdir(parents=True, exist_ok=True)
    np.savetxt(directory / f"A_K{K}_T{T}_{suffix}.txt", A, fmt="%.16f")  # DD:63-66
    np.savetxt(directory / f"B_K{K}_T{T}_{suffix}.txt", B, fmt="%.16f")
    np.savetxt(directory / f"Pi_K{K}_T{T}_{suffix}.txt", pi, fmt="%.16f", newline=" ")
    np.savetxt(directory / f"ob_K{K}_T{T}_{suffix}.txt", ob, fmt="%d", newline=" ")


def main():
    out_dir = Path(__file__).resolve().parent
    work = Path(tempfile.mkdtemp(prefix="flashv_golden_dag_"))
    try:
        A, B, pi, ob = dag_hmm()
        assert A.shape == (K, K) and np.allclose(np.tril(A), 0) and A[K - 1].sum() == 0
        data = work / "data"
        write_dag_text(data, A, B, pi, ob)
        for kind in ("A", "B", "Pi", "ob"):  # the names the reference programs open (F:51)
            shutil.copy(data / f"{kind}_K{K}_T{T}_DAG.txt", build_ref.data_file(data, kind, K, T, PROB_NAME))
        A32 = oracle.read_floats(data / f"A_K{K}_T{T}_DAG.txt", K * K).reshape(K, K)
        B32 = oracle.read_floats(data / f"B_K{K}_T{T}_DAG.txt", K * M).reshape(K, M)
        Pi32 = oracle.read_floats(data / f"Pi_K{K}_T{T}_DAG.txt", K)
        cases = []
        for N in FLASH_NS:
            r = build_ref.run(build_ref.build("FLASH", K, M, T, PROB_NAME, N, out_dir=work / "bin"), work)
            cases.append((0, 0, N, 0, r["memory"], r["path"]))
        for N, Bw in BS_CFGS:
            r = build_ref.run(build_ref.build("FLASH_BS", K, M, T, PROB_NAME, N, Bw, out_dir=work / "bin"), work)
            cases.append((1, 0, N, Bw, r["memory"], r["path"]))
        np.savez_compressed(
            out_dir / "dag_k96.npz",
            A=A32, B=B32, Pi=Pi32, obs=ob[None, :], A64=A, B64=B, Pi64=pi,
            prob=np.float64(PROB_NAME), seed=np.int64(SEED),
            case_prog=np.array([c[0] for c in cases], np.int32), case_seq=np.array([c[1] for c in cases], np.int32),
            case_N=np.array([c[2] for c in cases], np.int32), case_B=np.array([c[3] for c in cases], np.int32),
            case_memory=np.array([c[4] for c in cases], np.int64), case_path=np.array([c[5] for c in cases], np.int32),
        )
        print(f"dag_k96: {len(cases)} cases; path[:16] = {cases[0][5][:16]}; -1 entries in BS cases:",
              [int((np.array(c[5]) < 0).sum()) for c in cases if c[0] == 1])
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
