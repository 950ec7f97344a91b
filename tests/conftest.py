"""Shared fixtures: marker registration, the package loader (the directory name
flash-viterbi_b200 is not an importable identifier) and the golden vectors."""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    name = "flash_viterbi_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = ROOT / "flash-viterbi_b200"
    spec = importlib.util.spec_from_file_location(name, pkg_dir / "__init__.py", submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def fv():
    mod = load_pkg()
    if not mod.LIB_PATH.exists():
        mod.build()
    return mod


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle

    oracle.build()
    return oracle


# dag_k96: a DAG-structured HMM in data_script_dag.py's construction (tests/golden/make_golden_dag.py)
GOLDEN_NAMES = ["hmm_k64", "hmm_k37", "hmm_k128", "hmm_k257", "dag_k96"]


def load_golden(name):
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    return {k: g[k] for k in g.files}


def golden_cases(name):
    g = load_golden(name)
    for c in range(len(g["case_prog"])):
        yield {
            "prog": int(g["case_prog"][c]), "seq": int(g["case_seq"][c]), "N": int(g["case_N"][c]),
            "B": int(g["case_B"][c]), "memory": int(g["case_memory"][c]), "path": g["case_path"][c],
            "ob": g["obs"][int(g["case_seq"][c])],
        }


@pytest.fixture(scope="session")
def gpu_ctx(fv):
    ctx = fv.Context(0)
    yield ctx
    ctx.close()


def random_hmm(K, M, p, seed):
    """Random sparse HMM in the style of the reference generator (data_script.py:5-49): Binomial(K,p)
    out-edges per row with U(0.01,1) weights, row-normalised; B ~ U(0.1,1) row-normalised; pi uniform.
    Values are rounded through '%.16f' text the way the reference's files are, then to float32."""
    rng = np.random.RandomState(seed)
    A = np.zeros((K, K))
    for s in range(K):
        n = max(1, rng.binomial(K, p))
        idx = rng.choice(K, size=n, replace=False)
        A[s, idx] = rng.uniform(0.01, 1, size=n)
    A /= A.sum(axis=1, keepdims=True)
    B = rng.uniform(0.1, 1, (K, M))
    B /= B.sum(axis=1, keepdims=True)
    Pi = np.full(K, 1.0 / K)
    r = lambda x: np.round(x, 16).astype(np.float32)
    return r(A), r(B), r(Pi)
