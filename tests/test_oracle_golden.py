"""The oracle (oracle/flashv_oracle.c) against vectors produced by RUNNING the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, golden_cases, load_golden

# reference FLASH path for hmm_k64 sequence 0, identical for N in {1,2,4,8,30}
APPENDIX_A_PATH = """24 13 24 13 24 13 24 13 24 13 24 13 47 36 13 24 13 9 3 56 5 20 37 5 38 13 9 29 53 2 25 20 37 5 38 13 24 25 20 37 5 20 63 10 3 36 13 9 38 62 20 37 5 38 53 2 32 36 13 24 13 24 13 24 13 24 13 24 32 38 13 24 13 24 13 9 38 13 24 13 24 13 24 13 24 32 12 38 13 24 13 47 28 36 13 24 13 9 9 29 5 38 13 24 13 24 13 24 13 47 36 13 24 13 24 13 24 13 47 62 23 23 13 24 13 9 29 5 38 13 24 63 49 50 57 24 13 24 13 24 36 13 24 13 9 56 5 38 13 24 13 24 13 24 13 24 13 24 13 24 13 24 13 24 13 24 13 9 9 29 53 37 36 13 24 13 24 32 63 49 30 5 38 13 24 13 9 38 13 24 13 24 13 24 13 47 62 20 63 49 31 20 37 5 20 37 5 38 13 24 13 24 32 38 13 47 62 37 5 20 37 3 1 5 38 13 24 13 9 56 23 13 24 32 12 38 13 24 13 9 29 5 38 13 47 36 13 24 13 24 13 24 13 24 13 24"""


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_reference_outputs(oracle_mod, name):
    g = load_golden(name)
    m = oracle_mod.OracleModel(g["A"], g["B"], g["Pi"])
    n = 0
    for case in golden_cases(name):
        if case["prog"] == 0:
            path, score, mem = m.flash(case["ob"], case["N"])
        else:
            path, score, mem = m.flash_bs(case["ob"], case["N"], case["B"])
        assert np.array_equal(path, case["path"]), (name, case["prog"], case["N"], case["B"])
        assert mem == case["memory"], (name, case["prog"], case["N"], case["B"])
        n += 1
    assert n == len(g["case_prog"]) and n > 0


def test_appendix_a_known_answer(oracle_mod):
    g = load_golden("hmm_k64")
    want = np.array(APPENDIX_A_PATH.split(), dtype=np.int32)
    m = oracle_mod.OracleModel(g["A"], g["B"], g["Pi"])
    for N in (1, 2, 4, 8, 30):
        path, score, _ = m.flash(g["obs"][0], N)
        assert np.array_equal(path, want)
        assert float(score) == -1388.0970458984375
    # and the reference binary itself said the same
    for case in golden_cases("hmm_k64"):
        if case["prog"] == 0 and case["seq"] == 0 and case["N"] in (1, 2, 4, 8, 30):
            assert np.array_equal(case["path"], want)


def test_goldens_cover_beam_dropouts():
    # FLASH-BS paths with -1 entries (S:73-86) must be part of the pinned set
    n = sum(int((c["path"] == -1).any()) for name in GOLDEN_NAMES for c in golden_cases(name) if c["prog"] == 1)
    assert n >= 5


def test_memory_formulas(oracle_mod):
    # SURVEY.md Appendix A "memory:" lines
    assert oracle_mod.flash_memory_bytes(64, 256, 1) == 1144
    assert oracle_mod.flash_memory_bytes(64, 256, 8) == 8368
    assert oracle_mod.flash_memory_bytes(64, 256, 30) == 31072
    assert oracle_mod.bs_memory_bytes(256, 8, 8) == 1904
    assert oracle_mod.bs_memory_bytes(256, 4, 32) == 3312
    assert oracle_mod.flash_memory_bytes(3965, 256, 8) == 507696
    assert oracle_mod.bs_memory_bytes(256, 8, 128) == 24944


def test_executed_steps_table(oracle_mod):
    # SURVEY.md §8d
    want = {(256, 1): 1793, (256, 8): 1288, (256, 16): 1040, (256, 32): 800, (256, 64): 576, (256, 127): 386,
            (1024, 8): 7176, (1024, 64): 4160, (1024, 256): 2304, (4096, 8): 36872, (4096, 256): 16640,
            (4096, 1024): 9216}
    for (T, N), s in want.items():
        assert oracle_mod.executed_steps(T, N) == s


def test_oracle_rejects_broken_domain(oracle_mod):
    g = load_golden("hmm_k37")
    m = oracle_mod.OracleModel(g["A"], g["B"], g["Pi"])
    ob = g["obs"][0][:60]
    with pytest.raises(ValueError):
        m.flash(ob, 30)  # T == 2N: the reference leaves Ans[] entries unset


def test_step_and_decode_agree(oracle_mod):
    # the exposed single-step pieces reproduce the N=1 root pass end state
    g = load_golden("hmm_k37")
    m = oracle_mod.OracleModel(g["A"], g["B"], g["Pi"])
    ob = g["obs"][1]
    d = m.init(-1, ob[0])
    for j in range(1, len(ob)):
        d, psi = m.step(d, ob[j])
    path, score, _ = m.flash(ob, 1)
    assert int(np.argmax(d)) == path[-1] and np.float32(d.max()) == score


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_lean_oracle_equals_literal(oracle_mod, name):
    """The edge-list form of the oracle (used for K=32768 shapes) is the same function as the literal
    loops: identical start vectors, delta/psi of single steps (dead columns included) and decoded paths."""
    g = load_golden(name)
    full = oracle_mod.OracleModel(g["A"], g["B"], g["Pi"])
    lean = oracle_mod.OracleModel(g["A"], g["B"], g["Pi"], lean=True)
    rng = np.random.RandomState(7)
    K, M = full.K, full.M
    for prev in (-1, 0, K - 1):
        assert np.array_equal(full.init(prev, 1).view(np.uint32), lean.init(prev, 1).view(np.uint32))
    d = full.init(-1, 0)
    d[rng.randint(0, K, K // 3)] = -np.inf
    for o in range(min(M, 4)):
        d1, p1 = full.step(d, o)
        d2, p2 = lean.step(d, o)
        assert np.array_equal(d1.view(np.uint32), d2.view(np.uint32)) and np.array_equal(p1, p2)
        d = d1
    for case in golden_cases(name):
        if case["prog"] != 0:
            continue
        path, score, mem = lean.flash(case["ob"], case["N"])
        assert np.array_equal(path, case["path"]) and mem == case["memory"]
        assert np.float32(score).view(np.uint32) == np.float32(full.flash(case["ob"], case["N"])[1]).view(np.uint32)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_vanilla_oracle_matches_reference_baseline(oracle_mod, name):
    """the restatement of the reference's vanilla Viterbi baseline against the outputs of the
    baseline program itself (tests/golden/make_golden_vanilla.py), and the sanity relation README:71 claims —
    here FLASH and vanilla decode the same path on every frozen instance (not guaranteed in general: the two
    programs add in a different order, "vanilla Viterbi.c":140 vs F:170)."""
    from conftest import ROOT

    v = np.load(ROOT / "tests" / "golden" / "vanilla.npz")
    g = load_golden(name)
    om = oracle_mod.OracleModel(g["A"], g["B"], g["Pi"])
    for si, ob in enumerate(g["obs"]):
        path, score, mem = om.vanilla(ob)
        assert np.array_equal(path, v[f"{name}__{si}__path"]) and mem == int(v[f"{name}__{si}__memory"])
        assert np.array_equal(path, om.flash(ob, 1)[0])
