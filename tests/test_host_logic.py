"""Host-side logic of libflashv.so that needs no GPU: the task tree, the reference's memory
formulas, the text loader, and the arithmetic claim the trellis kernels rest on."""
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


@pytest.mark.parametrize("T,N", [(256, 1), (256, 2), (256, 3), (256, 8), (256, 64), (256, 127), (61, 30), (61, 31),
                                 (100, 49), (40, 19), (2, 1), (3, 1), (5, 2), (7, 3), (1024, 256), (4096, 1024)])
def test_task_list_matches_oracle(fv, oracle_mod, T, N):
    tasks, fp, mids = fv.task_list(T, N)
    otasks, ofp, omids = oracle_mod.task_list(T, N)
    assert fp == ofp and mids == omids
    assert tasks == otasks  # same FIFO order as the reference's queue (F:284-304)
    assert len(tasks) == (T - N if fp else T - 1)
    assert fv.executed_steps(T, N) == oracle_mod.executed_steps(T, N)
    # every time index is decided exactly once: N-way pass (N entries) + one per task
    decided = ([T - 1] + mids) if fp else [T - 1]
    decided += [(l + r) >> 1 for l, r in tasks]
    assert sorted(decided) == list(range(T))


def test_task_list_rejects_broken_domain(fv):
    for T, N in [(1, 1), (60, 30), (16, 8), (10, 0)]:
        with pytest.raises(fv.FlashvError):
            fv.task_list(T, N)


def test_memory_formulas_match_reference(fv, oracle_mod):
    for K, T, N in [(64, 256, 1), (64, 256, 8), (64, 256, 30), (3965, 256, 8), (512, 1024, 256), (37, 61, 31)]:
        assert fv.memory_bytes(K, T, N) == oracle_mod.flash_memory_bytes(K, T, N)
    for T, N, B in [(256, 8, 8), (256, 4, 32), (256, 8, 128), (256, 1, 128), (61, 30, 6)]:
        assert fv.bs_memory_bytes(T, N, B) == oracle_mod.bs_memory_bytes(T, N, B)
    assert fv.memory_bytes(64, 256, 8) == 8368 and fv.bs_memory_bytes(256, 8, 128) == 24944


def test_text_loader_is_fscanf_compatible(fv, oracle_mod, tmp_path):
    rng = np.random.RandomState(7)
    vals = np.concatenate([rng.uniform(0, 1, 3996), [0.0, 1.0, 1e-17, 0.5, 0.1, 2.0 ** -24 + 2.0 ** -49]])
    p = tmp_path / "A.txt"
    np.savetxt(p, vals.reshape(-1, 6), fmt="%.16f")
    a = fv.read_floats_text(p, vals.size)
    b = oracle_mod.read_floats(p, vals.size)  # literal fscanf("%f"), F:85-91
    assert a.tobytes() == b.tobytes()
    q = tmp_path / "ob.txt"
    ints = rng.randint(0, 50, 300)
    np.savetxt(q, ints, fmt="%d", newline=" ")
    assert np.array_equal(fv.read_ints_text(q, 300), oracle_mod.read_ints(q, 300))
    with pytest.raises(IOError):
        fv.read_floats_text(p, vals.size + 5)


def test_parallel_parse_and_binary_cache(fv, oracle_mod, tmp_path):
    """Large float files are parsed in parallel pieces and may be served from a binary side-car
    (SURVEY 8f-2); both must give exactly the values of a serial fscanf("%f") walk, and the cache
    must notice when the text changes."""
    import os
    import time

    rng = np.random.RandomState(3)
    n = (1 << 20) + 4321  # above the parallel-parse threshold
    vals = rng.uniform(0, 1, n) * (rng.uniform(0, 1, n) < 0.3)
    p = tmp_path / "A_big.txt"
    with open(p, "w") as f:
        for r in range(0, n, 997):  # ragged rows, like np.savetxt of a matrix with a remainder
            f.write(" ".join("%.16f" % v for v in vals[r:r + 997]) + "\n")
    want = np.array([float("%.16f" % v) for v in vals]).astype(np.float32)
    a = fv.read_floats_text(p, n)
    assert a.tobytes() == want.tobytes()
    assert oracle_mod.read_floats(p, 5000).tobytes() == a[:5000].tobytes()  # the literal fscanf walk agrees
    with pytest.raises(IOError):
        fv.read_floats_text(p, n + 1)
    c1 = fv.read_floats_cached(p, n)
    assert (tmp_path / "A_big.txt.f32cache").exists() and c1.tobytes() == a.tobytes()
    c2 = fv.read_floats_cached(p, n)  # served from the side-car
    assert c2.tobytes() == a.tobytes()
    time.sleep(0.01)
    with open(p, "r+") as f:  # same size, different content, newer mtime
        f.write("0.5000000000000000")
    os.utime(p, None)
    c3 = fv.read_floats_cached(p, n)
    assert c3[0] == np.float32(0.5) and c3[1:].tobytes() == a[1:].tobytes()


def _ord(x):
    b = x.view(np.int32).astype(np.int64)
    return np.where(b >= 0, b, -(b & 0x7FFFFFFF))


def test_window_filter_bound():
    """The float estimate (tmp+d)+(float)logA is within 2 float steps of the reference's
    (float)((double)(tmp+d)+logA) when all three operands are <= 0, so a 4-step window around the
    largest estimate contains the exact first-argmax (trellis_common.cuh)."""
    rng = np.random.RandomState(11)
    worst = 0
    for scale in (1.0, 30.0, 3000.0, 2.0e4):
        n = 400000
        tmp = np.float32(-rng.uniform(0.5, 6, n))
        d = np.float32(-rng.uniform(0, scale, n))
        la = np.log(rng.uniform(1e-6, 1.0, n).astype(np.float32).astype(np.float64))
        pre = (tmp + d).astype(np.float32)
        exact = (pre.astype(np.float64) + la).astype(np.float32)
        est = (pre + la.astype(np.float32)).astype(np.float32)
        worst = max(worst, int(np.abs(_ord(exact) - _ord(est)).max()))
    assert worst <= 2
    # columns: filtered first-argmax == exact first-argmax, also with ties and -inf entries
    for trial in range(300):
        K = 257
        tmp = np.float32(-rng.uniform(0.5, 6))
        d = np.float32(-rng.uniform(0, 50, K)) * np.float32(rng.choice([1, 40, 400]))
        A = rng.uniform(0.01, 1, K) * (rng.uniform(0, 1, K) < 0.4)
        A = (A / max(A.sum(), 1e-9)).astype(np.float32)
        if trial % 3 == 0:
            A[rng.randint(0, K, 40)] = A[rng.randint(0, K)]  # force repeated values
        with np.errstate(divide="ignore"):
            la = np.log(A.astype(np.float64))
        pre = (tmp + d).astype(np.float32)
        exact = (pre.astype(np.float64) + la).astype(np.float32)
        est = (pre + la.astype(np.float32)).astype(np.float32)
        best, arg = np.float32(-3.4028234663852886e38), -1
        for k in range(K):
            if exact[k] > best:
                best, arg = exact[k], k
        top = est.max()
        if not top > np.float32(-3.4028234663852886e38):
            assert arg == -1
            continue
        cand = np.nonzero(_ord(est) >= _ord(np.array([top], np.float32))[0] - 4)[0]
        fb, fa = np.float32(-3.4028234663852886e38), -1
        for k in cand:
            if exact[k] > fb:
                fb, fa = exact[k], int(k)
        assert (fb, fa) == (best, arg)


def _sum_first_threshold(top, tmp):
    """filter_threshold() of flash_group.cu in numpy float32."""
    c = np.float32(top + tmp)
    e = (int(np.array([c], np.float32).view(np.int32)[0]) >> 23) & 0xFF
    G = np.array([max(e - 23, 1) << 23], np.int32).view(np.float32)[0]
    return np.float32(np.float64(top) - 12.0 * np.float64(G))  # one rounding, like the FFMA


def test_sum_first_filter_bound():
    """The group engine keeps the maximum of m2 = delta[k] + (float)logA[k][i] (two operations per
    update; the emission term tmp is applied afterwards).  Its window — m2 >= top - 12 G, G the float
    spacing at |top + tmp| — must contain the reference's first-argmax of
    (float)((double)(float)(tmp + delta[k]) + logA), also when |tmp| dwarfs |delta + logA|, near
    powers of two, with ties and with -inf entries (flash_group.cu: filter_threshold)."""
    rng = np.random.RandomState(23)
    NEG = np.float32(-3.4028234663852886e38)
    widest = 0
    for trial in range(1500):
        K = 193
        kind = trial % 5
        tmp = np.float32(-rng.uniform(0.0, 6))
        scale = float(rng.choice([1, 40, 400, 4000]))
        d = (np.float32(-rng.uniform(0, 5, K)) - np.float32(rng.uniform(0, 10) * scale)).astype(np.float32)
        A = rng.uniform(0.01, 1, K) * (rng.uniform(0, 1, K) < 0.4)
        A = (A / max(A.sum(), 1e-9)).astype(np.float32)
        if kind == 1:
            tmp = np.float32(-rng.uniform(20, 90))  # tmp much larger than delta + logA
            d = np.float32(-rng.uniform(0, 1.0, K))
        if kind == 2:  # sums straddling a power of two
            pw = np.float32(2.0 ** rng.randint(1, 13))
            d = (-pw - np.float32(rng.uniform(-3, 3, K))).astype(np.float32)
        if kind == 3:
            A[rng.randint(0, K, 60)] = A[rng.randint(0, K)]  # repeated values: exact ties
            d[:] = d[0]
        if kind == 4:
            d[rng.randint(0, K, 20)] = NEG  # dead sources
        with np.errstate(divide="ignore"):
            la = np.log(A.astype(np.float64))
        pre = (tmp + d).astype(np.float32)
        with np.errstate(over="ignore", invalid="ignore"):
            exact = (pre.astype(np.float64) + la).astype(np.float32)
            m2 = (d + la.astype(np.float32)).astype(np.float32)
        best, arg = NEG, -1
        for k in range(K):
            if exact[k] > best:
                best, arg = exact[k], k
        top = m2.max()
        if not top > NEG:
            assert arg == -1
            continue
        thr = _sum_first_threshold(top, tmp)
        cand = np.nonzero(m2 >= thr)[0]
        widest = max(widest, len(cand))
        fb, fa = NEG, -1
        for k in cand:
            if exact[k] > fb:
                fb, fa = exact[k], int(k)
        assert (fb, fa) == (best, arg), (trial, kind)
    assert widest >= 2  # the ties did produce multi-candidate windows


def _half_filter_threshold(top, tmp, c):
    """The window of scan16_fetch() (flash_persistent.cu) in numpy: float arithmetic, rounded DOWN to half."""
    top, tmp, c = np.float32(top), np.float32(tmp), np.float32(c)
    if not top >= np.float32(-30000.0):
        return np.float16(-np.inf)
    atop = np.abs(top)
    W = np.float32(np.float32(2.1) * np.float32(2.0 ** -10) * atop + np.float32(2.0 ** -20)
                   + np.float32(2.0 ** -20) * (np.abs(tmp) + np.abs(c) + atop))
    t32 = np.float32(top - W)
    t16 = np.float16(t32)
    if np.float32(t16) > t32:
        t16 = np.nextafter(t16, np.float16(-np.inf))
    return t16


def test_half_precision_filter_bound():
    """k_flash_persist16 keeps, per chain q = k & 255, the maximum of
        est_k = fl16( fl16(max(delta[k] - c, -60000)) + fl16(log A[k][i]) ),   c = max(delta),
    and evaluates exactly every element of every chain whose maximum reaches top - W.  The reference's
    first-argmax of (float)((double)(float)(tmp + delta[k]) + logA) — and every source tying with it — must sit
    in such a chain: across magnitudes, near powers of two, with ties, dead sources, -inf edges, a huge spread
    of delta (the clamp) and emission terms much larger than the rest."""
    rng = np.random.RandomState(31)
    NEG = np.float32(-3.4028234663852886e38)
    widest, slow = 0, 0
    for trial in range(1500):
        K = 700
        kind = trial % 7
        tmp = np.float32(-rng.uniform(0.0, 6))
        base = np.float32(-rng.uniform(0, 10) * float(rng.choice([1, 40, 400, 4000])))
        d = (np.float32(-rng.uniform(0, 8, K)) + base).astype(np.float32)
        A = rng.uniform(0.01, 1, K) * (rng.uniform(0, 1, K) < 0.4)
        A = (A / max(A.sum(), 1e-9)).astype(np.float32)
        if kind == 1:
            tmp = np.float32(-rng.uniform(20, 90))
        if kind == 2:  # normalised sums straddling a power of two
            pw = np.float32(2.0 ** rng.randint(-2, 9))
            d = (base - np.float32(rng.uniform(0, 1, K)) * pw).astype(np.float32)
            d[rng.randint(0, K)] = base
        if kind == 3:
            A[rng.randint(0, K, 200)] = A[rng.randint(0, K)]  # repeated values: exact ties
            d[:] = d[0]
        if kind == 4:
            d[rng.randint(0, K, 80)] = NEG  # dead sources
        if kind == 5:  # every source with an edge lies far below the best one: clamp / distrust regime
            far = rng.uniform(0, 1, K) < 0.97
            d = np.where(far, d - np.float32(rng.choice([2.0e4, 4.0e4, 7.0e4, 3.0e5])), d).astype(np.float32)
            A = np.where(far, A, 0).astype(np.float32)
        if kind == 6:
            A = (A * (rng.uniform(0, 1, K) < 0.02)).astype(np.float32)  # nearly no edges
        with np.errstate(divide="ignore"):
            la = np.log(A.astype(np.float64))
        pre = (tmp + d).astype(np.float32)
        with np.errstate(over="ignore", invalid="ignore"):
            exact = (pre.astype(np.float64) + la).astype(np.float32)
        best, arg = NEG, -1
        for k in range(K):
            if exact[k] > best:
                best, arg = exact[k], k
        c = d.max()
        if not c > NEG:
            assert arg == -1
            continue
        a16 = np.maximum((d - c).astype(np.float32), np.float32(-60000.0)).astype(np.float16)
        with np.errstate(over="ignore"):
            b16 = la.astype(np.float16)  # double -> half, one rounding
            est = (a16.astype(np.float32) + b16.astype(np.float32)).astype(np.float16)  # == correctly rounded half add
        assert not np.isnan(est.astype(np.float32)).any()
        top = est.max()
        if not top > np.float16(-np.inf):
            assert arg == -1, "an all -inf estimate column must be dead"
            continue
        thr = _half_filter_threshold(top, tmp, c)
        slow += int(np.isinf(np.float32(thr)))
        chain_max = np.full(256, -np.inf, np.float16)
        np.maximum.at(chain_max, np.arange(K) & 255, est)
        chains = np.nonzero(chain_max >= thr)[0]
        cand = np.nonzero(np.isin(np.arange(K) & 255, chains))[0]
        widest = max(widest, len(chains))
        fb, fa = NEG, -1
        for k in cand:
            if exact[k] > fb:
                fb, fa = exact[k], int(k)
        assert (fb, fa) == (best, arg), (trial, kind)
        ties = np.nonzero(exact == best)[0] if arg >= 0 else []
        assert all((k & 255) in chains for k in ties)
    assert widest >= 2 and slow >= 1  # multi-chain windows and the distrust regime both occurred


def test_half_filter_normaliser_is_a_bound():
    """k_flash_persist16 normalises the arriving vector by c = fl32(fl64(fl32(max tmp + max delta_prev) + max log A)),
    derived one step ahead; the filter's error analysis needs c >= every entry of the vector (no cancellation).
    All roundings in the reference's chain are monotone, so the bound must hold exactly, in float arithmetic."""
    rng = np.random.RandomState(41)
    for trial in range(200):
        K = 150
        A = rng.uniform(0.01, 1, (K, K)) * (rng.uniform(0, 1, (K, K)) < rng.choice([0.1, 0.5, 1.0]))
        A = (A / np.maximum(A.sum(axis=1, keepdims=True), 1e-9)).astype(np.float32)
        with np.errstate(divide="ignore"):
            la = np.log(A.astype(np.float64))
        tmp = np.float32(-rng.uniform(0, 6, K))
        d = (np.float32(-rng.uniform(0, 8, K)) * np.float32(rng.choice([1, 30, 900]))).astype(np.float32)
        if trial % 4 == 0:
            d[rng.randint(0, K, K // 3)] = np.float32(-3.4028234663852886e38)
        pre = (tmp[None, :] + d[:, None]).astype(np.float32)  # [k][i]
        with np.errstate(over="ignore", invalid="ignore"):
            cand = (pre.astype(np.float64) + la).astype(np.float32)
        nxt = cand.max(axis=0)
        bound = np.float32(np.float64(np.float32(tmp.max() + d.max())) + la.max())
        assert (nxt <= bound).all(), trial
        # ... and it is tight enough to be useful: within the spread of one step's terms
        live = nxt > np.float32(-3.0e38)
        if live.any() and np.isfinite(bound):
            assert bound - nxt[live].max() < 40.0


def test_header_declares_only_exported_symbols(fv):
    text = (ROOT / "include" / "flashv.h").read_text()
    declared = sorted(set(re.findall(r"\b(flashv_[a-z_A-Z0-9]+)\s*\(", text)))
    assert declared == sorted(fv.ABI_SYMBOLS)
    lib = fv.lib()
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_gpu(fv):
    import ctypes as C

    h = C.c_void_p()
    rc = fv.lib().flashv_ctx_create(0, None, C.byref(h))
    if rc == 0:  # a GPU is present: nothing to check here
        fv.lib().flashv_ctx_destroy(h)
        return
    assert rc == fv.ERR_CUDA and b"no CPU path" in fv.lib().flashv_last_error()


def test_product_never_touches_the_oracle():
    pkg = ROOT / "flash-viterbi_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.c")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + \
            list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("Makefile")):
        t = f.read_text()
        assert "flashv_oracle" not in t and "from oracle" not in t and "import oracle" not in t, f
