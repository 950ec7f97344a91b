"""ctypes binding of the CPU oracle (oracle/flashv_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libflashv_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile the C restatement (gcc, seconds)."""
    src_newer = (not _LIB_PATH.exists()) or any(
        (_HERE / f).stat().st_mtime > _LIB_PATH.stat().st_mtime for f in ("flashv_oracle.c", "flashv_oracle.h")
    )
    if force or src_newer:
        subprocess.run(["make", "-C", str(_HERE), "-s", "libflashv_oracle.so"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.fvo_model_create.restype = vp
        L.fvo_model_create.argtypes = [C.c_int, C.c_int, fp, fp, fp]
        L.fvo_model_create_lean.restype = vp
        L.fvo_model_create_lean.argtypes = [C.c_int, C.c_int, fp, fp, fp]
        L.fvo_model_free.argtypes = [vp]
        L.fvo_flash_decode.argtypes = [vp, ip, C.c_int, C.c_int, ip, fp, ip]
        L.fvo_bs_decode.argtypes = [vp, ip, C.c_int, C.c_int, C.c_int, ip, fp, ip]
        L.fvo_vanilla_decode.argtypes = [vp, ip, C.c_int, ip, fp, ip]
        L.fvo_flash_step.argtypes = [vp, fp, C.c_int, fp, ip]
        L.fvo_flash_init.argtypes = [vp, C.c_int, C.c_int, fp]
        L.fvo_bs_score_step.argtypes = [vp, fp, ip, C.c_int, C.c_int, fp, ip]
        L.fvo_bs_heap_replay.argtypes = [C.c_int, C.c_int, fp, ip, fp, ip, ip]
        L.fvo_task_list.argtypes = [C.c_int, C.c_int, ip, ip, ip, ip]
        L.fvo_executed_steps.restype = C.c_long
        L.fvo_executed_steps.argtypes = [C.c_int, C.c_int]
        L.fvo_flash_memory_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.fvo_bs_memory_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.fvo_read_floats.restype = C.c_long
        L.fvo_read_floats.argtypes = [C.c_char_p, C.c_long, fp]
        L.fvo_read_ints.restype = C.c_long
        L.fvo_read_ints.argtypes = [C.c_char_p, C.c_long, ip]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleModel:
    """HMM held the way the reference's VIT struct holds it (float32 A, B, Pi)."""

    def __init__(self, A, B, Pi, lean=False):
        """lean=True keeps only the non-zero transitions (fvo_model_create_lean): FLASH decodes only,
        for shapes like K=32768 whose dense double tables would not fit the host."""
        self.A = np.ascontiguousarray(A, dtype=np.float32)
        self.B = np.ascontiguousarray(B, dtype=np.float32)
        self.Pi = np.ascontiguousarray(Pi, dtype=np.float32)
        self.K, self.M = self.B.shape
        self.lean = bool(lean)
        assert self.A.shape == (self.K, self.K) and self.Pi.shape == (self.K,)
        create = lib().fvo_model_create_lean if lean else lib().fvo_model_create
        self._h = create(self.K, self.M, _f(self.A), _f(self.B), _f(self.Pi))
        if not self._h:
            raise MemoryError("fvo_model_create")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fvo_model_free(self._h)
            self._h = None

    def flash(self, ob, N):
        ob = np.ascontiguousarray(ob, dtype=np.int32)
        T = ob.shape[0]
        path = np.empty(T, np.int32)
        score, mem = C.c_float(), C.c_int()
        rc = lib().fvo_flash_decode(self._h, _i(ob), T, N, _i(path), C.byref(score), C.byref(mem))
        if rc != 0:
            raise ValueError(f"oracle: unsupported (T={T}, N={N})")
        return path, np.float32(score.value), mem.value

    def vanilla(self, ob):
        """The reference's vanilla Viterbi baseline (its own arithmetic): (path, score, memory)."""
        ob = np.ascontiguousarray(ob, dtype=np.int32)
        T = ob.shape[0]
        path = np.empty(T, np.int32)
        score, mem = C.c_float(), C.c_int()
        rc = lib().fvo_vanilla_decode(self._h, _i(ob), T, _i(path), C.byref(score), C.byref(mem))
        if rc != 0:
            raise ValueError("oracle: the vanilla decoder's path runs through a dead column (outside the reference's domain)")
        return path, np.float32(score.value), mem.value

    def flash_bs(self, ob, N, Bw):
        assert not self.lean, "the lean oracle model decodes FLASH only"
        ob = np.ascontiguousarray(ob, dtype=np.int32)
        T = ob.shape[0]
        path = np.empty(T, np.int32)
        score, mem = C.c_float(), C.c_int()
        rc = lib().fvo_bs_decode(self._h, _i(ob), T, N, Bw, _i(path), C.byref(score), C.byref(mem))
        if rc != 0:
            raise ValueError(f"oracle: unsupported (T={T}, N={N}, B={Bw})")
        return path, np.float32(score.value), mem.value

    def init(self, prev_state, o):
        d = np.empty(self.K, np.float32)
        lib().fvo_flash_init(self._h, int(prev_state), int(o), _f(d))
        return d

    def step(self, d_in, o):
        d_in = np.ascontiguousarray(d_in, dtype=np.float32)
        d_out = np.empty(self.K, np.float32)
        psi = np.empty(self.K, np.int32)
        lib().fvo_flash_step(self._h, _f(d_in), int(o), _f(d_out), _i(psi))
        return d_out, psi

    def bs_score_step(self, hval, hstate, o):
        hval = np.ascontiguousarray(hval, dtype=np.float32)
        hstate = np.ascontiguousarray(hstate, dtype=np.int32)
        score = np.empty(self.K, np.float32)
        arg = np.empty(self.K, np.int32)
        lib().fvo_bs_score_step(self._h, _f(hval), _i(hstate), hval.shape[0], int(o), _f(score), _i(arg))
        return score, arg


def heap_replay(score, payload, Bw):
    score = np.ascontiguousarray(score, dtype=np.float32)
    payload = np.ascontiguousarray(payload, dtype=np.int32)
    hv = np.empty(Bw, np.float32)
    hs = np.empty(Bw, np.int32)
    hp = np.empty(Bw, np.int32)
    lib().fvo_bs_heap_replay(score.shape[0], Bw, _f(score), _i(payload), _f(hv), _i(hs), _i(hp))
    return hv, hs, hp


def task_list(T, N):
    L = np.zeros(T + 2, np.int32)
    R = np.zeros(T + 2, np.int32)
    mids = np.zeros(max(N, 1), np.int32)
    fp = C.c_int()
    n = lib().fvo_task_list(T, N, _i(L), _i(R), C.byref(fp), _i(mids))
    return [(int(L[q]), int(R[q])) for q in range(n)], bool(fp.value), [int(x) for x in mids[: max(N - 1, 0)]] if fp.value else []


def executed_steps(T, N):
    return int(lib().fvo_executed_steps(T, N))


def flash_memory_bytes(K, T, N):
    return int(lib().fvo_flash_memory_bytes(K, T, N))


def bs_memory_bytes(T, N, Bw):
    return int(lib().fvo_bs_memory_bytes(T, N, Bw))


def read_floats(path, n):
    out = np.empty(n, np.float32)
    got = lib().fvo_read_floats(os.fsencode(str(path)), n, _f(out))
    if got != n:
        raise IOError(f"{path}: wanted {n} floats, got {got}")
    return out


def read_ints(path, n):
    out = np.empty(n, np.int32)
    got = lib().fvo_read_ints(os.fsencode(str(path)), n, _i(out))
    if got != n:
        raise IOError(f"{path}: wanted {n} ints, got {got}")
    return out


def set_threads(n: int) -> None:
    """OpenMP threads of the oracle's loops (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass
