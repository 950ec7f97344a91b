"""flash-viterbi_b200 — Python binding of libflashv.so (the C ABI of include/flashv.h).

This is a thin ctypes mirror used by tests/ and bench.py; the product is the shared library and
the reference-shaped host programs under host/.  There is no CPU path: every call goes to the
CUDA library or raises.  The directory name is not a valid Python identifier; load it with

    import importlib.util, sys
    spec = importlib.util.spec_from_file_location(
        "flash_viterbi_b200", "<repo>/flash-viterbi_b200/__init__.py",
        submodule_search_locations=["<repo>/flash-viterbi_b200"])
    mod = importlib.util.module_from_spec(spec); sys.modules[spec.name] = mod; spec.loader.exec_module(mod)

(tests/conftest.py and __graft_entry__.py do exactly that).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libflashv.so"

ENGINE_AUTO, ENGINE_STEP, ENGINE_PERSISTENT, ENGINE_SPARSE = 0, 1, 2, 3
OK, ERR_ARG, ERR_DOMAIN, ERR_CUDA, ERR_NOMEM, ERR_STATE = 0, -1, -2, -3, -4, -5

# every symbol include/flashv.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = [
    "flashv_last_error", "flashv_version",
    "flashv_ctx_create", "flashv_ctx_destroy", "flashv_ctx_stream", "flashv_ctx_sync", "flashv_ctx_sm_count",
    "flashv_model_create", "flashv_model_destroy", "flashv_model_K", "flashv_model_M", "flashv_model_prep_ms",
    "flashv_read_floats_text", "flashv_read_ints_text",
    "flashv_decode", "flashv_bs_decode", "flashv_decode_batch", "flashv_bs_decode_batch", "flashv_vanilla_decode",
    "flashv_plan_create", "flashv_plan_destroy", "flashv_plan_upload", "flashv_plan_run",
    "flashv_plan_download", "flashv_plan_report",
    "flashv_plan_shard_init", "flashv_plan_shard_buffers", "flashv_plan_shard_ipc_handle",
    "flashv_plan_shard_set_peer", "flashv_plan_shard_open_peer",
    "flashv_model_create_rows", "flashv_model_rows_handle", "flashv_model_pull_rows", "flashv_model_pull_rows_from",
    "flashv_model_finish", "flashv_shard_count", "flashv_decode_batch_shard", "flashv_bs_decode_batch_shard",
    "flashv_mgpu_create", "flashv_mgpu_destroy", "flashv_mgpu_world", "flashv_mgpu_ctx", "flashv_mgpu_model",
    "flashv_mgpu_model_create", "flashv_mgpu_decode_batch", "flashv_mgpu_bs_decode_batch", "flashv_mgpu_decode",
    "flashv_read_floats_cached", "flashv_trellis_init", "flashv_trellis_step", "flashv_trellis_step_columns_dev", "flashv_bs_score_step", "flashv_bs_heap_replay",
    "flashv_task_list", "flashv_executed_steps", "flashv_memory_bytes", "flashv_bs_memory_bytes",
]


class Report(C.Structure):
    _fields_ = [
        ("decode_ms", C.c_double), ("first_pass_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("executed_steps", C.c_longlong), ("device_bytes", C.c_longlong),
        ("memory_bytes", C.c_int), ("first_pass", C.c_int), ("n_tasks", C.c_int), ("n_levels", C.c_int),
        ("kernel_launches", C.c_int), ("engine", C.c_int),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class FlashvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"flashv error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> Path:
    """Compile libflashv.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(PKG_DIR / "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libflashv.so failed:\n" + (res.stdout or "")[-2000:] + (res.stderr or "")[-4000:])
    return LIB_PATH


_lib = None


def lib():
    """The loaded library.  Fails loudly if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `make -C {PKG_DIR / 'csrc'}` (or __graft_entry__.build())")
    L = C.CDLL(str(LIB_PATH))
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)
    rp = C.POINTER(Report)
    L.flashv_last_error.restype = C.c_char_p
    L.flashv_version.restype = C.c_char_p
    L.flashv_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.flashv_ctx_destroy.argtypes = [vp]
    L.flashv_ctx_destroy.restype = None
    L.flashv_ctx_stream.argtypes = [vp]
    L.flashv_ctx_stream.restype = vp
    L.flashv_ctx_sync.argtypes = [vp]
    L.flashv_ctx_sm_count.argtypes = [vp]
    L.flashv_model_create.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, C.POINTER(vp)]
    L.flashv_model_destroy.argtypes = [vp]
    L.flashv_model_destroy.restype = None
    L.flashv_model_K.argtypes = [vp]
    L.flashv_model_M.argtypes = [vp]
    L.flashv_model_prep_ms.argtypes = [vp]
    L.flashv_model_prep_ms.restype = C.c_double
    L.flashv_read_floats_text.argtypes = [C.c_char_p, C.c_long, fp]
    L.flashv_read_floats_text.restype = C.c_long
    L.flashv_read_floats_cached.argtypes = [C.c_char_p, C.c_long, fp]
    L.flashv_read_floats_cached.restype = C.c_long
    L.flashv_read_ints_text.argtypes = [C.c_char_p, C.c_long, ip]
    L.flashv_read_ints_text.restype = C.c_long
    L.flashv_decode.argtypes = [vp, ip, C.c_int, C.c_int, ip, fp, rp]
    L.flashv_bs_decode.argtypes = [vp, ip, C.c_int, C.c_int, C.c_int, ip, fp, rp]
    L.flashv_vanilla_decode.argtypes = [vp, ip, C.c_int, ip, fp, rp]
    L.flashv_decode_batch.argtypes = [vp, ip, C.c_int, C.c_int, C.c_int, ip, fp, rp]
    L.flashv_bs_decode_batch.argtypes = [vp, ip, C.c_int, C.c_int, C.c_int, C.c_int, ip, fp, rp]
    L.flashv_plan_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.flashv_plan_destroy.argtypes = [vp]
    L.flashv_plan_destroy.restype = None
    L.flashv_plan_upload.argtypes = [vp, ip]
    L.flashv_plan_run.argtypes = [vp]
    L.flashv_plan_download.argtypes = [vp, ip, fp]
    L.flashv_plan_report.argtypes = [vp, rp]
    L.flashv_plan_shard_init.argtypes = [vp, C.c_int, C.c_int]
    L.flashv_plan_shard_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.flashv_plan_shard_ipc_handle.argtypes = [vp, vp]
    L.flashv_plan_shard_set_peer.argtypes = [vp, C.c_int, C.c_int, vp]
    L.flashv_plan_shard_open_peer.argtypes = [vp, C.c_int, vp]
    L.flashv_model_create_rows.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, C.c_int, C.c_int, C.POINTER(vp)]
    L.flashv_model_rows_handle.argtypes = [vp, vp]
    L.flashv_model_pull_rows.argtypes = [vp, C.c_int, vp]
    L.flashv_model_pull_rows_from.argtypes = [vp, C.c_int, vp]
    L.flashv_model_finish.argtypes = [vp]
    L.flashv_shard_count.argtypes = [C.c_int, C.c_int, C.c_int]
    L.flashv_decode_batch_shard.argtypes = [vp, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, fp, rp]
    L.flashv_mgpu_create.argtypes = [C.c_int, ip, C.POINTER(vp)]
    L.flashv_mgpu_destroy.argtypes = [vp]
    L.flashv_mgpu_destroy.restype = None
    L.flashv_mgpu_world.argtypes = [vp]
    L.flashv_mgpu_ctx.argtypes = [vp, C.c_int]
    L.flashv_mgpu_ctx.restype = vp
    L.flashv_mgpu_model.argtypes = [vp, C.c_int]
    L.flashv_mgpu_model.restype = vp
    L.flashv_mgpu_model_create.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp]
    L.flashv_mgpu_decode_batch.argtypes = [vp, ip, C.c_int, C.c_int, C.c_int, ip, fp, rp]
    L.flashv_mgpu_decode.argtypes = [vp, ip, C.c_int, C.c_int, ip, fp, rp]
    L.flashv_bs_decode_batch_shard.argtypes = [vp, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, fp, rp]
    L.flashv_mgpu_bs_decode_batch.argtypes = [vp, ip, C.c_int, C.c_int, C.c_int, C.c_int, ip, fp, rp]
    L.flashv_trellis_init.argtypes = [vp, C.c_int, C.c_int, fp]
    L.flashv_trellis_step.argtypes = [vp, fp, C.c_int, fp, ip, C.c_int]
    L.flashv_trellis_step_columns_dev.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.flashv_bs_score_step.argtypes = [vp, fp, ip, C.c_int, C.c_int, fp, ip]
    L.flashv_bs_heap_replay.argtypes = [vp, fp, C.c_int, C.c_int, fp, ip]
    L.flashv_task_list.argtypes = [C.c_int, C.c_int, ip, ip, ip, ip]
    L.flashv_executed_steps.argtypes = [C.c_int, C.c_int]
    L.flashv_executed_steps.restype = C.c_longlong
    L.flashv_memory_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.flashv_bs_memory_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise FlashvError(rc, lib().flashv_last_error().decode(errors="replace"))


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


# ---- host-side pure functions ----------------------------------------------------------------
def task_list(T, N):
    """(tasks in queue order, first_pass, mids) — F:284-304 / F:349-359."""
    L = np.zeros(max(T, 1), np.int32)
    R = np.zeros(max(T, 1), np.int32)
    mids = np.zeros(max(N, 1), np.int32)
    fp = C.c_int32()
    n = lib().flashv_task_list(T, N, _i(L), _i(R), C.byref(fp), _i(mids))
    if n < 0:
        _check(n)
    return [(int(L[q]), int(R[q])) for q in range(n)], bool(fp.value), [int(x) for x in mids[: N - 1]] if fp.value else []


def executed_steps(T, N):
    v = lib().flashv_executed_steps(T, N)
    if v < 0:
        _check(int(v))
    return int(v)


def memory_bytes(K, T, N):
    return int(lib().flashv_memory_bytes(K, T, N))


def bs_memory_bytes(T, N, B):
    return int(lib().flashv_bs_memory_bytes(T, N, B))


def read_floats_text(path, n):
    out = np.empty(n, np.float32)
    got = lib().flashv_read_floats_text(os.fsencode(str(path)), n, _f(out))
    if got != n:
        raise IOError(f"{path}: wanted {n} floats, got {got}")
    return out


def read_floats_cached(path, n):
    """read_floats_text through the binary side-car <path>.f32cache (written on the first parse)."""
    out = np.empty(n, np.float32)
    got = lib().flashv_read_floats_cached(os.fsencode(str(path)), n, _f(out))
    if got != n:
        raise IOError(f"{path}: wanted {n} floats, got {got}")
    return out


def read_ints_text(path, n):
    out = np.empty(n, np.int32)
    got = lib().flashv_read_ints_text(os.fsencode(str(path)), n, _i(out))
    if got != n:
        raise IOError(f"{path}: wanted {n} ints, got {got}")
    return out


# ---- device objects --------------------------------------------------------------------------
class Context:
    """One GPU + one stream.  stream: a raw cudaStream_t value (e.g. torch's .cuda_stream) or None."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        _check(lib().flashv_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = device

    @property
    def stream(self):
        return lib().flashv_ctx_stream(self._h)

    @property
    def sm_count(self):
        return lib().flashv_ctx_sm_count(self._h)

    def sync(self):
        _check(lib().flashv_ctx_sync(self._h))

    def heap_replay(self, score, B):
        """generate_state_heap() over a score vector (S:167-211): (heap values, heap states) in array order."""
        score = np.ascontiguousarray(score, np.float32)
        hv = np.empty(B, np.float32)
        hs = np.empty(B, np.int32)
        _check(lib().flashv_bs_heap_replay(self._h, _f(score), score.shape[0], B, _f(hv), _i(hs)))
        return hv, hs

    def close(self):
        if self._h:
            lib().flashv_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model:
    """Device-resident log tables of one HMM (replaces create_vit(), F:97-107)."""

    def __init__(self, ctx: Context, A, B, Pi):
        A = np.ascontiguousarray(A, np.float32)
        B = np.ascontiguousarray(B, np.float32)
        Pi = np.ascontiguousarray(Pi, np.float32)
        K, M = B.shape
        if A.shape != (K, K) or Pi.shape != (K,):
            raise ValueError("A must be [K][K], B [K][M], Pi [K]")
        self.ctx, self.K, self.M = ctx, K, M
        self._h = C.c_void_p()
        _check(lib().flashv_model_create(ctx._h, K, M, _f(A), _f(B), _f(Pi), C.byref(self._h)))

    @classmethod
    def create_rows(cls, ctx: "Context", A, B, Pi, rank, world):
        """Model creation shared by the ranks of a box: this rank's rows of host logarithms only; fetch the
        rest with pull_rows() after a barrier, then finish() after another (include/flashv.h).  A may be a
        read-only mapping shared by the ranks (only this rank's rows are read)."""
        self = cls.__new__(cls)
        A = np.asarray(A)
        assert A.dtype == np.float32 and A.flags.c_contiguous
        B = np.ascontiguousarray(B, np.float32)
        Pi = np.ascontiguousarray(Pi, np.float32)
        K, M = B.shape
        self.ctx, self.K, self.M = ctx, K, M
        self._h = C.c_void_p()
        _check(lib().flashv_model_create_rows(ctx._h, K, M, _f(A), _f(B), _f(Pi), rank, world, C.byref(self._h)))
        return self

    def rows_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib().flashv_model_rows_handle(self._h, buf))
        return buf.raw

    def pull_rows(self, peer_rank, handle: bytes):
        _check(lib().flashv_model_pull_rows(self._h, peer_rank, C.create_string_buffer(handle, 64)))

    def pull_rows_from(self, peer_rank, peer: "Model"):
        _check(lib().flashv_model_pull_rows_from(self._h, peer_rank, peer._h))

    def finish(self):
        _check(lib().flashv_model_finish(self._h))

    @property
    def prep_ms(self):
        return lib().flashv_model_prep_ms(self._h)

    def decode_batch_shard(self, obs, N, rank, world, paths=None, scores=None, B=0):
        """This rank's share (rows rank, rank+world, ...) of obs[total][T], left in place in paths/scores."""
        obs = np.ascontiguousarray(obs, np.int32)
        total, T = obs.shape
        if paths is None:
            paths = np.full((total, T), -2, np.int32)
        if scores is None:
            scores = np.zeros(total, np.float32)
        rep = Report()
        if B > 0:
            _check(lib().flashv_bs_decode_batch_shard(self._h, _i(obs), total, T, N, B, rank, world, _i(paths), _f(scores), C.byref(rep)))
        else:
            _check(lib().flashv_decode_batch_shard(self._h, _i(obs), total, T, N, rank, world, _i(paths), _f(scores), C.byref(rep)))
        return paths, scores, rep

    def decode(self, ob, N):
        """calc() of FLASH (F:338-368): (path[T], score, report)."""
        ob = np.ascontiguousarray(ob, np.int32)
        path = np.empty(ob.shape[0], np.int32)
        score, rep = C.c_float(), Report()
        _check(lib().flashv_decode(self._h, _i(ob), ob.shape[0], N, _i(path), C.byref(score), C.byref(rep)))
        return path, np.float32(score.value), rep

    def vanilla_decode(self, ob):
        """The reference's vanilla Viterbi baseline (sanity path, its own arithmetic): (path[T], score, report)."""
        ob = np.ascontiguousarray(ob, np.int32)
        path = np.empty(ob.shape[0], np.int32)
        score, rep = C.c_float(), Report()
        _check(lib().flashv_vanilla_decode(self._h, _i(ob), ob.shape[0], _i(path), C.byref(score), C.byref(rep)))
        return path, np.float32(score.value), rep

    def bs_decode(self, ob, N, B):
        """calc() of FLASH-BS (S:548-577)."""
        ob = np.ascontiguousarray(ob, np.int32)
        path = np.empty(ob.shape[0], np.int32)
        score, rep = C.c_float(), Report()
        _check(lib().flashv_bs_decode(self._h, _i(ob), ob.shape[0], N, B, _i(path), C.byref(score), C.byref(rep)))
        return path, np.float32(score.value), rep

    def decode_batch(self, obs, N, B=0):
        obs = np.ascontiguousarray(obs, np.int32)
        batch, T = obs.shape
        paths = np.empty((batch, T), np.int32)
        scores = np.empty(batch, np.float32)
        rep = Report()
        if B > 0:
            _check(lib().flashv_bs_decode_batch(self._h, _i(obs), batch, T, N, B, _i(paths), _f(scores), C.byref(rep)))
        else:
            _check(lib().flashv_decode_batch(self._h, _i(obs), batch, T, N, _i(paths), _f(scores), C.byref(rep)))
        return paths, scores, rep

    def trellis_init(self, prev_state, o):
        d = np.empty(self.K, np.float32)
        _check(lib().flashv_trellis_init(self._h, int(prev_state), int(o), _f(d)))
        return d

    def trellis_step(self, delta_in, o, engine=ENGINE_STEP):
        delta_in = np.ascontiguousarray(delta_in, np.float32)
        d = np.empty(self.K, np.float32)
        psi = np.empty(self.K, np.int32)
        _check(lib().flashv_trellis_step(self._h, _f(delta_in), int(o), _f(d), _i(psi), engine))
        return d, psi

    def trellis_step_columns_dev(self, delta_in_ptr, o, col_begin, col_end, delta_out_ptr, psi_out_ptr):
        """One step over the destination states [col_begin, col_end), device pointers, asynchronous."""
        _check(lib().flashv_trellis_step_columns_dev(self._h, C.c_void_p(delta_in_ptr), int(o), int(col_begin), int(col_end),
                                                     C.c_void_p(delta_out_ptr), C.c_void_p(psi_out_ptr)))

    def bs_score_step(self, hval, hstate, o):
        hval = np.ascontiguousarray(hval, np.float32)
        hstate = np.ascontiguousarray(hstate, np.int32)
        score = np.empty(self.K, np.float32)
        arg = np.empty(self.K, np.int32)
        _check(lib().flashv_bs_score_step(self._h, _f(hval), _i(hstate), hval.shape[0], int(o), _f(score), _i(arg)))
        return score, arg

    def close(self):
        if self._h:
            lib().flashv_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    """Staged decode: upload / run / download separately (what bench.py times)."""

    def __init__(self, model: Model, T, N, batch=1, B=0, engine=ENGINE_AUTO):
        self.model, self.T, self.N, self.batch, self.B = model, T, N, batch, B
        self._h = C.c_void_p()
        _check(lib().flashv_plan_create(model._h, T, N, batch, B, engine, C.byref(self._h)))

    def upload(self, obs):
        obs = np.ascontiguousarray(obs, np.int32)
        assert obs.size == self.batch * self.T
        self._keep = obs
        _check(lib().flashv_plan_upload(self._h, _i(obs)))

    def upload_ptr(self, host_ptr):
        """Upload from a caller-owned host buffer (e.g. pinned torch tensor .data_ptr())."""
        _check(lib().flashv_plan_upload(self._h, C.cast(C.c_void_p(host_ptr), C.POINTER(C.c_int32))))

    def run(self):
        _check(lib().flashv_plan_run(self._h))

    def download(self):
        paths = np.empty((self.batch, self.T), np.int32)
        scores = np.empty(self.batch, np.float32)
        _check(lib().flashv_plan_download(self._h, _i(paths), _f(scores)))
        return paths, scores

    def download_ptr(self, path_ptr, score_ptr):
        _check(lib().flashv_plan_download(self._h, C.cast(C.c_void_p(path_ptr), C.POINTER(C.c_int32)),
                                          C.cast(C.c_void_p(score_ptr), C.POINTER(C.c_float))))

    # ---- state sharding across GPUs (flashv_plan_shard_*) -------------------------------------
    def shard_init(self, rank, world):
        _check(lib().flashv_plan_shard_init(self._h, rank, world))

    def shard_buffers(self):
        """(base address, bytes) of the region this plan's peers store into."""
        d, n = C.c_void_p(), C.c_size_t()
        _check(lib().flashv_plan_shard_buffers(self._h, C.byref(d), C.byref(n)))
        return d.value, n.value

    def shard_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib().flashv_plan_shard_ipc_handle(self._h, buf))
        return buf.raw

    def shard_set_peer(self, peer_rank, peer_device, region_ptr):
        _check(lib().flashv_plan_shard_set_peer(self._h, peer_rank, peer_device, C.c_void_p(region_ptr)))

    def shard_open_peer(self, peer_rank, handle: bytes):
        buf = C.create_string_buffer(handle, 64)
        _check(lib().flashv_plan_shard_open_peer(self._h, peer_rank, buf))

    def report(self):
        rep = Report()
        _check(lib().flashv_plan_report(self._h, C.byref(rep)))
        return rep

    def close(self):
        if self._h:
            lib().flashv_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_count(total, rank, world):
    return int(lib().flashv_shard_count(total, rank, world))


class MultiGpu:
    """Every GPU of the box from one host process (flashv_mgpu_*): replicated model built once with the
    host logarithms split over the devices' threads, batches sharded b mod G, single sequences state-sharded."""

    def __init__(self, ndev, devices=None):
        self._h = C.c_void_p()
        dv = np.ascontiguousarray(devices, np.int32) if devices is not None else None
        _check(lib().flashv_mgpu_create(ndev, _i(dv) if dv is not None else None, C.byref(self._h)))
        self.world = int(lib().flashv_mgpu_world(self._h))
        self.K = self.M = 0

    def model_create(self, A, B, Pi):
        A = np.ascontiguousarray(A, np.float32)
        B = np.ascontiguousarray(B, np.float32)
        Pi = np.ascontiguousarray(Pi, np.float32)
        self.K, self.M = B.shape
        _check(lib().flashv_mgpu_model_create(self._h, self.K, self.M, _f(A), _f(B), _f(Pi)))

    def prep_ms(self, rank=0):
        return lib().flashv_model_prep_ms(lib().flashv_mgpu_model(self._h, rank))

    def decode_batch(self, obs, N, B=0):
        obs = np.ascontiguousarray(obs, np.int32)
        batch, T = obs.shape
        paths = np.empty((batch, T), np.int32)
        scores = np.empty(batch, np.float32)
        rep = Report()
        if B > 0:
            _check(lib().flashv_mgpu_bs_decode_batch(self._h, _i(obs), batch, T, N, B, _i(paths), _f(scores), C.byref(rep)))
        else:
            _check(lib().flashv_mgpu_decode_batch(self._h, _i(obs), batch, T, N, _i(paths), _f(scores), C.byref(rep)))
        return paths, scores, rep

    def decode(self, ob, N):
        ob = np.ascontiguousarray(ob, np.int32)
        path = np.empty(ob.shape[0], np.int32)
        score, rep = C.c_float(), Report()
        _check(lib().flashv_mgpu_decode(self._h, _i(ob), ob.shape[0], N, _i(path), C.byref(score), C.byref(rep)))
        return path, np.float32(score.value), rep

    def close(self):
        if self._h:
            lib().flashv_mgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
