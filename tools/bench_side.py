"""Side measurements of bench.py: BASELINE.json configs 4 and 5 on the GPUs of one box, one process per
GPU (torchrun), reported under `other_configs` beside the headline line.

  config 4   batched FLASH, K=512 T=1024, 8192 sequences of one HMM: sequence b -> rank b mod G
             (flashv_decode_batch_shard), no data-path collective, paths gathered on rank 0
  config 5   one sequence, K=32768 T=4096: pass 0 state-sharded (per-step delta / backpointer exchange with
             in-kernel peer stores over NVLink), tree levels spread over the ranks; the model's host
             logarithms are split over the ranks and exchanged over NVLink; buffers shared by cudaIpc

Every function is collective: all ranks call it with the same arguments.  Control-plane exchanges (handles,
flags, times) go through a gloo group so that a rank whose CUDA context died cannot wedge the others in
NCCL; the data path has no collective at all.  Parity is against the CPU oracle (test infrastructure,
imported here only as the checker): rank 0 decodes, every rank compares its own result.
"""
from __future__ import annotations

import os
import time
from pathlib import Path

import numpy as np

M = 50
SEED = 1


class Ctrl:
    """Control plane of the side measurements: CPU (gloo) collectives, or no-ops on one rank."""

    def __init__(self, dist, rank, world):
        self.dist, self.rank, self.world = dist, rank, world
        self.group = dist.new_group(backend="gloo") if dist is not None and world > 1 else None

    def barrier(self):
        if self.group is not None:
            self.dist.barrier(group=self.group)

    def gather_objects(self, obj):
        if self.group is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def bcast_object(self, obj):
        if self.group is None:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0, group=self.group)
        return box[0]

    def fmax(self, x):
        return max(self.gather_objects(float(x)))

    def all_true(self, ok):
        return all(self.gather_objects(bool(ok)))


def _bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------------------------
# config 4
# ------------------------------------------------------------------------------------------------------
def config4_batch_sharded(fv, ctx, ctrl, torch, stream, total=8192, K=512, T=1024, N=32, runs=2, check=3):
    """8192 sequences over the ranks (strong scaling: the batch is fixed), through the C-ABI call with HOST
    buffers (flashv_decode_batch_shard: H2D of the rank's observations and D2H of its paths inside the
    timed region), max over ranks.  Every rank checks `check` of its sequences against the oracle; rank 0
    gathers all paths and checks the row bookkeeping."""
    import gen_hmm
    from oracle import oracle

    rank, world = ctrl.rank, ctrl.world
    A, B, Pi = gen_hmm.make_hmm(K, M, 0.253, SEED)
    f = gen_hmm.as_reference_floats
    A, B, Pi = f(A), f(B), f(Pi)
    model = fv.Model(ctx, A, B, Pi)
    obs = gen_hmm.observation_batch(total, T, M, 1000)
    mine = fv.shard_count(total, rank, world)
    paths = np.full((total, T), -2, np.int32)
    scores = np.zeros(total, np.float32)
    model.decode_batch_shard(obs, N, rank, world, paths, scores)  # warm-up: plan, pinned staging
    e2e, dev = [], []
    launches = 0
    for _ in range(runs):
        ctrl.barrier()
        t0 = time.perf_counter()
        _, _, rep = model.decode_batch_shard(obs, N, rank, world, paths, scores)
        e2e.append(ctrl.fmax(time.perf_counter() - t0))
        dev.append(ctrl.fmax(rep.decode_ms))
        launches += rep.kernel_launches
    # parity: a few of this rank's sequences against the oracle, then the gather
    om = oracle.OracleModel(A, B, Pi)
    ok = True
    for q in range(min(check, mine)):
        b = rank + (q * max(1, mine // max(1, check))) * world
        want, wscore, _ = om.flash(obs[b], N)
        ok &= bool(np.array_equal(paths[b], want)) and bool(_bits(scores[b]) == _bits(wscore))
    rows = ctrl.gather_objects((rank, paths[rank::world].copy()))
    if rank == 0:
        full = np.full((total, T), -2, np.int32)
        for r, part in rows:
            full[r::world] = part
        ok &= bool((full >= 0).all()) and bool(np.array_equal(full[rank::world], paths[rank::world]))
    # weak scaling (the north star's batched target): the config's 8192 sequences on EVERY GPU
    weak = None
    if world > 1:
        wobs = gen_hmm.observation_batch(total, T, M, 1000 + total * rank)
        wpaths, wscores, _ = model.decode_batch(wobs, N)  # warm-up: the plan for this batch size
        we2e, wdev = [], []
        for _ in range(runs):
            ctrl.barrier()
            t0 = time.perf_counter()
            wpaths, wscores, wrep = model.decode_batch(wobs, N)
            we2e.append(ctrl.fmax(time.perf_counter() - t0))
            wdev.append(ctrl.fmax(wrep.decode_ms))
        wok = True
        for b in (0, total // 2, total - 1):
            want, wscore, _ = om.flash(wobs[b], N)
            wok &= bool(np.array_equal(wpaths[b], want)) and bool(_bits(wscores[b]) == _bits(wscore))
        ok &= wok
        wcanon = world * total * float(K) * K * T
        weak = {"scaling": "weak", "sequences_per_gpu": total, "sequences_total": total * world, "ms_per_batch_device": min(wdev),
                "ms_per_batch_e2e": min(we2e) * 1e3, "value": wcanon / (min(wdev) * 1e-3) / 1e9, "e2e_value": wcanon / min(we2e) / 1e9,
                "parity": ctrl.all_true(wok), "note": "efficiency = ms at 1 GPU / ms here (per-GPU work fixed); the strong-scaling line above "
                "leaves a GPU less than one wave of the group kernel (296 groups of 8 sequences) from 4 GPUs on"}
    # the same batch at the largest segment count the reference accepts at T=1024: every task is one or two steps
    # long and a task's last step needs one column, so the tree costs almost nothing
    n_big = 511
    bp, bsc, _ = model.decode_batch_shard(obs, n_big, rank, world)
    ctrl.barrier()
    bp, bsc, brep = model.decode_batch_shard(obs, n_big, rank, world)
    big_ms = ctrl.fmax(brep.decode_ms)
    bok = True
    for q in range(min(2, mine)):
        b = rank + q * world
        want, wscore, _ = om.flash(obs[b], n_big)
        bok &= bool(np.array_equal(bp[b], want)) and bool(_bits(bsc[b]) == _bits(wscore))
    ok &= bok
    parity = ctrl.all_true(ok)
    steps = model_steps = rep.executed_steps
    model.close()
    e2e_s, dev_ms = min(e2e), min(dev)
    canon = total * float(K) * K * T
    clk_hz = 1.965e9
    fp32_peak = world * 148 * 128 * clk_hz
    executed = steps * float(K) * K * total
    # what the kernels actually run: a task's last step needs one destination column (K updates, not K^2) unless
    # the pass is full-range, so every queue task skips one K^2 step
    n_tasks = rep.n_tasks
    executed_run = (steps - n_tasks) * float(K) * K * total
    return {
        "workload": f"batched FLASH K={K} T={T} N={N}: {total} sequences of one HMM, sequence b on GPU b mod {world} "
                    f"(flashv_decode_batch_shard), paths gathered on rank 0; strong scaling",
        "n_gpus": world, "sequences_total": total, "sequences_per_gpu": mine if world == 1 else [fv.shard_count(total, r, world) for r in range(world)],
        "ms_per_batch_device": dev_ms, "ms_per_batch_e2e": e2e_s * 1e3,
        "value": canon / (dev_ms * 1e-3) / 1e9, "e2e_value": canon / e2e_s / 1e9, "unit": "G trellis-updates/s (K^2*T/s)",
        "executed_steps_per_sequence": int(model_steps), "gpu_launches": launches,
        "h2d_bytes_per_batch": int(total * T * 4), "d2h_bytes_per_batch": int(total * T * 4 + total * 4),
        "roofline": {"bound": "fp32 pipe (SURVEY 8d: 3 lane-operations per executed update), all GPUs",
                     "achieved": executed * 3 / (dev_ms * 1e-3) / 1e12, "peak": fp32_peak / 1e12, "unit": "T lane-op/s",
                     "frac": executed * 3 / (dev_ms * 1e-3) / fp32_peak,
                     "frac_on_executed_work": executed_run * 3 / (dev_ms * 1e-3) / fp32_peak,
                     "note": f"frac counts the reference's {int(steps)} steps per sequence; the kernels run {int(steps - n_tasks)} K^2 steps "
                             f"(the last step of each of the {int(n_tasks)} queue tasks is one column) at 2.2 issued instructions per update"},
        "parity": parity, "parity_checked": f"{check} sequences per rank vs the CPU oracle (path + score bits); all {total} rows gathered and complete",
        "N511": {"segments_N": n_big, "ms_per_batch_device": big_ms, "value": canon / (big_ms * 1e-3) / 1e9, "executed_steps_per_sequence": int(brep.executed_steps),
                 "parity": ctrl.all_true(bok)},
        "weak_scaling": weak,
    }


# ------------------------------------------------------------------------------------------------------
# config 5
# ------------------------------------------------------------------------------------------------------
def shared_hmm(ctrl, K, prob, seed):
    """A[K][K] float32 in /dev/shm, generated once per box by rank 0 (the reference generator's numbers,
    gen_hmm.transition_matrix_f32_into) and mapped read-only by every rank — and by later bench runs on
    the same box."""
    import gen_hmm

    path = Path("/dev/shm") / f"flashv_hmm_K{K}_p{prob}_s{seed}.f32"
    done = Path(str(path) + ".ok")
    gen_s = 0.0
    if ctrl.rank == 0 and not done.exists():
        t0 = time.time()
        out = np.lib.format.open_memmap(str(path) + ".npy", mode="w+", dtype=np.float32, shape=(K, K))
        gen_hmm.transition_matrix_f32_into(out, K, prob, seed)
        out.flush()
        del out
        os.replace(str(path) + ".npy", str(path))
        done.write_text("ok")
        gen_s = time.time() - t0
    ctrl.barrier()
    A = np.load(str(path), mmap_mode="r")
    assert A.shape == (K, K) and A.dtype == np.float32
    return A, gen_s


def sharded_model(fv, ctx, ctrl, A, B, Pi):
    """Model creation shared by the ranks: own rows of host logarithms, the rest over NVLink (cudaIpc)."""
    rank, world = ctrl.rank, ctrl.world
    t0 = time.time()
    model = fv.Model.create_rows(ctx, A, B, Pi, rank, world)
    handles = ctrl.gather_objects(model.rows_handle())  # also the barrier: every rank's rows are on its device
    for q in range(world):
        if q != rank:
            model.pull_rows(q, handles[q])
    ctrl.barrier()  # peers may still be reading this rank's table
    model.finish()
    return model, ctrl.fmax(time.time() - t0)


class ShardedPlan:
    """A state-sharded plan connected to its peers through cudaIpc handles."""

    def __init__(self, fv, model, ctx, ctrl, T, N):
        self.ctx, self.ctrl = ctx, ctrl
        self.plan = fv.Plan(model, T, N, 1, 0, fv.ENGINE_PERSISTENT)
        self.plan.shard_init(ctrl.rank, ctrl.world)
        if ctrl.world > 1:
            hs = ctrl.gather_objects(self.plan.shard_ipc_handle())
            for q in range(ctrl.world):
                if q != ctrl.rank:
                    self.plan.shard_open_peer(q, hs[q])

    def decode(self, ob):
        """One collective decode: (path, score, report).  Sync + barrier first — the sharded run's contract."""
        self.plan.upload(ob)
        self.ctx.sync()
        self.ctrl.barrier()
        self.plan.run()
        paths, scores = self.plan.download()
        return paths[0], scores[0], self.plan.report()

    def close(self):
        self.ctrl.barrier()  # nobody unmaps a region a peer may still be storing into
        self.plan.close()


def config5_state_sharded(fv, ctx, ctrl, K=32768, T=4096, N=2047, prob=0.112, runs=2, prefix_T=64, prefix_N=(31, 8),
                          mid=(512, 8), oracle_threads=None):
    import gen_hmm
    from oracle import oracle

    rank, world = ctrl.rank, ctrl.world
    out = {"workload": f"FLASH K={K} T={T} single sequence (p={prob}, data_script distribution, seed {SEED}), destination states of "
                       f"pass 0 sharded over {world} GPU(s) with in-kernel peer stores, tree levels spread over the ranks; cudaIpc, no NCCL on the data path",
           "n_gpus": world}
    A, gen_s = shared_hmm(ctrl, K, prob, SEED)
    B = gen_hmm.as_reference_floats(gen_hmm.emission_matrix(K, M, SEED))
    Pi = gen_hmm.as_reference_floats(np.full(K, 1 / K))
    os.environ.setdefault("FLASHV_NO_SPARSE", "1")  # the edge lists (2.4 GB, unused by the dense engines) are skipped at this size
    model, prep_s = sharded_model(fv, ctx, ctrl, A, B, Pi)
    out["hmm_generation_s"] = gen_s
    out["model_prep_s"] = prep_s
    out["model_prep"] = f"host logarithms of K/{world} rows per rank, other rows pulled over NVLink, layouts built on the device"
    ob = gen_hmm.observations(T, M, 1000)

    def timed(Tn, Nn, nruns):
        sp = ShardedPlan(fv, model, ctx, ctrl, Tn, Nn)
        path, score, rep = sp.decode(ob[:Tn])  # warm-up
        best, best_fp, launches = None, None, 0
        for _ in range(nruns):
            path, score, rep = sp.decode(ob[:Tn])
            d, fp = ctrl.fmax(rep.decode_ms), ctrl.fmax(rep.first_pass_ms)
            launches = rep.kernel_launches
            if best is None or d < best:
                best, best_fp = d, fp
        sp.close()
        return path, score, best, best_fp, rep.executed_steps, launches

    # the config itself
    path, score, ms, fp_ms, steps, launches = timed(T, N, runs)
    same = ctrl.gather_objects((path.tobytes(), float(score)))
    out.update({
        "T": T, "segments_N": N, "ms_per_decode": ms, "first_pass_ms": fp_ms, "executed_steps": int(steps), "gpu_launches": launches,
        "value": float(K) * K * T / (ms * 1e-3) / 1e9, "unit": "G trellis-updates/s (K^2*T/s)",
        "us_per_step_first_pass": fp_ms * 1e3 / (T - 1),
        "algorithmic_GBps_first_pass_all_gpus": (T - 1) * float(K) * K * 4 / (fp_ms * 1e-3) / 1e9,
        "all_ranks_same_path": all(s == same[0] for s in same),
    })
    # the tree levels spread over the ranks: a mid-N shape where the levels are most of the work
    if mid:
        Tm, Nm = mid
        pm, sm, msm, fpm, stm, _ = timed(Tm, Nm, 1)
        out["mid_N"] = {"T": Tm, "segments_N": Nm, "ms_per_decode": msm, "first_pass_ms": fpm, "levels_ms": msm - fpm,
                        "executed_steps": int(stm), "note": "prefix of the same sequence; levels_ms shrinks with the GPU count because rank r runs every world-th task of a level (the same spread-level path the N=8 prefix parity check decodes)"}
    # parity at this K: a prefix the CPU oracle finishes in seconds, decoded by the same sharded engine
    oracle.set_threads(oracle_threads or os.cpu_count() or 1)
    t0 = time.time()
    om = oracle.OracleModel(A, B, Pi, lean=True) if rank == 0 else None
    ok, detail = True, []
    for Np in prefix_N:
        want = None
        if rank == 0:
            w, ws, _ = om.flash(ob[:prefix_T], Np)
            want = (w, float(ws))
        want = ctrl.bcast_object(want)
        sp = ShardedPlan(fv, model, ctx, ctrl, prefix_T, Np)
        got, sc, _ = sp.decode(ob[:prefix_T])
        sp.close()
        good = bool(np.array_equal(got, want[0])) and bool(_bits(sc) == _bits(np.float32(want[1])))
        ok &= good
        detail.append(f"T={prefix_T} N={Np}")
    del om
    out["parity"] = ctrl.all_true(ok)
    out["parity_checked"] = "every rank's path and score bits vs the CPU oracle (edge-list form) on prefixes " + ", ".join(detail)
    out["oracle_s"] = time.time() - t0
    model.close()
    return out
