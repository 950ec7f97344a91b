#!/usr/bin/env python3
"""bench.py — headline benchmark of the FLASH decode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--segments P]

One "step" = one FLASH decode (flashv_decode semantics, calc() of the reference) of one synthetic
sequence of the workload BASELINE.json quotes the metric on: K=3965 states, M=50 symbols, T=256,
transition density p=0.112 (the reference driver's own p, src/run.py:13), HMM drawn by the
reference generator's distribution (flash-viterbi_b200/host/gen_hmm.py, seed 1).  With --gpus N
every rank decodes its own sequence on its own GPU (independent sequences shard with no data-path
collective: weak scaling); value = canonical trellis updates of all ranks / max-over-ranks time.

  value      canonical G trellis-updates/s (K^2*T / t), observations already resident in HBM,
             timed with CUDA events on the stream the kernels run on, L2 flushed between decodes
  e2e        same metric through the C-ABI one-call decode with HOST buffers (H2D of the
             observations and D2H of the path inside the timed region); e2e_cold adds the one-time
             model preparation (host libm log tables + upload + layouts), i.e. the reference
             program's own shape: one model, one sequence
  roofline   the full-length pass kernel (k_flash_persist): algorithmic bytes = (T-1)*K^2*4 B per
             launch / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs (SURVEY 8d's
             convention); `physical` holds the on-chip limits that actually bound it
  cpu_baseline  the UNMODIFIED reference binary (oracle/_ref, built by oracle/build_ref.py) timed on
             this box's host cores on a bounded sample (same K, T=34), or the oracle port
  other_configs  BASELINE configs 3, 4 and 5 (FLASH-BS; 8192 batched sequences sharded b mod G;
             K=32768 state-sharded over the ranks through cudaIpc), each with a parity flag
             against the CPU oracle — at EVERY --gpus N (tools/bench_side.py)

--impl reference times the reference's own CPU implementation of the same config (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "flash-viterbi_b200" / "host"))
sys.path.insert(0, str(ROOT / "tools"))

K, M, T, PROB, SEED = 3965, 50, 256, 0.112, 1
SAMPLE_T = 34  # bounded sample of the workload for the cpu_baseline leg of our arm
METRIC = "flash_decode_canonical_trellis_updates_per_s"
UNIT = "G trellis-updates/s (K^2*T/s)"
WORKLOAD = f"FLASH Viterbi K={K} T={T} single sequence (M={M}, p={PROB}, data_script distribution, seed {SEED})"
SIDE_DEADLINE_S = 480  # the side measurements may not hold the headline line back longer than this


def env_int(name, default):
    return int(os.environ.get(name, default))


def executed_steps(T_, N_):
    """S(T,N) of SURVEY 8d, enumerated from the reference's queue logic (F:129-136, F:284-304, F:342)."""
    first = N_ > 2 and T_ >= 2 * N_
    total, level = 0, []
    if first:
        gap, extra, at, lo = (T_ - 1) // N_, (T_ - 1) % N_, 0, 0
        for _ in range(N_ - 1):
            at += gap + (1 if extra else 0)
            extra -= 1 if extra else 0
            level.append((lo, at))
            lo = at + 1
        level.append((lo, T_ - 1))
        total = T_ - 1
    else:
        level = [(0, T_ - 1)]
    while level:
        nxt = []
        for lo, hi in level:
            total += hi - lo
            if hi <= lo + 1:
                continue
            mid = (lo + hi) >> 1
            nxt.append((lo, mid))
            if hi > mid + 1:
                nxt.append((mid + 1, hi))
        level = nxt
    return total


def config_dict(segments):
    """The workload both arms are measured on (identical in both JSON lines)."""
    return {"workload": WORKLOAD, "K": K, "T": T, "M": M, "segments_N": segments, "executed_steps": executed_steps(T, segments),
            "sequences_per_gpu": 1, "sharding": "independent sequences per GPU, no collective",
            "l2": "flushed between decodes (256 MiB fill); the 62.9 MB log table is meant to stay on chip / L2-resident within a decode"}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed regions run (one streaming process,
    a sample every 5 ms — the timed regions are only tens of milliseconds long)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "5"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)  # nvidia-smi needs a moment before its first sample; the timed regions are short
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            for ln in out.splitlines():
                parts = [x.strip() for x in ln.split(",")]
                if len(parts) >= 7:
                    rows.append(parts)
        num = lambda x: x.replace(".", "", 1).isdigit()
        sm = sorted(float(r[0]) for r in rows if num(r[0]))
        pw = [float(r[2]) for r in rows if num(r[2])]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max((float(r[1]) for r in rows if num(r[1])), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


def synthetic(rank):
    import gen_hmm

    A, B, Pi = gen_hmm.make_hmm(K, M, PROB, SEED)
    f = gen_hmm.as_reference_floats
    ob = gen_hmm.observations(T, M, 1000 + rank)
    return f(A), f(B), f(Pi), ob, (A, B, Pi)


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference(raw, T_run, N, runs, warm):
    """Time the unmodified reference binary on the first T_run observations of the workload with
    MAX_THREADS = N.  Returns (list of seconds, cores used, kind, description)."""
    import gen_hmm
    from oracle import build_ref, oracle

    cores = os.cpu_count() or 1
    name = build_ref.binary_name("FLASH", K, M, T_run, PROB, N)
    binary = build_ref.OUT_DIR / name
    if not binary.exists() and build_ref.reference_available():
        binary = build_ref.build("FLASH", K, M, T_run, PROB, N)
    ob = gen_hmm.observations(T, M, 1000)[:T_run]
    times = []
    if binary.exists():
        work = Path(tempfile.mkdtemp(prefix="flashv_ref_"))
        try:
            A, B, Pi = raw
            gen_hmm.write_text(work / "data", K, T_run, PROB, A, B, Pi, ob)
            for it in range(warm + runs):
                r = build_ref.run(binary, work, timeout=900)
                if it >= warm:
                    times.append(r["time"])  # the program's own "time:" line: calc() only, F:375-378
        finally:
            shutil.rmtree(work, ignore_errors=True)
        kind = "reference"
        what = (f"unmodified FLASH_Viterbi_multithread.c built per src/run.py:54 (gcc -g), K={K}, T={T_run}"
                f"{'' if T_run == T else ' prefix of the workload'}, MAX_THREADS={N}, its own 'time:' line")
        used = min(N, cores)
    else:
        A, B, Pi, _, _ = synthetic(0)
        oracle.set_threads(cores)
        om = oracle.OracleModel(A, B, Pi)
        for it in range(warm + runs):
            t0 = time.perf_counter()
            om.flash(ob, N)
            if it >= warm:
                times.append(time.perf_counter() - t0)
        kind = "port"
        what = (f"oracle/flashv_oracle.c (gcc -O2, OpenMP over destination states, log tables hoisted), K={K}, "
                f"T={T_run}, N={N}; reference binary not prebuilt in oracle/_ref")
        used = cores
    return times, used, kind, what


def run_reference_arm(args):
    """The reference's own CPU implementation on OUR config: full T=256, MAX_THREADS = our segment count.
    One run takes ~45 s on the GPU box's host (the N-way pass is single-threaded by design, F:347), so
    the run count is capped at 2 and there is no warm-up."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    runs = max(1, min(args.steps, 2))
    _, _, _, _, raw = synthetic(0)
    t_all = time.perf_counter()
    times, cores, kind, what = cpu_reference(raw, T, args.segments, runs, 0)
    mean = sum(times) / len(times)
    value = K * K * T / mean / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": 0, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 storage, f64 log/add (reference arithmetic)", "data": "synthetic",
        "config": config_dict(args.segments),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": what + f"; the whole config, {len(times)} run(s) (requested steps {args.steps}: capped, one run is {mean:.0f} s)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# side measurements on the headline model (reported beside the headline line, never as `value`)
# ------------------------------------------------------------------------------------------------
def sequential_steps(fv, T_, N_):
    """Trellis steps on the critical path of one decode: the N-way pass plus the longest task of every
    level of the task tree (tasks of a level run side by side on the device)."""
    tasks, first_pass, mids = fv.task_list(T_, N_)
    if first_pass:
        bounds = [0] + [m + 1 for m in mids]
        level = [(lo, hi) for lo, hi in zip(bounds, list(mids) + [T_ - 1])]
    else:
        level = [(0, T_ - 1)]
    total = T_ - 1 if first_pass else 0
    while level:
        total += max(r - l for l, r in level)
        nxt = []
        for l, r in level:
            if r <= l + 1:
                continue
            mid = (l + r) >> 1
            nxt.append((l, mid))
            if r > mid + 1:
                nxt.append((mid + 1, r))
        level = nxt
    return total


def headline_side_configs(fv, ctx, model, ob, stream, torch, om, om_full):
    """FLASH at the other segment counts, the opt-in sparse engine and FLASH-BS (config 3) on the headline
    model and sequence, each with a parity flag against the oracle."""
    bits = lambda x: np.asarray(x, np.float32).view(np.uint32)

    def timed(plan, runs):
        plan.run()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(runs):
                plan.run()
            e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / runs

    also, out = {}, {}
    # N=1 is the reference driver's own setting (src/run.py:14: MAX_THREADS 1: no N-way pass, a 9-level tree)
    for n_other, engine, key in ((1, fv.ENGINE_AUTO, "N=1"), (8, fv.ENGINE_AUTO, "N=8"), (64, fv.ENGINE_AUTO, "N=64")):
        try:
            p2 = fv.Plan(model, T, n_other, 1, 0, engine)
            p2.upload(ob[None, :])
            ms = timed(p2, 3)
            paths, scores = p2.download()
            want, wscore, _ = om.flash(ob, n_other)
            rep = p2.report()
            also[key] = {"ms_per_decode": ms, "value": K * K * T / (ms * 1e-3) / 1e9, "executed_steps": rep.executed_steps,
                         "kernel_launches": rep.kernel_launches,
                         "parity": bool(np.array_equal(paths[0], want) and bits(scores[0]) == bits(wscore))}
            p2.close()
        except Exception as e:
            also[key] = f"failed: {e}"

    # the headline decode on the opt-in sparse engine (in-edge lists resident in shared memory): same
    # results, but not the dense table's bytes, so it is reported here and never as `value`
    try:
        ps = fv.Plan(model, T, 127, 1, 0, fv.ENGINE_SPARSE)
        ps.upload(ob[None, :])
        ms = timed(ps, 5)
        rs = ps.report()
        paths, scores = ps.download()
        want, wscore, _ = om.flash(ob, 127)
        out["flash_sparse_engine_N127"] = {"workload": WORKLOAD + ", N=127, ENGINE_SPARSE (edges with A[k][i] == 0 skipped: they can never win)",
                                           "ms_per_decode": ms, "value": K * K * T / (ms * 1e-3) / 1e9, "unit": UNIT,
                                           "first_pass_ms": rs.first_pass_ms, "us_per_step_first_pass": rs.first_pass_ms * 1e3 / (T - 1),
                                           "parity": bool(np.array_equal(paths[0], want) and bits(scores[0]) == bits(wscore))}
        ps.close()
    except Exception as e:
        out["flash_sparse_engine_N127"] = f"failed: {e}"

    # config 3: FLASH-BS, same model and sequence (latency-bound: report ms and us per sequential step);
    # (N=1, B=32) is the reference driver's own setting (src/run.py:10-16), whose path holds -1 entries
    for n_seg, beam in ((8, 128), (127, 128), (1, 32)):
        key = f"flash_bs_B{beam}_N{n_seg}"
        try:
            p3 = fv.Plan(model, T, n_seg, 1, beam, fv.ENGINE_AUTO)
            p3.upload(ob[None, :])
            ms = timed(p3, 3)
            paths, scores = p3.download()
            want, wscore, _ = om_full.flash_bs(ob, n_seg, beam)
            seq_steps = sequential_steps(fv, T, n_seg)
            out[key] = {"workload": f"FLASH-BS K={K} T={T} B={beam} N={n_seg}, single sequence", "ms_per_decode": ms,
                        "sequential_steps": seq_steps, "us_per_sequential_step": ms * 1e3 / seq_steps,
                        "bound": "latency (K x B dependent double reads + beam selection per step)",
                        "dropped_out_entries": int((want < 0).sum()),
                        "parity": bool(np.array_equal(paths[0], want) and bits(scores[0]) == bits(wscore))}
            p3.close()
        except Exception as e:
            out[key] = f"failed: {e}"
    return also, out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from __graft_entry__ import load_pkg

    fv = load_pkg()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    import shard

    def max_over_ranks(x):
        return shard.max_over_ranks(x, dist, dev)

    A, B, Pi, ob, raw = synthetic(rank)
    stream = torch.cuda.Stream(device=dev)
    ctx = fv.Context(local, stream.cuda_stream)
    # cold, one-shot: what a reference-shaped program pays for ONE sequence (model preparation included)
    t0 = time.perf_counter()
    model = fv.Model(ctx, A, B, Pi)
    t_model = time.perf_counter() - t0
    cold_path, cold_score, _ = model.decode(ob, args.segments)
    t_cold = time.perf_counter() - t0
    engine = {"auto": fv.ENGINE_AUTO, "step": fv.ENGINE_STEP, "persistent": fv.ENGINE_PERSISTENT, "sparse": fv.ENGINE_SPARSE}[args.engine]
    plan = fv.Plan(model, T, args.segments, 1, 0, engine)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    ob_pinned = torch.from_numpy(ob.copy()).pin_memory()
    plan.upload_ptr(ob_pinned.data_ptr())
    for _ in range(args.warmup):
        plan.run()
    ctx.sync()
    paths, scores = plan.download()

    # ---- value: kernels only, observations resident, CUDA events on the launching stream -------
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    fp_ms, launches = [], 0
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            flush.fill_(k & 0xFF)  # evict the log table from L2 between decodes
            starts[k].record(stream)
            plan.run()
            ends[k].record(stream)
            rep = plan.report()  # syncs on the run's end event
            fp_ms.append(rep.first_pass_ms)
            launches += rep.kernel_launches
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    dev_ms = max_over_ranks(dev_ms)
    value = world * K * K * T * args.steps / (dev_ms * 1e-3) / 1e9

    # ---- e2e: the one-call C-ABI decode with host buffers ---------------------------------------
    for _ in range(min(args.warmup, 3)):
        model.decode(ob, args.segments)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_path, e2e_score, e2e_rep = model.decode(ob, args.segments)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop()
    e2e_value = world * K * K * T * args.steps / e2e_s / 1e9
    t_cold = max_over_ranks(t_cold)

    rep = plan.report()
    peak, peak_src = peaks()
    fp_mean = sum(fp_ms) / len(fp_ms)
    algo_bytes = (T - 1) * K * K * 4.0
    achieved = algo_bytes / (fp_mean * 1e-3) / 1e9 if fp_mean > 0 else None
    traffic, traffic_src = None, None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists():
        try:
            tj = json.loads(tfile.read_text())
            traffic, traffic_src = tj.get("k_flash_persist_dram_bytes_per_launch"), tj.get("source")
        except Exception:
            traffic = None
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    # which sweep the pass kernel runs: models up to 4096 states take the half-precision filter (k_flash_persist16)
    half_filter = rep.engine == 2 and K <= 4096 and os.environ.get("FLASHV_F16", "1") != "0"
    Kp16 = (K + 255) // 256 * 256
    physical = None
    if fp_mean > 0 and half_filter:
        lane_instr = (T - 1) * float(K) * K * 1.0  # HADD2 + HMNMX2 per two updates
        tmem_bytes = (K / 148.0) * Kp16 * 2
        smem_bytes = 14 * Kp16 * 2
        step_s = fp_mean * 1e-3 / (T - 1)
        physical = {
            "note": "the (half)log A table lives in tensor memory for the whole launch and HBM is touched only by the exact re-evaluation "
                    "(one 128-byte chain of doubles per column and step), so the HBM convention exceeds 1; these are the on-chip limits of the pass kernel",
            "f16x2_issue_frac": lane_instr / (fp_mean * 1e-3) / (148 * 128 * sm_hz),
            "f16x2_issue_basis": "1 lane-instruction per update (HADD2 and HMNMX2 each cover two updates) against 148 SMs x 128 lanes x SM clock",
            "us_per_step": fp_mean * 1e3 / (T - 1),
            "tensor_memory_bytes_per_step_per_sm": int(tmem_bytes),
            "shared_memory_bytes_per_step_per_sm": int(smem_bytes),
            "on_chip_bytes_per_clk_per_sm_over_the_whole_step": (tmem_bytes + smem_bytes) / step_s / sm_hz,
            "operand_path_basis": "table slice (2 B per update) read from tensor memory, 14 warps' half-precision delta reads from shared memory; the sweep itself is "
                                  "about 1.0 us of the step (profiles/r02_persist16_phase_trace.txt: ~110 B/clk per SM out of tensor memory), the rest is the hand-over "
                                  "(publish -> L2 -> poll, one CTA barrier) and the exact evaluation (one HBM round trip)",
        }
    elif fp_mean > 0:
        fp32_lane_ops = (T - 1) * float(K) * K * 3
        physical = {
            "note": "the table is served from tensor memory + shared memory / L2, not HBM (see traffic), so the HBM convention can exceed 1; "
                    "these are the on-chip limits of the pass kernel",
            "fp32_issue_frac": fp32_lane_ops / (fp_mean * 1e-3) / (148 * 128 * sm_hz),
            "fp32_issue_basis": "3 lane-operations per update (FADD, FADD, FMNMX) against 148 SMs x 128 lanes x SM clock",
            "us_per_step": fp_mean * 1e3 / (T - 1),
            "operand_bytes_per_step_per_sm": int((K / 148.0) * K * 4 + 14 * K * 4),
            "operand_path_frac": ((K / 148.0) * K * 4 + 14 * K * 4) / (fp_mean * 1e-3 / (T - 1)) / ((128 + 64) * sm_hz),
            "operand_path_basis": "table slice + 14 warps' delta reads per step per SM against shared memory 128 B/clk + tensor memory ~64 B/clk",
        }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("f16 filter" if half_filter else "f32 filter") + ", f64 exact evaluation (reference arithmetic, bit-exact)", "data": "synthetic",
        "config": config_dict(args.segments),
        "engine": {1: "step", 2: "persistent", 3: "sparse"}.get(rep.engine),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(T * 4), "d2h_bytes_per_step": int(T * 4 + 4),
                "ms_per_step": e2e_s * 1e3 / args.steps,
                "note": "model resident (amortised over the sequences of one HMM); e2e_cold is the one-model-one-sequence shape of the reference program"},
        "e2e_cold": {"ms": t_cold * 1e3, "value": world * K * K * T / t_cold / 1e9, "unit": UNIT, "model_create_ms": t_model * 1e3,
                     "includes": "flashv_model_create (host libm log tables of all K^2 entries, upload, device layouts) + the first flashv_decode "
                                 "(plan creation, H2D, kernels, D2H); the reference's timed calc() likewise includes every log() call (F:170)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": ("k_flash_persist16" if half_filter else "k_flash_persist") + " (full-length pass, one launch)" if rep.engine == 2 else "k_flash_step x (T-1)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "ms_per_launch": fp_mean,
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "note": "SURVEY 8d convention: 4 B per executed update of the full-length pass only ((T-1)*K^2 updates per launch); the tree "
                             "levels are excluded because one-column last steps and shared table reads skip work the convention would count",
                     "physical": physical},
        "model_prep_ms": model.prep_ms,
    }

    # ---- parity of the timed decodes + CPU baseline (rank 0, one GPU) -----------------------------
    om = om_full = None
    if rank == 0 and not args.no_cpu:
        try:
            from oracle import oracle

            oracle.set_threads(os.cpu_count() or 1)
            om = oracle.OracleModel(A, B, Pi, lean=True)  # the edge-list form (== the literal loops, tests/test_oracle_golden.py)
            om_full = oracle.OracleModel(A, B, Pi)        # FLASH-BS needs the literal tables
            want, wscore, _ = om.flash(ob, args.segments)
            same = lambda p, s: bool(np.array_equal(p, want) and np.float32(s).view(np.uint32) == np.float32(wscore).view(np.uint32))
            line["parity"] = same(paths[0], scores[0]) and same(e2e_path, e2e_score) and same(cold_path, cold_score)
        except Exception as e:  # the checker failing must not hide the measurement
            line["parity"] = f"unchecked: {e}"
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            n_cpu = 16 if (os.cpu_count() or 1) >= 16 else 8
            times, cores, kind, what = cpu_reference(raw, SAMPLE_T, n_cpu, 1, 0)
            mean = sum(times) / len(times)
            line["cpu_baseline"] = {"value": K * K * SAMPLE_T / mean / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": what + f"; {mean:.2f} s (the whole config: bench.py --impl reference)"}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {e}"}

    # ---- side measurements; a deadline keeps them from holding the headline line back -------------
    printed = threading.Lock()

    def emit():
        if printed.acquire(blocking=False) and rank == 0:
            print(json.dumps(line), flush=True)

    def deadline():
        line.setdefault("other_configs", {})["deadline"] = f"side measurements cut off after {SIDE_DEADLINE_S} s"
        emit()
        os._exit(0)

    timer = threading.Timer(SIDE_DEADLINE_S, deadline)
    timer.daemon = True
    timer.start()
    extras = {}
    if not args.no_extras:
        if rank == 0 and om is not None:
            try:
                line["other_segment_counts"], extras = headline_side_configs(fv, ctx, model, ob, stream, torch, om, om_full)
            except Exception as e:
                extras = {"headline_side_configs": f"failed: {e}"}
        om = om_full = None
        import bench_side

        ctrl = bench_side.Ctrl(dist, rank, world)
        for key, fn in (("config4_batch_sharded", lambda: bench_side.config4_batch_sharded(fv, ctx, ctrl, torch, stream)),
                        ("config5_state_sharded", lambda: bench_side.config5_state_sharded(fv, ctx, ctrl))):
            if os.environ.get("FLASHV_BENCH_SKIP", "").find(key[:7]) >= 0:
                continue
            t0 = time.time()
            try:
                res = fn()
                res["wall_s"] = time.time() - t0
                ok = True
            except Exception as e:
                res, ok = f"failed on rank {rank}: {type(e).__name__}: {e}", False
            # a rank that failed must not leave the others waiting inside the next collective
            oks = ctrl.gather_objects((ok, res if not ok else None))
            if not all(o for o, _ in oks):
                res = "; ".join(str(r) for o, r in oks if not o)
                extras[key] = res
                break
            extras[key] = res
    line["other_configs"] = extras
    timer.cancel()
    emit()
    plan.close()
    model.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--segments", type=int, default=127,
                    help="MAX_THREADS of the reference = segment count N (127 is the largest the reference handles at T=256: "
                         "T == 2N is broken there); N=1, 8 and 64 are timed beside it")
    ap.add_argument("--engine", default="auto", choices=["auto", "step", "persistent", "sparse"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the parity check and the CPU baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (other segment counts, FLASH-BS, configs 4 and 5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
