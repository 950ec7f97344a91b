"""Multi-GPU forms of the path against the oracle, through the C ABI:

  * one process driving all GPUs (flashv_mgpu_*: what a C host program uses) — also runs on a 1-GPU box,
    where it still exercises model creation in parts, the shard bookkeeping and the gather;
  * one process per GPU with cudaIpc handles exchanged over torch.distributed (the torchrun form bench.py
    uses): model rows pulled over NVLink, batch sharding, the state-sharded pass 0 and the task-tree levels
    spread over the ranks.  Needs 2 GPUs; skipped otherwise."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, random_hmm

pytestmark = pytest.mark.gpu


def _bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def _device_count():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("K,T,N,world", [(300, 40, 4, 2), (3965, 24, 3, 2), (8200, 10, 2, 2), (1000, 30, 1, 4),
                                         (520, 200, 8, 2), (640, 129, 5, 4)])
def test_state_sharded_decode(fv, oracle_mod, K, T, N, world):
    """Destination states of pass 0 sharded over `world` GPUs (per-step delta exchange with in-kernel peer
    stores), tree levels spread over the ranks with the Ans[] exchange between levels.  One process drives
    all GPUs here; every rank must end with the reference's path."""
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    A, B, Pi = random_hmm(K, 6, 0.01 if K > 5000 else (0.05 if K > 2000 else 0.2), 81)
    om = oracle_mod.OracleModel(A, B, Pi)
    ob = np.random.RandomState(81).randint(0, 6, T).astype(np.int32)
    want, wscore, _ = om.flash(ob, N)
    ctxs = [fv.Context(r) for r in range(world)]
    models = [fv.Model(c, A, B, Pi) for c in ctxs]
    plans = [fv.Plan(m, T, N, 1, 0, fv.ENGINE_PERSISTENT) for m in models]
    for r, p in enumerate(plans):
        p.shard_init(r, world)
    bufs = [p.shard_buffers() for p in plans]
    for r, p in enumerate(plans):
        for q in range(world):
            if q != r:
                p.shard_set_peer(q, q, bufs[q][0])
    for rep in range(3):  # several runs: the run tags must keep them apart
        for p in plans:
            p.upload(ob)
        for c in ctxs:
            c.sync()
        for p in plans:
            p.run()  # asynchronous: the ranks' kernels wait for each other's slices
        for r, p in enumerate(plans):
            paths, scores = p.download()
            assert np.array_equal(paths[0], want), (rep, r, np.nonzero(paths[0] != want)[0][:6])
            assert _bits(scores[0]) == _bits(wscore)
    for p in plans:
        p.close()
    for m in models:
        m.close()
    for c in ctxs:
        c.close()


def test_one_process_all_gpus(fv, oracle_mod):
    """flashv_mgpu_*: model built once (host logarithms split over the devices' threads, rows exchanged over
    NVLink), a batch sharded b mod G and gathered, one sequence state-sharded — all against the oracle."""
    world = max(1, min(_device_count(), 8))
    K, M, T = 333, 7, 48
    A, B, Pi = random_hmm(K, M, 0.2, 91)
    om = oracle_mod.OracleModel(A, B, Pi)
    rng = np.random.RandomState(91)
    obs = rng.randint(0, M, (13, T)).astype(np.int32)
    g = fv.MultiGpu(world)
    assert g.world == world
    g.model_create(A, B, Pi)
    for N in (3, 1):
        paths, scores, rep = g.decode_batch(obs, N)
        for b in range(len(obs)):
            want, wscore, _ = om.flash(obs[b], N)
            assert np.array_equal(paths[b], want), (N, b)
            assert _bits(scores[b]) == _bits(wscore)
        assert rep.kernel_launches > 0
    paths, scores, _ = g.decode_batch(obs, 3, B=9)  # FLASH-BS batches shard the same way
    for b in range(len(obs)):
        want, wscore, _ = om.flash_bs(obs[b], 3, 9)
        assert np.array_equal(paths[b], want) and _bits(scores[b]) == _bits(wscore), b
    for N in (5, 2):
        path, score, rep = g.decode(obs[0], N)
        want, wscore, _ = om.flash(obs[0], N)
        assert np.array_equal(path, want) and _bits(score) == _bits(wscore), N
    # a second model replaces the first (plans of the old one are dropped with it)
    A2, B2, Pi2 = random_hmm(200, M, 0.3, 92)
    g.model_create(A2, B2, Pi2)
    om2 = oracle_mod.OracleModel(A2, B2, Pi2)
    path, score, _ = g.decode(obs[1], 4)
    want, wscore, _ = om2.flash(obs[1], 4)
    assert np.array_equal(path, want) and _bits(score) == _bits(wscore)
    g.close()


def test_model_created_in_parts_equals_whole(fv, oracle_mod, gpu_ctx):
    """flashv_model_create_rows on one device standing in for 3 ranks: every part computes its rows, the
    assembled table must decode (dense, sparse and FLASH-BS engines read layouts built from it) like a
    model created in one go."""
    K, M, T = 301, 5, 40
    A, B, Pi = random_hmm(K, M, 0.15, 93)
    parts = [fv.Model.create_rows(gpu_ctx, A, B, Pi, r, 3) for r in range(3)]
    with pytest.raises(fv.FlashvError) as e:
        fv.Plan(parts[0], T, 2, 1, 0, fv.ENGINE_AUTO)  # not finished yet
    assert e.value.code == fv.ERR_STATE
    for q in (1, 2):
        parts[0].pull_rows_from(q, parts[q])
    parts[0].finish()
    om = oracle_mod.OracleModel(A, B, Pi)
    ob = np.random.RandomState(93).randint(0, M, T).astype(np.int32)
    for eng in (fv.ENGINE_PERSISTENT, fv.ENGINE_STEP, fv.ENGINE_SPARSE):
        plan = fv.Plan(parts[0], T, 4, 1, 0, eng)
        plan.upload(ob)
        plan.run()
        paths, scores = plan.download()
        plan.close()
        want, wscore, _ = om.flash(ob, 4)
        assert np.array_equal(paths[0], want) and _bits(scores[0]) == _bits(wscore), eng
    got, score, _ = parts[0].bs_decode(ob, 3, 16)
    want, wscore, _ = om.flash_bs(ob, 3, 16)
    assert np.array_equal(got, want) and _bits(score) == _bits(wscore)
    for m in parts:
        m.close()


# ---- one process per GPU: the cudaIpc form ------------------------------------------------------------
def _ipc_worker(rank, world, port, K, M, T, seed, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import load_pkg, random_hmm as rh
    from oracle import oracle

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fv = load_pkg()
        torch.cuda.set_device(rank)
        A, B, Pi = rh(K, M, 0.1, seed)
        om = oracle.OracleModel(A, B, Pi)
        rng = np.random.RandomState(seed)
        obs = rng.randint(0, M, (9, T)).astype(np.int32)
        ctx = fv.Context(rank)

        def all_gather_bytes(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out

        # model created in parts: own rows, handles, barrier, pull, barrier, finish
        model = fv.Model.create_rows(ctx, A, B, Pi, rank, world)
        handles = all_gather_bytes(model.rows_handle())
        dist.barrier()
        for q in range(world):
            if q != rank:
                model.pull_rows(q, handles[q])
        dist.barrier()
        model.finish()

        # batch sharding: rows in place, gathered with the transport at hand (here: reduce MAX over -2 fill)
        paths, scores, rep = model.decode_batch_shard(obs, 3, rank, world)
        tp = torch.from_numpy(paths)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        ok_batch = all(np.array_equal(tp.numpy()[b], om.flash(obs[b], 3)[0]) for b in range(len(obs)))

        # state sharding through cudaIpc handles
        ok_shard = True
        for N in (6, 2):
            plan = fv.Plan(model, T, N, 1, 0, fv.ENGINE_PERSISTENT)
            plan.shard_init(rank, world)
            hs = all_gather_bytes(plan.shard_ipc_handle())
            for q in range(world):
                if q != rank:
                    plan.shard_open_peer(q, hs[q])
            want, wscore, _ = om.flash(obs[0], N)
            for _ in range(2):
                plan.upload(obs[0])
                ctx.sync()
                dist.barrier()  # the contract: every rank idle before any rank runs
                plan.run()
                got, sc = plan.download()
                ok_shard &= bool(np.array_equal(got[0], want)) and bool(np.float32(sc[0]).view(np.uint32) == np.float32(wscore).view(np.uint32))
            dist.barrier()  # nobody closes its region while a peer may still be storing into it
            plan.close()
        ret[rank] = (ok_batch, ok_shard)
        dist.barrier()
        model.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_two_processes_ipc(fv):
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    with mp.Manager() as mgr:
        ret = mgr.dict()
        port = 29700 + (os.getpid() % 2000)
        mp.spawn(_ipc_worker, args=(2, port, 777, 6, 60, 95, ret), nprocs=2, join=True)
        assert dict(ret) == {0: (True, True), 1: (True, True)}
