"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sequences shard b mod G, no data-path
collective, paths gathered on rank 0, times reduced with MAX.  The per-rank decode is played by the
oracle here (this is a CPU test of the plumbing); on the GPU box the same functions wrap
flashv_decode_batch (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden

sys.path.insert(0, str(ROOT / "flash-viterbi_b200" / "host"))


def _worker(rank, world, port, obs, N, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import shard
        from oracle import oracle

        g = load_golden("hmm_k37")
        om = oracle.OracleModel(g["A"], g["B"], g["Pi"])
        mine = shard.rank_sequences(len(obs), world, rank)
        local = np.stack([om.flash(obs[b], N)[0] for b in mine]) if len(mine) else np.zeros((0, obs.shape[1]), np.int32)
        slowest = shard.max_over_ranks(1.0 + rank, dist)
        full = shard.gather_paths(local, len(obs), dist)
        if rank == 0:
            ret["paths"] = full
            ret["slowest"] = slowest
    finally:
        dist.destroy_process_group()


def test_sequences_shard_over_two_ranks():
    import shard

    assert list(shard.rank_sequences(7, 2, 0)) == [0, 2, 4, 6] and list(shard.rank_sequences(7, 2, 1)) == [1, 3, 5]
    assert sorted(np.concatenate([shard.rank_sequences(9, 4, r) for r in range(4)])) == list(range(9))
    from conftest import load_pkg

    fv = load_pkg()  # the C ABI's count is the length of that index list (pure host function: runs without a GPU)
    for total, world in ((7, 2), (9, 4), (3, 8), (8192, 8), (0, 3)):
        for r in range(world):
            assert fv.shard_count(total, r, world) == len(shard.rank_sequences(total, world, r))
    g = load_golden("hmm_k37")
    rng = np.random.RandomState(5)
    obs = rng.randint(0, g["B"].shape[1], (5, 40)).astype(np.int32)
    from oracle import oracle

    om = oracle.OracleModel(g["A"], g["B"], g["Pi"])
    want = np.stack([om.flash(o, 3)[0] for o in obs])
    with mp.Manager() as mgr:
        ret = mgr.dict()
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_worker, args=(2, port, obs, 3, ret), nprocs=2, join=True)
        assert np.array_equal(ret["paths"], want)
        assert ret["slowest"] == 2.0


def _ctrl_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, str(ROOT / "tools"))
        import bench_side
        import gen_hmm

        ctrl = bench_side.Ctrl(dist, rank, world)
        got = ctrl.gather_objects(("rank", rank))
        bc = ctrl.bcast_object({"x": 7} if rank == 0 else None)
        slowest = ctrl.fmax(10.0 - rank)
        agree = (ctrl.all_true(True), ctrl.all_true(rank == 0))
        # the HMM of config 5 at toy size: rank 0 writes it into /dev/shm, every rank maps the same bytes
        K, prob, seed = 40, 0.3, 4242 + os.getppid() % 1000
        A, gen_s = bench_side.shared_hmm(ctrl, K, prob, seed)
        want = gen_hmm.as_reference_floats(gen_hmm.transition_matrix(K, prob, seed))
        same = bool(np.array_equal(np.asarray(A), want))
        ctrl.barrier()
        if rank == 0:
            for suffix in ("", ".ok"):
                try:
                    os.unlink(f"/dev/shm/flashv_hmm_K{K}_p{prob}_s{seed}.f32{suffix}")
                except OSError:
                    pass
        ret[rank] = (got, bc, slowest, agree, same, gen_s > 0)
    finally:
        dist.destroy_process_group()


def test_bench_control_plane_two_ranks():
    """tools/bench_side.py's control plane (what bench.py --gpus N uses beside the data path: handle exchange,
    flags, max-over-ranks times) and the shared-memory HMM of config 5, on two gloo ranks without a GPU."""
    with mp.Manager() as mgr:
        ret = mgr.dict()
        port = 29900 + (os.getpid() % 2000)
        mp.spawn(_ctrl_worker, args=(2, port, ret), nprocs=2, join=True)
        for rank in (0, 1):
            got, bc, slowest, agree, same, generated = ret[rank]
            assert got == [("rank", 0), ("rank", 1)] and bc == {"x": 7} and slowest == 10.0
            assert agree == (True, False) and same
            assert generated == (rank == 0)
