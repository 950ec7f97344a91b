"""Host-side sharding of independent sequences over the GPUs of one box ("batched
sequences"): sequence b goes to rank b mod G, HMM tables are replicated, there is no data-path
collective — only the result gather and the max-over-ranks clock use torch.distributed."""
from __future__ import annotations

import numpy as np


def rank_sequences(total: int, world: int, rank: int) -> np.ndarray:
    """Indices of the sequences rank `rank` decodes (b mod G == rank): the rows flashv_decode_batch_shard
    fills in place; their number is flashv_shard_count(total, rank, world)."""
    return np.arange(rank, total, world, dtype=np.int64)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """The slowest rank's time: every multi-GPU number is reported against it."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_paths(local_paths: np.ndarray, total: int, dist=None, device=None) -> np.ndarray | None:
    """Reassemble [total][T] paths on rank 0 from the per-rank shards (rows in rank_sequences order)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_paths
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    T = local_paths.shape[1]
    per = (total + world - 1) // world
    buf = torch.full((per, T), -2, dtype=torch.int32, device=device)
    buf[: local_paths.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_paths, np.int32)).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    full = np.empty((total, T), np.int32)
    for r in range(world):
        idx = rank_sequences(total, world, r)
        full[idx] = out[r][: len(idx)].cpu().numpy()
    return full
