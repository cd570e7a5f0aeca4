"""Multi-GPU plumbing for the batched path: streams are independent units (no cross-stream arithmetic anywhere
in DSP::process), so rank r of G owns a contiguous block of streams, keeps their state resident
on its own GPU, and the only collective is the optional gather of the output blocks.

torch.distributed is plumbing only: NCCL over NVLink/NVSwitch for CUDA tensors, gloo for the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_bounds(n_streams: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced partition: the first (n_streams % world_size) ranks carry one extra stream."""
    if world_size < 1 or not (0 <= rank < world_size) or n_streams < 0:
        raise ValueError("bad partition arguments")
    base, rem = divmod(n_streams, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_shard(n_streams: int, world_size: int) -> int:
    return -(-n_streams // world_size)


def gather_streams(local, n_streams: int, group=None, dst: Optional[int] = None):
    """All-gather (dst None) or gather-to-dst of per-rank output blocks [B_local, n] -> [n_streams, n].

    Uneven shards are padded to the largest shard for the collective and trimmed afterwards.  Returns the
    full tensor on every rank (all-gather) or on `dst` only (None elsewhere).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(n_streams, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} streams, expected {hi - lo}")
    pad = max_shard(n_streams, world)
    buf = local
    if local.shape[0] != pad:
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    buf = buf.contiguous()
    if dst is None:
        chunks = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(chunks, buf, group=group)
    else:
        chunks = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, chunks, dst=dst, group=group)
        if rank != dst:
            return None
    parts = []
    for r in range(world):
        rlo, rhi = shard_bounds(n_streams, world, r)
        parts.append(chunks[r][: rhi - rlo])
    return torch.cat(parts, dim=0)
