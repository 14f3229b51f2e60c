"""Clip-level data parallelism: the only multi-GPU mechanism the path needs.

Every (batch, channel) row of every Activation1d call is independent and so is every clip, so
ranks never exchange data inside the hot path.  One process per GPU (torchrun); clips are dealt
round-robin; a single collective at the very end gathers finished waveforms (NCCL on GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Indices of the clips rank `rank` vocodes: rank, rank + W, rank + 2W, ..."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if n_items < 0:
        raise ValueError("n_items must be >= 0")
    return list(range(rank, n_items, world_size))


def gather_waveforms(local: torch.Tensor, n_items: int, rank: int, world_size: int, group=None) -> torch.Tensor:
    """All-gather per-rank results [n_local, ...] back into clip order [n_items, ...].

    Ranks may own different counts (n_items not divisible by world_size): shards are padded to the
    largest count for the collective and the padding is dropped afterwards.
    """
    import torch.distributed as dist

    mine = shard_indices(n_items, rank, world_size)
    if local.shape[0] != len(mine):
        raise ValueError(f"rank {rank} holds {local.shape[0]} items, expected {len(mine)}")
    if world_size == 1:
        return local
    n_max = (n_items + world_size - 1) // world_size
    padded = local
    if local.shape[0] < n_max:
        pad = torch.zeros((n_max - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    padded = padded.contiguous()
    # NCCL has no 16-bit integer type: interleaved stereo int16 PCM [n, T, 2] travels as one int32 per sample pair
    as_pairs = padded.dtype == torch.int16 and padded.dim() >= 1 and padded.shape[-1] % 2 == 0
    wire = padded.view(torch.int32) if as_pairs else padded
    bufs = [torch.empty_like(wire) for _ in range(world_size)]
    dist.all_gather(bufs, wire, group=group)
    if as_pairs:
        bufs = [b.view(torch.int16) for b in bufs]
    out = torch.empty((n_items,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world_size):
        idx = shard_indices(n_items, r, world_size)
        if idx:
            out[idx] = bufs[r][: len(idx)]
    return out


def split_batch(items: Sequence, rank: int, world_size: int) -> list:
    return [items[i] for i in shard_indices(len(items), rank, world_size)]
