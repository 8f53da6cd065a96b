"""Satellite-channel sharding across the GPUs of one box (one process per GPU).

The reference has no multi-GPU code; the layout follows north_star / SURVEY.md 8(e):
channels are independent given the same signal block, so they are partitioned across
ranks; each integration period's signal block is broadcast over NCCL (NVLink 5 / NVSwitch)
and only the tiny per-satellite accumulators are gathered back.  torch.distributed is
plumbing here (rendezvous + NCCL/gloo collectives); the compute on every rank is libgat.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (first n%w ranks get one more)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sample_ranges(n_samples: int, world_size: int, align: int = 4) -> list[tuple[int, int]]:
    """Sample sharding (include/gat.h gat_set_sample_origin): [(first sample, length)] of every rank's contiguous range of
    a block, balanced, every start a multiple of `align` samples (16-byte aligned rows for the TMA descriptors when the
    range is bound in place).  The decomposition for few channels per GPU: no signal crosses NVLink, the partial sums do."""
    cuts = [min(n_samples, (n_samples * r // world_size) // align * align) for r in range(world_size)] + [n_samples]
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(world_size)]


def shard_channels(channels: Sequence, world_size: int, rank: int, keep_bands_together: bool = True):
    """Returns (indices, shard).  With keep_bands_together, channels are first ordered by
    system id so a GPU tends to need only one band's signal (SURVEY 8e 'Partitioning')."""
    order = list(range(len(channels)))
    if keep_bands_together:
        order.sort(key=lambda i: (channels[i].system.system_id, i))
    lo, hi = shard_bounds(len(order), world_size, rank)
    idx = order[lo:hi]
    return idx, [channels[i] for i in idx]


def broadcast_signal(re, im, src: int = 0, group=None, async_op: bool = False):
    """Broadcast one signal block (two planes) from `src`.  With NCCL this is one fused
    group over NVLink; returns the work handles when async_op (double buffering)."""
    h1 = dist.broadcast(re, src=src, group=group, async_op=async_op)
    h2 = dist.broadcast(im, src=src, group=group, async_op=async_op)
    return (h1, h2) if async_op else None


def gather_outputs(local_re, local_im, counts: Sequence[int], group=None):
    """All-gather the per-rank accumulators [K_r, L, M] into [K, L, M] on every rank.
    Shards may differ by one channel, so each is padded to max(counts)."""
    world = dist.get_world_size(group)
    kmax = max(counts)
    pad = kmax - local_re.shape[0]
    if pad:
        z = local_re.new_zeros((pad,) + tuple(local_re.shape[1:]))
        local_re = torch.cat([local_re, z])
        local_im = torch.cat([local_im, z])
    both = torch.stack([local_re, local_im]).contiguous()          # [2, kmax, L, M]
    out = [torch.empty_like(both) for _ in range(world)]
    dist.all_gather(out, both, group=group)
    re = torch.cat([o[0, :c] for o, c in zip(out, counts)])
    im = torch.cat([o[1, :c] for o, c in zip(out, counts)])
    return re, im


def sharded_correlate(channels: Sequence, correlate_fn: Callable[[Sequence], tuple], group=None,
                      keep_bands_together: bool = True):
    """Run `correlate_fn(shard) -> (re, im)` tensors [K_r, L, M] on this rank's shard and
    return the full [K, L, M] accumulators in the ORIGINAL channel order on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    all_idx = [shard_channels(channels, world, r, keep_bands_together)[0] for r in range(world)]
    idx, shard = shard_channels(channels, world, rank, keep_bands_together)
    re, im = correlate_fn(shard)
    g_re, g_im = gather_outputs(re, im, [len(i) for i in all_idx], group)
    perm = np.concatenate([np.asarray(i, dtype=np.int64) for i in all_idx])
    inv = torch.as_tensor(np.argsort(perm), device=g_re.device)
    return g_re.index_select(0, inv), g_im.index_select(0, inv)


def gather_setup(engine, elems_per_rank: int, group=None):
    """Create this rank's gather buffer, exchange the CUDA IPC handles over torch.distributed and map
    every peer's buffer.  After this, engine.correlate_batch(..., gather=True) stores its accumulators
    into all ranks' buffers from the kernel epilogue (NVLink peer stores) and engine.gather_wait()
    orders the stream behind every rank's arrival flag -- no collective launch per step."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    handle = engine.gather_create(world, rank, elems_per_rank)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    engine.gather_connect(handles)
    dist.barrier(group=group)


def ring_setup(engine, n_slots: int, n_samples: int, n_ants: int, group=None):
    """Create this rank's share of the signal ring (include/gat.h gat_ring_*), exchange the CUDA IPC handles over
    torch.distributed and map every owner's memory.  Afterwards the engine's slots 0 .. n_slots-1 are blocks whose
    samples live on all ranks; a correlate call on any rank gathers its tiles over NVLink inside the kernel."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    handle = engine.ring_create(world, rank, n_slots, n_samples, n_ants)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    engine.ring_connect(handles)
    dist.barrier(group=group)
