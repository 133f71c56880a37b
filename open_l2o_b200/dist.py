"""Multi-GPU plumbing for the coordinate-sharded meta-step.

Coordinates are independent given theta (DM/networks.py:251-271), so each rank owns a contiguous slice of the flat
coordinate arena and the forward unroll needs no data-path collective.  The only exchange per outer step is one
all-reduce (SUM) of the packed fp64 buffer [dtheta_net0 | dtheta_net1 | ... | fx_0..fx_T]; every rank then applies the
identical TF-Adam update.  Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of n coordinates for `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack(dtheta: Dict[str, torch.Tensor], fx: torch.Tensor) -> torch.Tensor:
    return torch.cat([dtheta[k].reshape(-1).double() for k in dtheta] + [fx.reshape(-1).double()])


def unpack(packed: torch.Tensor, dtheta: Dict[str, torch.Tensor], fx: torch.Tensor):
    off = 0
    for k in dtheta:
        n = dtheta[k].numel()
        dtheta[k].copy_(packed[off:off + n].reshape(dtheta[k].shape))
        off += n
    return packed[off:off + fx.numel()].reshape(fx.shape)


def allreduce_meta_grad(dtheta: Dict[str, torch.Tensor], fx: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of every net's dtheta; returns the summed fx.  One collective per outer step."""
    import torch.distributed as dist
    packed = pack(dtheta, fx)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return unpack(packed, dtheta, fx)


def allgather_shards(local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """Reassemble a flat [n] tensor whose `shard_range` slices live on the ranks (slices differ by at most one element,
    so every rank pads to the largest slice and one equal-size all-gather suffices).  Used by the sharded
    HierarchicalRNN step to republish the updated optimizee parameters."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cap = (n + world - 1) // world
    lo, hi = shard_range(n, rank, world)
    assert local.numel() == hi - lo, (local.numel(), lo, hi)
    buf = torch.zeros(cap, dtype=local.dtype, device=local.device)
    buf[:hi - lo].copy_(local.reshape(-1))
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = []
    for r in range(world):
        l, h = shard_range(n, r, world)
        parts.append(out[r][:h - l])
    return torch.cat(parts)
