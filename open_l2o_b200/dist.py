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
