"""Sharding of independent problems across ranks (one process per GPU, torch.distributed).

The batched sweep (config C4) partitions trivially: rank r of W solves a contiguous slice of
the problem indices with no data-path collective; only scalar timings / counters are reduced
(max over ranks for time, sum for work), as bench.py's contract requires.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) of a contiguous, balanced partition of `total` problems over `world` ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def weak_slice(per_rank: int, rank: int) -> Tuple[int, int]:
    """Weak scaling: every rank owns `per_rank` problems; rank r gets indices [r*B, (r+1)*B)."""
    return rank * per_rank, (rank + 1) * per_rank


class Reducer:
    """max / sum of host scalars over the ranks of a torch.distributed group (NCCL or gloo)."""

    def __init__(self, dist=None, device=None):
        self.dist, self.device = dist, device

    def _reduce(self, x: float, op):
        if self.dist is None or not self.dist.is_initialized() or self.dist.get_world_size() == 1:
            return float(x)
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=self.device or "cpu")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX if self.dist is not None else None)

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM if self.dist is not None else None)

    def gather_concat(self, arr):
        """All ranks' arrays (same shape) concatenated along axis 0, on every rank."""
        if self.dist is None or not self.dist.is_initialized() or self.dist.get_world_size() == 1:
            return arr
        import numpy as np
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device or "cpu")
        out = [torch.empty_like(t) for _ in range(self.dist.get_world_size())]
        self.dist.all_gather(out, t)
        return np.concatenate([o.cpu().numpy() for o in out], axis=0)
