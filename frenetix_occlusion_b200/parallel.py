"""Trajectory sharding across GPUs: contiguous split of the bundle over the ranks, the
(tiny) agent table replicated, and ONE all-gather of the per-trajectory result vectors.

One process per GPU; ``torch.distributed`` (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU
tests) is plumbing only -- there is no other exchange step on this path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_size(n_total: int, world: int) -> int:
    return (n_total + world - 1) // world


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the trajectory axis owned by ``rank`` (last shards may be short/empty)."""
    per = shard_size(n_total, world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


class ResultGatherer:
    """Pre-allocated buffers for the single collective of the path."""

    def __init__(self, n_total: int, summary_k: int, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.per = shard_size(n_total, self.world)
        self.lo, self.hi = shard_bounds(n_total, self.world, self.rank)
        # one packed row per trajectory: [valid, flags, summary...] as float32 -> a single all-gather
        self.k = summary_k + 2
        self.local = torch.zeros((self.per, self.k), dtype=torch.float32, device=device)
        self.full = torch.empty((self.per * self.world, self.k), dtype=torch.float32, device=device)

    def gather(self, valid: torch.Tensor, summary: torch.Tensor, flags: Optional[torch.Tensor] = None):
        """all-gather the shard results; returns (valid[N] uint8, summary[N,K] f32, flags[N] int32) views."""
        n = self.hi - self.lo
        self.local[:n, 0] = valid[:n].to(torch.float32)
        self.local[:n, 1] = flags[:n].to(torch.float32) if flags is not None else 0
        self.local[:n, 2:] = summary[:n]
        if self.world > 1:
            if self.full.is_cuda:
                dist.all_gather_into_tensor(self.full, self.local, group=self.group)
            else:   # gloo has no all_gather_into_tensor for every dtype/layout: use the list form
                parts = list(self.full.view(self.world, self.per, self.k).unbind(0))
                dist.all_gather(parts, self.local, group=self.group)
        else:
            self.full.copy_(self.local)
        out = self.full[:self.n_total]
        return out[:, 0].to(torch.uint8), out[:, 2:], out[:, 1].to(torch.int32)
