"""Multi-GPU plumbing for the check path: one process per GPU (torch.distributed).

The path shards by CHECK: every rank holds a replica of the snapshot (100 M tuples are
~2 GB, far below one GPU's HBM) and answers its own slice of the batch; there is no
data-path collective. The only communication is the
all-gather that returns the answers to the caller's order, plus the max-reduce of the
timings in bench.py. Works with backend "nccl" (GPU tensors) and "gloo" (CPU tensors;
used by the world_size-2 tests, where the evaluator is a stub).
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced slice of n items for this rank."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedChecker:
    """check_bulk over a process group: slice, answer locally, all-gather the answers.

    evaluate: callable(items ndarray) -> uint8 ndarray; on GPU ranks this is
    `engine.check_bulk`. Every rank must call check_bulk with the same items.
    """

    def __init__(self, evaluate: Callable[[np.ndarray], np.ndarray], group=None, device: str | None = None):
        self.evaluate = evaluate
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")

    def check_bulk(self, items: np.ndarray) -> np.ndarray:
        n = items.size
        lo, hi = shard_bounds(n, self.rank, self.world)
        mine = self.evaluate(items[lo:hi]) if hi > lo else np.empty(0, dtype=np.uint8)
        width = -(-n // self.world)  # every rank contributes a fixed-width block
        block = torch.zeros(width, dtype=torch.uint8)
        block[: hi - lo] = torch.from_numpy(np.ascontiguousarray(mine))
        block = block.to(self.device)
        gathered = [torch.empty_like(block) for _ in range(self.world)]
        dist.all_gather(gathered, block, group=self.group)
        out = np.empty(n, dtype=np.uint8)
        for r, t in enumerate(gathered):
            rlo, rhi = shard_bounds(n, r, self.world)
            out[rlo:rhi] = t[: rhi - rlo].cpu().numpy()
        return out


def max_over_ranks(value: float, group=None, device: str = "cpu") -> float:
    """Timings are reported as the max over ranks (bench.py contract)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
