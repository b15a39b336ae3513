"""Sharding of a batch over ranks / devices: contiguous index ranges, no data-path collective.

The reference's only parallelism on this path is one process per sample (simulate_gt.py:64-69); samples are
independent, so rank r of R simply takes [floor(r n / R), floor((r+1) n / R)) — the same rule
`pdeb_eval_batch_multi` uses inside one process.  For mixed batches `cost_balanced_bounds` cuts by estimated work
(steps x elements) instead of by count.  `gather_results` is the one optional collective: an all_gather of the
8-byte results (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return n * rank // world, n * (rank + 1) // world


def cost_balanced_bounds(cost: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous cuts of a cost array so that every shard carries ~1/world of the total."""
    c = np.cumsum(np.asarray(cost, dtype=np.float64))
    total = c[-1] if len(c) else 0.0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(c, total * r / world, side="left")))
    cuts.append(len(c))
    cuts = np.maximum.accumulate(cuts)
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def batch_cost(batch) -> np.ndarray:
    """Work estimate per sample: time steps x mesh rows (x1.5 for waves: three state vectors)."""
    return batch.steps().astype(np.float64) * batch.n_elem * np.where(batch.kind == 1, 1.5, 1.0)


def gather_results(local, sizes: List[int], group=None):
    """all_gather of per-rank result tensors of different lengths -> one tensor in global sample order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if len(sizes) != world:
        raise ValueError("sizes must list every rank's shard length")
    m = max(sizes)
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    out = [torch.empty(m, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:k] for o, k in zip(out, sizes)])
