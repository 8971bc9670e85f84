"""Multi-GPU plumbing of the hot path: regions are independent units, so ranks take contiguous
blocks of regions balanced by aligned-base count and the only collective is one all-gather of the per-candidate
predictions at the end (preceded by a tiny count all-gather for the ragged sizes).  Backend-agnostic
(`nccl` on GPUs, `gloo` in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_regions(work: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous blocks [begin, end) of regions per rank, balanced by `work` (e.g. aligned bases per region).
    Contiguity keeps rank-major order == genomic order after the gather."""
    n = int(work.shape[0])
    if n == 0:
        return [(0, 0)] * world
    csum = np.concatenate([[0], np.cumsum(work.astype(np.float64))])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(csum, target, side="left"))
        k = min(max(k, cuts[-1]), n)
        # choose the nearer of k-1 / k to the target
        if k > cuts[-1] and abs(csum[k - 1] - target) <= abs(csum[min(k, n)] - target):
            k -= 1
        cuts.append(max(k, cuts[-1]))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_predictions(probs, n_valid: int, world: int):
    """All-gather of per-candidate predictions with ragged counts.  `probs` is a [cap, C] tensor whose first
    `n_valid` rows are valid.  Returns (all_probs [sum n, C] in rank-major order, counts list)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return probs[:n_valid], [n_valid]
    dev = probs.device
    cnt = torch.tensor([n_valid], dtype=torch.int64, device=dev)
    cnts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnts, cnt)
    counts = [int(c) for c in cnts.tolist()]
    m = max(counts)
    if m == 0:
        return probs[:0], counts
    mine = probs[:m]
    if mine.shape[0] < m:                      # capacity smaller than the largest shard: pad
        pad = torch.zeros((m - mine.shape[0],) + tuple(probs.shape[1:]), dtype=probs.dtype, device=dev)
        mine = torch.cat([mine, pad])
    allp = torch.empty((world * m,) + tuple(probs.shape[1:]), dtype=probs.dtype, device=dev)   # concatenated layout
    dist.all_gather_into_tensor(allp, mine.contiguous())
    return torch.cat([allp[r * m:r * m + counts[r]] for r in range(world)]), counts
