"""Sharding of a batch of independent blocks across ranks.

Blocks (pages / fragments / streams) never reference each other, so a multi-GPU run is a
static partition of the block index range: rank g owns [g*B/G, (g+1)*B/G).  There is NO
data-path collective.  The only thing that crosses ranks is the per-block size index
(4 bytes per block): `gather_sizes` collects it on rank 0, which turns it into the packed
output offsets with an exclusive prefix sum -- the block_compressor-style index
(/root/reference/block_compressor.c:298-335).
"""
from __future__ import annotations

import numpy as np


def block_range(n_blocks: int, rank: int, world: int):
    """Contiguous, balanced [first, last) of `n_blocks` for `rank` of `world`."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return n_blocks * rank // world, n_blocks * (rank + 1) // world


def byte_balanced_ranges(block_bytes: np.ndarray, world: int):
    """Contiguous ranges balanced by BYTES for variable-sized blocks: list of (first, last)."""
    total = int(block_bytes.sum())
    cum = np.concatenate([[0], np.cumsum(block_bytes.astype(np.int64))])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(len(block_bytes))
    cuts = [min(max(c, cuts[i - 1] if i else 0), len(block_bytes)) for i, c in enumerate(cuts)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_sizes(local_sizes, rank: int, world: int, group=None):
    """All ranks call; rank 0 gets the concatenated int32 size index (others get None).
    Uses torch.distributed.gather_object-free tensor gather: sizes only, never payload."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local_sizes.clone()
    counts = [torch.zeros(1, dtype=torch.int64, device=local_sizes.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local_sizes.numel()], dtype=torch.int64, device=local_sizes.device),
                    group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    padded = torch.zeros(m, dtype=local_sizes.dtype, device=local_sizes.device)
    padded[: local_sizes.numel()] = local_sizes
    bucket = [torch.zeros(m, dtype=local_sizes.dtype, device=local_sizes.device) for _ in range(world)]
    dist.all_gather(bucket, padded, group=group)
    if rank != 0:
        return None
    return torch.cat([b[:c] for b, c in zip(bucket, counts)])


def packed_offsets(sizes) -> np.ndarray:
    """Exclusive prefix sum of the gathered size index -> int64 offsets, length n+1."""
    s = np.asarray(sizes, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(s)])
