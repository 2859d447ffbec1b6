"""Multi-GPU sharding of a batch: ciphertexts are independent (reference src/cipher.rs:180-185, :227-237;
common.rs touches only its two operands), so values are split by index into contiguous ranges, one per rank,
with the keys replicated and NO collective on the data path.  torch.distributed is used only for the timing
barrier / max-over-ranks and, optionally, to gather small plaintext results."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of rank `rank`; sizes differ by at most one; concatenation order == rank order."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard(values: np.ndarray, masks: np.ndarray, rank: int, world: int, bits: int, mask_bytes: int):
    """The slice of plaintexts and of the subset-mask stream (value-major, bit-minor) that belongs to `rank`."""
    lo, hi = shard_range(values.shape[0], rank, world)
    m = np.ascontiguousarray(masks, dtype=np.uint8).reshape(values.shape[0], bits * mask_bytes)
    return values[lo:hi], m[lo:hi].reshape(-1)


def max_over_ranks(seconds: float, device=None) -> float:
    """Timing rule of bench.py: the job's time is the slowest rank's."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_plaintexts(local: np.ndarray) -> List[np.ndarray]:
    """Rank-ordered list of every rank's decrypted values (small: 4 B per u32)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out: List = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return out
