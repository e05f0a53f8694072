"""Batch sharding of the op lists across the GPUs of one box: every named op is independent per batch
item, so rank r of W takes rows [r*B/W, (r+1)*B/W) of a global batch (or, for weak scaling, its own fixed-size
batch), weights are replicated, and the only data-path collective is one all-gather of the outputs."""
from __future__ import annotations

from typing import Tuple


def shard_range(rank: int, world: int, batch: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `batch` items: the first `batch % world` ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_layout(world: int, per_rank_shape) -> tuple:
    """Shape of the all-gathered output buffer: rank-major concatenation along the batch axis."""
    return (world * per_rank_shape[0],) + tuple(per_rank_shape[1:])


def all_gather_outputs(dist, out_local, gather_buf):
    """One collective per step: `gather_buf[r*b:(r+1)*b] = out of rank r` (equal per-rank batch)."""
    dist.all_gather_into_tensor(gather_buf, out_local)
    return gather_buf
