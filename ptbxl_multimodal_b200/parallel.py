"""Host-side plan of the data-parallel step (one process per GPU).

The reference is single-device; under data parallel every rank holds a replica, trains on its slice of the
global batch, and the ranks exchange gradients once per step.  ecgb200 does not all-reduce: the flat fp32
parameter space is split into `world` contiguous, 16-byte aligned shards; rank r sums everybody's gradients
for shard r, applies AdamW there (its Adam moments exist only on r) and writes the new parameters into every
replica (csrc/dp_fused.cu).  This module holds the arithmetic both the engine and the tests rely on."""
from __future__ import annotations

from typing import List, Tuple

DP_MAX_WORLD = 8          # csrc/dp_fused.cu: peer pointer table size
DP_ALIGN = 4 * 840        # lcm(4 * w for w in 1..8): float4 shards for ANY world size up to 8


def padded_size(n: int) -> int:
    """Length of the flat parameter / gradient / moment buffers for `n` real parameters."""
    if n <= 0:
        raise ValueError("n must be positive")
    return (n + DP_ALIGN - 1) // DP_ALIGN * DP_ALIGN


def shard_bounds(n_pad: int, world: int) -> List[Tuple[int, int]]:
    """[lo, hi) element range each rank owns.  Equal, contiguous, 16-byte aligned, covering [0, n_pad)."""
    if not (1 <= world <= DP_MAX_WORLD):
        raise ValueError(f"world size must be in 1..{DP_MAX_WORLD}")
    if n_pad % (4 * world):
        raise ValueError("padded size must be a multiple of 4 * world (use padded_size())")
    per = n_pad // world
    return [(r * per, (r + 1) * per) for r in range(world)]


def split_batch(global_batch: int, world: int) -> int:
    """Per-rank batch: the global batch must divide evenly (every rank replays the same captured graph)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} does not divide over {world} ranks")
    return global_batch // world
