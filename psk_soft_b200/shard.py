"""Channel-bank sharding across the GPUs of one box.

Every channel is an independent demodulator (reference: one psk_soft_i instance each, private
state cpp/psk_soft.h:66-86), so a bank shards by channel with NO collective: rank r owns a
contiguous channel range, runs its own pskd bank, and only per-rank counters / timings are
reduced (max over ranks) by the caller.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def channel_ranges(n_channels: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) per rank, sizes differing by at most one."""
    base, extra = divmod(n_channels, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def balanced_ranges(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges with near-equal summed cost (mixed banks: cost ~ samples per channel,
    mildly ~ 1/samplesPerBaud).  Greedy split at the running-cost quantiles."""
    total = float(sum(costs))
    n = len(costs)
    out, lo, acc, r = [], 0, 0.0, 0
    for i, c in enumerate(costs):
        acc += c
        remaining_ranks = world - r - 1
        if remaining_ranks > 0 and (acc >= total * (r + 1) / world or n - (i + 1) == remaining_ranks):
            out.append((lo, i + 1))
            lo = i + 1
            r += 1
    out.append((lo, n))
    while len(out) < world:
        out.append((n, n))
    return out
