"""Partitioning of independent units (batch rows, STFT channels) across ranks.

kofft's batched entry points loop over independent rows (`batch`, src/fft.rs:2156-2164) and
STFT channels/frames are independent as well (src/stft.rs:246-262), so multi-GPU execution is
one process per GPU, each transforming its own contiguous block of units with its own replicated
twiddle tables: there is no data-path collective.  These helpers are the whole protocol.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `total` units owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(x, rank: int, world: int):
    """View of this rank's rows of a [units, ...] array or tensor."""
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def owner_of(unit: int, total: int, world: int) -> int:
    """Rank that owns `unit` under shard_range."""
    base, extra = divmod(total, world)
    cut = extra * (base + 1)
    if unit < cut:
        return unit // (base + 1)
    return extra + (unit - cut) // base if base else world - 1
