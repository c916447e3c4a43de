"""Frame sharding across ranks (one process per GPU).  The attack layer needs no collective:
every frame — indeed every 16x16 MCU — is independent, so rank r simply owns the
contiguous frame range [r*N/W, (r+1)*N/W) and its own Philox sub-stream."""
from __future__ import annotations

from typing import Tuple


def frame_shard(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Half-open [start, stop) of the frames owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(x, rank: int, world: int, dim: int = 0):
    """View of this rank's frames of a [N, ...] batch (or [B,3,T,H,W] clip with dim=2)."""
    a, b = frame_shard(x.shape[dim], rank, world)
    return x.narrow(dim, a, b - a)
