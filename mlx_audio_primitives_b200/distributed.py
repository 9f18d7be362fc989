"""Multi-GPU: clips shard across ranks with no data-path exchange; the single collective of the
hot path is a one-float all-reduce(MAX) of the spectrogram peak, needed because the reference's
``ref=max`` and ``top_db`` use the maximum over the WHOLE batch (convert.py:42-58, mfcc.py:260)."""
from __future__ import annotations

import torch

_group = None
_enabled = False


def enable(group=None) -> None:
    """Make power_to_db / mfcc reduce their peak over ``group`` (default: the world group)."""
    global _group, _enabled
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _group, _enabled = group, True


def disable() -> None:
    global _group, _enabled
    _group, _enabled = None, False


def is_enabled() -> bool:
    return _enabled


def all_reduce_max_(peak: torch.Tensor) -> torch.Tensor:
    """In-place MAX all-reduce of a 1-element tensor when sharding is enabled; exact and
    order-independent, so results do not depend on the GPU count."""
    if _enabled:
        import torch.distributed as dist
        dist.all_reduce(peak, op=dist.ReduceOp.MAX, group=_group)
    return peak


def shard_bounds(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous batch split: rank r owns clips [lo, hi)."""
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
