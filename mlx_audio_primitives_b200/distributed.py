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


class PeakExchange:
    """One-float MAX exchange over peer memory (include/mlxa_cuda.h: mlxa_peak_exchange) for the ranks of the
    enabled group: 2*world uint64 slots per rank in torch symmetric memory (peer-mapped over NVLink /
    NVSwitch), a local ticket counter and an epoch that every rank advances once per producer launch.  The
    mel kernel's last CTA publishes, the dB kernel collects: no NCCL call, no extra launch.  ``create``
    returns None (callers keep the NCCL all-reduce) when the group has one rank or symmetric memory is
    unavailable."""

    def __init__(self, slots, handle, ticket, rank, world):
        import ctypes as C

        class _Desc(C.Structure):
            _fields_ = [("peer_slots", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32),
                        ("epoch", C.c_uint32), ("ticket", C.c_void_p)]
        self._slots, self._handle, self._ticket = slots, handle, ticket  # keep the allocations alive
        self._desc = _Desc(int(handle.buffer_ptrs_dev), rank, world, 0, ticket.data_ptr())
        self._byref = C.byref(self._desc)

    @classmethod
    def create(cls, device) -> "PeakExchange | None":
        import os
        if not _enabled or os.environ.get("MLXA_NO_PEER_EXCHANGE"):
            return None
        import torch.distributed as dist
        group = _group if _group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or dist.get_backend(group) != "nccl":
            return None
        try:
            import torch.distributed._symmetric_memory as symm
            slots = symm.empty(2 * world, dtype=torch.int64, device=device)
            slots.zero_()
            handle = symm.rendezvous(slots, group.group_name)
            ok = torch.ones(1, device=device)
        except Exception:  # no peer access / allocator unsupported: every rank must agree to fall back
            slots = handle = None
            ok = torch.zeros(1, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also orders the zeroing before any publish
        if float(ok) < 1.0:
            return None
        return cls(slots, handle, torch.zeros(1, dtype=torch.int32, device=device), rank, world)

    def next_epoch(self) -> None:
        """Call once per producer launch, before it (same count on every rank)."""
        self._desc.epoch += 1

    @property
    def ref(self):
        """ctypes by-reference handle of the descriptor for the C-ABI calls."""
        return self._byref


def shard_bounds(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous batch split: rank r owns clips [lo, hi)."""
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
