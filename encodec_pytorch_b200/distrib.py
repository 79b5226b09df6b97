"""Collective plumbing of the RVQ path (one process per GPU, ``torch.distributed``).

Mirrors the two helpers of the reference's ``distrib.py`` that the quantizer
is meant to use -- ``all_reduce`` (distrib.py:32-34) for the EMA statistics
and ``broadcast_tensors`` (distrib.py:55-68) after k-means init -- whose call
sites the reference left commented out (core_vq.py:157, :175).  Frames are
sharded over ranks; these are the only exchanges on the path.
Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist


# Cross-rank synchronisation of the codebook buffers in training.  OFF by default: the reference runs DDP with
# ``broadcast_buffers=False`` and left both call sites commented out (core_vq.py:157, :175), so every rank keeps its own
# EMA statistics -- a drop-in must not add blocking collectives (or change checkpoints) behind the caller's back.
# ``sync_buffers(True)`` turns on what BASELINE.json's north_star describes: one all-reduce of the packed EMA statistics
# per training forward and a rank-0 broadcast after k-means init, so that N ranks on N frame shards equal one rank on the
# whole batch.  Every rank must then run the same sequence of training forwards.
_SYNC = {"enabled": False, "group": None}


def sync_buffers(enabled: bool = True, group=None) -> None:
    """Enable / disable the EMA all-reduce and the k-means-init broadcast (process-wide switch)."""
    _SYNC["enabled"] = bool(enabled)
    _SYNC["group"] = group


def sync_enabled() -> bool:
    return _SYNC["enabled"] and is_distributed()


def rank() -> int:
    """distrib.py:14-18."""
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def world_size() -> int:
    """distrib.py:21-25."""
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def is_distributed() -> bool:
    """distrib.py:28-29."""
    return world_size() > 1


def all_reduce(tensor: torch.Tensor, op=None):
    """distrib.py:32-34: in-place SUM over ranks; no-op in a single process."""
    if is_distributed():
        return dist.all_reduce(tensor, dist.ReduceOp.SUM if op is None else op, group=_SYNC["group"])
    return None


def _check_count(tensors: tp.List[torch.Tensor]) -> None:
    """distrib.py:41-52: every rank must broadcast the same number of tensors (deadlock guard)."""
    if not is_distributed() or not tensors:
        return
    n = torch.tensor([len(tensors)], device=tensors[0].device, dtype=torch.long)
    all_reduce(n)
    if int(n.item()) != len(tensors) * world_size():
        raise RuntimeError(f"Mismatch in number of params: ours is {len(tensors)}, "
                           "at least one worker has a different one.")


def broadcast_tensors(tensors: tp.Iterable[torch.Tensor], src: int = 0, check: bool = True) -> None:
    """distrib.py:55-68: broadcast the floating-point tensors from ``src`` (async, then wait).  ``check=False`` skips the
    count check (an all-reduce + host read) where the list is structurally the same on every rank."""
    if not is_distributed():
        return
    floats = [t for t in tensors if torch.is_floating_point(t) or torch.is_complex(t)]
    if check:
        _check_count(floats)
    handles = [dist.broadcast(t.data, src=src, async_op=True, group=_SYNC["group"]) for t in floats]
    for h in handles:
        h.wait()


def all_reduce_stats(flat: torch.Tensor) -> torch.Tensor:
    """Sum the packed EMA statistics (``[n_q*K]`` counts followed by ``[n_q*K*D]`` per-code sums,
    one buffer so a single collective covers all stages) over the frame shards of all ranks."""
    all_reduce(flat)
    return flat


def shard_frames(batch: int, world: tp.Optional[int] = None, r: tp.Optional[int] = None) -> tp.Tuple[int, int]:
    """Contiguous split of ``batch`` items over ranks (frames are independent, so any split is
    valid; batch items keep ``[B, D, T]`` slices contiguous).  Returns ``(start, stop)``."""
    world = world_size() if world is None else world
    r = rank() if r is None else r
    base, rem = divmod(batch, world)
    start = r * base + min(r, rem)
    return start, start + base + (1 if r < rem else 0)
