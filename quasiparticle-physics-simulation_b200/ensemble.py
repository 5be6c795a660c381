"""Parameter-sweep ensembles (BASELINE configs[4], SURVEY.md section 8e: "replicas only").

The runs of an ensemble are independent calls of ``run_2d_crank_nicolson``; there is nothing to exchange, so the
ensemble is dealt out round-robin over the ranks of a ``torchrun`` job (one process per GPU) or, in a single process,
over the devices it is given, and every member runs the unchanged single-GPU path.  Results come back in the order
of the input list on every rank (``all_gather_object``), or on the calling process when there is no process group.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Iterable, Sequence

from .solver import run_2d_crank_nicolson


def member_indices(n_members: int, world: int, rank: int) -> list[int]:
    """Members of rank ``rank``: i = rank, rank + world, ... (run times within a sweep vary smoothly with the
    parameters, so interleaving balances the ranks)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    return list(range(rank, int(n_members), world))


def run_ensemble(members: Sequence[dict[str, Any]], *, device: int | None = None, group=None,
                 reduce: Callable[[tuple], Any] | None = None, runner: Callable[..., tuple] | None = None,
                 gather: bool = True) -> list[Any]:
    """Run every keyword dict of ``members`` through ``run_2d_crank_nicolson`` (or ``runner``).

    Under ``torch.distributed`` (initialised process group) each rank runs its share on ``device`` (default: its
    LOCAL_RANK) and, when ``gather`` is set, every rank receives the full list.  ``reduce`` maps the 6-tuple a run
    returns to what should travel (default: ``(times, mass)``; whole frame stacks of 64 runs are rarely wanted on
    every rank).  Without a process group all members run here, one after the other.
    """
    runner = runner or run_2d_crank_nicolson
    reduce = reduce or (lambda out: (out[0], out[2]))
    world, rank = 1, 0
    dist = None
    try:
        import torch.distributed as _dist

        if _dist.is_available() and _dist.is_initialized():
            dist = _dist
            world, rank = dist.get_world_size(group), dist.get_rank(group)
    except ImportError:
        pass
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if dist is not None else 0
    mine = member_indices(len(members), world, rank)
    local = {}
    for i in mine:
        kw = dict(members[i])
        kw.setdefault("device", device)
        local[i] = reduce(runner(**kw))
    if dist is None or not gather:
        return [local.get(i) for i in range(len(members))]
    parts: list[Any] = [None] * world
    dist.all_gather_object(parts, local, group=group)
    merged: dict[int, Any] = {}
    for p in parts:
        merged.update(p)
    return [merged[i] for i in range(len(members))]


def parameter_grid(base: dict[str, Any], **axes: Iterable[Any]) -> list[dict[str, Any]]:
    """Cartesian product of the given keyword axes over a base keyword dict (first axis slowest)."""
    out = [dict(base)]
    for name, values in axes.items():
        out = [{**kw, name: v} for kw in out for v in values]
    return out
