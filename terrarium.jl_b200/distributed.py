"""Multi-GPU plumbing: one process per GPU, one integrator per process, contiguous column ranges.

Columns are laterally independent (the reference only has d/dz operators, ``src/Terrarium.jl:27``), so the time
step needs no halo exchange and no collective.  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is
used only for what crosses ranks: global diagnostic reductions and output gathers.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.distributed as dist

_SUM = ("energy", "water", "nan_count", "ncol")
_MIN = ("t_min", "sat_min")
_MAX = ("t_max", "sat_max")


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def reduce_diagnostics(local: Dict[str, float]) -> Dict[str, float]:
    """All-reduce the per-rank ``trm_diag`` values: budgets and counts add up, extrema take min / max."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    dev = _device()
    out = dict(local)
    for names, op in ((_SUM, dist.ReduceOp.SUM), (_MIN, dist.ReduceOp.MIN), (_MAX, dist.ReduceOp.MAX)):
        t = torch.tensor([local[n] for n in names], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        out.update(zip(names, t.tolist()))
    return out


def gather_field(integrator, name: str) -> np.ndarray:
    """Gather one field of every rank's column range on all ranks, in global column order (output cadence only)."""
    local = getattr(integrator.state, name).numpy()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    parts = [None] * world
    dist.all_gather_object(parts, (integrator.col0, local))
    parts.sort(key=lambda p: p[0])
    return np.concatenate([p[1] for p in parts], axis=-1)


def max_over_ranks(value: float) -> float:
    """Device-side timing of a multi-GPU step is the maximum over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
