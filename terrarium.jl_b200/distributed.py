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


class _DeviceArray:
    """Minimal ``__cuda_array_interface__`` carrier for a buffer borrowed from the library (``trm_field_view``)."""

    def __init__(self, ptr: int, shape, strides, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "strides": tuple(strides), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}   # (torch refuses the read-only flag)


def field_tensor(integrator, name: str) -> torch.Tensor:
    """Zero-copy ``torch`` view ``[rows, ncol]`` (treat it as read-only: the library is not told about writes) of a field in the library's device memory (rows = layers,
    faces or 1). The view is valid until the integrator is closed; synchronise (``integrator.synchronize()``) first."""
    import ctypes as C
    from . import _abi as abi
    lib, h = integrator._lib, integrator._h
    ptr, ld, rows = C.c_void_p(), C.c_int64(), C.c_int32()
    lib.check(lib.field_view(h, abi.FIELD_IDS[name], C.byref(ptr), C.byref(ld), C.byref(rows)), f"field_view({name})")
    nf = np.dtype(integrator.nf)
    arr = _DeviceArray(ptr.value, (rows.value, ld.value), (ld.value * nf.itemsize, nf.itemsize), nf.str)
    return torch.as_tensor(arr, device=torch.device("cuda", torch.cuda.current_device()))[:, :integrator.ncol]


def gather_field_device(integrator, name: str) -> torch.Tensor:
    """NCCL all-gather of one field straight from the library's device buffers: every rank receives the field in global
    column order as a device tensor ``[rows, ncol_global]`` (no host staging). Column ranges may differ by one column
    between ranks, so every rank contributes a buffer padded to the widest range."""
    local = field_tensor(integrator, name)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.clone()
    world = dist.get_world_size()
    counts = [integrator.grid.partition(r, world) for r in range(world)]
    width = max(c1 - c0 for c0, c1 in counts)
    send = torch.zeros((local.shape[0], width), dtype=local.dtype, device=local.device)
    send[:, :local.shape[1]] = local
    recv = torch.empty((world,) + tuple(send.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send)
    return torch.cat([recv[r, :, :c1 - c0] for r, (c0, c1) in enumerate(counts)], dim=1)


def bind_to_gpu_numa(device_index: int) -> list:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA node), so that
    the pinned host buffers of the per-step forcing uploads / result downloads are allocated next to the GPU's PCIe
    root. Returns the CPU list (empty if NVML or the affinity call is unavailable: nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []


def max_over_ranks(value: float) -> float:
    """Device-side timing of a multi-GPU step is the maximum over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
