"""Loader of the product library ``csrc/libterrarium_b200.so`` (hand written sm_100a CUDA + C ABI).

There is deliberately no fallback: if the shared object has not been built (``__graft_entry__.build()``
or ``make -C terrarium.jl_b200/csrc``) loading raises, and if no CUDA device is usable ``trm_create``
returns ``TRM_ERR_NO_DEVICE`` which surfaces as :class:`TerrariumError`.
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

from . import _abi as abi

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
# TRM_LIB selects an alternative build of the same library (kernel tuning experiments); default is the in-tree build
LIB_PATH = os.environ.get("TRM_LIB") or os.path.join(CSRC, "libterrarium_b200.so")


@lru_cache(maxsize=1)
def cuda_library() -> abi.BoundLibrary:
    if not os.path.exists(LIB_PATH):
        raise abi.TerrariumError(
            abi.TRM_ERR_NO_DEVICE,
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). This package has no CPU fallback.")
    return abi.BoundLibrary(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL), "trm_")
