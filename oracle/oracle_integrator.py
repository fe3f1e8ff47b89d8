"""TEST INFRASTRUCTURE ONLY -- drives the CPU oracle (oracle/libterrarium_oracle.so) through the same
host logic as the product integrator, so that a test can build one scenario and run it twice.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package never does (terrarium.jl_b200/_lib.py loads the CUDA library only).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from functools import lru_cache

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import terrarium_jl_b200 as trm  # noqa: E402
from terrarium_jl_b200 import _abi as abi  # noqa: E402

ORACLE_PATH = os.path.join(HERE, "libterrarium_oracle.so")


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "terrarium_oracle.cpp")
    hdr = os.path.join(ROOT, "include", "terrarium_b200.h")
    stale = (not os.path.exists(ORACLE_PATH)
             or os.path.getmtime(ORACLE_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return ORACLE_PATH


def build_native_oracle() -> str:
    """The oracle compiled with -march=native ON THIS HOST (bench.py's CPU baseline); falls back to the portable build."""
    try:
        subprocess.run(["make", "-C", HERE, "-s", "native"], check=True, capture_output=True, timeout=300)
        path = os.path.join(HERE, "libterrarium_oracle_native.so")
        C.CDLL(path)
        return path
    except Exception:
        return build_oracle()


@lru_cache(maxsize=1)
def oracle_library() -> abi.BoundLibrary:
    path = os.environ.get("TERRARIUM_ORACLE_LIB") or build_oracle()
    lib = abi.BoundLibrary(C.CDLL(path), "orc_", skip=abi.DEVICE_ONLY)
    lib.cdll.orc_array_passes.restype = C.c_int64
    lib.cdll.orc_array_passes.argtypes = [C.c_void_p]
    lib.cdll.orc_num_threads.restype = C.c_int
    return lib


class OracleIntegrator(trm.ModelIntegrator):
    """Same constructor as ``ModelIntegrator``; every ABI call goes to the CPU oracle instead."""

    @classmethod
    def _library(cls):
        return oracle_library()

    def array_passes(self) -> int:
        return int(self._lib.cdll.orc_array_passes(self._h))


def oracle_initialize(model, timestepper, inputs=None, *, boundary_conditions=None, initializers=None, partition=None,
                      math="faithful"):
    # `math` is accepted for signature parity with trm.initialize and ignored: the oracle has one arithmetic
    return OracleIntegrator(model, timestepper, inputs, boundary_conditions, initializers, partition)
