"""terrarium.jl_b200 -- B200-native (sm_100a) per-column land time-step behind Terrarium.jl's API.

The directory name contains a dot, so the package is imported through the ``terrarium_jl_b200``
shim at the repository root (``import terrarium_jl_b200 as trm``).
"""
from . import _abi as abi
from . import netcdf4
from ._abi import TerrariumError
from .grids import (B200, ColumnGrid, ColumnRingGrid, ExponentialSpacing, PrescribedSpacing, UniformSpacing,
                    get_spacing, num_layers)
from .models import *  # noqa: F401,F403  (configuration types mirror the reference's exported names)
from .integrator import (Field, ModelIntegrator, StateVariables, current_time, get_steps, initialize, interior, run, set_,
                         timestep)
from .simulation import (AveragedTimeInterval, Callback, FieldTimeSeries, IterationInterval, NetCDFWriter, Simulation, TimeInterval,
                         run_simulation)
from ._lib import LIB_PATH, cuda_library

# `terrarium_jl_b200.distributed` (torch.distributed helpers) is imported on demand: it pulls in torch

__version__ = "0.1.0"
