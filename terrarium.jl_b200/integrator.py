"""``ModelIntegrator``: the host side of ``initialize`` / ``timestep!`` / ``run!`` over the C ABI.

Mirrors ``src/timesteppers/model_integrator.jl`` of the reference:

* ``initialize(model, timestepper, inputs...; boundary_conditions, initializers)``  (:145-161)
* ``timestep(integrator, dt; finalize=True)``   = ``timestep!``                    (:124-131)
* ``run(integrator; steps | period, dt)``        = ``run!``                        (:72-88)
* ``integrator.state.<name>`` / ``interior(field)`` / ``set_(field, value)``       (state_variables.jl:476-489)

Every numerical operation happens inside ``libterrarium_b200.so`` (hand written sm_100a kernels).
There is no CPU implementation in this package: if the library or a GPU is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Dict, Optional

import numpy as np

from . import _abi as abi
from ._lib import cuda_library
from .models import (BoundaryCondition, ForwardEuler, Heun, LandModel, RasterInputSource, Sinusoid, SoilModel, TimeSeries, build_config,
                     default_dt, merge_boundary_conditions)


class Clock:
    def __init__(self, integ):
        self._i = integ

    def _get(self):
        t, it = C.c_double(), C.c_int64()
        self._i._lib.check(self._i._lib.get_clock(self._i._h, C.byref(t), C.byref(it)), "get_clock")
        return t.value, it.value

    @property
    def time(self):
        return self._get()[0]

    @property
    def iteration(self):
        return self._get()[1]


class Field:
    """A state variable living in device memory; ``numpy()`` copies it to the host as ``[layer, column]``
    (layer 0 = bottom cell, like ``interior(field)[i, 1, :]`` in the reference) or ``[column]`` for 2-D fields."""

    def __init__(self, integ, name: str):
        self._i, self.name, self.id = integ, name, abi.FIELD_IDS[name]

    @property
    def shape(self):
        nz, nc = self._i.nz, self._i.ncol
        if self.name in abi.FIELDS_3D:
            return (nz, nc)
        if self.name in abi.FIELDS_FACE:
            return (nz + 1, nc)
        return (nc,)

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self._i.nf)
        lib = self._i._lib
        lib.check(lib.get_field(self._i._h, self.id, out.ctypes.data_as(C.c_void_p), out.size), f"get_field({self.name})")
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def set(self, value):
        """``set!(field, value)``: number, array or function ``(x, z)`` (``(x,)`` for 2-D fields)."""
        arr = self._i._evaluate_initializer(value, self.shape)
        lib = self._i._lib
        lib.check(lib.set_field(self._i._h, self.id, arr.ctypes.data_as(C.c_void_p), arr.size), f"set_field({self.name})")

    # -- ColumnRingGrid conversions (src/grids/column_ring_grid.jl:102-149), scatter / gather on the device ------
    def to_ring(self, fill_value=np.nan) -> np.ndarray:
        """``RingGrids.Field(field, grid; fill_value)``: the field on the full ring grid, ``fill_value`` off the mask.
        On a multi-rank partition every rank fills the points of its own column range."""
        integ = self._i
        nring = integ._ensure_ring_index()
        rows = 1 if len(self.shape) == 1 else self.shape[0]
        out = np.empty((rows, nring), dtype=integ.nf)
        lib = integ._lib
        lib.check(lib.get_field_ring(integ._h, self.id, out.ctypes.data_as(C.c_void_p), out.size, float(fill_value)), f"get_field_ring({self.name})")
        return out[0] if len(self.shape) == 1 else out

    def set_from_ring(self, ring_field) -> None:
        """``Field(ring_field, grid)``: copy the masked points of a ring-grid array ``[nring]`` / ``[rows, nring]``."""
        integ = self._i
        nring = integ._ensure_ring_index()
        rows = 1 if len(self.shape) == 1 else self.shape[0]
        arr = np.ascontiguousarray(np.broadcast_to(np.asarray(ring_field, dtype=integ.nf).reshape(rows, -1), (rows, nring)), dtype=integ.nf)
        lib = integ._lib
        lib.check(lib.set_field_ring(integ._h, self.id, arr.ctypes.data_as(C.c_void_p), arr.size), f"set_field_ring({self.name})")

    def __repr__(self):
        return f"Field({self.name}, shape={self.shape}, {np.dtype(self._i.nf).name}) on device"


def interior(field: Field) -> np.ndarray:
    return field.numpy()


def set_(field: Field, value) -> None:
    field.set(value)


class StateVariables:
    def __init__(self, integ):
        object.__setattr__(self, "_i", integ)
        object.__setattr__(self, "clock", Clock(integ))

    def __getattr__(self, name):
        if name == "inputs":   # state.inputs.<name> (state_variables.jl:38-41)
            return _Inputs(self._i)
        if name in abi.PRESCRIBED_FLUX_INPUTS and self._i._prescribed_flux_input(name):
            return InputField(self._i, abi.PRESCRIBED_FLUX_INPUTS[name])   # an input variable under the prescribed scheme
        if name in abi.FIELD_IDS:
            return Field(self._i, name)
        try:
            return InputField(self._i, self._i._input_id(name))
        except KeyError:
            raise AttributeError(name) from None


class _Inputs:
    def __init__(self, integ):
        self._i = integ

    def __getattr__(self, name):
        try:
            return InputField(self._i, self._i._input_id(name))
        except KeyError:
            raise AttributeError(name) from None


class InputField:
    """An input variable (``state.inputs``); ``set`` replaces its source."""

    def __init__(self, integ, input_id: int):
        self._i, self.id = integ, input_id

    def set(self, value):
        self._i._set_input(self.id, value)

    def numpy(self) -> np.ndarray:
        """The input field as of the last ``update_inputs!`` (start of the last step)."""
        out = np.empty(self._i.ncol, dtype=self._i.nf)
        lib = self._i._lib
        lib.check(lib.get_input(self._i._h, self.id, out.ctypes.data_as(C.c_void_p), out.size), "get_input")
        return out


class ModelIntegrator:
    @classmethod
    def _library(cls) -> abi.BoundLibrary:
        return cuda_library()  # raises if libterrarium_b200.so is missing: no fallback

    def __init__(self, model, timestepper, inputs=None, boundary_conditions=None, initializers=None,
                 partition=None, math: str = "faithful"):
        self.model, self.timestepper = model, timestepper
        grid = model.grid
        self.grid, self.nf, self.nz = grid, grid.nf, grid.Nz
        rank, world = partition if partition is not None else (0, 1)
        self.partition = (int(rank), int(world))
        self.col0, self.col1 = grid.partition(rank, world)
        self.ncol = self.col1 - self.col0
        self.ncol_global = grid.Nc
        self._lib = self._library()
        self._h = abi._H()
        self._host_callbacks: Dict[int, Callable] = {}
        self._bc_inputs: Dict[str, int] = {}
        self._keep = []

        cfg, zbuf = build_config(model, timestepper, self.ncol, self.col0, grid.arch.device, math)
        bcs = merge_boundary_conditions(boundary_conditions or {})
        pending = []
        next_user = abi.TRM_IN_USER0
        for var, sides in bcs.items():
            for side, bc in sides.items():
                if not isinstance(bc, BoundaryCondition):
                    raise TypeError(f"boundary condition for {var}.{side} must be a BoundaryCondition")
                cfg.bc[bc.slot].kind = bc.kind
                if bc.kind == abi.TRM_BC_DEFAULT:
                    continue
                if next_user >= abi.TRM_IN_USER0 + abi.TRM_NUM_USER_INPUTS:
                    raise ValueError("too many boundary condition inputs")
                cfg.bc[bc.slot].input = next_user
                # no value: the boundary value is the input variable `bc.name` (`var(name, XY())`, soil_model_bcs.jl:17),
                # zero until an InputSource of that name provides it
                pending.append((next_user, 0.0 if bc.value is None else bc.value))
                if bc.name:
                    self._bc_inputs[bc.name] = next_user
                next_user += 1
        self._lib.check(self._lib.create(C.byref(cfg), C.byref(self._h)), "create")
        self._cfg = cfg
        for input_id, value in pending:
            self._set_input(input_id, value)
        if inputs is not None and not isinstance(inputs, dict):   # initialize(model, ts, InputSource(...), ...) style
            sources = [inputs] if hasattr(inputs, "name") else list(inputs)
            inputs = {src.name: src.value for src in sources}
        for name, value in (inputs or {}).items():
            if name in self._bc_inputs:       # a boundary condition input, e.g. PrescribedSurfaceTemperature("Tair")
                self._set_input(self._bc_inputs[name], value)
                continue
            self._set_input(self._input_id(name), value)
        self.state = StateVariables(self)
        self.clock = self.state.clock
        self.initializers = dict(initializers or {})
        self.initialize_state()

    def _input_id(self, name: str) -> int:
        """Input slot of a named input variable. With a prescribed flux scheme (PrescribedRadiativeFluxes /
        PrescribedTurbulentFluxes) the fluxes are INPUT variables that carry the names the diagnosed schemes give their
        auxiliary fields (radiative_fluxes.jl:19-23, turbulent_fluxes.jl:13-16)."""
        if name in self._bc_inputs:
            return self._bc_inputs[name]
        if name in abi.INPUT_IDS:
            return abi.INPUT_IDS[name]
        if name in abi.PRESCRIBED_FLUX_INPUTS and self._prescribed_flux_input(name):
            return abi.PRESCRIBED_FLUX_INPUTS[name]
        raise KeyError(f"unknown input variable {name!r}")

    def _prescribed_flux_input(self, name: str) -> bool:
        cfg = self._cfg
        radiative = name in ("surface_shortwave_up", "surface_longwave_up")
        return bool(cfg.radiative if radiative else cfg.turbulent) and cfg.model == abi.TRM_MODEL_LAND

    # ------------------------------------------------------------------------------------------
    def initialize_state(self):
        """``initialize!(integrator)`` (model_integrator.jl:96-109): user/model initializers, then the
        process initialisation inside the library (closures at t0)."""
        # reset!(integrator.state) (model_integrator.jl:98): a re-initialisation starts from zeroed fields and clock
        self._lib.check(self._lib.reset(self._h), "reset")
        fields = dict(self.model.initializer.fields(self.grid))
        # model initializer first, then user field initializers? The reference evaluates the user
        # initializers (:105) *before* the model initializer (:107), so the model initializer wins.
        merged = dict(self.initializers)
        merged.update(fields)
        for name, value in merged.items():
            Field(self, name).set(value)
        self._lib.check(self._lib.initialize(self._h), "initialize")
        return self

    def _ensure_ring_index(self) -> int:
        """Hand the ring positions of this rank's columns to the library once (ColumnRingGrid only)."""
        if not hasattr(self.grid, "mask"):
            raise TypeError("ring-grid conversions need a ColumnRingGrid")
        if not getattr(self, "_ring_ready", False):
            idx = np.ascontiguousarray(np.flatnonzero(self.grid.mask)[self.col0:self.col1], dtype=np.int64)
            self._lib.check(self._lib.set_ring_index(self._h, idx.ctypes.data_as(C.POINTER(C.c_int64)), int(self.grid.npoints)), "set_ring_index")
            self._ring_ready = True
        return int(self.grid.npoints)

    def _local(self, a, shape):
        """Slice a per-column array given for the global domain down to this rank's column range."""
        a = np.asarray(a)
        if a.ndim and a.shape[-1] == self.ncol_global and self.ncol_global != self.ncol:
            a = a[..., self.col0:self.col1]
        return a

    def _evaluate_initializer(self, value, shape) -> np.ndarray:
        nf = self.nf
        if callable(value):
            x = self.grid.xnodes()[self.col0:self.col1].astype(np.float64)
            if len(shape) == 2:
                z = (self.grid.znodes_center() if shape[0] == self.nz else self.grid.znodes_face()).astype(np.float64)
                try:
                    # functions may close over per-column arrays of the *global* domain: slice to this rank's range
                    v = self._local(np.asarray(value(x[None, :], z[:, None]), dtype=np.float64), shape)
                    v = np.broadcast_to(v, shape)
                except Exception:
                    v = np.array([[value(xi, zi) for xi in x] for zi in z], dtype=np.float64)
            else:
                try:
                    v = np.broadcast_to(self._local(np.asarray(value(x), dtype=np.float64), shape), shape)
                except Exception:
                    v = np.array([value(xi) for xi in x], dtype=np.float64)
            return np.ascontiguousarray(v, dtype=nf)
        v = self._local(value, shape)
        return np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=nf), shape), dtype=nf)

    def _percol(self, v) -> np.ndarray:
        v = self._local(v, (self.ncol,))
        return np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=self.nf), (self.ncol,)), dtype=self.nf)

    def _set_input(self, input_id: int, value):
        lib, h = self._lib, self._h
        self._host_callbacks.pop(input_id, None)
        if isinstance(value, Sinusoid):
            m, a, p = self._percol(value.mean), self._percol(value.amp), self._percol(value.phase)
            lib.check(lib.set_input_sinusoid(h, input_id, m.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                             p.ctypes.data_as(C.c_void_p), float(value.period), float(value.lo), float(value.hi)),
                      "set_input_sinusoid")
        elif isinstance(value, TimeSeries):
            times = np.ascontiguousarray(value.times, dtype=np.float64)
            vals = self._local(value.values, None)
            vals = np.ascontiguousarray(np.broadcast_to(np.asarray(vals, dtype=self.nf), (times.size, self.ncol)), dtype=self.nf)
            lib.check(lib.set_input_table(h, input_id, int(times.size), times.ctypes.data_as(C.POINTER(C.c_double)),
                                          vals.ctypes.data_as(C.c_void_p)), "set_input_table")
        elif isinstance(value, RasterInputSource):
            if not hasattr(self.grid, "mask"):
                raise TypeError("RasterInputSource needs a ColumnRingGrid (the raster lives on its ring grid)")
            idx = np.flatnonzero(self.grid.mask)[self.col0:self.col1]   # idxmap = findall(mask), this rank's columns
            vals = np.asarray(value.values)
            if value.times is None:   # static raster: copied once (initialize_from_raster!, TerrariumRastersExt.jl:70-73)
                v = np.ascontiguousarray(vals.reshape(-1)[idx], dtype=self.nf)
                lib.check(lib.set_input_field(h, input_id, v.ctypes.data_as(C.c_void_p)), "set_input_field")
            else:
                times = np.ascontiguousarray(np.asarray(value.times, dtype=np.float64) - float(value.reftime))
                tab = np.ascontiguousarray(vals.reshape(times.size, -1)[:, idx], dtype=self.nf)
                lib.check(lib.set_input_raster(h, input_id, int(times.size), times.ctypes.data_as(C.POINTER(C.c_double)),
                                               tab.ctypes.data_as(C.c_void_p)), "set_input_raster")
        elif callable(value):
            # function valued BC (x, t): evaluated on the host before every step (slow path, kept for
            # API compatibility with e.g. examples/simulations/soil_heat_global.jl:72-93)
            self._host_callbacks[input_id] = value
            self._push_callback(input_id, 0.0)
        elif np.ndim(value) == 0:
            lib.check(lib.set_input_const(h, input_id, float(value)), "set_input_const")
        else:
            v = self._percol(value)
            lib.check(lib.set_input_field(h, input_id, v.ctypes.data_as(C.c_void_p)), "set_input_field")

    def _evaluate_callback(self, input_id, t):
        f = self._host_callbacks[input_id]
        x = self.grid.xnodes()[self.col0:self.col1].astype(self.nf)
        try:
            v = np.broadcast_to(np.asarray(f(x, self.nf(t)), dtype=np.float64), (self.ncol,))
        except Exception:
            v = np.array([f(xi, self.nf(t)) for xi in x], dtype=np.float64)
        return np.ascontiguousarray(v, dtype=self.nf)

    def _push_callback(self, input_id, t, dt=None):
        """Hand a function valued boundary condition f(x, t) to the library for the step starting at `t`: its values at t,
        and -- for Heun, whose second stage re-evaluates the function at the stage clock (heun.jl:53) -- at t + dt."""
        v = self._evaluate_callback(input_id, t)
        if dt is None:
            self._lib.check(self._lib.set_input_field(self._h, input_id, v.ctypes.data_as(C.c_void_p)), "set_input_field")
            return
        v1 = self._evaluate_callback(input_id, self.nf(t) + self.nf(dt))   # tick!(clock, dt) in the clock's number format
        self._lib.check(self._lib.set_input_field_pair(self._h, input_id, v.ctypes.data_as(C.c_void_p), v1.ctypes.data_as(C.c_void_p)),
                        "set_input_field_pair")

    # ------------------------------------------------------------------------------------------
    def step(self, dt: float, nsteps: int = 1):
        """``nsteps`` x ``timestep!(integrator, dt; finalize=false)``."""
        lib = self._lib
        if self._host_callbacks:
            heun = isinstance(self.timestepper, Heun)
            for _ in range(int(nsteps)):
                t = self.clock.time
                for input_id in self._host_callbacks:
                    self._push_callback(input_id, t, float(dt) if heun else None)
                lib.check(lib.step(self._h, float(dt), 1), "step")
        else:
            lib.check(lib.step(self._h, float(dt), int(nsteps)), "step")

    def compute_auxiliary(self):
        self._lib.check(self._lib.compute_auxiliary(self._h), "compute_auxiliary")

    # -- coupled-model exchange through mapped host memory (trm_bind_host_io) --------------------------------------
    def bind_host_io(self, input_name: Optional[str], field_name: Optional[str], nslots: int = 4, write_combined: bool = True):
        """Bind page-locked host rings for a per-step exchange with a host-side coupler (the coupling flow of
        examples/simulations/speedy_dry_land.jl:45-68): the step from iteration ``k`` reads input ``input_name`` from
        ``ring_in[k % nslots]`` and stores field ``field_name`` of the new state in ``ring_out[k % nslots]``; the stage
        kernel accesses the host memory directly. Returns ``(ring_in, ring_out)`` as numpy views ``[nslots, ncol]``."""
        lib, h = self._lib, self._h
        nbytes = int(nslots) * self.ncol * np.dtype(self.nf).itemsize

        def ring(flags):
            q = C.c_void_p()
            lib.check(lib.host_alloc_ex(nbytes, flags, C.byref(q)), "host_alloc_ex")
            self._pinned = getattr(self, "_pinned", []) + [q]
            buf = (C.c_char * nbytes).from_address(q.value)
            return q, np.frombuffer(buf, dtype=self.nf).reshape(int(nslots), self.ncol)

        in_id, p_in, a_in = -1, C.c_void_p(), None
        if input_name is not None:
            in_id = self._bc_inputs[input_name] if input_name in self._bc_inputs else abi.INPUT_IDS[input_name]
            self._host_callbacks.pop(in_id, None)
            p_in, a_in = ring(1 if write_combined else 0)
        out_id, p_out, a_out = -1, C.c_void_p(), None
        if field_name is not None:
            out_id = abi.FIELD_IDS[field_name]
            p_out, a_out = ring(0)
            a_out[...] = 0
        lib.check(lib.bind_host_io(h, in_id, p_in, out_id, p_out, int(nslots)), "bind_host_io")
        return a_in, a_out

    def host_io_wait(self, iteration: int):
        self._lib.check(self._lib.host_io_wait(self._h, int(iteration)), "host_io_wait")

    def step_async(self, dt: float, nsteps: int = 1):
        self._lib.check(self._lib.step_async(self._h, float(dt), int(nsteps)), "step_async")

    def compute_tendencies(self):
        self._lib.check(self._lib.compute_tendencies(self._h), "compute_tendencies")

    def synchronize(self):
        self._lib.check(self._lib.sync(self._h), "sync")

    def diagnostics(self) -> Dict[str, float]:
        d = abi.trm_diag()
        self._lib.check(self._lib.diagnostics(self._h, C.byref(d)), "diagnostics")
        return {n: getattr(d, n) for n, _ in abi.trm_diag._fields_}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.destroy(self._h)
            self._h = abi._H()
            for q in getattr(self, "_pinned", []):
                self._lib.host_free(q)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __repr__(self):
        return (f"Integrator of {type(self.model).__name__} with {type(self.timestepper).__name__}\n"
                f"├── Current time: {self.clock.time}\n├── columns [{self.col0}, {self.col1}) of {self.ncol_global}, Nz={self.nz}")


# ---------------------------------------------------------------------------------------------
# reference API
# ---------------------------------------------------------------------------------------------
def initialize(model, timestepper, inputs: Optional[Dict[str, Any]] = None, *, boundary_conditions=None,
               initializers=None, partition=None, math: str = "faithful") -> ModelIntegrator:
    """``initialize(model, timestepper, inputs...; boundary_conditions, initializers)``."""
    return ModelIntegrator(model, timestepper, inputs, boundary_conditions, initializers, partition, math)


def current_time(integrator) -> float:
    return integrator.clock.time


def convert_dt(dt) -> float:
    """``convert_dt`` (utils/utils.jl:17-18): numbers are seconds, ``datetime.timedelta`` is converted."""
    return float(dt.total_seconds()) if hasattr(dt, "total_seconds") else float(dt)


def timestep(integrator, dt=None, finalize: bool = True) -> None:
    """``timestep!(integrator, dt; finalize = true)``."""
    dt = default_dt(integrator.timestepper) if dt is None else convert_dt(dt)
    integrator.step(dt, 1)
    if finalize:
        integrator.compute_auxiliary()


def get_steps(steps, period, dt) -> int:
    if steps is None and period is None:
        raise ValueError("either `steps` or `period` must be specified")
    if steps is not None and period is not None:
        raise ValueError("both `steps` and `period` cannot be specified")
    if steps is not None:
        return int(steps)
    return int(convert_dt(period) // dt)


def run(integrator, steps=None, period=None, dt=None):
    """``run!(integrator; steps, period, Δt)``: one C-ABI call for all steps, then ``compute_auxiliary!``."""
    dt = default_dt(integrator.timestepper) if dt is None else convert_dt(dt)
    n = get_steps(steps, period, dt)
    integrator.step(dt, n)
    integrator.compute_auxiliary()
    return integrator
