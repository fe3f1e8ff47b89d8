"""``Simulation`` / output writers / callbacks around a ``ModelIntegrator``.

The reference inherits these from Oceananigans (``src/timesteppers/model_integrator.jl:45-66``,
``docs/src/running/time_stepping.md:86-180``): ``Simulation(integrator; stop_time, Δt)``, ``sim.output_writers[:name] =
JLD2Writer(integrator, fields; filename, schedule)``, ``sim.callbacks[:name] = Callback(f, schedule)``, ``run!(sim)``,
``FieldTimeSeries(file, name)``. This is the host mirror of that surface for the B200 library:

* between two scheduled events the steps go to the library in ONE call (``trm_step_async``), so the stage kernels run
  back to back exactly as in ``run!``;
* a snapshot is taken with ``trm_get_field_async`` into page-locked buffers (``trm_host_alloc``) and written to disk one
  event later, so the device-to-host copy and the file write overlap the following steps;
* the file format is NetCDF-3 (``scipy.io.netcdf_file``; JLD2 is a Julia format and not available here).

Semantics kept from Oceananigans: writers and callbacks fire once at the start of ``run`` and then whenever their
schedule actuates; with ``align_time_step`` (default) the last step before a ``TimeInterval`` event is shortened to land
on it; ``compute_auxiliary!`` runs before anything observes the state (every step when ``finalize_every_step``, which is
the reference behaviour -- ``time_step!`` forwards to ``timestep!(integrator, Δt)`` with ``finalize = true`` -- and matters
for the vegetated LandModel, whose stomatal conductance reads the net assimilation of the previous evaluation).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Sequence

import numpy as np

from . import _abi as abi
from .integrator import Field, ModelIntegrator, convert_dt, default_dt


# ---------------------------------------------------------------------------------------------
# schedules (Oceananigans.Utils)
# ---------------------------------------------------------------------------------------------
@dataclass
class TimeInterval:
    interval: float

    def __post_init__(self):
        self.interval = convert_dt(self.interval)
        self._next = self.interval

    def reset(self, t0: float):
        self._next = t0 + self.interval

    def next_time(self) -> Optional[float]:
        return self._next

    def actuate(self, time: float, iteration: int) -> bool:
        if time >= self._next - 1e-9 * max(1.0, abs(self._next)):
            while time >= self._next - 1e-9 * max(1.0, abs(self._next)):
                self._next += self.interval
            return True
        return False


class AveragedTimeInterval(TimeInterval):
    """``AveragedTimeInterval(interval)`` (window = interval, stride = 1): the writer stores the time average over the
    window that ends at each output time, ``sum_k field(t_k) dt_k / sum_k dt_k`` with the field after every step
    (Oceananigans' ``WindowedTimeAverage`` accumulates the same right-endpoint sum). The sum lives on the device
    (``trm_accumulate``); the record written at the start of the run is the instantaneous field."""
    averaged = True


@dataclass
class IterationInterval:
    interval: int
    offset: int = 0

    def reset(self, t0: float):
        pass

    def next_time(self) -> Optional[float]:
        return None

    def steps_until(self, iteration: int) -> int:
        return self.interval - ((iteration - self.offset) % self.interval)

    def actuate(self, time: float, iteration: int) -> bool:
        return (iteration - self.offset) % self.interval == 0


@dataclass
class Callback:
    func: Callable
    schedule: object = None

    def __post_init__(self):
        if self.schedule is None:
            self.schedule = IterationInterval(1)


# ---------------------------------------------------------------------------------------------
# output
# ---------------------------------------------------------------------------------------------
class _Pinned:
    """A page-locked host array obtained from the library (``trm_host_alloc``)."""

    def __init__(self, lib, shape, dtype):
        self._lib, self.shape, self.dtype = lib, tuple(shape), np.dtype(dtype)
        self._ptr = C.c_void_p()
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        lib.check(lib.host_alloc(nbytes, C.byref(self._ptr)), "host_alloc")
        buf = (C.c_char * nbytes).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            self._lib.host_free(self._ptr)
            self._ptr = C.c_void_p()


class NetCDFWriter:
    """``JLD2Writer(integrator, outputs; filename, schedule, overwrite_existing, including)`` with a NetCDF-3 file.

    ``outputs``: mapping ``name -> Field`` (or field names). Every record appends along the unlimited ``time`` dimension;
    3-D fields are ``(time, z, column)`` with layer 0 = bottom cell, z-face fields ``(time, zf, column)``, 2-D fields
    ``(time, column)``. ``including=("grid",)`` stores the vertical coordinates and, on a ``ColumnRingGrid``, the ring position
    (and longitude / latitude) of every column, so that ``FieldTimeSeries.ring(i)`` can rebuild the global map."""

    def __init__(self, integrator: ModelIntegrator, outputs, filename: str, schedule, overwrite_existing: bool = False,
                 including: Sequence[str] = ("grid",)):
        from scipy.io import netcdf_file
        if isinstance(outputs, (list, tuple)):
            outputs = {n: n for n in outputs}
        self.integrator = integrator
        self.fields: Dict[str, Field] = {k: (v if isinstance(v, Field) else Field(integrator, v)) for k, v in outputs.items()}
        # one file per rank of a partitioned run (each rank holds its own column range): <name>_rank<r>.<ext>
        rank_world = getattr(integrator, "partition", None)
        if rank_world is not None and rank_world[1] > 1:
            stem, ext = os.path.splitext(filename)
            filename = f"{stem}_rank{rank_world[0]}{ext}"
        self.filename, self.schedule = filename, schedule
        if os.path.exists(filename) and not overwrite_existing:
            raise FileExistsError(f"{filename} exists (pass overwrite_existing=True)")
        f = netcdf_file(filename, "w")
        nz, nc = integrator.nz, integrator.ncol
        f.createDimension("time", None)
        f.createDimension("z", nz); f.createDimension("zf", nz + 1); f.createDimension("column", nc)
        self._time = f.createVariable("time", "f8", ("time",))
        self._time.units = "s"
        f.column_range_start = np.int32(integrator.col0)     # [col0, col1) of the global column axis held by this file
        f.column_range_stop = np.int32(integrator.col1)
        f.columns_global = np.int32(integrator.ncol_global)
        if "grid" in including:
            zc = f.createVariable("z", "f8", ("z",)); zc[:] = integrator.grid.znodes_center().astype(np.float64)
            zf = f.createVariable("zf", "f8", ("zf",)); zf[:] = integrator.grid.znodes_face().astype(np.float64)
            grid = integrator.grid
            if hasattr(grid, "mask"):   # ColumnRingGrid: where each column sits on the ring grid (column_ring_grid.jl:46-54)
                idx = np.flatnonzero(grid.mask)[integrator.col0:integrator.col1]
                ri = f.createVariable("ring_index", "i4", ("column",)); ri[:] = idx.astype(np.int32)
                ri.long_name = "0-based position of the column in the ring grid"
                f.ring_points = np.int32(grid.npoints)
                for cname, coord in (("lon", grid.lon), ("lat", grid.lat)):
                    if coord is not None:
                        cv = f.createVariable(cname, "f8", ("column",)); cv[:] = coord[idx]
                        cv.units = "radians"
        code = "f8" if np.dtype(integrator.nf) == np.float64 else "f4"
        self._vars = {}
        for name, fld in self.fields.items():
            dims = {2: ("time", "z" if fld.shape[0] == nz else "zf", "column"), 1: ("time", "column")}[len(fld.shape)]
            self._vars[name] = f.createVariable(name, code, dims)
        self._file, self._nrec = f, 0
        self._pending = None       # (time, {name: pinned buffer}) downloaded asynchronously, not yet written
        self._buffers = []         # two sets of pinned buffers, used alternately
        self._async = hasattr(integrator._lib, "get_field_async") and hasattr(integrator._lib, "host_alloc")
        self.averaged = bool(getattr(schedule, "averaged", False))
        self._elapsed = 0.0        # length of the averaging window accumulated so far

    def accumulate(self, dt: float):
        """Add ``dt * field`` of every output to its device-side accumulator (time-averaged writers, after each step)."""
        lib, h = self.integrator._lib, self.integrator._h
        for n, f in self.fields.items():
            lib.check(lib.accumulate(h, f.id, float(dt)), f"accumulate({n})")
        self._elapsed += float(dt)

    # -- snapshot now (enqueue), write later -------------------------------------------------
    def snapshot(self, time: float):
        integ, lib = self.integrator, self.integrator._lib
        if self.averaged and self._elapsed > 0:
            self.flush()   # (the instantaneous first record may still be in flight)
            arrays = {}
            for n, f in self.fields.items():
                out = np.empty(f.shape, dtype=integ.nf)
                lib.check(lib.get_accumulated(integ._h, f.id, out.ctypes.data_as(C.c_void_p), out.size, 1.0 / self._elapsed, 1), f"get_accumulated({n})")
                arrays[n] = out
            self._elapsed = 0.0
            self._write(time, arrays)
            return
        if not self._async:   # engine without the asynchronous entry points (the CPU checker in tests)
            self._write(time, {n: f.numpy() for n, f in self.fields.items()})
            return
        self.flush()          # the previous snapshot has had a whole chunk of steps to arrive
        if len(self._buffers) < 2:
            self._buffers.append({n: _Pinned(lib, f.shape, integ.nf) for n, f in self.fields.items()})
        bufs = self._buffers[self._nrec % 2] if len(self._buffers) == 2 else self._buffers[-1]
        for n, f in self.fields.items():
            lib.check(lib.get_field_async(integ._h, f.id, C.c_void_p(bufs[n]._ptr.value), int(np.prod(f.shape))), f"get_field_async({n})")
        self._pending = (time, bufs)

    def flush(self):
        if self._pending is not None:
            self.integrator._lib.check(self.integrator._lib.sync(self.integrator._h), "sync")
            time, bufs = self._pending
            self._write(time, {n: b.array for n, b in bufs.items()})
            self._pending = None

    def _write(self, time, arrays):
        i = self._nrec
        self._time[i] = time
        for n, a in arrays.items():
            self._vars[n][i] = a
        self._nrec += 1

    def close(self):
        self.flush()
        if self._file is not None:
            self._file.close()
            self._file = None
        for bufs in self._buffers:
            for b in bufs.values():
                b.free()
        self._buffers = []


class FieldTimeSeries:
    """``FieldTimeSeries(file, name)``: the saved records of one field, ``series[i]`` = record ``i`` (``-1`` = last)."""

    def __init__(self, filename: str, name: str):
        from scipy.io import netcdf_file
        with netcdf_file(filename, "r", mmap=False) as f:
            self.times = np.array(f.variables["time"][:], dtype=np.float64)
            self.data = np.array(f.variables[name][:])
            self.z = np.array(f.variables["z"][:]) if "z" in f.variables else None
            self.ring_index = np.array(f.variables["ring_index"][:], dtype=np.int64) if "ring_index" in f.variables else None
            self.ring_points = int(getattr(f, "ring_points", 0))
        self.name = name

    def __len__(self):
        return self.times.size

    def __getitem__(self, i):
        return self.data[i]

    def ring(self, i, fill_value=np.nan) -> np.ndarray:
        """Record ``i`` on the full ring grid (``RingGrids.Field(field, grid; fill_value)``, column_ring_grid.jl:102-125):
        ``[..., ring point]`` with ``fill_value`` at the points without a column."""
        if self.ring_index is None:
            raise ValueError("the file was not written from a ColumnRingGrid (no ring_index)")
        rec = np.asarray(self.data[i])
        out = np.full(rec.shape[:-1] + (self.ring_points,), fill_value, dtype=rec.dtype)
        out[..., self.ring_index] = rec
        return out


# ---------------------------------------------------------------------------------------------
# simulation
# ---------------------------------------------------------------------------------------------
class Simulation:
    """``Simulation(integrator; Δt, stop_time, stop_iteration)`` with ``output_writers`` and ``callbacks`` dictionaries."""

    def __init__(self, integrator: ModelIntegrator, dt=None, stop_time=None, stop_iteration=None, align_time_step: bool = True,
                 finalize_every_step: Optional[bool] = None):
        self.model = self.integrator = integrator
        self.dt = default_dt(integrator.timestepper) if dt is None else convert_dt(dt)
        self.stop_time = None if stop_time is None else convert_dt(stop_time)
        self.stop_iteration = stop_iteration
        if self.stop_time is None and stop_iteration is None:
            raise ValueError("Simulation needs stop_time or stop_iteration")
        self.align_time_step = align_time_step
        if finalize_every_step is None:
            # the reference's Simulation calls timestep!(...; finalize = true) every step (model_integrator.jl:64,125-131).
            # compute_auxiliary! is not idempotent for ANY LandModel: the surface block reads the stored skin temperature and
            # writes a new one (ImplicitSkinTemperature, skin_temperature.jl:62-80); the vegetated model additionally reads the
            # net assimilation of the previous evaluation. Only a SoilModel, whose auxiliaries are never read back, may batch
            # steps without changing the trajectory.
            from .models import LandModel
            finalize_every_step = isinstance(integrator.model, LandModel)
        self.finalize_every_step = finalize_every_step
        self.output_writers: Dict[str, NetCDFWriter] = {}
        self.callbacks: Dict[str, Callback] = {}
        self.steps_taken = 0

    # the steps that can go to the library in one call: up to the next event of any schedule / the stop criterion
    def _chunk(self, t: float, it: int):
        n_max = math.inf
        t_next = math.inf
        if self.stop_iteration is not None:
            n_max = min(n_max, self.stop_iteration - it)
        if self.stop_time is not None:
            t_next = min(t_next, self.stop_time)
        for obj in list(self.output_writers.values()) + list(self.callbacks.values()):
            sch = obj.schedule
            if isinstance(sch, IterationInterval):
                n_max = min(n_max, sch.steps_until(it))
            else:
                nt = sch.next_time()
                if nt is not None:
                    t_next = min(t_next, nt)
        return n_max, t_next

    def _done(self, t: float, it: int) -> bool:
        if self.stop_iteration is not None and it >= self.stop_iteration:
            return True
        return self.stop_time is not None and t >= self.stop_time - 1e-9 * max(1.0, abs(self.stop_time))

    def _observe(self, t: float, it: int, first: bool = False):
        due_w = [w for w in self.output_writers.values() if first or w.schedule.actuate(t, it)]
        due_c = [c for c in self.callbacks.values() if first or c.schedule.actuate(t, it)]
        if (due_w or due_c) and not self.finalize_every_step and not any(w.averaged for w in self.output_writers.values()):
            self.integrator.compute_auxiliary()
        for w in due_w:
            w.snapshot(t)
        for c in due_c:
            c.func(self)

    def run(self):
        integ = self.integrator
        t, it = integ.clock.time, integ.clock.iteration
        if self._done(t, it):
            return self   # (like Oceananigans: re-initialise the integrator before running again)
        for obj in list(self.output_writers.values()) + list(self.callbacks.values()):
            obj.schedule.reset(t)
        integ.compute_auxiliary()
        self._observe(t, it, first=True)
        lib = integ._lib
        averaging = [w for w in self.output_writers.values() if w.averaged]
        single = self.finalize_every_step or bool(averaging)   # one step at a time, auxiliaries current after each

        def after_step(dt_step):
            if single:
                integ.compute_auxiliary()
            for w in averaging:
                w.accumulate(dt_step)

        use_async = hasattr(lib, "step_async") and not integ._host_callbacks and not single
        while not self._done(t, it):
            n_max, t_next = self._chunk(t, it)
            n_time = math.floor((t_next - t) / self.dt + 1e-9) if math.isfinite(t_next) else math.inf
            n_full = int(min(n_max, n_time))   # (one of the two is finite: the simulation has a stop criterion)
            if n_full >= 1:
                if single:
                    for _ in range(n_full):
                        integ.step(self.dt, 1)
                        after_step(self.dt)
                elif use_async:
                    lib.check(lib.step_async(integ._h, float(self.dt), n_full), "step_async")
                else:
                    integ.step(self.dt, n_full)
                self.steps_taken += n_full
            elif self.align_time_step and math.isfinite(t_next) and t_next - t > 1e-9 * max(1.0, abs(t_next)):
                integ.step(t_next - t, 1)   # shortened step that lands on the scheduled time
                after_step(t_next - t)
                self.steps_taken += 1
            else:   # an unaligned event time with alignment switched off: step over it
                integ.step(self.dt, 1)
                after_step(self.dt)
                self.steps_taken += 1
            t, it = integ.clock.time, integ.clock.iteration
            self._observe(t, it)
        integ.synchronize()
        for w in self.output_writers.values():
            w.flush()
        if not single:
            integ.compute_auxiliary()
        return self

    def close(self):
        for w in self.output_writers.values():
            w.close()


def run_simulation(sim: Simulation) -> Simulation:
    """``run!(sim)``."""
    return sim.run()
