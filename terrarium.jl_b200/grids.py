"""Vertical discretisations and column grids (host side mirror of ``src/grids`` of the reference).

* ``UniformSpacing`` / ``ExponentialSpacing`` / ``PrescribedSpacing`` follow
  ``src/grids/vertical_discretization.jl:30-93``;
* ``ColumnGrid`` follows ``src/grids/column_grid.jl:9-39``: ``z_faces = [-reverse(cumsum(dz)); 0]``
  where the thickness list runs top -> bottom, so face/cell index 0 is the BOTTOM of the column;
* ``ColumnRingGrid`` follows ``src/grids/column_ring_grid.jl:37-59``: the columns are the ``True``
  points of a mask over a ring grid, kept in ring order (a 1-D index).  RingGrids.jl itself is not
  available here, so the ring grid is represented by its mask (and optional lon/lat vectors).

A grid also owns the column partition used for multi-GPU runs: ``grid.partition(rank, world)`` returns
the contiguous column range of one rank (columns never exchange data, so there is no halo).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np


class B200:
    """Architecture tag: the only architecture of this package (there is no CPU fallback)."""

    def __init__(self, device: int = 0):
        self.device = int(device)

    def __repr__(self):
        return f"B200(device={self.device})"


def _round_sigdigits(x: float, sig: int) -> float:
    """Julia ``round(x; sigdigits=sig)`` (base/floatfuncs.jl ``_round_sigdigits``/``_round_digits``)."""
    if x == 0 or not math.isfinite(x):
        return x
    h = math.floor(math.log10(abs(x))) + 1
    digits = sig - h
    if digits >= 0:
        sc = 10.0 ** digits
        return float(np.rint(x * sc) / sc)
    sc = 10.0 ** (-digits)
    return float(np.rint(x / sc) * sc)


@dataclass
class UniformSpacing:
    dz: float = 0.1
    N: int = 100

    def thicknesses(self) -> np.ndarray:
        return np.full(self.N, float(self.dz))


@dataclass
class ExponentialSpacing:
    dz_min: float = 0.05
    dz_max: float = 100.0
    N: int = 50
    sig: Optional[int] = 3

    def thicknesses(self) -> np.ndarray:
        assert self.N > 1, "number of grid points for exponential spacing must be > 1"
        l0, ln = math.log2(self.dz_min), math.log2(self.dz_max)
        out = []
        for i in range(1, self.N + 1):
            li = l0 + (i - 1) * (ln - l0) / (self.N - 1)
            v = 2.0 ** li
            out.append(v if self.sig is None else _round_sigdigits(v, self.sig))
        return np.asarray(out, dtype=np.float64)


@dataclass
class PrescribedSpacing:
    dz: Sequence[float]

    @property
    def N(self) -> int:
        return len(self.dz)

    def thicknesses(self) -> np.ndarray:
        return np.asarray(self.dz, dtype=np.float64)


def num_layers(spacing) -> int:
    return int(spacing.N)


def get_spacing(spacing) -> np.ndarray:
    return spacing.thicknesses()


def z_faces_from_spacing(spacing, nf) -> np.ndarray:
    """``convert.(NF, vcat(-reverse(cumsum(z_thick)), 0))`` -- returned as float64 holding NF values."""
    dz = get_spacing(spacing)  # top -> bottom, Float64 like the (Float64 typed) spacing structs
    faces = np.concatenate([-np.cumsum(dz)[::-1], [0.0]])
    return faces.astype(nf).astype(np.float64)


class ColumnGrid:
    """A set of laterally independent vertical columns (``x`` = column index, ``z`` = vertical)."""

    def __init__(self, *args, num_columns: Optional[int] = None):
        # ColumnGrid(arch, NF, vert, n) | ColumnGrid(arch, vert, n) | ColumnGrid(vert, n)   (column_grid.jl:20-38)
        args = list(args)
        arch = args.pop(0) if args and isinstance(args[0], B200) else B200()
        is_vert = lambda a: hasattr(a, "thicknesses")
        nf = args.pop(0) if args and not is_vert(args[0]) and not isinstance(args[0], (int, np.integer)) else np.float64
        vert = args.pop(0) if args and is_vert(args[0]) else ExponentialSpacing()
        n = args.pop(0) if args else 1
        if num_columns is not None:
            n = num_columns
        self.arch = arch
        self.nf = np.dtype(nf).type
        self.vert = vert
        self.num_columns = int(n)
        if self.num_columns < 1:
            raise ValueError("num_columns must be >= 1")
        self.Nz = num_layers(self.vert)
        self.z_faces = z_faces_from_spacing(self.vert, self.nf)

    # -- geometry helpers (Oceananigans znodes) ------------------------------------------------
    @property
    def Nc(self) -> int:
        return self.num_columns

    def znodes_center(self) -> np.ndarray:
        f = self.z_faces.astype(self.nf)
        return ((f[1:] + f[:-1]) / 2).astype(self.nf)

    def znodes_face(self) -> np.ndarray:
        return self.z_faces.astype(self.nf)

    def dz(self) -> np.ndarray:
        f = self.z_faces.astype(self.nf)
        return (f[1:] - f[:-1]).astype(self.nf)

    def xnodes(self) -> np.ndarray:
        """x-node of every column: (i - 1/2)/Nc on x in (0, 1) [OCN]."""
        i = np.arange(1, self.Nc + 1, dtype=np.float64)
        return ((i - 0.5) / self.Nc).astype(self.nf)

    def partition(self, rank: int, world: int):
        """Contiguous, near-equal column range [c0, c1) of ``rank`` out of ``world``."""
        base, rem = divmod(self.Nc, world)
        c0 = rank * base + min(rank, rem)
        return c0, c0 + base + (1 if rank < rem else 0)

    def __repr__(self):
        return f"{type(self).__name__}{{{np.dtype(self.nf).name}}}(Nc={self.Nc}, Nz={self.Nz}) on {self.arch}"


class ColumnRingGrid(ColumnGrid):
    """Global grid of independent columns at the ``True`` points of ``mask`` (ring order)."""

    def __init__(self, *args, mask=None, lon=None, lat=None):
        # ColumnRingGrid(arch, NF, vert, mask) | (arch, vert, mask) | (vert, mask); NF defaults to Float32
        args = list(args)
        if mask is None:
            mask = args.pop()  # last positional argument
        head = [a for a in args]
        if not any((not isinstance(a, B200)) and (not hasattr(a, "thicknesses")) for a in head):
            # no explicit number format: reference default is Float32 (column_ring_grid.jl:66-71)
            idx = 1 if head and isinstance(head[0], B200) else 0
            head.insert(idx, np.float32)
        mask = np.asarray(mask, dtype=bool).ravel()
        super().__init__(*head, num_columns=int(mask.sum()))
        self.mask = mask
        self.npoints = mask.size
        self.lon = None if lon is None else np.asarray(lon, dtype=np.float64).ravel()
        self.lat = None if lat is None else np.asarray(lat, dtype=np.float64).ravel()

    @classmethod
    def from_land_sea_mask(cls, *args, path: str, variable: str = "lsm", threshold: float = 0.5):
        """``ColumnRingGrid(arch, NF, spacing, rings, land_sea_frac .> 0.5)`` from a land fraction raster file as in
        ``examples/simulations/soil_heat_global.jl:29-38``: the raster (``[time = 1,] lat, lon``, north to south) is
        converted to the grid's number format first, its storage order is the ring order of a full grid
        (``FullGaussianGrid(Matrix(raster), input_as = Matrix)``), points with a land fraction above ``threshold``
        become columns; ``lon`` / ``lat`` of the ring points are kept in radians as ``RingGrids.get_lonlats`` gives them
        (:39). ``args`` = ``([arch,] [NF,] spacing)``; reads NetCDF-3 and NetCDF-4 files."""
        from .models import RasterInputSource
        nf = next((a for a in args if not isinstance(a, B200) and not hasattr(a, "thicknesses")), np.float32)
        import warnings
        from . import netcdf4
        lon = lat = None
        src = RasterInputSource.from_netcdf(path, variable)
        frac = np.asarray(src.values, dtype=np.float64)
        if frac.ndim == 2:
            if frac.shape[0] != 1:
                raise ValueError(f"{path}:{variable} has {frac.shape[0]} time slices; a land-sea mask has one (dropdims over Ti)")
            frac = frac[0]
        if netcdf4.is_hdf5(path):
            with netcdf4.File(path) as f:
                dims = f.variables[variable].dimensions[-2:]
                if all(d in f.variables for d in dims):
                    la, lo = (np.asarray(f.variables[d].read(), dtype=np.float64) for d in dims)
                    lat, lon = np.deg2rad(np.repeat(la, lo.size)), np.deg2rad(np.tile(lo, la.size))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mask = frac.astype(nf) > nf(threshold)
        return cls(*args, mask=mask, lon=lon, lat=lat)

    def xnodes(self) -> np.ndarray:
        """x in (1, Nh): x_i = 1 + (i - 1/2)(Nh - 1)/Nh so that round(x_i) == i (column_ring_grid.jl:54)."""
        n = self.Nc
        i = np.arange(1, n + 1, dtype=np.float64)
        return (1 + (i - 0.5) * (n - 1) / n).astype(self.nf)

    # masked-column <-> full ring conversion (column_ring_grid.jl:102-149)
    def to_ring(self, field: np.ndarray, fill_value=np.nan) -> np.ndarray:
        field = np.asarray(field)
        if field.ndim == 1:
            out = np.full(self.npoints, fill_value, dtype=field.dtype)
            out[self.mask] = field
            return out
        out = np.full((field.shape[0], self.npoints), fill_value, dtype=field.dtype)
        out[:, self.mask] = field
        return out

    def from_ring(self, ring_field: np.ndarray) -> np.ndarray:
        ring_field = np.asarray(ring_field)
        return ring_field[..., self.mask]

    def masked_lonlat(self):
        return self.lon[self.mask], self.lat[self.mask]
