"""Host-side mirror of the reference's model / process configuration types for the hot path.

Same names, keyword arguments and defaults as the Julia constructors they mirror (file:line below,
relative to the reference root), reduced to what the per-column time-step consumes.  These objects
hold *configuration only*; all arithmetic runs in the CUDA library behind the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Optional, Sequence, Union

import numpy as np

from . import _abi as abi
from .grids import ColumnGrid


# ---------------------------------------------------------------------------------------------
# physical constants and soil parameterisations
# ---------------------------------------------------------------------------------------------
@dataclass
class PhysicalConstants:  # src/processes/physical_constants.jl:9-51
    rho_w: float = 1000.0
    rho_i: float = 916.2
    rho_a: float = 1.293
    c_a: float = 1005.7
    Lsl: float = 3.34e5
    Llg: float = 2.257e6
    Lsg: float = 2.834e6
    g: float = 9.80665
    Tref: float = 273.15
    sigma: float = 5.6704e-8
    kappa: float = 0.4
    eps: float = 0.622
    R_a: float = 287.058
    C_mass: float = 12.0


@dataclass
class SoilThermalConductivities:  # soil_thermal_properties.jl:13-24
    water: float = 0.57
    ice: float = 2.2
    air: float = 0.025
    mineral: float = 3.8
    organic: float = 0.25


@dataclass
class SoilHeatCapacities:  # soil_thermal_properties.jl:34-45
    water: float = 4.2e6
    ice: float = 1.9e6
    air: float = 0.00125e6
    mineral: float = 2.0e6
    organic: float = 2.5e6


@dataclass
class SoilThermalProperties:  # soil_thermal_properties.jl:56-76 (InverseQuadratic + FreeWater only)
    conductivities: SoilThermalConductivities = field(default_factory=SoilThermalConductivities)
    heat_capacities: SoilHeatCapacities = field(default_factory=SoilHeatCapacities)


@dataclass
class SoilEnergyBalance:  # soil_energy.jl:23-44 (ExplicitTwoPhaseHeatConduction + SoilEnergyTemperatureClosure)
    thermal_properties: SoilThermalProperties = field(default_factory=SoilThermalProperties)


@dataclass
class VanGenuchten:  # FreezeCurves.jl SWRC (soil_hydraulic_closures.jl:95-97,115-118)
    alpha: float = 1.0
    n: float = 2.0
    theta_res: float = 0.0


@dataclass
class BrooksCorey:  # FreezeCurves.jl SWRC
    psi_s: float = 0.01
    lam: float = 0.2
    theta_res: float = 0.0


@dataclass
class UnsatKLinear:  # soil_hydraulic_properties.jl:163-182
    pass


@dataclass
class UnsatKVanGenuchten:  # soil_hydraulic_properties.jl:196-221
    impedance: float = 7.0


@dataclass
class ConstantSoilHydraulics:  # soil_hydraulic_properties.jl:62-91
    swrc: Any = field(default_factory=BrooksCorey)
    unsat_hydraulic_cond: Any = field(default_factory=UnsatKLinear)
    sat_hydraulic_cond: float = 1.0e-5
    field_capacity: float = 0.25
    wilting_point: float = 0.05


@dataclass
class SoilHydraulicsSURFEX(ConstantSoilHydraulics):  # soil_hydraulic_properties.jl:112-156 (same K path)
    # field capacity and wilting point follow the clay content of the (homogeneous) soil texture
    wilting_point_coef: float = 37.13e-3
    field_capacity_coef: float = 89.0e-3
    field_capacity_exp: float = 0.35

    def capacities(self, texture, nf):
        """``(field_capacity, wilting_point)`` evaluated in the number format ``nf`` like the reference kernels (:145-156)."""
        clay100 = nf(texture.clay) * nf(100)
        return (float(nf(self.field_capacity_coef) * clay100 ** nf(self.field_capacity_exp)),
                float(nf(self.wilting_point_coef) * np.sqrt(clay100)))


class NoFlow:  # soil_hydrology.jl:13
    pass


class RichardsEq:  # soil_hydrology_rre.jl:18
    pass


@dataclass
class SoilHydrology:  # soil_hydrology.jl:21-51
    vertical_flow: Any = field(default_factory=NoFlow)
    hydraulic_properties: Any = field(default_factory=SoilHydraulicsSURFEX)
    vwc_forcing: Optional[float] = None  # constant source/sink [1/s] (test/soil/soil_hydrology_tests.jl:191-233)


@dataclass
class ConstantSoilPorosity:  # soil_porosity.jl:7-13
    mineral_porosity: float = 0.49
    organic_porosity: float = 0.9


@dataclass
class SoilTexture:  # stratigraphy/soil_texture.jl:6-20 (named textures: :35-39)
    sand: float = 1.0
    clay: float = 0.0
    silt: Optional[float] = None

    def __post_init__(self):
        if self.silt is None:
            self.silt = 1.0 - self.sand - self.clay
        for v in (self.sand, self.silt, self.clay):
            if not -1e-12 <= v <= 1.0 + 1e-12:
                raise ValueError("sand, silt and clay fractions must lie in [0, 1]")
        if abs(self.sand + self.silt + self.clay - 1.0) > 1e-8:
            raise ValueError("sand, silt, and clay fractions must sum to unity")

    @classmethod
    def named(cls, name: str):
        return cls(**{"sand": dict(sand=1.0, silt=0.0, clay=0.0), "silt": dict(sand=0.0, silt=1.0, clay=0.0),
                      "clay": dict(sand=0.0, silt=0.0, clay=1.0), "sandyclay": dict(sand=0.5, silt=0.0, clay=0.5),
                      "siltyclay": dict(sand=0.0, silt=0.5, clay=0.5)}[name])


@dataclass
class HomogeneousStratigraphy:  # homogeneous_strat.jl:8-23
    porosity: ConstantSoilPorosity = field(default_factory=ConstantSoilPorosity)
    texture: SoilTexture = field(default_factory=SoilTexture)


@dataclass
class ConstantSoilCarbonDensity:  # constant_soil_carbon.jl:10-16
    rho_soc: float = 0.0
    rho_org: float = 1300.0


@dataclass
class SoilEnergyWaterCarbon:  # soil_coupled.jl:7-35
    strat: HomogeneousStratigraphy = field(default_factory=HomogeneousStratigraphy)
    energy: SoilEnergyBalance = field(default_factory=SoilEnergyBalance)
    hydrology: SoilHydrology = field(default_factory=SoilHydrology)
    biogeochem: ConstantSoilCarbonDensity = field(default_factory=ConstantSoilCarbonDensity)


# ---------------------------------------------------------------------------------------------
# surface processes (bare ground)
# ---------------------------------------------------------------------------------------------
@dataclass
class ConstantAlbedo:  # surface_energy/albedo.jl:20-27
    albedo: float = 0.3
    emissivity: float = 0.97


@dataclass
class ImplicitSkinTemperature:  # skin_temperature.jl:52-55
    kappa_s: float = 2.0


@dataclass
class PrescribedSkinTemperature:  # skin_temperature.jl:12-15
    kappa_s: float = 2.0


class PrescribedAlbedo:  # albedo.jl:7-14: albedo and emissivity are the input variables `albedo` / `emissivity`
    pass


class DiagnosedRadiativeFluxes:  # radiative_fluxes.jl:72-76
    pass


class PrescribedRadiativeFluxes:  # radiative_fluxes.jl:13-23: inputs `surface_shortwave_up` / `surface_longwave_up`
    pass


class DiagnosedTurbulentFluxes:  # turbulent_fluxes.jl:20-28
    pass


class PrescribedTurbulentFluxes:  # turbulent_fluxes.jl:9-16: inputs `sensible_heat_flux` / `latent_heat_flux`
    pass


@dataclass
class SurfaceEnergyBalance:  # surface_energy_balance.jl:9-38
    skin_temperature: Any = field(default_factory=ImplicitSkinTemperature)
    albedo: Any = field(default_factory=ConstantAlbedo)
    radiative_fluxes: Any = field(default_factory=DiagnosedRadiativeFluxes)
    turbulent_fluxes: Any = field(default_factory=DiagnosedTurbulentFluxes)


@dataclass
class DirectSurfaceRunoff:  # runoff/direct_surface_runoff.jl:15-18
    tau_r: float = 3600.0


class SoilMoistureResistanceFactor:  # evapotranspiration/ground_resistance_factor.jl:32-57 (Lee & Pielke 1992)
    pass


@dataclass
class ConstantEvaporationResistanceFactor:  # ground_resistance_factor.jl:6-11
    factor: float = 1.0


@dataclass
class BareGroundEvaporation:  # evapotranspiration/bare_ground_evaporation.jl:12-20
    # a number / ConstantEvaporationResistanceFactor, or SoilMoistureResistanceFactor()
    ground_resistance_factor: Any = 1.0


class NoCanopyInterception:  # canopy_interception.jl:7
    pass


@dataclass
class PALADYNCanopyInterception:  # canopy_interception.jl:33-45
    alpha_int: float = 0.2
    k_ext: float = 0.5
    w_can_max: float = 2.0e-4
    tau_w: float = 86400.0


@dataclass
class PALADYNCanopyEvapotranspiration:  # canopy_evapotranspiration.jl:28-44
    C_can: float = 0.006
    ground_resistance_factor: Any = 1.0  # number / ConstantEvaporationResistanceFactor / SoilMoistureResistanceFactor()


@dataclass
class SurfaceHydrology:  # surface_hydrology.jl:11-34 ; defaults are filled by LandModel (land_model.jl:114-125)
    canopy_interception: Any = None
    evapotranspiration: Any = None
    surface_runoff: DirectSurfaceRunoff = field(default_factory=DirectSurfaceRunoff)


# ---- vegetation processes (src/processes/vegetation/*.jl), one PFT (needleleaf tree defaults) ----
@dataclass
class LUEPhotosynthesis:  # photosynthesis.jl:19-67
    tau25: float = 2600.0
    Kc25: float = 30.0
    Ko25: float = 3.0e4
    q10_tau: float = 0.57
    q10_Kc: float = 2.1
    q10_Ko: float = 1.2
    alpha_leaf: float = 0.17
    alpha_a: float = 0.5
    alpha_C3: float = 0.08
    cq: float = 4.6e-6
    k_ext: float = 0.5
    T_CO2_high: float = 42.0
    T_CO2_low: float = -4.0
    T_photos_high: float = 30.0
    T_photos_low: float = 15.0
    theta_r: float = 0.7


@dataclass
class MedlynStomatalConductance:  # stomatal_conductance.jl:18-25
    g1: float = 2.3
    g_min: float = 0.5


@dataclass
class PALADYNAutotrophicRespiration:  # autotrophic_respiration.jl:16-25
    cn_sapwood: float = 330.0
    cn_root: float = 29.0
    aws: float = 10.0


class PALADYNPhenology:  # phenology.jl:14-15
    pass


@dataclass
class PALADYNCarbonDynamics:  # carbon_dynamics.jl:18-39
    SLA: float = 10.0
    awl: float = 2.0
    LAI_min: float = 1.0
    LAI_max: float = 6.0
    gamma_L: float = 0.3
    gamma_R: float = 0.3
    gamma_S: float = 0.05


@dataclass
class PALADYNVegetationDynamics:  # vegetation_dynamics.jl:14-20
    nu_seed: float = 0.001
    gamma_v_min: float = 0.002


@dataclass
class StaticExponentialRootDistribution:  # root_distribution.jl:25-31
    a: float = 7.0
    b: float = 2.0


class FieldCapacityLimitedPAW:  # plant_available_water.jl:17
    pass


@dataclass
class VegetationCarbon:  # vegetation_carbon.jl:6-66
    photosynthesis: LUEPhotosynthesis = field(default_factory=LUEPhotosynthesis)
    stomatal_conductance: MedlynStomatalConductance = field(default_factory=MedlynStomatalConductance)
    autotrophic_respiration: PALADYNAutotrophicRespiration = field(default_factory=PALADYNAutotrophicRespiration)
    phenology: PALADYNPhenology = field(default_factory=PALADYNPhenology)
    carbon_dynamics: PALADYNCarbonDynamics = field(default_factory=PALADYNCarbonDynamics)
    vegetation_dynamics: PALADYNVegetationDynamics = field(default_factory=PALADYNVegetationDynamics)
    root_distribution: StaticExponentialRootDistribution = field(default_factory=StaticExponentialRootDistribution)
    plant_available_water: FieldCapacityLimitedPAW = field(default_factory=FieldCapacityLimitedPAW)


@dataclass
class PrescribedAtmosphere:  # prescribed_atmosphere.jl:44-87 (ConstantAerodynamics, aerodynamics.jl:6-9)
    altitude: float = 10.0
    min_windspeed: float = 0.01
    C_h: float = 1.2e-3


# ---------------------------------------------------------------------------------------------
# initializers (src/models/soil/soil_model_init.jl, src/initializers.jl)
# ---------------------------------------------------------------------------------------------
class DefaultInitializer:
    def fields(self, grid) -> Dict[str, Any]:
        return {}


@dataclass
class ConstantSoilTemperature:
    T0: float = 0.0

    def fields(self, grid):
        return {"temperature": self.T0}


@dataclass
class QuasiThermalSteadyState:
    T0: float = 0.0
    Qgeo: float = 0.02
    k_eff: float = 1.0

    def fields(self, grid):
        return {"temperature": lambda x, z: self.T0 - self.Qgeo / self.k_eff * z}


@dataclass
class ConstantSaturation:
    sat: float = 1.0

    def fields(self, grid):
        return {"saturation_water_ice": self.sat}


@dataclass
class SaturationWaterTable:
    vadose_zone_saturation: float = 0.5
    water_table_depth: float = 5.0

    def fields(self, grid):
        # as coded (soil_model_init.jl:150): z <= +depth is always true for z <= 0 -> fully saturated
        return {"saturation_water_ice": lambda x, z: np.where(z <= self.water_table_depth, 1.0, self.vadose_zone_saturation)}


@dataclass
class SoilInitializer:  # soil_model_init.jl:6-36 ; order hydrology, biogeochem, energy
    energy: Any = field(default_factory=QuasiThermalSteadyState)
    hydrology: Any = field(default_factory=SaturationWaterTable)

    def fields(self, grid):
        out = {}
        out.update(self.hydrology.fields(grid))
        out.update(self.energy.fields(grid))
        return out


# ---------------------------------------------------------------------------------------------
# time-dependent per-column values: what a boundary condition or an input may be made of
# ---------------------------------------------------------------------------------------------
@dataclass
class Sinusoid:
    """``clamp(mean + amp*sin(2*pi*t/period - phase), lo, hi)`` per column, evaluated on the device.

    The device-resident form of the function valued boundary condition of
    ``examples/simulations/soil_heat_global.jl:72-93``."""
    mean: Union[float, np.ndarray]
    amp: Union[float, np.ndarray]
    phase: Union[float, np.ndarray] = 0.0
    period: float = 86400.0
    lo: float = -np.inf
    hi: float = np.inf


@dataclass
class TimeSeries:
    """Snapshots ``values[nt, ncol]`` at ``times[nt]`` (FieldTimeSeriesInputSource, input_sources.jl:142-171)."""
    times: Sequence[float]
    values: np.ndarray


@dataclass
class RasterInputSource:
    """Time-varying (or static) raster on the ring grid of a ``ColumnRingGrid`` (``InputSource(grid, raster)``,
    ext/TerrariumRastersExt/TerrariumRastersExt.jl:22-56): ``values[nt, nring]`` at ``times[nt]`` seconds relative to
    ``reftime`` (or ``values[nring]`` without a time axis). The masked points are gathered to the columns
    (``idxmap = findall(mask)``, :44) and evaluated on the device with the extension's update rule (:96-121)."""
    values: np.ndarray
    times: Optional[Sequence[float]] = None
    reftime: float = 0.0

    @classmethod
    def from_netcdf(cls, path: str, variable: str, time: str = "time", reftime: Optional[float] = 0.0, decode_times: bool = False,
                    time_range: Optional[tuple] = None):
        """Read ``variable[time, ...]`` from a NetCDF file; the trailing dimensions are flattened in storage order to the
        ring-grid points (a full Gaussian / lon-lat raster stored ``[lat, lon]`` north to south is in RingGrids order).
        NetCDF-3 (classic / 64-bit offset) goes through ``scipy.io.netcdf_file``, NetCDF-4 / HDF5 through the decoder in
        ``netcdf4.py``. ``_FillValue`` / ``missing_value`` become NaN, ``scale_factor`` / ``add_offset`` are applied.
        ``decode_times`` converts the time axis to seconds with the CF ``units`` attribute (``"hours since ..."``);
        ``reftime = None`` takes the first time of the axis (``default_reftime``, TerrariumRastersExt.jl:130-133).
        ``time_range = (i0, i1)`` reads only the snapshots ``i0:i1`` (NetCDF-4: only the chunks they cover are decoded)."""
        from . import netcdf4
        if netcdf4.is_hdf5(path):
            with netcdf4.File(path) as f:
                var = f.variables[variable]
                has_time = bool(var.dimensions) and var.dimensions[0] == time and time in f.variables
                i0, i1 = time_range if (has_time and time_range is not None) else (0, None)
                data = var.scaled(np.float64, i0, i1)
                times = np.array(f.variables[time].read(i0, i1), dtype=np.float64) if has_time else None
                units = f.variables[time].attrs.get("units") if has_time else None
        else:
            from scipy.io import netcdf_file
            with netcdf_file(path, "r", mmap=False) as f:
                var = f.variables[variable]
                raw = np.array(var[:])
                data = raw.astype(np.float64)
                for key in ("_FillValue", "missing_value"):
                    if hasattr(var, key):
                        data[raw == np.asarray(getattr(var, key)).reshape(-1)[0]] = np.nan
                scale, offset = getattr(var, "scale_factor", 1.0), getattr(var, "add_offset", 0.0)
                data = data * scale + offset
                has_time = time in f.variables and var.dimensions and var.dimensions[0] == time
                times = np.array(f.variables[time][:], dtype=np.float64) if has_time else None
                if has_time and time_range is not None:
                    data, times = data[time_range[0]:time_range[1]], times[time_range[0]:time_range[1]]
                units = getattr(f.variables[time], "units", None) if has_time else None
                units = units.decode() if isinstance(units, bytes) else units
        if has_time and decode_times:
            times = times * cf_time_unit_seconds(units)
        if has_time and reftime is None:
            reftime = float(times[0])
        data = data.reshape(data.shape[0], -1) if has_time else data.reshape(-1)
        return cls(values=data, times=times, reftime=0.0 if reftime is None else reftime)


def raster_from_netcdf_files(paths: Sequence[str], variable: str, time: str = "time", reftime: Optional[float] = None) -> "RasterInputSource":
    """One time-varying raster from several files that continue each other in time (one file per month or year, as
    ERA5-Land is distributed): the time axes are decoded to seconds with each file's own CF units — they must share
    the epoch — concatenated and checked to be strictly increasing."""
    parts = [RasterInputSource.from_netcdf(p, variable, time=time, decode_times=True, reftime=0.0) for p in paths]
    if not parts or any(p.times is None for p in parts):
        raise ValueError("raster_from_netcdf_files needs files with a time axis")
    times = np.concatenate([np.asarray(p.times, dtype=np.float64) for p in parts])
    if np.any(np.diff(times) <= 0):
        raise ValueError("the time axes of the files do not continue each other (different epochs or overlapping files?)")
    values = np.concatenate([p.values for p in parts], axis=0)
    return RasterInputSource(values=values, times=times, reftime=float(times[0]) if reftime is None else reftime)


def cf_time_unit_seconds(units: Optional[str]) -> float:
    """Seconds per unit of a CF time axis (``"<unit> since <epoch>"``)."""
    unit = (units or "seconds").split()[0].lower()
    table = {"s": 1.0, "sec": 1.0, "secs": 1.0, "second": 1.0, "seconds": 1.0, "min": 60.0, "minute": 60.0, "minutes": 60.0,
             "h": 3600.0, "hr": 3600.0, "hour": 3600.0, "hours": 3600.0, "d": 86400.0, "day": 86400.0, "days": 86400.0}
    if unit not in table:
        raise ValueError(f"unsupported CF time unit {units!r}")
    return table[unit]


def InputSource(grid, data, name: Optional[str] = None, times=None, reftime: float = 0.0):
    """``InputSource(grid, field_or_raster; name = ...)`` (input_sources.jl:81-171, TerrariumRastersExt.jl:39-56): a named
    input for ``initialize(model, timestepper, inputs...)``. ``data`` is a number / per-column array / ``Sinusoid`` /
    ``TimeSeries`` (field sources), or an array on the ring grid of a ``ColumnRingGrid`` (optionally with ``times``)."""
    if name is None:
        raise ValueError("InputSource needs the name of the input variable it provides")
    if isinstance(data, (RasterInputSource, TimeSeries, Sinusoid)):   # e.g. RasterInputSource.from_netcdf(...)
        return NamedInput(name, data)
    ring = hasattr(grid, "mask") and isinstance(data, np.ndarray) and data.shape[-1] == grid.npoints and grid.npoints != grid.Nc
    value = RasterInputSource(values=data, times=times, reftime=reftime) if ring else (TimeSeries(times, data) if times is not None else data)
    return NamedInput(name, value)


@dataclass
class NamedInput:
    name: str
    value: Any


@dataclass
class BoundaryCondition:
    kind: int           # abi.TRM_BC_*
    slot: int           # abi.TRM_BC_<field>_<side>
    value: Any          # number | per-column array | Sinusoid | TimeSeries | callable (x, t)
    name: Optional[str] = None


def ValueBoundaryCondition(value):
    return ("value", value)


def FluxBoundaryCondition(value):
    return ("flux", value)


def GradientBoundaryCondition(value):
    return ("gradient", value)


# aliases of src/models/soil/soil_model_bcs.jl:6-40 -------------------------------------------
def PrescribedSurfaceTemperature(name: str, value=None) -> Dict[str, Dict[str, BoundaryCondition]]:
    return {"temperature": {"top": BoundaryCondition(abi.TRM_BC_VALUE, abi.TRM_BC_TEMPERATURE_TOP, value, name)}}


def PrescribedBottomTemperature(name: str, value=None):
    return {"temperature": {"bottom": BoundaryCondition(abi.TRM_BC_VALUE, abi.TRM_BC_TEMPERATURE_BOTTOM, value, name)}}


def GroundHeatFlux(value):
    return {"internal_energy": {"top": BoundaryCondition(abi.TRM_BC_FLUX, abi.TRM_BC_ENERGY_TOP, value, "ground_heat_flux")}}


def GeothermalHeatFlux(value):
    return {"internal_energy": {"bottom": BoundaryCondition(abi.TRM_BC_FLUX, abi.TRM_BC_ENERGY_BOTTOM, value, "geothermal_heat_flux")}}


def InfiltrationFlux(value):
    return {"saturation_water_ice": {"top": BoundaryCondition(abi.TRM_BC_FLUX, abi.TRM_BC_SATURATION_TOP, value, "infiltration")}}


def ImpermeableBoundary():
    return {"saturation_water_ice": {"bottom": BoundaryCondition(abi.TRM_BC_DEFAULT, abi.TRM_BC_SATURATION_BOTTOM, 0.0)}}


def FreeDrainage():
    return {"pressure_head": {"bottom": BoundaryCondition(abi.TRM_BC_GRADIENT, abi.TRM_BC_PRESSURE_BOTTOM, 0.0)}}


_SLOTS = {
    ("temperature", "top"): abi.TRM_BC_TEMPERATURE_TOP, ("temperature", "bottom"): abi.TRM_BC_TEMPERATURE_BOTTOM,
    ("internal_energy", "top"): abi.TRM_BC_ENERGY_TOP, ("internal_energy", "bottom"): abi.TRM_BC_ENERGY_BOTTOM,
    ("saturation_water_ice", "top"): abi.TRM_BC_SATURATION_TOP, ("saturation_water_ice", "bottom"): abi.TRM_BC_SATURATION_BOTTOM,
    ("pressure_head", "top"): abi.TRM_BC_PRESSURE_TOP, ("pressure_head", "bottom"): abi.TRM_BC_PRESSURE_BOTTOM,
}
_KINDS = {"value": abi.TRM_BC_VALUE, "flux": abi.TRM_BC_FLUX, "gradient": abi.TRM_BC_GRADIENT}


def merge_boundary_conditions(*bcs) -> Dict[str, Dict[str, BoundaryCondition]]:
    """Recursive merge, later arguments win (boundary_conditions.jl:18)."""
    out: Dict[str, Dict[str, BoundaryCondition]] = {}
    for b in bcs:
        for var, sides in (b or {}).items():
            for side, bc in sides.items():
                if isinstance(bc, tuple):  # (kind, value) from Value/Flux/GradientBoundaryCondition
                    bc = BoundaryCondition(_KINDS[bc[0]], _SLOTS[(var, side)], bc[1])
                out.setdefault(var, {})[side] = bc
    return out


# ---------------------------------------------------------------------------------------------
# models
# ---------------------------------------------------------------------------------------------
@dataclass
class SoilModel:  # src/models/soil/soil_model.jl:9-27
    grid: ColumnGrid
    soil: SoilEnergyWaterCarbon = field(default_factory=SoilEnergyWaterCarbon)
    constants: PhysicalConstants = field(default_factory=PhysicalConstants)
    initializer: Any = field(default_factory=DefaultInitializer)
    # unpinned Oceananigans semantics: what the never filled z-halo of the
    # auxiliary saturation field holds under NoFlow hydrology. "zero" | "copy"
    sat_halo: str = "zero"


_DEFAULT_VEGETATION = object()


@dataclass
class LandModel:  # src/models/coupled/land_model.jl:10-44
    grid: ColumnGrid
    # reference default: VegetationCarbon(NF) (land_model.jl:26) ; vegetation=None is the bare-ground LandModel
    vegetation: Any = _DEFAULT_VEGETATION
    # default_soil: immobile soil water without vegetation, RichardsEq with (land_model.jl:111-112)
    soil: Any = None
    surface_energy_balance: SurfaceEnergyBalance = field(default_factory=SurfaceEnergyBalance)
    # default_surface_hydrology (land_model.jl:118-125)
    surface_hydrology: Any = None
    atmosphere: PrescribedAtmosphere = field(default_factory=PrescribedAtmosphere)
    constants: PhysicalConstants = field(default_factory=PhysicalConstants)
    initializer: Any = field(default_factory=DefaultInitializer)
    sat_halo: str = "zero"

    def __post_init__(self):
        if self.vegetation is _DEFAULT_VEGETATION:
            self.vegetation = VegetationCarbon()
        veg = self.vegetation is not None
        if veg and not isinstance(self.vegetation, VegetationCarbon):
            raise TypeError(f"unsupported vegetation {type(self.vegetation).__name__}")
        if self.soil is None:
            self.soil = SoilEnergyWaterCarbon(hydrology=SoilHydrology(vertical_flow=RichardsEq())) if veg else SoilEnergyWaterCarbon()
        if self.surface_hydrology is None:
            self.surface_hydrology = SurfaceHydrology()
        sh = self.surface_hydrology
        if sh.canopy_interception is None:
            sh.canopy_interception = PALADYNCanopyInterception() if veg else NoCanopyInterception()
        if sh.evapotranspiration is None:
            sh.evapotranspiration = PALADYNCanopyEvapotranspiration() if veg else BareGroundEvaporation()
        canopy = isinstance(sh.canopy_interception, PALADYNCanopyInterception), isinstance(sh.evapotranspiration, PALADYNCanopyEvapotranspiration)
        if veg != canopy[0] or veg != canopy[1]:
            raise NotImplementedError("built combinations: vegetation=None with BareGroundEvaporation + NoCanopyInterception, "
                                      "VegetationCarbon with PALADYNCanopyInterception + PALADYNCanopyEvapotranspiration")


# ---------------------------------------------------------------------------------------------
# time steppers (src/timesteppers/forward_euler.jl:6-17, heun.jl:6-20)
# ---------------------------------------------------------------------------------------------
@dataclass
class ForwardEuler:
    dt: float = 300.0


@dataclass
class Heun:
    dt: float = 300.0


def default_dt(ts) -> float:
    return ts.dt


def is_adaptive(ts) -> bool:
    return False


# ---------------------------------------------------------------------------------------------
# model -> trm_config
# ---------------------------------------------------------------------------------------------
def build_params(model) -> abi.trm_params:
    p = abi.trm_params()
    soil, c = model.soil, model.constants
    p.mineral_porosity = soil.strat.porosity.mineral_porosity
    p.organic_porosity = soil.strat.porosity.organic_porosity
    p.rho_soc, p.rho_org = soil.biogeochem.rho_soc, soil.biogeochem.rho_org
    k, h = soil.energy.thermal_properties.conductivities, soil.energy.thermal_properties.heat_capacities
    for i, n in enumerate(("water", "ice", "air", "mineral", "organic")):
        p.kappa[i] = getattr(k, n)
        p.heatcap[i] = getattr(h, n)
    p.rho_w, p.Lsl, p.Llg, p.rho_a, p.c_a = c.rho_w, c.Lsl, c.Llg, c.rho_a, c.c_a
    p.Tref, p.sigma, p.eps_mw = c.Tref, c.sigma, c.eps
    hp = soil.hydrology.hydraulic_properties
    p.K_sat = hp.sat_hydraulic_cond
    vg, bc = VanGenuchten(), BrooksCorey()
    if isinstance(hp.swrc, VanGenuchten):
        vg = hp.swrc
    elif isinstance(hp.swrc, BrooksCorey):
        bc = hp.swrc
    else:
        raise TypeError(f"unsupported SWRC {type(hp.swrc).__name__}")
    p.vg_alpha, p.vg_n, p.bc_psis, p.bc_lambda = vg.alpha, vg.n, bc.psi_s, bc.lam
    p.theta_res = hp.swrc.theta_res
    p.impedance = getattr(hp.unsat_hydraulic_cond, "impedance", 7.0)
    p.vwc_forcing = soil.hydrology.vwc_forcing or 0.0
    # surface defaults; overwritten for LandModel
    p.albedo, p.emissivity, p.kappa_skin, p.C_h, p.min_windspeed, p.tau_r, p.evap_beta = 0.3, 0.97, 2.0, 1.2e-3, 0.01, 3600.0, 1.0
    if isinstance(model, LandModel):
        seb, sh, atm = model.surface_energy_balance, model.surface_hydrology, model.atmosphere
        if isinstance(seb.albedo, ConstantAlbedo):   # (PrescribedAlbedo: per-column inputs instead)
            p.albedo, p.emissivity = seb.albedo.albedo, seb.albedo.emissivity
        p.kappa_skin = seb.skin_temperature.kappa_s
        p.C_h, p.min_windspeed = atm.C_h, atm.min_windspeed
        p.tau_r = sh.surface_runoff.tau_r
        gr = sh.evapotranspiration.ground_resistance_factor
        p.evap_beta = 1.0 if isinstance(gr, SoilMoistureResistanceFactor) else float(getattr(gr, "factor", gr))
    # vegetated LandModel
    fc, wp = hp.field_capacity, hp.wilting_point
    if isinstance(hp, SoilHydraulicsSURFEX):
        fc, wp = hp.capacities(soil.strat.texture, np.dtype(model.grid.nf).type)
    d = dict(field_capacity=fc, wilting_point=wp, C_mass=c.C_mass)
    veg = getattr(model, "vegetation", None) or VegetationCarbon()
    ph, sc, ar, cd, vd, rd = (veg.photosynthesis, veg.stomatal_conductance, veg.autotrophic_respiration, veg.carbon_dynamics,
                              veg.vegetation_dynamics, veg.root_distribution)
    for n in ("tau25", "Kc25", "Ko25", "q10_tau", "q10_Kc", "q10_Ko", "alpha_leaf", "alpha_a", "alpha_C3", "cq", "k_ext",
              "T_CO2_high", "T_CO2_low", "T_photos_high", "T_photos_low", "theta_r"):
        d[n] = getattr(ph, n)
    d.update(g1=sc.g1, g_min=sc.g_min, cn_sapwood=ar.cn_sapwood, cn_root=ar.cn_root, aws=ar.aws,
             SLA=cd.SLA, awl=cd.awl, LAI_min=cd.LAI_min, LAI_max=cd.LAI_max, gamma_L=cd.gamma_L, gamma_R=cd.gamma_R, gamma_S=cd.gamma_S,
             nu_seed=vd.nu_seed, gamma_v_min=vd.gamma_v_min, root_a=rd.a, root_b=rd.b)
    ci, et = PALADYNCanopyInterception(), PALADYNCanopyEvapotranspiration()
    if isinstance(model, LandModel) and model.vegetation is not None:
        ci, et = model.surface_hydrology.canopy_interception, model.surface_hydrology.evapotranspiration
    d.update(alpha_int=ci.alpha_int, k_ext_can=ci.k_ext, w_can_max=ci.w_can_max, tau_w=ci.tau_w, C_can=et.C_can)
    for n in abi.VEGETATION_PARAMS:
        setattr(p, n, float(d[n]))
    return p


def build_config(model, timestepper, ncol: int, col0: int = 0, device: int = 0, math: str = "faithful"):
    """Translate a model + timestepper into the POD ``trm_config`` (BC slots are filled by the caller).

    Returns ``(config, z_faces_buffer)``; the buffer must stay alive until ``trm_create`` returned."""
    grid = model.grid
    cfg = abi.trm_config()
    cfg.abi_version = abi.TRM_ABI_VERSION
    cfg.dtype = abi.dtype_code(grid.nf)
    cfg.ncol, cfg.col0, cfg.nz, cfg.device = int(ncol), int(col0), int(grid.Nz), int(device)
    cfg.model = abi.TRM_MODEL_LAND if isinstance(model, LandModel) else abi.TRM_MODEL_SOIL
    if isinstance(timestepper, Heun):
        cfg.timestepper = abi.TRM_HEUN
    elif isinstance(timestepper, ForwardEuler):
        cfg.timestepper = abi.TRM_EULER
    else:
        raise TypeError(f"unsupported timestepper {type(timestepper).__name__}")
    hyd = model.soil.hydrology
    cfg.hydrology = abi.TRM_RICHARDS if isinstance(hyd.vertical_flow, RichardsEq) else abi.TRM_NOFLOW
    hp = hyd.hydraulic_properties
    cfg.swrc = abi.TRM_SWRC_VANGENUCHTEN if isinstance(hp.swrc, VanGenuchten) else abi.TRM_SWRC_BROOKSCOREY
    if isinstance(hp.unsat_hydraulic_cond, UnsatKVanGenuchten):
        if not isinstance(hp.swrc, VanGenuchten):
            raise TypeError("UnsatKVanGenuchten requires a VanGenuchten SWRC (soil_hydraulic_properties.jl:203-206)")
        cfg.unsat_k = abi.TRM_UNSATK_VANGENUCHTEN
    else:
        cfg.unsat_k = abi.TRM_UNSATK_LINEAR
    cfg.sat_halo = abi.TRM_HALO_COPY if model.sat_halo == "copy" else abi.TRM_HALO_ZERO
    cfg.skin = abi.TRM_SKIN_IMPLICIT
    if isinstance(model, LandModel) and isinstance(model.surface_energy_balance.skin_temperature, PrescribedSkinTemperature):
        cfg.skin = abi.TRM_SKIN_PRESCRIBED
    cfg.math = abi.TRM_MATH_FAST if math == "fast" else abi.TRM_MATH_FAITHFUL
    if isinstance(model, LandModel) and isinstance(model.surface_hydrology.evapotranspiration.ground_resistance_factor, SoilMoistureResistanceFactor):
        cfg.ground_resistance = abi.TRM_GROUND_RES_SOIL_MOISTURE
    cfg.vegetation = abi.TRM_VEG_CARBON if isinstance(model, LandModel) and model.vegetation is not None else abi.TRM_VEG_NONE
    if isinstance(model, LandModel):
        seb = model.surface_energy_balance
        cfg.albedo_kind = abi.TRM_ALBEDO_PRESCRIBED if isinstance(seb.albedo, PrescribedAlbedo) else abi.TRM_ALBEDO_CONSTANT
        cfg.radiative = abi.TRM_RADIATIVE_PRESCRIBED if isinstance(seb.radiative_fluxes, PrescribedRadiativeFluxes) else abi.TRM_RADIATIVE_DIAGNOSED
        cfg.turbulent = abi.TRM_TURBULENT_PRESCRIBED if isinstance(seb.turbulent_fluxes, PrescribedTurbulentFluxes) else abi.TRM_TURBULENT_DIAGNOSED
    zbuf = np.ascontiguousarray(grid.z_faces, dtype=np.float64)
    import ctypes as C
    cfg.z_faces = zbuf.ctypes.data_as(C.POINTER(C.c_double))
    cfg.params = build_params(model)
    return cfg, zbuf
