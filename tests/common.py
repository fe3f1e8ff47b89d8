"""Shared helpers for the test-suite: one scenario, two engines.

``engine == "oracle"`` drives the CPU restatement (test infrastructure), ``engine == "cuda"`` drives the
product library through the same C ABI.  CUDA parametrisations carry the ``gpu`` marker.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import terrarium_jl_b200 as trm  # noqa: E402

ENGINES = [pytest.param("oracle", id="oracle"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
CUDA_MATH = [pytest.param("faithful", id="faithful"), pytest.param("fast", id="fast")]


def make(engine, model, ts, inputs=None, **kw):
    if engine == "oracle":
        import oracle_integrator as oi
        kw.pop("math", None)
        return oi.oracle_initialize(model, ts, inputs, **kw)
    return trm.initialize(model, ts, inputs, **kw)


def richards_soil(alpha=2.0, n=2.0, K_sat=1.0e-5, unsat="vg", vwc_forcing=None, **soil_kw):
    """ConstantSoilHydraulics(swrc=VanGenuchten(alpha, n), UnsatKVanGenuchten) + RichardsEq
    (test/soil/soil_hydrology_tests.jl:127-130, test/coupled_models/land_model_tests.jl:8-11)."""
    hp = trm.ConstantSoilHydraulics(
        swrc=trm.VanGenuchten(alpha=alpha, n=n),
        unsat_hydraulic_cond=trm.UnsatKVanGenuchten() if unsat == "vg" else trm.UnsatKLinear(),
        sat_hydraulic_cond=K_sat)
    hyd = trm.SoilHydrology(trm.RichardsEq(), hydraulic_properties=hp, vwc_forcing=vwc_forcing)
    return trm.SoilEnergyWaterCarbon(hydrology=hyd, **soil_kw)


def synthetic_columns(ncol, seed=20260101):
    """Per-column parameters of the synthetic benchmark."""
    rng = np.random.default_rng(seed)
    lat = rng.uniform(-np.pi / 2, np.pi / 2, ncol)
    lon = rng.uniform(0.0, 2 * np.pi, ncol)
    T0 = 20.0 - np.abs(40.0 * np.sin(lat))
    return lat, lon, T0


def synthetic_soil_case(engine, ncol, nf=np.float64, richards=True, heun=False, nz=30, math="faithful", dt=None):
    """BASELINE config 3/5: coupled soil energy + Richards with the sinusoidal surface temperature."""
    lat, lon, T0 = synthetic_columns(ncol)
    grid = trm.ColumnGrid(trm.B200(), nf, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=nz), ncol)
    soil = richards_soil() if richards else trm.SoilEnergyWaterCarbon()
    model = trm.SoilModel(grid, soil=soil)
    bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0))
    inits = {
        "temperature": lambda x, z: T0[None, :] - 0.05 * z,
        "saturation_water_ice": (lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x) if richards else 1.0,
    }
    ts = (trm.Heun if heun else trm.ForwardEuler)(dt=dt or (60.0 if richards else 300.0))
    return make(engine, model, ts, boundary_conditions=bcs, initializers=inits, math=math)


def synthetic_land_case(engine, ncol, nf=np.float64, heun=False, nz=30, math="faithful", dt=60.0, windspeed=3.0, richards=True):
    """BASELINE config 4 (bare ground): LandModel with the synthetic atmosphere of BASELINE.md section 5."""
    lat, lon, T0 = synthetic_columns(ncol)
    grid = trm.ColumnGrid(trm.B200(), nf, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=nz), ncol)
    # richards=False: the reference's default soil of LandModel(grid; vegetation = nothing), immobile soil water
    model = trm.LandModel(grid, soil=richards_soil(), vegetation=None) if richards else trm.LandModel(grid, vegetation=None)
    day = 86400.0
    # rain: 2e-8 m/s during the first 6 h of each day, as an hourly table over 3 days (flat outside)
    hours = np.arange(0, 73, dtype=np.float64)
    rain = np.where((hours % 24) < 6, 2.0e-8, 0.0)
    inputs = {
        "air_temperature": trm.Sinusoid(mean=T0, amp=8.0, phase=lon, period=day),
        "surface_shortwave_down": trm.Sinusoid(mean=0.0, amp=600.0, phase=lon, period=day, lo=0.0),
        "surface_longwave_down": 300.0,
        "specific_humidity": 0.005,
        "air_pressure": 101325.0,
        "windspeed": windspeed,
        "rainfall": trm.TimeSeries(hours * 3600.0, np.repeat(rain[:, None], ncol, axis=1)),
    }
    inits = {
        "temperature": lambda x, z: T0[None, :] - 0.05 * z,
        "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x,
        "skin_temperature": T0,
    }
    ts = (trm.Heun if heun else trm.ForwardEuler)(dt=dt)
    return make(engine, model, ts, inputs, initializers=inits, math=math)


def relerr(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else np.finfo(np.float64).tiny)))


def max_scaled_err(a, b):
    """max |a-b| / max|b| : relative error against the field's scale (robust where the field crosses 0)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), np.finfo(np.float64).tiny))


def pointwise_relerr(a, b, floor_frac=1.0e-6):
    """max |a-b| / max(|b|, floor_frac * max|b|): pointwise relative error with an explicit floor, so that a cell whose
    value is small against the field's scale is still held to (1 / floor_frac) x the scaled tolerance."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), np.finfo(np.float64).tiny)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor_frac * scale)))
