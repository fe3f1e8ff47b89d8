"""BASELINE configs 2 and 3 on the reference's own grids: ``ColumnRingGrid`` over the ERA5-Land N72 / N145 land masks
(``tests/golden/era5_land_masks.npz``, frozen from ``/root/reference/inputs`` by ``tests/golden/make_land_masks.py``),
set up as ``examples/simulations/soil_heat_global.jl`` does (latitude climatology, 0.05 K/m profile, daily surface
temperature cycle shifted by longitude)."""
import datetime
import os

import numpy as np
import pytest

from common import ENGINES, make, max_scaled_err, richards_soil, trm

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "era5_land_masks.npz")
COLUMNS = {"N72": 14017, "N145": 56951}


def era5_land_grid(name, nf, spacing):
    d = np.load(FIXTURE)
    lat, lon = d[f"{name}_lat"], d[f"{name}_lon"]
    mask = np.unpackbits(d[f"{name}_bits"])[:lat.size * lon.size].astype(bool)
    # ring order of a full grid: rings north to south, longitude fastest; radians as RingGrids.get_lonlats (:39)
    return trm.ColumnRingGrid(trm.B200(), nf, spacing, mask, lon=np.deg2rad(np.tile(lon, lat.size)), lat=np.deg2rad(np.repeat(lat, lon.size)))


def soil_heat_global(engine, name, nf, soil=None, stepper=trm.ForwardEuler, math="faithful", saturation=None):
    grid = era5_land_grid(name, nf, trm.ExponentialSpacing(N=30))
    lon, lat = grid.masked_lonlat()
    T0 = 20.0 - np.abs(40.0 * np.sin(lat))                        # mean_annual_temperature, soil_heat_global.jl:51
    bc = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0))   # :72-93
    inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z}  # :59-64
    if saturation is not None:
        inits["saturation_water_ice"] = saturation
    model = trm.SoilModel(grid, soil=soil) if soil is not None else trm.SoilModel(grid)
    return grid, T0, make(engine, model, stepper(), boundary_conditions=bc, initializers=inits, math=math)


def test_mask_fixture_geometry():
    for name, ncol in COLUMNS.items():
        grid = era5_land_grid(name, np.float32, trm.ExponentialSpacing(N=30))
        n = int(name[1:])
        assert grid.Nc == ncol and grid.npoints == (2 * n) * (4 * n) and grid.Nz == 30
        lon, lat = grid.masked_lonlat()
        assert np.all(np.diff(lat) <= 0) and lat[0] > 1.4 and lat[-1] < -1.1     # Greenland ... Antarctica
        # Antarctica: every point of the southernmost ring is land; the northernmost ring is ocean
        assert grid.mask[-4 * n:].all() and not grid.mask[:4 * n].any()


@pytest.mark.parametrize("engine", ENGINES)
def test_soil_heat_global_n72(engine):
    """Config 2 as the example runs it: Float32, two default time steps, then 12 h at 600 s (:108-113)."""
    grid, T0, integ = soil_heat_global(engine, "N72", np.float32)
    assert grid.Nc == COLUMNS["N72"]
    Tg0 = integ.state.ground_temperature.numpy().copy()
    assert np.allclose(Tg0.reshape(-1), T0 + 0.05 * 0.025, atol=1e-4)    # top cell centre at z = -0.025 m
    trm.timestep(integ)
    trm.timestep(integ)
    trm.run(integ, period=datetime.timedelta(hours=12), dt=600.0)
    assert integ.clock.time == 2 * 300.0 + 43200.0 and integ.clock.iteration == 74
    T = integ.state.temperature.numpy()
    Tg = integ.state.ground_temperature.numpy().reshape(-1)
    assert T.dtype == np.float32 and np.isfinite(T).all()
    assert Tg.min() > -30.5 and Tg.max() < 30.5 and np.abs(Tg - Tg0.reshape(-1)).max() > 1.0
    assert np.array_equal(T[-1].reshape(-1), Tg)                           # ground_temperature is the top layer
    deep = T[0].reshape(-1)                                               # the bottom cell (z ~ -180 m) has not moved
    assert np.allclose(deep, (T0 - 0.05 * grid.znodes_center()[0]).astype(np.float32), atol=1e-3)
    # surface map on the full ring grid (RingGrids.Field(..., grid), :104, :116): ocean points stay NaN
    ring = grid.to_ring(Tg)
    assert ring.shape == (41472,) and np.isnan(ring).sum() == 41472 - grid.Nc and np.array_equal(grid.from_ring(ring), Tg)


@pytest.mark.gpu
@pytest.mark.parametrize("math", ["faithful", "fast"])
def test_soil_heat_global_n72_parity(math):
    """Config 2, Float64: library against the oracle after the example's 74 steps (north_star tolerance 1e-9)."""
    runs = []
    for engine in ("oracle", "cuda"):
        _, _, integ = soil_heat_global(engine, "N72", np.float64, math=math)
        trm.timestep(integ)
        trm.timestep(integ)
        trm.run(integ, period=43200.0, dt=600.0)
        runs.append(integ)
    o, c = runs
    for name in ("temperature", "internal_energy", "liquid_water_fraction"):
        assert max_scaled_err(getattr(c.state, name).numpy(), getattr(o.state, name).numpy()) <= 1e-9, name
    assert np.array_equal(c.state.saturation_water_ice.numpy(), o.state.saturation_water_ice.numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("stepper", [trm.ForwardEuler, trm.Heun], ids=["euler", "heun"])
def test_soil_energy_richards_n145_parity(stepper):
    """Config 3: coupled soil energy + Richards hydrology on the N145 land columns, Float64, 60 steps of 60 s."""
    runs = []
    for engine in ("oracle", "cuda"):
        grid, _, integ = soil_heat_global(engine, "N145", np.float64, soil=richards_soil(), stepper=stepper,
                                          saturation=lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x)
        assert grid.Nc == COLUMNS["N145"]
        integ.step(60.0, 60)
        integ.compute_auxiliary()
        runs.append(integ)
    o, c = runs
    for name in ("temperature", "internal_energy", "saturation_water_ice", "pressure_head", "water_table"):
        assert max_scaled_err(getattr(c.state, name).numpy(), getattr(o.state, name).numpy()) <= 1e-9, name
    # column water is conserved (no-flux boundaries): thickness-weighted saturation, test/soil/soil_hydrology_tests.jl:171-188
    dz = grid.dz()[:, None]
    sat = c.state.saturation_water_ice.numpy().reshape(grid.Nz, -1)
    want = (np.minimum(1.0, 0.5 - 0.1 * grid.znodes_center())[:, None] * dz).sum(axis=0)
    assert np.allclose((sat * dz).sum(axis=0) + c.state.surface_excess_water.numpy().reshape(-1) / 0.49, want, rtol=1e-12)


@pytest.mark.parametrize("engine", ENGINES)
def test_soil_heat_global_era5_flow(engine, tmp_path):
    """``examples/simulations/soil_heat_global_era5.jl:13-48`` end to end: land mask and a 2 m temperature raster read from
    NetCDF-4 files, ``InputSource(grid, raster; name = :Tair)`` passed to ``initialize``, ``PrescribedSurfaceTemperature(:Tair)``
    without a value, the first raster slice as initial surface temperature. (The ERA5 raster is not part of the reference
    checkout: a packed, deflated stand-in on the N72 grid is written here.)"""
    from hdf5_writer import Writer
    d = np.load(FIXTURE)
    lat, lon = d["N72_lat"], d["N72_lon"]
    land = np.unpackbits(d["N72_bits"])[:lat.size * lon.size].astype(bool).reshape(lat.size, lon.size)
    hours = np.arange(0, 25, 6)
    truth = (288.0 - 30.0 * np.abs(np.sin(np.deg2rad(lat)))[None, :, None]
             + 8.0 * np.sin(2 * np.pi * hours[:, None, None] / 24.0 - np.deg2rad(lon)[None, None, :]))
    scale, offset = 80.0 / 65000.0, 270.0
    packed = np.round((truth - offset) / scale).astype("<i2")
    packed[:, ~land] = -32767                                    # ERA5-Land: missing over the ocean
    mask_path, t2m_path = str(tmp_path / "lsm.nc"), str(tmp_path / "t2m.nc")
    w = Writer(2)
    w.add("time", np.array([998520], dtype="<i4"), layout="contiguous", attrs={"units": "hours since 1900-01-01 00:00:0.0"})
    w.add("lat", lat, layout="contiguous")
    w.add("lon", lon, layout="contiguous")
    w.add("lsm", land[None].astype("<f8"), chunks=(1, 144, 288), dims=("time", "lat", "lon"), attrs={"_FillValue": np.float64(-32767.0)})
    w.save(mask_path)
    w = Writer(1)
    w.add("valid_time", (hours * 3600 + 1672531200).astype("<i8"), layout="contiguous", attrs={"units": "seconds since 1970-01-01"})
    w.add("latitude", lat, layout="contiguous")
    w.add("longitude", lon, layout="contiguous")
    w.add("t2m", packed, chunks=(1, 72, 144), shuffle=True, deflate=5, dims=("valid_time", "latitude", "longitude"),
          attrs={"scale_factor": np.float64(scale), "add_offset": np.float64(offset), "_FillValue": np.int16(-32767), "units": "K"})
    w.save(t2m_path)

    grid = trm.ColumnRingGrid.from_land_sea_mask(np.float64, trm.ExponentialSpacing(N=30), path=mask_path)
    assert grid.Nc == COLUMNS["N72"]
    raster = trm.RasterInputSource.from_netcdf(t2m_path, "t2m", time="valid_time", decode_times=True, reftime=None)
    assert raster.values.shape == (5, 41472) and np.isnan(raster.values[:, ~grid.mask]).all() and np.isfinite(raster.values[:, grid.mask]).all()
    raster.values = raster.values - 273.15
    forcing = trm.InputSource(grid, raster, name="Tair")
    Tsurf0 = raster.values[0][grid.mask]
    inits = {"temperature": lambda x, z: Tsurf0[None, :] - 0.02 * z, "saturation_water_ice": 1.0}
    integ = make(engine, trm.SoilModel(grid), trm.ForwardEuler(), forcing, initializers=inits,
                 boundary_conditions=trm.PrescribedSurfaceTemperature("Tair"))
    trm.timestep(integ, 120.0)
    trm.run(integ, period=3 * 3600.0, dt=120.0)    # last update_inputs! at t = 3 h: halfway between two slices
    want = (0.5 * (packed[0].astype(np.float64) + packed[1]) * scale + offset - 273.15)[land]
    np.testing.assert_allclose(integ.state.Tair.numpy().reshape(-1), want, rtol=1e-12, atol=1e-12)
    T = integ.state.temperature.numpy()
    assert np.isfinite(T).all() and np.abs(T[-1].reshape(-1) - Tsurf0).max() > 0.5


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config 1: the quick-start column (README.md:85-95) and its freeze-thaw variant
# (examples/simulations/soil_heat_column.jl:13-43), Float32 as written, Float64 for the parity leg
# ---------------------------------------------------------------------------------------------------------------
def quick_start(engine, nf, freeze_thaw, math="faithful"):
    grid = trm.ColumnGrid(trm.B200(), nf, trm.ExponentialSpacing(N=10), 1)
    if freeze_thaw:
        init = trm.SoilInitializer(energy=trm.QuasiThermalSteadyState(T0=-1.0), hydrology=trm.ConstantSaturation(sat=1.0))
        model = trm.SoilModel(grid, initializer=init)
    else:
        model = trm.SoilModel(grid)       # DefaultInitializer: T = 0 and the saturation auxiliary left at zero (dry conduction)
    integ = make(engine, model, trm.ForwardEuler(), boundary_conditions=trm.PrescribedSurfaceTemperature("T_ub", 1.0), math=math)
    if freeze_thaw:
        trm.timestep(integ)
        trm.timestep(integ)
        trm.run(integ, period=datetime.timedelta(days=3))
    else:
        trm.run(integ, period=datetime.timedelta(days=10))
    return grid, integ


@pytest.mark.parametrize("engine", ENGINES)
def test_quick_start_column(engine):
    grid, integ = quick_start(engine, np.float32, freeze_thaw=False)
    assert integ.clock.time == 864000.0 and integ.clock.iteration == 2880          # default dt = 300 s (forward_euler.jl:8)
    T = integ.state.temperature.numpy().reshape(-1)
    assert T.dtype == np.float32 and T.shape == (10,)
    assert np.all(integ.state.saturation_water_ice.numpy() == 0) and np.all(integ.state.liquid_water_fraction.numpy() == 1)
    assert np.all(np.diff(T) > 0) and 0.0 < T[0] and 0.97 < T[-1] < 1.0            # warming from the top, bottom cell first in memory
    assert np.all(integ.state.internal_energy.numpy() > 0)
    assert np.allclose(integ.state.T_ub.numpy(), 1.0)


@pytest.mark.parametrize("engine", ENGINES)
def test_freeze_thaw_column(engine):
    grid, integ = quick_start(engine, np.float32, freeze_thaw=True)
    assert integ.clock.time == 600.0 + 3 * 86400.0
    T = integ.state.temperature.numpy().reshape(-1)
    liq = integ.state.liquid_water_fraction.numpy().reshape(-1)
    U = integ.state.internal_energy.numpy().reshape(-1)
    # T0 - Qgeo / k * z with z < 0: the deepest cells start above 0 degC (thawed), the middle of the column is frozen,
    # and three days of +1 degC at the surface thaw the top cell and move the front into the second
    assert liq[-1] == 1.0 and T[-1] > 0 and liq[0] == 1.0 and T[0] > 0
    frozen = liq == 0
    assert frozen.any() and np.all(T[frozen] < 0)
    front = (liq > 0) & (liq < 1)
    assert front.any() and np.all(T[front] == 0.0)                                 # a cell holding a thaw front sits at 0 degC
    Lvol = 1000.0 * 3.34e5 * 0.49                                                  # rho_w * Lsl * porosity (saturated)
    assert np.all(U[liq == 0] < -Lvol * 0.999) and np.all(U[liq == 1] >= 0)


@pytest.mark.gpu
@pytest.mark.parametrize("freeze_thaw", [False, True], ids=["readme", "freeze-thaw"])
@pytest.mark.parametrize("math", ["faithful", "fast"])
def test_quick_start_parity(freeze_thaw, math):
    (_, o), (_, c) = quick_start("oracle", np.float64, freeze_thaw), quick_start("cuda", np.float64, freeze_thaw, math=math)
    for name in ("temperature", "internal_energy", "liquid_water_fraction"):
        assert max_scaled_err(getattr(c.state, name).numpy(), getattr(o.state, name).numpy()) <= 1e-9, name
    (_, o), (_, c) = quick_start("oracle", np.float32, freeze_thaw), quick_start("cuda", np.float32, freeze_thaw, math=math)
    assert max_scaled_err(c.state.temperature.numpy(), o.state.temperature.numpy()) <= 2e-4    # Float32 as the README runs it


@pytest.mark.parametrize("engine", ENGINES)
def test_speedy_dry_land_coupling_flow(engine):
    """``examples/simulations/speedy_dry_land.jl:45-86`` from the land side: a SoilModel on a full ring grid whose surface
    temperature is the input variable ``air_temperature`` (``PrescribedSurfaceTemperature(:air_temperature)`` without a value,
    ``InputSource(grid, field)``), overwritten by the atmosphere before every coupling interval
    (``set!(state.inputs.air_temperature, Tair)``), the land advanced with ``run!(period = dt_atm, dt = 300)`` and the top
    layer handed back; checked against the same integrator stepped by hand with ``timestep!``."""
    nring = 48 * 96                                              # FullGaussianGrid(24)
    grid = trm.ColumnRingGrid(trm.B200(), np.float32, trm.ExponentialSpacing(N=30, dz_min=0.05), np.ones(nring, dtype=bool))
    assert grid.Nc == nring
    rng = np.random.default_rng(5)

    def build():
        model = trm.SoilModel(grid, initializer=trm.SoilInitializer())
        forcing = trm.InputSource(grid, np.zeros(nring, dtype=np.float32), name="air_temperature")
        return make(engine, model, trm.ForwardEuler(), forcing, boundary_conditions=trm.PrescribedSurfaceTemperature("air_temperature"))

    coupled, manual = build(), build()
    Tsoil = coupled.state.temperature.numpy()[-1].reshape(-1) + np.float32(273.15)     # Speedy.initialize!, :39-41
    assert Tsoil.shape == (nring,) and np.all(Tsoil > 273.0)
    dt_atm = 1800.0
    for k in range(4):
        Tair_kelvin = (285.0 + 5.0 * rng.standard_normal(nring)).astype(np.float32)
        coupled.state.inputs.air_temperature.set(Tair_kelvin - np.float32(273.15))     # :56-57
        trm.run(coupled, period=dt_atm, dt=300.0)                                       # :60
        Tsurf = coupled.state.temperature.numpy()[-1].reshape(-1)                       # :65
        manual.state.air_temperature.set(Tair_kelvin - np.float32(273.15))
        for _ in range(6):
            trm.timestep(manual, 300.0)
        np.testing.assert_allclose(Tsurf, manual.state.ground_temperature.numpy().reshape(-1), rtol=0, atol=1e-5)
        assert np.array_equal(coupled.state.inputs.air_temperature.numpy(), Tair_kelvin - np.float32(273.15))
    assert coupled.clock.time == 4 * dt_atm and np.isfinite(coupled.state.temperature.numpy()).all()
    assert np.abs(coupled.state.temperature.numpy()[-1].reshape(-1) + 273.15 - Tsoil).max() > 0.5


@pytest.mark.parametrize("engine", ENGINES)
def test_soil_heat_global_simulation_output_on_the_ring_grid(engine, tmp_path):
    """``soil_heat_global.jl:117-123``: the integrator wrapped in a ``Simulation`` (dt = 600 s) with a scheduled writer; the
    saved surface temperature is put back on the full Gaussian grid (``RingGrids.Field(..., grid)``, :113)."""
    grid, T0, integ = soil_heat_global(engine, "N72", np.float32)
    sim = trm.Simulation(integ, dt=600.0, stop_time=datetime.timedelta(hours=6))
    path = str(tmp_path / "global.nc")
    sim.output_writers["surface"] = trm.NetCDFWriter(integ, ["temperature", "ground_temperature"], filename=path, schedule=trm.TimeInterval(7200.0))
    sim.run(); sim.close()
    Tg = trm.FieldTimeSeries(path, "ground_temperature")
    assert list(Tg.times) == [0.0, 7200.0, 14400.0, 21600.0] and Tg.ring_points == 41472
    assert np.array_equal(Tg[-1].reshape(-1), integ.state.ground_temperature.numpy().reshape(-1))
    ring = Tg.ring(-1)
    assert ring.shape[-1] == 41472 and np.isnan(ring).sum() == 41472 - grid.Nc
    assert np.array_equal(ring.reshape(-1)[grid.mask], Tg[-1].reshape(-1)) and np.array_equal(ring.reshape(-1), grid.to_ring(Tg[-1].reshape(-1)), equal_nan=True)
    T = trm.FieldTimeSeries(path, "temperature")
    assert T.ring(1).shape == (30, 41472) and np.array_equal(T.ring(1, fill_value=-999.0)[:, grid.mask], T[1])
    from scipy.io import netcdf_file
    with netcdf_file(path, "r", mmap=False) as f:
        lon, lat = grid.masked_lonlat()
        assert np.array_equal(f.variables["lon"][:], lon) and np.array_equal(f.variables["lat"][:], lat)
        assert np.array_equal(f.variables["ring_index"][:], np.flatnonzero(grid.mask))
