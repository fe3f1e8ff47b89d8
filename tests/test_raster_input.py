"""RasterInputSource (ext/TerrariumRastersExt/TerrariumRastersExt.jl:22-121): rasters on the ring grid
of a ColumnRingGrid, gathered to the masked columns and interpolated in time on the device with the extension's rule."""
import numpy as np
import pytest

from common import ENGINES, make, trm


def _ring_case(engine, source, nf=np.float64):
    rng = np.random.default_rng(3)
    nring = 40
    mask = rng.uniform(size=nring) > 0.4
    grid = trm.ColumnRingGrid(trm.B200(), nf, trm.ExponentialSpacing(dz_max=1.0, N=8), mask)
    model = trm.SoilModel(grid)
    bcs = trm.PrescribedSurfaceTemperature("T_ub", source)
    integ = make(engine, model, trm.ForwardEuler(dt=100.0), boundary_conditions=bcs, initializers={"temperature": 1.0, "saturation_water_ice": 0.5})
    return integ, mask


def _expected(values, times, reftime, mask, t):
    """update_from_raster!, TerrariumRastersExt.jl:96-121, restated with numpy searchsorted."""
    tt = np.asarray(times, dtype=np.float64) - reftime
    x = values[:, mask]
    right = int(np.searchsorted(tt, t, side="left")) + 1    # first(searchsorted(...)), 1-based
    left = int(np.searchsorted(tt, t, side="right"))        # last(searchsorted(...)), 1-based
    if left >= 1 and right <= tt.size:
        x1, x2 = x[left - 1], x[right - 1]
        dt = tt[right - 1] - tt[left - 1]
        return x1 + (t - tt[left - 1]) * (x2 - x1) / dt if dt > 0 else x2
    return x[min(right, tt.size) - 1]


@pytest.mark.parametrize("engine", ENGINES)
def test_time_varying_raster_drives_the_surface_temperature(engine):
    rng = np.random.default_rng(4)
    times = np.array([1000.0, 1250.0, 1600.0, 2000.0])   # seconds on the raster's own axis ; reftime shifts it
    values = rng.uniform(-5.0, 15.0, (times.size, 40))
    src = trm.RasterInputSource(values=values, times=times, reftime=1000.0)
    integ, mask = _ring_case(engine, src)
    assert integ.ncol == int(mask.sum())
    # the input field after a step holds the value of the last update_inputs! (time of the step's start):
    # before the axis (none here: t = 0 is the first node), on a node, between nodes, beyond the end
    for nsteps, t_eval in ((1, 0.0), (2, 200.0), (1, 300.0), (3, 600.0), (6, 1200.0)):
        integ.step(100.0, nsteps)
        got = integ.state.T_ub.numpy()
        np.testing.assert_allclose(got, _expected(values, times, 1000.0, mask, t_eval), rtol=1e-15, atol=0, err_msg=str(t_eval))
    # t = 250 s is the second node: the node value itself
    integ2, _ = _ring_case(engine, trm.RasterInputSource(values=values, times=times, reftime=750.0))
    integ2.step(100.0, 1)
    np.testing.assert_array_equal(integ2.state.T_ub.numpy(), values[0][mask])   # t = 0 lies before the axis: flat
    integ2.step(100.0, 5)   # last update at t = 500 = node 2 (1250 - 750)
    np.testing.assert_array_equal(integ2.state.T_ub.numpy(), values[1][mask])


@pytest.mark.parametrize("engine", ENGINES)
def test_static_raster_and_netcdf3_reader(engine, tmp_path):
    from scipy.io import netcdf_file
    rng = np.random.default_rng(5)
    nlat, nlon, nt = 5, 8, 3
    data = rng.uniform(0.0, 10.0, (nt, nlat, nlon)).astype(np.float32)
    path = str(tmp_path / "forcing.nc")
    with netcdf_file(path, "w") as f:
        f.createDimension("time", nt); f.createDimension("lat", nlat); f.createDimension("lon", nlon)
        tv = f.createVariable("time", "f8", ("time",)); tv[:] = [0.0, 3600.0, 7200.0]
        v = f.createVariable("t2m", "f4", ("time", "lat", "lon")); v[:] = data
        v.scale_factor = 2.0; v.add_offset = -1.0
        s = f.createVariable("orography", "f4", ("lat", "lon")); s[:] = data[0]
    src = trm.RasterInputSource.from_netcdf(path, "t2m")
    assert src.values.shape == (nt, nlat * nlon) and np.allclose(src.values, data.reshape(nt, -1) * 2.0 - 1.0)
    integ, mask = _ring_case(engine, src)
    integ.step(100.0, 19)   # last update at t = 1800 s: halfway between the first two snapshots
    want = 0.5 * (src.values[0] + src.values[1])[mask]
    np.testing.assert_allclose(integ.state.T_ub.numpy(), want, rtol=1e-14)
    # a raster without a time axis is copied once
    static = trm.RasterInputSource.from_netcdf(path, "orography")
    assert static.times is None
    integ, mask = _ring_case(engine, static)
    integ.step(100.0, 3)
    np.testing.assert_array_equal(integ.state.T_ub.numpy(), data[0].reshape(-1).astype(np.float64)[mask])


@pytest.mark.gpu
def test_raster_forcing_parity_with_oracle():
    rng = np.random.default_rng(6)
    times = np.linspace(0.0, 86400.0, 25)
    values = 5.0 + 8.0 * np.sin(2 * np.pi * times / 86400.0)[:, None] + rng.normal(0.0, 1.0, (25, 40))
    runs = []
    for engine in ("cuda", "oracle"):
        for nf in (np.float64, np.float32):
            integ, _ = _ring_case(engine, trm.RasterInputSource(values=values, times=times), nf=nf)
            integ.step(100.0, 500)
            runs.append(integ.state.temperature.numpy())
    assert np.max(np.abs(runs[0] - runs[2])) <= 1e-11 * np.max(np.abs(runs[2]))
    assert np.max(np.abs(runs[1] - runs[3])) <= 2e-5 * np.max(np.abs(runs[3]))


@pytest.mark.parametrize("engine", ENGINES)
def test_input_source_constructor(engine):
    """initialize(model, timestepper, InputSource(grid, data; name)...) -- field sources and ring-grid rasters."""
    rng = np.random.default_rng(8)
    mask = rng.uniform(size=30) > 0.3
    grid = trm.ColumnRingGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_max=1.0, N=8), mask)
    land = trm.LandModel(grid, vegetation=None)
    Ta_ring = rng.uniform(-5.0, 20.0, (3, 30))
    srcs = [trm.InputSource(grid, Ta_ring, name="air_temperature", times=[0.0, 600.0, 1200.0]),
            trm.InputSource(grid, 2.5, name="windspeed"),
            trm.InputSource(grid, rng.uniform(100.0, 400.0, int(mask.sum())), name="surface_shortwave_down")]
    integ = make(engine, land, trm.ForwardEuler(dt=300.0), srcs, initializers={"temperature": 3.0})
    integ.step(300.0, 2)   # last update_inputs! at t = 300: halfway between the first two snapshots
    np.testing.assert_allclose(integ.state.air_temperature.numpy(), 0.5 * (Ta_ring[0] + Ta_ring[1])[mask], rtol=1e-14)
    assert np.all(integ.state.windspeed.numpy() == 2.5)
    assert np.all(np.isfinite(integ.state.ground_heat_flux.numpy()))


def test_netcdf3_packed_fill_value_time_units_and_range(tmp_path):
    """The NetCDF-3 leg of ``from_netcdf`` treats packing, missing values, CF time units and ``time_range`` like the
    NetCDF-4 leg (tests/test_netcdf4.py)."""
    from scipy.io import netcdf_file
    nt, nlat, nlon = 6, 3, 4
    packed = (np.arange(nt * nlat * nlon, dtype=np.int16).reshape(nt, nlat, nlon) - 30).astype(np.int16)
    packed[2, 1, 1] = -32767
    path = str(tmp_path / "era5_classic.nc")
    with netcdf_file(path, "w") as f:
        f.createDimension("time", nt); f.createDimension("latitude", nlat); f.createDimension("longitude", nlon)
        tv = f.createVariable("time", "i4", ("time",)); tv[:] = 1000 + np.arange(nt); tv.units = "hours since 1900-01-01 00:00:00.0"
        v = f.createVariable("t2m", "i2", ("time", "latitude", "longitude")); v[:] = packed
        v.scale_factor = 0.5; v.add_offset = 280.0; v._FillValue = np.int16(-32767); v.missing_value = np.int16(-32767)
    src = trm.RasterInputSource.from_netcdf(path, "t2m", decode_times=True, reftime=None, time_range=(1, 5))
    want = packed[1:5].astype(np.float64) * 0.5 + 280.0
    want[1, 1, 1] = np.nan
    np.testing.assert_array_equal(src.values, want.reshape(4, -1))
    assert np.array_equal(src.times, (1001 + np.arange(4)) * 3600.0) and src.reftime == 1001 * 3600.0
    with pytest.raises(ValueError, match="unsupported CF time unit"):
        trm.cf_time_unit_seconds("fortnights since 1900-01-01")
    assert trm.cf_time_unit_seconds("days since 1970-01-01") == 86400.0 and trm.cf_time_unit_seconds(None) == 1.0
