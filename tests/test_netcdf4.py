"""NetCDF-4 / HDF5 decoder of the forcing-ingestion row: files written by ``tests/hdf5_writer.py`` in the
netCDF-4 layouts, and — where the reference checkout is present (this container, not the GPU box) — the reference's own
ERA5-Land land-sea masks (``inputs/era5-land_land_sea_mask_N72.nc``, read by ``examples/simulations/soil_heat_global.jl:30``)."""
import os

import numpy as np
import pytest

from common import ENGINES, make, trm
from hdf5_writer import Writer

netcdf4 = trm.netcdf4
REF_INPUTS = "/root/reference/inputs"


def _era5_like(path, header_version, nt=5, nlat=6, nlon=9, dtype=">i2", seed=0):
    """A packed ERA5-style variable (int16, scale / offset, fill value) chunked with edge chunks, shuffle + deflate."""
    rng = np.random.default_rng(seed)
    truth = 250.0 + 50.0 * rng.random((nt, nlat, nlon))
    scale, offset = 50.0 / 60000.0, 275.0
    packed = np.round((truth - offset) / scale).astype(dtype)
    packed[1, 2, 3] = -32767
    w = Writer(header_version, attrs={"Conventions": "CF-1.6"})
    w.add("time", (np.arange(nt) * 6 + 876576).astype("<i4"), layout="contiguous", attrs={"units": "hours since 1900-01-01 00:00:00.0"})
    w.add("latitude", np.linspace(75.0, -75.0, nlat).astype("<f4"), layout="compact")
    w.add("longitude", np.linspace(0.0, 320.0, nlon), chunks=(4,), deflate=4)
    w.add("t2m", packed, chunks=(2, 4, 4), deflate=6, shuffle=True, fletcher32=(header_version == 2), dims=("time", "latitude", "longitude"),
          attrs={"scale_factor": np.float64(scale), "add_offset": np.float64(offset), "_FillValue": np.dtype(dtype).type(-32767), "units": "K"})
    w.add("z", rng.random((nlat, nlon)).astype(">f8"), chunks=(nlat, nlon), shuffle=True, deflate=1, dims=("latitude", "longitude"))
    w.save(path)
    want = packed.astype(np.float64) * scale + offset
    want[1, 2, 3] = np.nan
    return packed, want


@pytest.mark.parametrize("header_version", [1, 2])
@pytest.mark.parametrize("dtype", [">i2", "<i2", "<u1", "<f4"])
def test_decoder_layouts_filters_and_types(tmp_path, header_version, dtype):
    path = str(tmp_path / "era5.nc")
    packed, want = _era5_like(path, header_version, dtype=dtype if dtype != "<u1" else "<i2")
    if dtype == "<u1":   # an unshuffled single-byte variable in its own file
        w = Writer(header_version)
        w.add("flag", np.arange(35, dtype="u1").reshape(5, 7), chunks=(2, 3), shuffle=True, deflate=9)
        w.save(path)
        with netcdf4.File(path) as f:
            assert np.array_equal(f.variables["flag"][:], np.arange(35, dtype="u1").reshape(5, 7))
        return
    assert netcdf4.is_hdf5(path)
    with netcdf4.File(path) as f:
        assert set(f.variables) == {"time", "latitude", "longitude", "t2m", "z"}
        assert f.attrs["Conventions"] == "CF-1.6"
        v = f.variables["t2m"]
        assert v.shape == packed.shape and v.dtype == np.dtype(dtype).newbyteorder("=")
        assert v.dimensions == ("time", "latitude", "longitude") and f.variables["z"].dimensions == ("latitude", "longitude")
        assert f.variables["time"].dimensions == ("time",)
        assert np.array_equal(v[:], packed.astype(v.dtype))
        assert v.attrs["units"] == "K" and v.attrs["scale_factor"] == pytest.approx(50.0 / 60000.0, rel=1e-15)
        np.testing.assert_array_equal(v.scaled(), want)
        assert np.array_equal(f.variables["time"][:], np.arange(5) * 6 + 876576)
        assert np.array_equal(f.variables["latitude"][:], np.linspace(75.0, -75.0, 6).astype("f4"))
        assert np.array_equal(f.variables["longitude"][:], np.linspace(0.0, 320.0, 9))


def test_decoder_rejects_what_it_does_not_decode(tmp_path):
    path = str(tmp_path / "plain.nc")
    with open(path, "wb") as f:
        f.write(b"CDF\x01" + b"\0" * 64)
    assert not netcdf4.is_hdf5(path)
    with pytest.raises(ValueError):
        netcdf4.File(path)
    w = Writer(2)
    w.add("x", np.arange(8.0), chunks=(8,), deflate=1)
    w.save(path)
    raw = bytearray(open(path, "rb").read())
    i = raw.index(bytes([2, 1]) + b"\x01\x00\x01\x00\x01\x00")   # filter pipeline message: deflate -> an unknown filter id
    raw[i + 2:i + 4] = (4).to_bytes(2, "little")   # szip
    open(path, "wb").write(bytes(raw))
    with pytest.raises(NotImplementedError, match="filter 4"):
        netcdf4.File(path).variables["x"][:]


@pytest.mark.parametrize("engine", ENGINES)
def test_raster_input_from_netcdf4_file(engine, tmp_path):
    """``RasterInputSource.from_netcdf`` on a packed, compressed NetCDF-4 file drives the surface temperature with the
    Rasters-extension update rule (TerrariumRastersExt.jl:96-121): halfway between two six-hourly snapshots."""
    path = str(tmp_path / "era5.nc")
    nt, nlat, nlon = 5, 6, 9
    _, want = _era5_like(path, 2, nt, nlat, nlon)
    want = np.nan_to_num(want - 273.15, nan=0.0)
    src = trm.RasterInputSource.from_netcdf(path, "t2m", decode_times=True, reftime=None)
    assert src.values.shape == (nt, nlat * nlon) and src.reftime == 876576 * 3600.0
    assert np.array_equal(np.diff(src.times), np.full(nt - 1, 21600.0))
    src.values = np.nan_to_num(src.values - 273.15, nan=0.0)
    mask = np.zeros(nlat * nlon, dtype=bool)
    mask[::2] = True
    grid = trm.ColumnRingGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_max=1.0, N=8), mask)
    bcs = trm.PrescribedSurfaceTemperature("T_ub", src)
    integ = make(engine, trm.SoilModel(grid), trm.ForwardEuler(dt=100.0), boundary_conditions=bcs,
                 initializers={"temperature": 1.0, "saturation_water_ice": 0.5})
    integ.step(100.0, 109)   # last update at t = 10 800 s: halfway between the first two snapshots
    np.testing.assert_allclose(integ.state.T_ub.numpy(), (0.5 * (want[0] + want[1])).reshape(-1)[mask], rtol=1e-13)


needs_reference = pytest.mark.skipif(not os.path.isdir(REF_INPUTS), reason="reference checkout not present (GPU box)")


@needs_reference
@pytest.mark.parametrize("name, shape, ncol", [("N72", (1, 144, 288), 14017), ("N145", (1, 290, 580), 56951)])
def test_reference_land_sea_masks(name, shape, ncol):
    """The masks of BASELINE configs 2-4: Gaussian grids N72 / N145, land fraction in [0, 1], and the number of columns
    ``sum(land_sea_frac .> 0.5)`` of ``soil_heat_global.jl:36-38``."""
    path = f"{REF_INPUTS}/era5-land_land_sea_mask_{name}.nc"
    with netcdf4.File(path) as f:
        v = f.variables["lsm"]
        assert v.shape == shape and v.dimensions == ("time", "lat", "lon") and v.attrs["standard_name"] == "land_binary_mask"
        x = v.scaled()
        assert np.isfinite(x).all() and x.min() == 0.0 and abs(x.max() - 1.0) < 1e-12
        lat = f.variables["lat"][:]
        assert lat[0] > 89.0 and np.allclose(lat, -lat[::-1]) and np.all(np.diff(lat) < 0)   # Gaussian latitudes, north first
        assert np.allclose(np.diff(f.variables["lon"][:]), 360.0 / shape[2])
        assert f.attrs["Conventions"] == "CF-1.6"
    grid = trm.ColumnRingGrid.from_land_sea_mask(np.float32, trm.ExponentialSpacing(N=30), path=path)
    assert grid.Nc == ncol and grid.npoints == shape[1] * shape[2] and grid.Nz == 30
    lon, lat = grid.masked_lonlat()
    assert lon.shape == (ncol,) and 0.0 <= lon.min() and lon.max() < 2 * np.pi and np.abs(lat).max() < np.pi / 2
    # ring order: latitude rings north to south, longitude fastest
    assert np.all(np.diff(grid.lat) <= 0) and grid.lon[1] > grid.lon[0]


@pytest.mark.parametrize("header_version", [1, 2])
def test_partial_reads_decode_only_the_covered_chunks(tmp_path, header_version, monkeypatch):
    """``var[i]`` / ``var[i0:i1]`` / ``from_netcdf(time_range=...)``: rows of the first axis without decoding the rest."""
    import zlib
    path = str(tmp_path / "era5.nc")
    packed, want = _era5_like(path, header_version, nt=7)
    calls = []
    real = zlib.decompress
    monkeypatch.setattr(netcdf4.zlib, "decompress", lambda raw: (calls.append(len(raw)), real(raw))[1])
    with netcdf4.File(path) as f:
        v = f.variables["t2m"]          # chunks (2, 4, 4) over (7, 6, 9): 4 x 2 x 3 chunks
        full = v[:]
        assert len(calls) == 24 and np.array_equal(full, packed.astype(full.dtype))
        del calls[:]
        assert np.array_equal(v[3], full[3]) and len(calls) == 6
        del calls[:]
        assert np.array_equal(v[2:5, 1:3], full[2:5, 1:3]) and len(calls) == 12     # rows 2..4 touch chunk rows 1 and 2
        assert np.array_equal(v[-1], full[-1]) and np.array_equal(v[5:5], full[5:5]) and np.array_equal(v[::2], full[::2])
        np.testing.assert_array_equal(v.scaled(first=1, last=3), want[1:3])
        assert np.array_equal(f.variables["time"][1:3], (np.arange(7) * 6 + 876576)[1:3])        # contiguous
        assert np.array_equal(f.variables["latitude"][2:], np.linspace(75.0, -75.0, 6).astype("f4")[2:])   # compact
        with pytest.raises(IndexError):
            v[7]
    src = trm.RasterInputSource.from_netcdf(path, "t2m", decode_times=True, reftime=None, time_range=(2, 6))
    assert src.values.shape == (4, 54) and src.reftime == (876576 + 12) * 3600.0
    np.testing.assert_array_equal(src.values, want[2:6].reshape(4, -1))
    with pytest.raises(ValueError):      # closed with the context manager
        f.variables["t2m"].read()


def test_decoder_round_trips_random_layouts(tmp_path):
    """Random shapes, chunk shapes (edge chunks, more chunks than one B-tree leaf), types, byte orders and filter pipelines
    written by the test-side writer come back bit-exact, whole and by row range."""
    from hypothesis import given, settings, strategies as st

    counter = [0]

    @settings(max_examples=60, deadline=None)
    @given(shape=st.lists(st.integers(1, 9), min_size=1, max_size=3), data=st.data(),
           dtype=st.sampled_from(["<i2", ">i2", "<i4", ">u4", "<u1", "<f4", ">f8", "<i8"]),
           shuffle=st.booleans(), deflate=st.integers(0, 9), fletcher=st.booleans(), hv=st.sampled_from([1, 2]))
    def check(shape, data, dtype, shuffle, deflate, fletcher, hv):
        chunks = tuple(data.draw(st.integers(1, s)) for s in shape)
        rng = np.random.default_rng(counter[0])
        counter[0] += 1
        a = (rng.integers(0, 250, size=shape) if np.dtype(dtype).kind in "iu" else rng.normal(size=shape)).astype(dtype)
        path = str(tmp_path / f"r{counter[0]}.nc")
        w = Writer(hv)
        w.add("x", a, chunks=chunks, shuffle=shuffle, deflate=deflate, fletcher32=fletcher)
        w.save(path)
        with netcdf4.File(path) as f:
            v = f.variables["x"]
            assert v.shape == tuple(shape) and np.array_equal(v[:], a.astype(v.dtype))
            i0 = data.draw(st.integers(0, shape[0]))
            i1 = data.draw(st.integers(i0, shape[0]))
            assert np.array_equal(v[i0:i1], a[i0:i1].astype(v.dtype))
        os.remove(path)

    check()


def test_raster_from_several_files(tmp_path):
    """Monthly / yearly files concatenated along time (NetCDF-4 and NetCDF-3 mixed), epochs checked."""
    from scipy.io import netcdf_file
    a = np.arange(2 * 3 * 4, dtype="<f4").reshape(2, 3, 4)
    b = 100 + np.arange(3 * 3 * 4, dtype="<f4").reshape(3, 3, 4)
    p1, p2 = str(tmp_path / "m1.nc"), str(tmp_path / "m2.nc")
    w = Writer(2)
    w.add("time", np.array([0, 12], dtype="<i4"), layout="contiguous", attrs={"units": "hours since 2000-01-01"})
    w.add("y", np.arange(3.0), layout="contiguous")
    w.add("x", np.arange(4.0), layout="contiguous")
    w.add("t2m", a, chunks=(1, 3, 4), deflate=3, dims=("time", "y", "x"))
    w.save(p1)
    with netcdf_file(p2, "w") as f:
        f.createDimension("time", 3); f.createDimension("y", 3); f.createDimension("x", 4)
        tv = f.createVariable("time", "f8", ("time",)); tv[:] = [1440.0, 2160.0, 2880.0]; tv.units = "minutes since 2000-01-01"
        v = f.createVariable("t2m", "f4", ("time", "y", "x")); v[:] = b
    src = trm.raster_from_netcdf_files([p1, p2], "t2m")
    assert np.array_equal(src.times, np.array([0, 12, 24, 36, 48]) * 3600.0) and src.reftime == 0.0
    assert np.array_equal(src.values, np.concatenate([a, b]).reshape(5, 12))
    with pytest.raises(ValueError, match="continue each other"):
        trm.raster_from_netcdf_files([p2, p1], "t2m")
