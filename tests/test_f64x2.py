"""The Float64 pair kernel (two columns per thread, 16-byte accesses; csrc/euler2_kernel.cuh instantiated for double; runs on
request, TRM_F64X2=1 -- the one-column kernel is the faster one in Float64 and the default) against the one-column Float64 kernel
of the same math mode and against the Float64 oracle.

Float64 is the number format of the parity bar (BASELINE.json: max relative error 1e-9 after 1000 steps) and of the 10 M-column
benchmark. Both CUDA kernels evaluate the same fast-math formulas (reciprocal / rsqrt seeds + one third-order step) with a
different grouping of a few products, so they agree to rounding; the bar against the oracle is the one of tests/test_parity.py."""
import os

import numpy as np
import pytest

from common import make, max_scaled_err, pointwise_relerr, richards_soil, synthetic_columns, synthetic_land_case, synthetic_soil_case, trm

pytestmark = pytest.mark.gpu

FIELDS = ("temperature", "internal_energy", "saturation_water_ice", "liquid_water_fraction")


class pair_kernel:
    """Run the enclosed steps on the two-columns-per-thread Float64 kernel."""

    def __init__(self, value="1"):
        self.value = value

    def __enter__(self):
        os.environ["TRM_F64X2"] = self.value

    def __exit__(self, *exc):
        os.environ.pop("TRM_F64X2", None)


def both(build, nsteps, dt=60.0, chunks=1):
    a, b = build(), build()
    for _ in range(chunks):
        with pair_kernel():
            a.step(dt, nsteps)
        with pair_kernel("0"):   # (the default Brooks-Corey + linear soil runs the pair kernel unless it is switched off)
            b.step(dt, nsteps)
    return a, b


@pytest.mark.parametrize("ncol", [1, 2, 255, 1000 + 13])
def test_pair_equals_scalar_soil_richards(ncol):
    a, b = both(lambda: synthetic_soil_case("cuda", ncol, nf=np.float64, math="fast"), 100, chunks=2)
    for name in FIELDS + ("pressure_head", "water_table", "surface_excess_water"):
        x, y = getattr(a.state, name).numpy(), getattr(b.state, name).numpy()
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= 1.0e-12, (name, max_scaled_err(x, y))
    da, db = a.diagnostics(), b.diagnostics()
    assert da["water"] == pytest.approx(db["water"], rel=1e-13)


def test_pair_against_oracle_soil_richards_1000_steps():
    n = 512 + 5
    gpu = synthetic_soil_case("cuda", n, nf=np.float64, math="fast")
    cpu = synthetic_soil_case("oracle", n, nf=np.float64)
    with pair_kernel():
        gpu.step(60.0, 1000)
    cpu.step(60.0, 1000)
    for name in FIELDS + ("pressure_head",):
        x, y = getattr(gpu.state, name).numpy(), getattr(cpu.state, name).numpy()
        assert max_scaled_err(x, y) <= 1.0e-9, (name, max_scaled_err(x, y))
        assert pointwise_relerr(x, y) <= 1.0e-9, (name, pointwise_relerr(x, y))


def test_pair_heat_only():
    """BASELINE config 2 (soil_heat_global): immobile water."""
    n = 777
    a, b = both(lambda: synthetic_soil_case("cuda", n, nf=np.float64, richards=False, math="fast"), 200, dt=300.0)
    cpu = synthetic_soil_case("oracle", n, nf=np.float64, richards=False)
    cpu.step(300.0, 200)
    for name in ("temperature", "internal_energy", "liquid_water_fraction"):
        x, y, z = (getattr(s.state, name).numpy() for s in (a, b, cpu))
        assert max_scaled_err(x, y) <= 1.0e-12, name
        assert max_scaled_err(x, z) <= 1.0e-9, name


def test_pair_land_model():
    n = 301
    a, b = both(lambda: synthetic_land_case("cuda", n, nf=np.float64, math="fast", windspeed=0.5), 150)
    for name in FIELDS + ("pressure_head", "skin_temperature", "ground_heat_flux", "infiltration", "surface_excess_water"):
        x, y = getattr(a.state, name).numpy(), getattr(b.state, name).numpy()
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= 1.0e-11, (name, max_scaled_err(x, y))


def test_pair_vegetated_land_model():
    from test_vegetation import synthetic_vegetated_case
    n = 203
    a, b = both(lambda: synthetic_vegetated_case("cuda", n, nf=np.float64, math="fast"), 100)
    for name in FIELDS + ("carbon_vegetation", "canopy_water", "soil_moisture_limiting_factor", "transpiration", "ground_heat_flux"):
        x, y = getattr(a.state, name).numpy(), getattr(b.state, name).numpy()
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= 1.0e-10, (name, max_scaled_err(x, y))


def test_pair_negative_saturation_slow_path():
    """A strong sink drives layers negative in SOME columns: a pair may hold one column on the slow path and one on the
    fast path (soil_hydrology.jl:201-216)."""
    n = 97

    def build(engine):
        rng = np.random.default_rng(7)
        grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.1, N=20), n)
        model = trm.SoilModel(grid, soil=richards_soil(vwc_forcing=-2.0e-4))
        sat0 = rng.uniform(0.0, 0.05, (20, n))
        sat0[:, ::3] = 0.9   # every third column stays on the fast path
        return make(engine, model, trm.ForwardEuler(dt=60.0), initializers={"temperature": 5.0, "saturation_water_ice": sat0}, math="fast")

    a, b = both(lambda: build("cuda"), 1)
    cpu = build("oracle")
    cpu.step(60.0, 1)
    for name in FIELDS + ("water_table", "surface_excess_water"):
        x, y, z = (getattr(s.state, name).numpy() for s in (a, b, cpu))
        assert max_scaled_err(x, y) <= 1.0e-12, name
        assert max_scaled_err(x, z) <= 1.0e-10, name
    pa, pc = a.state.pressure_head.numpy(), cpu.state.pressure_head.numpy()
    assert np.array_equal(np.isneginf(pa), np.isneginf(pc)) and np.any(np.isneginf(pc))
    ok = np.isfinite(pc)
    assert max_scaled_err(pa[ok], pc[ok]) <= 1.0e-9


def test_pair_over_saturation_to_surface_excess():
    n = 65

    def build(engine):
        grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.1, N=12), n)
        model = trm.SoilModel(grid, soil=richards_soil(vwc_forcing=+4.0e-4))
        return make(engine, model, trm.ForwardEuler(dt=60.0), initializers={"temperature": 5.0, "saturation_water_ice": 0.97}, math="fast")

    a, b = both(lambda: build("cuda"), 10)
    cpu = build("oracle")
    cpu.step(60.0, 10)
    for name in FIELDS + ("pressure_head", "water_table", "surface_excess_water"):
        x, y, z = (getattr(s.state, name).numpy() for s in (a, b, cpu))
        assert max_scaled_err(x, y) <= 1.0e-12, name
        assert max_scaled_err(x, z) <= 1.0e-10, name
    assert np.all(cpu.state.surface_excess_water.numpy() > 0)


def test_pair_default_hydraulics_brooks_corey_linear():
    """The reference's DEFAULT hydraulics (Brooks-Corey lambda = 0.2 + linear conductivity, ConstantSoilHydraulics()) in Float64."""
    n = 515
    lat, lon, T0 = synthetic_columns(n)

    def build(engine):
        grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=30), n)
        soil = trm.SoilEnergyWaterCarbon(hydrology=trm.SoilHydrology(trm.RichardsEq()))
        model = trm.SoilModel(grid, soil=soil)
        bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0))
        inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z, "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
        return make(engine, model, trm.ForwardEuler(dt=60.0), boundary_conditions=bcs, initializers=inits, math="fast")

    a, b = both(lambda: build("cuda"), 20, chunks=2)
    cpu = build("oracle")
    cpu.step(60.0, 40)
    for name in FIELDS + ("pressure_head", "water_table"):
        x, y, z = (getattr(s.state, name).numpy() for s in (a, b, cpu))
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= 1.0e-11, (name, max_scaled_err(x, y))
        assert max_scaled_err(x, z) <= 1.0e-9, (name, max_scaled_err(x, z))
