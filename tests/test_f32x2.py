"""The packed Float32 kernel (two columns per thread, f32x2 arithmetic; csrc/euler2_kernel.cuh) against the one-column
Float32 kernel of the same math mode (TRM_F32X2=0) and against the Float32 oracle.

Float32 is the number format of the reference's global configurations (examples/simulations/soil_heat_global.jl:29-38) and of
its benchmark (test/benchmarks/gpu/soil_heat_hydrology_global.jl:40). The two CUDA kernels evaluate the same formulas with
different instruction selection (packed FMA where the scalar kernel lets the compiler contract), so they agree to a few
ulp of the field scale per step; the bar against the oracle is the one of the other Float32 parity tests."""
import os

import numpy as np
import pytest

from common import make, max_scaled_err, richards_soil, synthetic_columns, synthetic_land_case, synthetic_soil_case, trm

pytestmark = pytest.mark.gpu

FIELDS = ("temperature", "internal_energy", "saturation_water_ice", "liquid_water_fraction")


class scalar_kernel:
    """Run the enclosed steps on the one-column-per-thread kernel."""

    def __enter__(self):
        os.environ["TRM_F32X2"] = "0"

    def __exit__(self, *exc):
        os.environ.pop("TRM_F32X2", None)


def both(build, nsteps, dt=60.0, chunks=1):
    a, b = build(), build()
    for _ in range(chunks):
        a.step(dt, nsteps)
        with scalar_kernel():
            b.step(dt, nsteps)
    return a, b


@pytest.mark.parametrize("stepper", ["euler", "heun"])
@pytest.mark.parametrize("ncol", [1, 2, 255, 1000 + 13])
def test_packed_equals_scalar_soil_richards(stepper, ncol):
    a, b = both(lambda: synthetic_soil_case("cuda", ncol, nf=np.float32, heun=stepper == "heun", math="fast"), 100, chunks=2)
    for name in FIELDS + ("pressure_head", "water_table", "surface_excess_water"):
        x, y = getattr(a.state, name).numpy(), getattr(b.state, name).numpy()
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= 2.0e-6, (name, max_scaled_err(x, y))
    da, db = a.diagnostics(), b.diagnostics()
    assert da["water"] == pytest.approx(db["water"], rel=1e-6)


@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_packed_against_oracle_soil_richards(stepper):
    n = 512 + 5
    gpu = synthetic_soil_case("cuda", n, nf=np.float32, heun=stepper == "heun", math="fast")
    cpu = synthetic_soil_case("oracle", n, nf=np.float32, heun=stepper == "heun")
    gpu.step(60.0, 200)
    cpu.step(60.0, 200)
    for name in FIELDS:
        assert max_scaled_err(getattr(gpu.state, name).numpy(), getattr(cpu.state, name).numpy()) <= 5.0e-5, name


@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_packed_heat_only(stepper):
    """BASELINE config 2 (soil_heat_global): immobile water, both saturation-halo conventions."""
    n = 777
    a, b = both(lambda: synthetic_soil_case("cuda", n, nf=np.float32, richards=False, heun=stepper == "heun", math="fast"), 200, dt=300.0)
    cpu = synthetic_soil_case("oracle", n, nf=np.float32, richards=False, heun=stepper == "heun")
    cpu.step(300.0, 200)
    for name in ("temperature", "internal_energy", "liquid_water_fraction"):
        x, y, z = (getattr(s.state, name).numpy() for s in (a, b, cpu))
        assert max_scaled_err(x, y) <= 2.0e-6, name
        assert max_scaled_err(x, z) <= 5.0e-5, name


@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_packed_land_model(stepper):
    n = 301
    a, b = both(lambda: synthetic_land_case("cuda", n, nf=np.float32, heun=stepper == "heun", math="fast", windspeed=0.5), 150)
    for name in FIELDS + ("pressure_head", "skin_temperature", "ground_heat_flux", "infiltration", "surface_excess_water"):
        x, y = getattr(a.state, name).numpy(), getattr(b.state, name).numpy()
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= 5.0e-6, (name, max_scaled_err(x, y))


@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_packed_vegetated_land_model(stepper):
    from test_vegetation import synthetic_vegetated_case
    n = 203
    a, b = both(lambda: synthetic_vegetated_case("cuda", n, nf=np.float32, heun=stepper == "heun", math="fast"), 100)
    for name in FIELDS + ("carbon_vegetation", "canopy_water", "soil_moisture_limiting_factor", "transpiration", "ground_heat_flux"):
        x, y = getattr(a.state, name).numpy(), getattr(b.state, name).numpy()
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= 2.0e-5, (name, max_scaled_err(x, y))


def test_packed_negative_saturation_slow_path():
    """A strong sink drives layers negative in SOME columns: a pair may hold one column on the slow path and one on the
    fast path (soil_hydrology.jl:201-216)."""
    n = 97

    def build(engine):
        rng = np.random.default_rng(7)
        grid = trm.ColumnGrid(trm.B200(), np.float32, trm.UniformSpacing(dz=0.1, N=20), n)
        model = trm.SoilModel(grid, soil=richards_soil(vwc_forcing=-2.0e-4))
        sat0 = rng.uniform(0.0, 0.05, (20, n))
        sat0[:, ::3] = 0.9   # every third column stays on the fast path
        return make(engine, model, trm.ForwardEuler(dt=60.0), initializers={"temperature": 5.0, "saturation_water_ice": sat0}, math="fast")

    a, b = both(lambda: build("cuda"), 1)
    cpu = build("oracle")
    cpu.step(60.0, 1)
    for name in FIELDS + ("water_table", "surface_excess_water"):
        x, y, z = (getattr(s.state, name).numpy() for s in (a, b, cpu))
        assert max_scaled_err(x, y) <= 1.0e-6, name
        assert max_scaled_err(x, z) <= 1.0e-5, name
    pa, pc = a.state.pressure_head.numpy(), cpu.state.pressure_head.numpy()
    assert np.array_equal(np.isneginf(pa), np.isneginf(pc)) and np.any(np.isneginf(pc))
    ok = np.isfinite(pc)
    assert max_scaled_err(pa[ok], pc[ok]) <= 1.0e-4


def test_packed_over_saturation_to_surface_excess():
    n = 65

    def build(engine):
        grid = trm.ColumnGrid(trm.B200(), np.float32, trm.UniformSpacing(dz=0.1, N=12), n)
        model = trm.SoilModel(grid, soil=richards_soil(vwc_forcing=+4.0e-4))
        return make(engine, model, trm.ForwardEuler(dt=60.0), initializers={"temperature": 5.0, "saturation_water_ice": 0.97}, math="fast")

    a, b = both(lambda: build("cuda"), 10)
    cpu = build("oracle")
    cpu.step(60.0, 10)
    for name in FIELDS + ("pressure_head", "water_table", "surface_excess_water"):
        x, y, z = (getattr(s.state, name).numpy() for s in (a, b, cpu))
        assert max_scaled_err(x, y) <= 1.0e-6, name
        assert max_scaled_err(x, z) <= 1.0e-5, name
    assert np.all(cpu.state.surface_excess_water.numpy() > 0)


def test_packed_kernel_is_the_one_that_runs():
    """The Float32 fast-math step launches euler2_kernel (one launch per step, half as many blocks)."""
    if os.environ.get("TRM_WARP") == "1":
        pytest.skip("TRM_TEST_WARP=1: small domains run the warp-per-column kernel")
    gpu = synthetic_soil_case("cuda", 4096, nf=np.float32, math="fast")
    gpu.step(60.0, 2)          # first step reads the stored closure fields (one-column kernel), then the packed kernel
    l0 = gpu._lib.launch_count(gpu._h)
    gpu.step(60.0, 5)
    assert gpu._lib.launch_count(gpu._h) - l0 == 5


@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_packed_default_hydraulics_brooks_corey_linear(stepper):
    """The reference's DEFAULT hydraulics -- Brooks-Corey retention curve (lambda = 0.2) + linear conductivity,
    ConstantSoilHydraulics() -- are the soil of its own benchmark (test/benchmarks/gpu/soil_heat_hydrology_global.jl:46-56):
    the packed kernel has an instantiation for them (integer 1 / lambda by repeated multiplication)."""
    n = 515
    lat, lon, T0 = synthetic_columns(n)

    def build(engine):
        grid = trm.ColumnGrid(trm.B200(), np.float32, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=30), n)
        soil = trm.SoilEnergyWaterCarbon(hydrology=trm.SoilHydrology(trm.RichardsEq()))
        model = trm.SoilModel(grid, soil=soil)
        bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0))
        inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z, "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
        return make(engine, model, (trm.Heun if stepper == "heun" else trm.ForwardEuler)(dt=60.0), boundary_conditions=bcs, initializers=inits, math="fast")

    # (40 steps: with this steep retention curve the explicit scheme itself loses stability on the 5 cm top layers after ~60
    #  steps of 60 s from this unsaturated profile -- in the oracle exactly as on the GPU)
    a, b = both(lambda: build("cuda"), 20, chunks=2)
    cpu = build("oracle")
    cpu.step(60.0, 40)
    l0 = a._lib.launch_count(a._h)
    a.step(60.0, 1)
    if os.environ.get("TRM_WARP") != "1":
        assert a._lib.launch_count(a._h) - l0 == (2 if stepper == "heun" else 1)
    cpu.step(60.0, 1)
    with scalar_kernel():
        b.step(60.0, 1)
    for name in FIELDS + ("pressure_head", "water_table"):
        x, y, z = (getattr(s.state, name).numpy() for s in (a, b, cpu))
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= 5.0e-6, (name, max_scaled_err(x, y))
        assert max_scaled_err(x, z) <= 5.0e-5, (name, max_scaled_err(x, z))


def test_packed_heun_negative_saturation_stage_state():
    """Heun, recompute protocol in the pair kernels: in the SECOND step (the first one reads the stored closure fields on the
    one-column kernel) the stage state of two thirds of the columns goes negative; the pair kernel's stage 1 then forms and
    stores the stage state of exactly those columns and flags them -- a pair may hold one flagged and one rebuilt column --
    and its stage 2 reads them back. Same outcome as on the one-column kernel: the same columns end with a NaN saturation
    (tests/test_parity.py::test_heun_negative_saturation_stage_state), everything finite agrees."""
    n = 97

    def build():
        rng = np.random.default_rng(7)
        grid = trm.ColumnGrid(trm.B200(), np.float32, trm.UniformSpacing(dz=0.1, N=20), n)
        model = trm.SoilModel(grid, soil=richards_soil(vwc_forcing=-1.0e-4))
        sat0 = rng.uniform(0.014, 0.022, (20, n))
        sat0[:, ::3] = 0.9   # every third column stays on the fast path
        return make("cuda", model, trm.Heun(dt=60.0), initializers={"temperature": 5.0, "saturation_water_ice": sat0}, math="fast")

    a, b = both(build, 1, chunks=2)
    sa, sb = a.state.saturation_water_ice.numpy(), b.state.saturation_water_ice.numpy()
    broken = (~np.isfinite(sb)).any(axis=0)
    assert broken.sum() >= n // 2 and np.isfinite(sb[:, ::3]).all()
    assert np.array_equal((~np.isfinite(sa)).any(axis=0), broken)
    assert max_scaled_err(sa[:, ~broken], sb[:, ~broken]) <= 2.0e-6
    Ua, Ub = a.state.internal_energy.numpy(), b.state.internal_energy.numpy()
    assert np.isfinite(Ub).all() and max_scaled_err(Ua, Ub) <= 2.0e-6
