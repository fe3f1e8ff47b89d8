"""Small domains: the warp-per-column kernel (csrc/warp_kernel.cuh; lane = layer, several steps per launch, both Heun stages in
one launch) against the one-thread-per-column streaming kernels of the same math mode and against the oracle.

The library picks this kernel by itself for the SoilModel on up to 114 688 (Float32) / 65 536 (Float64) columns (TRM_WARP_COLS) with at most 31 layers; the
rest of the suite pins the streaming kernels (tests/conftest.py), this module switches the warp path on around its runs and
replays the parity / known-answer tests of the other modules through it."""
import os

import numpy as np
import pytest

import test_global_configs
import test_golden
import test_host_io
import test_parity
import test_raster_input
import test_simulation
import test_vegetation
from common import make, max_scaled_err, pointwise_relerr, richards_soil, synthetic_columns, synthetic_soil_case, trm

pytestmark = pytest.mark.gpu

FIELDS = ("temperature", "internal_energy", "saturation_water_ice", "liquid_water_fraction")
RFIELDS = FIELDS + ("pressure_head", "water_table", "surface_excess_water")


class warp_kernel:
    """Run the enclosed calls on the warp-per-column kernel (the library's default for small domains)."""

    def __enter__(self):
        self.old = os.environ.get("TRM_WARP")
        os.environ["TRM_WARP"] = "1"

    def __exit__(self, *exc):
        os.environ["TRM_WARP"] = self.old if self.old is not None else "0"


def launches(integ):
    return integ._lib.launch_count(integ._h)


@pytest.mark.parametrize("nf", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("math", ["faithful", "fast"])
@pytest.mark.parametrize("heun", [False, True], ids=["euler", "heun"])
@pytest.mark.parametrize("richards", [True, False], ids=["richards", "noflow"])
def test_warp_equals_streaming(richards, heun, math, nf):
    """200 steps in ONE launch against 200 (x 2 for Heun) launches of the streaming kernels."""
    n, dt = 257, (60.0 if richards else 300.0)
    a = synthetic_soil_case("cuda", n, nf=nf, richards=richards, heun=heun, math=math)
    b = synthetic_soil_case("cuda", n, nf=nf, richards=richards, heun=heun, math=math)
    l0 = launches(a)
    with warp_kernel():
        a.step(dt, 200)
    assert launches(a) - l0 == 1
    b.step(dt, 200)
    assert a.clock.time == b.clock.time and a.clock.iteration == b.clock.iteration
    # same formulas, same order; fast math contracts a few products differently. Float32 fast math: the streaming path is the
    # packed two-column kernel, which groups some products differently (tests/test_f32x2.py holds it to the same bar)
    tol = (1.0e-12 if nf == np.float64 else 2.0e-6) if math == "fast" else (1.0e-14 if nf == np.float64 else 1.0e-6)
    for name in (RFIELDS if richards else FIELDS):
        x, y = getattr(a.state, name).numpy(), getattr(b.state, name).numpy()
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= tol, (name, max_scaled_err(x, y))


@pytest.mark.parametrize("heun", [False, True], ids=["euler", "heun"])
def test_one_launch_equals_step_by_step(heun):
    """`nsteps` steps inside one launch == one launch per step (clock arithmetic, closure fields of intermediate states), bit for
    bit; the first step after initialize reads the stored closure fields in both."""
    n = 100
    with warp_kernel():
        a = synthetic_soil_case("cuda", n, heun=heun, math="fast")
        b = synthetic_soil_case("cuda", n, heun=heun, math="fast")
        a.step(60.0, 64)
        for _ in range(64):
            b.step(60.0, 1)
        a.step(60.0, 3)
        b.step(60.0, 3)
    for name in RFIELDS:
        assert np.array_equal(getattr(a.state, name).numpy(), getattr(b.state, name).numpy()), name
    assert a.clock.time == b.clock.time


@pytest.mark.parametrize("math", ["faithful", "fast"])
@pytest.mark.parametrize("heun", [False, True], ids=["euler", "heun"])
def test_warp_against_oracle_1000_steps(heun, math):
    n = 192 + 5
    cpu = synthetic_soil_case("oracle", n, heun=heun)
    cpu.step(60.0, 1000)
    with warp_kernel():
        gpu = synthetic_soil_case("cuda", n, heun=heun, math=math)
        l0 = launches(gpu)
        gpu.step(60.0, 1000)
        assert launches(gpu) - l0 == 1
    for name in FIELDS + ("pressure_head",):
        x, y = getattr(gpu.state, name).numpy(), getattr(cpu.state, name).numpy()
        assert max_scaled_err(x, y) <= 1.0e-9, (name, max_scaled_err(x, y))
        assert pointwise_relerr(x, y) <= 1.0e-9, (name, pointwise_relerr(x, y))


@pytest.mark.parametrize("nz", [2, 3, 12, 31])
def test_layer_counts(nz):
    """Every lane count up to the limit (lane nz is the halo cell); 32 layers and more fall back to the streaming kernels."""
    n = 33

    def build(engine, math="faithful"):
        grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.2, N=nz), n)
        model = trm.SoilModel(grid, soil=richards_soil())
        lat, lon, T0 = synthetic_columns(n)
        bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0))
        return make(engine, model, trm.Heun(dt=30.0), boundary_conditions=bcs,
                    initializers={"temperature": 2.0, "saturation_water_ice": 0.6}, math=math)

    cpu = build("oracle")
    cpu.step(30.0, 100)
    for math in ("faithful", "fast"):
        with warp_kernel():
            gpu = build("cuda", math)
            l0 = launches(gpu)
            gpu.step(30.0, 100)
            assert launches(gpu) - l0 == 1
        for name in FIELDS + ("pressure_head", "water_table"):
            x, y = getattr(gpu.state, name).numpy(), getattr(cpu.state, name).numpy()
            assert max_scaled_err(x, y) <= 1.0e-10, (nz, math, name, max_scaled_err(x, y))


def test_more_than_31_layers_run_the_streaming_kernels():
    n = 16
    with warp_kernel():
        a = synthetic_soil_case("cuda", n, nz=40, math="fast")
        l0 = launches(a)
        a.step(60.0, 5)
        assert launches(a) - l0 == 5
    cpu = synthetic_soil_case("oracle", n, nz=40)
    cpu.step(60.0, 5)
    assert max_scaled_err(a.state.temperature.numpy(), cpu.state.temperature.numpy()) <= 1e-11


def test_default_hydraulics_brooks_corey_linear():
    """The reference's default soil hydraulics (Brooks-Corey, lambda = 0.2, linear conductivity) -- the soil of its own benchmark
    (test/benchmarks/gpu/soil_heat_hydrology_global.jl) -- has a compile-time instantiation in fast math."""
    n = 130

    def build(engine, nf, math="faithful"):
        lat, lon, T0 = synthetic_columns(n)
        grid = trm.ColumnGrid(trm.B200(), nf, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=30), n)
        hyd = trm.SoilHydrology(trm.RichardsEq())
        model = trm.SoilModel(grid, soil=trm.SoilEnergyWaterCarbon(hydrology=hyd))
        bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0))
        inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z,
                 "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
        return make(engine, model, trm.ForwardEuler(dt=60.0), boundary_conditions=bcs, initializers=inits, math=math)

    # (40 steps: with this steep retention curve the explicit scheme itself loses stability on the 5 cm top layers after ~60
    #  steps of 60 s from this unsaturated profile -- in the oracle exactly as on the GPU, see tests/test_f32x2.py)
    for nf, tol in ((np.float64, 1.0e-9), (np.float32, 5.0e-5)):
        cpu = build("oracle", nf)
        cpu.step(60.0, 40)
        for math in ("faithful", "fast"):
            with warp_kernel():
                gpu = build("cuda", nf, math)
                gpu.step(60.0, 40)
            for name in FIELDS + ("pressure_head",):
                x, y = getattr(gpu.state, name).numpy(), getattr(cpu.state, name).numpy()
                assert max_scaled_err(x, y) <= tol, (nf, math, name, max_scaled_err(x, y))


# ---- the parity / known-answer tests of the other modules, replayed through the warp path ----
REPLAYS = [
    (test_parity.test_soil_energy_richards_1000_steps, dict(math="fast", stepper="heun")),
    (test_parity.test_heat_only_bit_exact, dict(nf=np.float64)),
    (test_parity.test_heat_only_bit_exact, dict(nf=np.float32)),
    (test_parity.test_heat_only_sinusoid_1000_steps, dict(math="faithful")),
    (test_parity.test_float32_soil_energy_richards, {}),
    (test_parity.test_negative_saturation_slow_path, {}),
    (test_parity.test_over_saturation_to_surface_excess, {}),
    (test_parity.test_heun_negative_saturation_stage_state, dict(math="faithful")),
    (test_parity.test_heun_negative_saturation_stage_state, dict(math="fast")),
    (test_parity.test_handles_are_independent_and_reusable, {}),
    (test_parity.test_async_pipeline_matches_blocking_calls, {}),
    (test_golden.test_energy_initialize_and_closure, dict(engine="cuda")),
    (test_golden.test_energy_to_temperature_branches, dict(engine="cuda")),
    (test_golden.test_heat_conduction_periodic_upper_bc, dict(engine="cuda", stepper="euler")),
    (test_golden.test_heat_conduction_periodic_upper_bc, dict(engine="cuda", stepper="heun")),
    (test_golden.test_heat_conduction_step_upper_bc, dict(engine="cuda")),
    (test_golden.test_adjust_saturation_profile, dict(engine="cuda")),
    (test_golden.test_richards_saturated_steady_state, dict(engine="cuda")),
    (test_golden.test_richards_variably_saturated, dict(engine="cuda", stepper="euler")),
    (test_golden.test_richards_variably_saturated, dict(engine="cuda", stepper="heun")),
    (test_golden.test_vwc_forcing, dict(engine="cuda")),
    (test_golden.test_euler_and_heun_stage_algebra, dict(engine="cuda")),
    (test_golden.test_run_on_ring_grid, dict(engine="cuda", stepper="heun")),
    (test_golden.test_function_valued_boundary_condition, dict(engine="cuda")),
    (test_global_configs.test_soil_heat_global_n72_parity, dict(math="fast")),
    (test_global_configs.test_soil_energy_richards_n145_parity, dict(stepper=trm.ForwardEuler)),
    (test_global_configs.test_soil_energy_richards_n145_parity, dict(stepper=trm.Heun)),
    (test_global_configs.test_quick_start_parity, dict(freeze_thaw=True, math="fast")),
    (test_global_configs.test_quick_start_parity, dict(freeze_thaw=False, math="faithful")),
    (test_global_configs.test_speedy_dry_land_coupling_flow, dict(engine="cuda")),
    (test_raster_input.test_time_varying_raster_drives_the_surface_temperature, dict(engine="cuda")),
    (test_raster_input.test_raster_forcing_parity_with_oracle, {}),
    (test_host_io.test_mapped_host_exchange_equals_copies, dict(heun=False, math="fast", nf=np.float64)),
    (test_host_io.test_mapped_host_exchange_equals_copies, dict(heun=True, math="faithful", nf=np.float32)),
    # LandModel: surface_kernel + one warp launch per step (both Heun stages in it; vegetated: stage-2 surface launch after it)
    (test_parity.test_land_model_1000_steps, dict(math="faithful", stepper="euler")),
    (test_parity.test_land_model_1000_steps, dict(math="fast", stepper="heun")),
    (test_parity.test_land_model_immobile_water_1000_steps, dict(math="fast", stepper="heun")),
    (test_parity.test_land_model_immobile_water_1000_steps, dict(math="faithful", stepper="euler")),
    (test_parity.test_land_model_baseline_forcing_until_blow_up, dict(math="fast", stepper="heun")),
    (test_golden.test_land_model_coupling_signs, dict(engine="cuda")),
    (test_golden.test_runoff_and_infiltration, dict(engine="cuda")),
    (test_golden.test_land_model_default_soil_is_immobile_water, dict(engine="cuda", stepper="heun")),
    (test_golden.test_prescribed_seb_schemes_parity, dict(math="fast")),
    (test_vegetation.test_vegetated_land_parity_1000_steps, dict(math="fast", heun=False)),
    (test_vegetation.test_vegetated_land_parity_1000_steps, dict(math="fast", heun=True)),
    (test_vegetation.test_vegetated_land_parity_1000_steps, dict(math="faithful", heun=True)),
    (test_vegetation.test_vegetated_land_default_parameters_parity, dict(heun=True)),
    (test_vegetation.test_vegetated_land_f32_and_noflow_soil, {}),
    (test_vegetation.test_vegetated_land_layer_counts_and_ragged_columns, dict(nz=7, ncol=129)),
    (test_vegetation.test_vegetated_land_layer_counts_and_ragged_columns, dict(nz=2, ncol=1)),
    (test_vegetation.test_user_write_between_steps_refreshes_the_soil_moisture_factor, {}),
    (test_vegetation.test_coupled_vegetation_soil_step, dict(engine="cuda", stepper="heun")),
    (test_vegetation.test_land_model_energy_and_water_budgets_close, dict(engine="cuda", vegetated=True)),
    (test_vegetation.test_land_model_energy_and_water_budgets_close, dict(engine="cuda", vegetated=False)),
    (test_host_io.test_mapped_host_exchange_land_model, {}),
    # randomised regimes (frozen / thawing / saturated / dry layers next to each other in one column = in one warp)
    (test_parity.test_fast_math_regimes_randomised, dict(seed=0)),
    (test_parity.test_fast_math_regimes_randomised, dict(seed=1)),
    (test_parity.test_fast_math_regimes_randomised, dict(seed=2)),
    (test_parity.test_fast_math_regimes_randomised, dict(seed=3)),
    (test_parity.test_fast_math_regimes_randomised, dict(seed=4)),
    (test_parity.test_fast_math_regimes_randomised, dict(seed=5)),
]


@pytest.mark.parametrize("fn,kw", REPLAYS, ids=[f"{f.__name__}-{'-'.join(str(getattr(v, '__name__', v)) for v in k.values())}" for f, k in REPLAYS])
def test_replay_through_the_warp_kernel(fn, kw):
    with warp_kernel():
        fn(**kw)


@pytest.mark.parametrize("vegetated", [False, True], ids=["bare", "vegetated"])
@pytest.mark.parametrize("heun", [False, True], ids=["euler", "heun"])
def test_land_model_warp_equals_streaming(heun, vegetated):
    """LandModel: surface_kernel + ONE warp launch per step (Heun: both stages in it, the stage-2 surface launch of the vegetated
    model after it) against the streaming kernels (two stage launches per Heun step)."""
    from common import synthetic_land_case
    n = 130

    def build():
        if vegetated:
            return test_vegetation.synthetic_vegetated_case("cuda", n, heun=heun, math="fast")
        return synthetic_land_case("cuda", n, heun=heun, math="fast", windspeed=0.5)

    a, b = build(), build()
    l0 = launches(a)
    with warp_kernel():
        a.step(60.0, 150)
    # per step: surface launch + warp launch (+ the stage-2 surface launch of the vegetated model under Heun)
    assert launches(a) - l0 == 150 * (2 + (1 if heun and vegetated else 0)) + (1 if vegetated else 0)
    b.step(60.0, 150)
    names = RFIELDS + ("skin_temperature", "ground_heat_flux", "infiltration")
    if vegetated:
        names += ("carbon_vegetation", "canopy_water", "vegetation_area_fraction", "soil_moisture_limiting_factor", "transpiration")
    for name in names:
        x, y = getattr(a.state, name).numpy(), getattr(b.state, name).numpy()
        assert np.all(np.isfinite(x)), name
        assert max_scaled_err(x, y) <= 1.0e-10, (name, max_scaled_err(x, y))


def _variant_cases():
    for m in test_parity.test_configuration_variants.pytestmark:
        if m.name == "parametrize" and m.args[0] == "kw":
            return [c if isinstance(c, dict) else c.values[0] for c in m.args[1]]
    raise AssertionError("test_configuration_variants lost its parametrisation")


@pytest.mark.parametrize("math", ["faithful", "fast"])
@pytest.mark.parametrize("kw", _variant_cases(), ids=lambda k: "-".join(f"{a}{getattr(b, '__name__', b)}" for a, b in k.items()))
def test_configuration_variants_through_the_warp_kernel(kw, math):
    """Every soil configuration variant of tests/test_parity.py (layer counts -- 64 and 128 layers fall back to the streaming
    kernels --, retention curves, conductivity schemes, boundary condition kinds), both math modes."""
    with warp_kernel():
        test_parity.test_configuration_variants(math=math, kw=kw)
