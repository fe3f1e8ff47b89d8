"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the BASELINE configurations.

Tolerance (BASELINE.json north_star): Float64, max relative error 1e-9 on temperature, internal energy
and saturation after 1000 steps, plus conservation of the column energy and water budgets.  Two measures
are asserted, both at that tolerance: the error against each field's scale, max|a-b| / max|b|, and the
POINTWISE relative error |a-b| / max(|b|, 1e-6 max|b|) -- temperature crosses zero (degC) and the internal
energy changes sign at the freezing point, so the pointwise measure needs a floor; with this one a cell whose
value is a millionth of the field's scale is still held to 1e-9 of its own value (observed: <= 4e-12 after 1000
steps in every configuration).  Heat-only runs with a constant surface temperature involve no
transcendental function and must be BIT-EXACT in the faithful math mode.
"""
import numpy as np
import pytest

from common import (make, max_scaled_err, pointwise_relerr, richards_soil, synthetic_columns, synthetic_land_case,
                    synthetic_soil_case, trm)

pytestmark = pytest.mark.gpu

FIELDS = ("temperature", "internal_energy", "saturation_water_ice", "liquid_water_fraction")
TOL = 1.0e-9


def compare(gpu, cpu, fields, tol, columns=None, pw_tol=None):
    """Both error measures of the module docstring, per field; `columns` restricts the GPU side to a column subset.
    The pointwise measure is asserted for the Float64 tolerances (at max(tol, 1e-9)); a Float32 field has only ~16 ulp
    between the floor of 1e-6 of its scale and the rounding of its largest values, so Float32 runs assert the scaled one."""
    worst = {}
    if pw_tol is None:
        pw_tol = max(tol, TOL) if tol <= 1.0e-8 else None
    for name in fields:
        a, b = getattr(gpu.state, name).numpy(), getattr(cpu.state, name).numpy()
        if columns is not None:
            a = a[..., columns]
        assert np.all(np.isfinite(a)), name
        worst[name] = (max_scaled_err(a, b), pointwise_relerr(a, b) if pw_tol is not None else 0.0)
    bad = {k: v for k, v in worst.items() if not (v[0] <= tol and (pw_tol is None or v[1] <= pw_tol))}
    assert not bad, (bad, tol, pw_tol)
    return worst


@pytest.mark.parametrize("math", ["faithful", "fast"])
@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_soil_energy_richards_1000_steps(math, stepper):
    """BASELINE config 3/5 at a size the oracle finishes in seconds."""
    n = 1536 + 7   # ragged: not a multiple of the block size
    gpu = synthetic_soil_case("cuda", n, heun=stepper == "heun", math=math)
    cpu = synthetic_soil_case("oracle", n, heun=stepper == "heun")
    d0 = gpu.diagnostics()
    for _ in range(4):
        gpu.step(60.0, 250)
        cpu.step(60.0, 250)
        compare(gpu, cpu, FIELDS + ("pressure_head", "surface_excess_water", "water_table"), TOL)
    assert gpu.clock.time == cpu.clock.time == 60000.0
    d1, dc = gpu.diagnostics(), cpu.diagnostics()
    assert d1["nan_count"] == 0
    # budgets: identical to the oracle's, and water is conserved (zero-flux boundaries)
    assert d1["energy"] == pytest.approx(dc["energy"], rel=1e-9)
    assert d1["water"] == pytest.approx(dc["water"], rel=1e-12)
    assert d1["water"] == pytest.approx(d0["water"], rel=1e-10)


@pytest.mark.parametrize("nf", [np.float64, np.float32])
def test_heat_only_bit_exact(nf):
    """NoFlow hydrology, constant per-column surface temperature: no transcendental function on the path,
    so the faithful build must reproduce the oracle bit for bit (both number formats)."""
    n = 700
    lat, lon, T0 = synthetic_columns(n)

    def build(engine):
        grid = trm.ColumnGrid(trm.B200(), nf, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=30), n)
        model = trm.SoilModel(grid)
        bcs = trm.PrescribedSurfaceTemperature("T_ub", T0 + 10.0 * np.sin(-lon))
        inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z, "saturation_water_ice": 1.0}
        return make(engine, model, trm.ForwardEuler(dt=300.0), boundary_conditions=bcs, initializers=inits, math="faithful")

    gpu, cpu = build("cuda"), build("oracle")
    gpu.step(300.0, 1000)
    cpu.step(300.0, 1000)
    for name in ("internal_energy", "temperature", "liquid_water_fraction"):
        a, b = getattr(gpu.state, name).numpy(), getattr(cpu.state, name).numpy()
        assert np.array_equal(a, b), (name, float(np.max(np.abs(a - b))))


@pytest.mark.parametrize("math", ["faithful", "fast"])
def test_heat_only_sinusoid_1000_steps(math):
    """BASELINE config 2 (soil_heat_global): heat only with the sinusoidal surface temperature."""
    gpu = synthetic_soil_case("cuda", 1000, richards=False, math=math)
    cpu = synthetic_soil_case("oracle", 1000, richards=False)
    e0 = gpu.diagnostics()["energy"]
    gpu.step(300.0, 1000)
    cpu.step(300.0, 1000)
    compare(gpu, cpu, FIELDS[:2] + FIELDS[3:], TOL)
    assert gpu.diagnostics()["energy"] == pytest.approx(cpu.diagnostics()["energy"], rel=1e-9)
    assert np.isfinite(e0)


LAND_FIELDS = FIELDS + ("pressure_head", "skin_temperature", "ground_heat_flux", "latent_heat_flux", "sensible_heat_flux",
                        "infiltration", "surface_runoff", "surface_excess_water")


@pytest.mark.parametrize("math", ["faithful", "fast"])
@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_land_model_1000_steps(math, stepper):
    """BASELINE config 4 (bare ground): surface energy balance + surface hydrology + soil.

    Wind speed 0.5 m/s: with the 3 m/s of BASELINE.md the as-coded sign convention of the ground heat flux
    (G = Rnet - Hs - Hl used as an upward Flux BC) makes the explicit skin/soil coupling
    diverge after ~100 steps of 60 s on the 5 cm top layer -- in the oracle exactly as on the GPU (see the
    60-step test below)."""
    n = 600
    gpu = synthetic_land_case("cuda", n, heun=stepper == "heun", math=math, windspeed=0.5)
    cpu = synthetic_land_case("oracle", n, heun=stepper == "heun", windspeed=0.5)
    for _ in range(4):
        gpu.step(60.0, 250)
        cpu.step(60.0, 250)
        compare(gpu, cpu, LAND_FIELDS, TOL)
    gpu.compute_auxiliary()
    cpu.compute_auxiliary()
    compare(gpu, cpu, ("hydraulic_conductivity", "skin_temperature", "surface_net_radiation", "evaporation_ground",
                       "surface_shortwave_up", "surface_longwave_up"), TOL)


@pytest.mark.parametrize("math", ["faithful", "fast"])
def test_land_model_baseline_forcing_60_steps(math):
    """The BASELINE.md synthetic atmosphere (V = 3 m/s) for the first simulated hour."""
    n = 600
    gpu = synthetic_land_case("cuda", n, math=math)
    cpu = synthetic_land_case("oracle", n)
    gpu.step(60.0, 60)
    cpu.step(60.0, 60)
    compare(gpu, cpu, LAND_FIELDS, TOL)


@pytest.mark.parametrize("math", ["faithful", "fast"])
@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_land_model_baseline_forcing_until_blow_up(math, stepper):
    """BASELINE config 4 with its own forcing (V = 3 m/s) for as long as the as-coded model stays finite. The explicit
    skin / top-layer coupling is unstable at this wind speed (G = Rnet - Hs - Hl used as an upward Flux BC on a 5 cm layer,
    SURVEY.md Appendix C): the oracle's state stops being finite after ~100 steps of 60 s. Both engines are run to the
    last step at which the oracle is finite everywhere; the instability amplifies rounding differences by the same factor
    that it amplifies the state (max |T_skin| grows from 23 to ~800 degC in the last 20 steps), so equality is asserted at
    1e-9 up to ten steps before that point and at 1e-6 at the last finite step; the CUDA path must stay finite as long as
    the oracle does."""
    n = 600
    cpu = synthetic_land_case("oracle", n, heun=stepper == "heun")
    last = 0
    snap = {}
    for i in range(1, 400):
        cpu.step(60.0, 1)
        fin = all(np.isfinite(getattr(cpu.state, f).numpy()).all() for f in ("temperature", "skin_temperature", "internal_energy", "saturation_water_ice"))
        if not fin:
            break
        last = i
    assert 60 < last < 399, last
    cpu = synthetic_land_case("oracle", n, heun=stepper == "heun")
    gpu = synthetic_land_case("cuda", n, heun=stepper == "heun", math=math)
    cpu.step(60.0, last - 10)
    gpu.step(60.0, last - 10)
    compare(gpu, cpu, LAND_FIELDS, TOL)
    cpu.step(60.0, 10)
    gpu.step(60.0, 10)
    assert gpu.clock.iteration == cpu.clock.iteration == last
    compare(gpu, cpu, LAND_FIELDS, 1.0e-6)


def test_float32_soil_energy_richards():
    """Float32 (the reference's default for global grids): rounding differences in powf/cbrtf are amplified
    by the number format; tolerance is a few hundred ulps of the field scale after 200 steps."""
    n = 512
    gpu = synthetic_soil_case("cuda", n, nf=np.float32)
    cpu = synthetic_soil_case("oracle", n, nf=np.float32)
    gpu.step(60.0, 200)
    cpu.step(60.0, 200)
    compare(gpu, cpu, FIELDS, 2.0e-5)


def test_negative_saturation_slow_path():
    """A strong sink drives layers negative: exercises the downward sweep of adjust_saturation_profile!
    (soil_hydrology.jl:201-216), which the kernel handles on its slow path.  One step only: a layer clipped to
    exactly zero saturation has psi_m = -Inf in the reference formulation, and the following step is NaN in the
    reference (and in the oracle) -- the -Inf pattern itself must match."""
    n = 96

    def build(engine, math):
        rng = np.random.default_rng(7)
        grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.1, N=20), n)
        model = trm.SoilModel(grid, soil=richards_soil(vwc_forcing=-2.0e-4))
        sat0 = rng.uniform(0.0, 0.05, (20, n))
        sat0[:, ::3] = 0.9   # every third column stays on the fast path
        return make(engine, model, trm.ForwardEuler(dt=60.0), initializers={"temperature": 5.0, "saturation_water_ice": sat0}, math=math)

    cpu = build("oracle", "faithful")
    cpu.step(60.0, 1)
    sat_c = cpu.state.saturation_water_ice.numpy()
    assert np.min(sat_c) == 0.0 and np.all(np.isfinite(sat_c))
    for math in ("faithful", "fast"):
        gpu = build("cuda", math)
        gpu.step(60.0, 1)
        compare(gpu, cpu, FIELDS + ("water_table", "surface_excess_water"), 1e-12)
        pg, pc = gpu.state.pressure_head.numpy(), cpu.state.pressure_head.numpy()
        assert np.array_equal(np.isneginf(pg), np.isneginf(pc)) and np.any(np.isneginf(pc))
        ok = np.isfinite(pc)
        assert max_scaled_err(pg[ok], pc[ok]) <= 1e-12


def test_over_saturation_to_surface_excess():
    """A strong source fills the column: upward sweep and hand-over to surface_excess_water."""
    n = 64

    def build(engine):
        grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.1, N=12), n)
        model = trm.SoilModel(grid, soil=richards_soil(vwc_forcing=+4.0e-4))
        return make(engine, model, trm.ForwardEuler(dt=60.0), initializers={"temperature": 5.0, "saturation_water_ice": 0.97})

    gpu, cpu = build("cuda"), build("oracle")
    gpu.step(60.0, 10)
    cpu.step(60.0, 10)
    compare(gpu, cpu, FIELDS + ("pressure_head", "water_table", "surface_excess_water"), 1e-12)
    assert np.all(cpu.state.surface_excess_water.numpy() > 0)


def test_full_size_properties():
    """BASELINE config 5 at its full size (10 M columns x 30 layers): properties that do not need the oracle --
    finite fields, water conservation, saturation bounds -- and agreement of the first 4096 columns with an
    oracle run of exactly those columns (columns are independent)."""
    n, sub = 10_000_000, 4096
    gpu = synthetic_soil_case("cuda", n, math="fast")
    d0 = gpu.diagnostics()
    gpu.step(60.0, 50)
    d1 = gpu.diagnostics()
    assert d1["nan_count"] == 0 and d1["ncol"] == n
    assert d1["water"] == pytest.approx(d0["water"], rel=1e-10)
    assert 0.0 <= d1["sat_min"] and d1["sat_max"] <= 1.0
    lat, lon, T0 = synthetic_columns(n)
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=30), sub)
    model = trm.SoilModel(grid, soil=richards_soil())
    bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0[:sub], amp=10.0, phase=lon[:sub], period=86400.0))
    inits = {"temperature": lambda x, z: T0[None, :sub] - 0.05 * z,
             "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
    cpu = make("oracle", model, trm.ForwardEuler(dt=60.0), boundary_conditions=bcs, initializers=inits)
    cpu.step(60.0, 50)
    for name in FIELDS:
        a = getattr(gpu.state, name).numpy()[:, :sub]
        assert max_scaled_err(a, getattr(cpu.state, name).numpy()) <= TOL, name


def test_full_size_strided_parity_1000_steps():
    """BASELINE config 5 at its full size, 1000 steps: eight slabs of 512 columns, one inside each column range of the
    8-GPU partition (unaligned offsets, the last one ending at the last column), against oracle runs of exactly those
    columns -- columns are independent, so the slabs must agree to the tolerance of the small-case tests."""
    n, slab, nranks = 10_000_000, 512, 8
    starts = [r * (n // nranks) + 777 * (r + 1) for r in range(nranks - 1)] + [n - slab]
    cols = np.concatenate([np.arange(s0, s0 + slab) for s0 in starts])
    gpu = synthetic_soil_case("cuda", n, math="fast")
    w0 = gpu.diagnostics()["water"]
    gpu.step(60.0, 1000)
    d1 = gpu.diagnostics()
    assert d1["nan_count"] == 0 and d1["water"] == pytest.approx(w0, rel=1e-10)
    lat, lon, T0 = synthetic_columns(n)
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=30), cols.size)
    model = trm.SoilModel(grid, soil=richards_soil())
    bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0[cols], amp=10.0, phase=lon[cols], period=86400.0))
    inits = {"temperature": lambda x, z: T0[None, cols] - 0.05 * z,
             "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
    cpu = make("oracle", model, trm.ForwardEuler(dt=60.0), boundary_conditions=bcs, initializers=inits)
    cpu.step(60.0, 1000)
    compare(gpu, cpu, FIELDS + ("pressure_head",), TOL, columns=cols)
    compare(gpu, cpu, ("water_table", "surface_excess_water"), TOL, columns=cols)


def test_async_pipeline_matches_blocking_calls():
    """trm_set_input_field_async / trm_step_async / trm_get_field_async (copy streams, double buffered input)
    give the same results as the blocking entry points, step for step."""
    import ctypes as C
    import torch
    n, steps = 5000, 12
    lat, lon, T0 = synthetic_columns(n)

    def build():
        grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=30), n)
        model = trm.SoilModel(grid, soil=richards_soil())
        bcs = trm.PrescribedSurfaceTemperature("T_ub", T0 + 0.0)
        inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z,
                 "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
        return make("cuda", model, trm.ForwardEuler(dt=60.0), boundary_conditions=bcs, initializers=inits, math="fast")

    forcing = [np.ascontiguousarray(T0 + 10.0 * np.sin(2 * np.pi * i * 60.0 / 86400.0 - lon)) for i in range(steps)]
    a, b = build(), build()
    gt_id, in_id = trm.abi.FIELD_IDS["ground_temperature"], a._bc_inputs["T_ub"]
    # blocking reference sequence
    ref = []
    for f in forcing:
        a.state.T_ub.set(f)
        a.step(60.0, 1)
        ref.append(a.state.ground_temperature.numpy())
    # asynchronous pipeline, nothing synchronised until the end
    pin_in = [torch.from_numpy(f).pin_memory() for f in forcing]
    pin_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in forcing]
    lib, h = b._lib, b._h
    for fi, fo in zip(pin_in, pin_out):
        lib.check(lib.set_input_field_async(h, in_id, C.c_void_p(fi.data_ptr())), "set_input_field_async")
        lib.check(lib.step_async(h, 60.0, 1), "step_async")
        lib.check(lib.get_field_async(h, gt_id, C.c_void_p(fo.data_ptr()), n), "get_field_async")
    lib.check(lib.sync(h), "sync")
    for r, o in zip(ref, pin_out):
        assert np.array_equal(r, o.numpy())
    assert np.array_equal(a.state.internal_energy.numpy(), b.state.internal_energy.numpy())
    assert a.clock.time == b.clock.time == steps * 60.0


# ---------------------------------------------------------------------------------------------
# configuration coverage: every process variant / boundary condition kind / size limit of the C ABI
# ---------------------------------------------------------------------------------------------
def _variant_case(engine, math, *, nz, ncol, nf=np.float64, swrc="vg", unsat="vg", n=2.0, heun=False, bcs="value", kernel_dt=60.0):
    rng = np.random.default_rng(11)
    T0 = rng.uniform(-8.0, 12.0, ncol)
    spacing = trm.ExponentialSpacing(dz_min=0.05, dz_max=20.0, N=nz) if nz > 2 else trm.PrescribedSpacing([0.1, 0.2])
    grid = trm.ColumnGrid(trm.B200(), nf, spacing, ncol)
    hp = trm.ConstantSoilHydraulics(
        swrc=trm.VanGenuchten(alpha=2.0, n=n) if swrc == "vg" else trm.BrooksCorey(psi_s=0.05, lam=0.4),
        unsat_hydraulic_cond=trm.UnsatKVanGenuchten() if unsat == "vg" else trm.UnsatKLinear(), sat_hydraulic_cond=2.0e-6)
    soil = trm.SoilEnergyWaterCarbon(hydrology=trm.SoilHydrology(trm.RichardsEq(), hydraulic_properties=hp))
    model = trm.SoilModel(grid, soil=soil)
    if bcs == "value":
        bc = trm.merge_boundary_conditions(trm.PrescribedSurfaceTemperature("T_ub", T0 + 3.0), trm.PrescribedBottomTemperature("T_lb", T0 - 1.0))
    elif bcs == "flux":
        bc = trm.merge_boundary_conditions(trm.GroundHeatFlux(-15.0 + 0 * T0), trm.GeothermalHeatFlux(0.05), trm.InfiltrationFlux(-1.0e-8))
    else:   # gradient on the pressure head at the bottom (FreeDrainage), default elsewhere
        bc = trm.merge_boundary_conditions(trm.FreeDrainage())
    zc = grid.znodes_center().astype(np.float64)
    sat0 = np.clip(0.35 + 0.5 * (zc[:, None] / zc[0]) + 0.1 * rng.uniform(-1, 1, (nz, ncol)), 0.05, 1.0)
    inits = {"temperature": T0[None, :] - 0.02 * zc[:, None], "saturation_water_ice": sat0}
    ts = (trm.Heun if heun else trm.ForwardEuler)(dt=kernel_dt)
    return make(engine, model, ts, boundary_conditions=bc, initializers=inits, math=math)


@pytest.mark.parametrize("kw", [
    dict(nz=2, ncol=1), dict(nz=3, ncol=33), dict(nz=128, ncol=257, kernel_dt=20.0), dict(nz=64, ncol=100, heun=True, kernel_dt=20.0),
    dict(nz=30, ncol=300, swrc="bc", unsat="lin"), dict(nz=30, ncol=300, n=1.6), dict(nz=30, ncol=300, n=3.0, heun=True),
    dict(nz=30, ncol=300, bcs="flux"), dict(nz=30, ncol=300, bcs="flux", heun=True), dict(nz=30, ncol=300, bcs="gradient"),
    dict(nz=30, ncol=300, nf=np.float32),
], ids=lambda k: "-".join(f"{a}{getattr(b, '__name__', b)}" for a, b in k.items()))
@pytest.mark.parametrize("math", ["faithful", "fast"])
def test_configuration_variants(math, kw):
    """Layer counts 2 ... TRM_MAX_NZ, single / ragged column counts, Brooks-Corey + linear K, general van Genuchten n,
    Value / Flux / Gradient boundary conditions, Heun, Float32: 200 steps against the oracle."""
    gpu, cpu = _variant_case("cuda", math, **kw), _variant_case("oracle", math, **kw)
    dt = kw.get("kernel_dt", 60.0)
    gpu.step(dt, 200)
    cpu.step(dt, 200)
    tol = 5.0e-5 if kw.get("nf") is np.float32 else TOL
    compare(gpu, cpu, FIELDS + ("pressure_head", "water_table", "surface_excess_water"), tol)
    gpu.compute_auxiliary(); cpu.compute_auxiliary()
    compare(gpu, cpu, ("hydraulic_conductivity",), tol)


def test_handles_are_independent_and_reusable():
    """Several handles on one device, interleaved stepping, destroy and recreate: results do not depend on it."""
    a = synthetic_soil_case("cuda", 300, math="fast")
    b = synthetic_soil_case("cuda", 300, math="fast")
    c = synthetic_soil_case("cuda", 300, math="fast", heun=True)
    for _ in range(5):
        a.step(60.0, 3); c.step(60.0, 2); b.step(60.0, 3)
    assert np.array_equal(a.state.internal_energy.numpy(), b.state.internal_energy.numpy())
    a.close()
    d = synthetic_soil_case("cuda", 300, math="fast")
    d.step(60.0, 15)
    assert np.array_equal(d.state.saturation_water_ice.numpy(), b.state.saturation_water_ice.numpy())
    with pytest.raises(trm.TerrariumError):
        d._lib.check(d._lib.step(d._h, 60.0, -1), "step")


@pytest.mark.parametrize("case", ["soil-euler", "soil-heun", "land-euler", "land-heun", "heat-euler"])
def test_generic_streaming_kernel(case, monkeypatch):
    """TRM_KERNEL=stream routes every stage through the generic register-streaming kernel (the fallback for fields of
    2^32 elements or more): same parity bar as the default shared-memory staged kernel."""
    monkeypatch.setenv("TRM_KERNEL", "stream")
    kind, stepper = case.split("-")
    heun = stepper == "heun"
    if kind == "land":
        gpu, cpu = synthetic_land_case("cuda", 333, heun=heun, math="fast", windspeed=0.5), synthetic_land_case("oracle", 333, heun=heun, windspeed=0.5)
        fields = LAND_FIELDS
    else:
        rich = kind == "soil"
        gpu, cpu = synthetic_soil_case("cuda", 333, richards=rich, heun=heun, math="fast"), synthetic_soil_case("oracle", 333, richards=rich, heun=heun)
        fields = FIELDS + (("pressure_head", "water_table") if rich else ())
    dt = 60.0 if kind != "heat" else 300.0
    gpu.step(dt, 300)
    cpu.step(dt, 300)
    compare(gpu, cpu, fields, TOL)


def test_ring_grid_scatter_gather():
    """ColumnRingGrid <-> ring grid conversions on the device (column_ring_grid.jl:102-149; reference test
    test/grids.jl:33-139) against the host mirror ColumnRingGrid.to_ring / from_ring, for 3-D, z-face and 2-D fields."""
    rng = np.random.default_rng(3)
    mask = rng.random(4000) > 0.6
    grid = trm.ColumnRingGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_min=0.05, dz_max=10.0, N=9), mask=mask)
    model = trm.SoilModel(grid, soil=richards_soil())
    T0 = rng.uniform(-5, 15, grid.Nc)
    integ = make("cuda", model, trm.ForwardEuler(dt=60.0), initializers={"temperature": T0[None, :] + np.zeros((9, 1)), "saturation_water_ice": 0.6})
    integ.step(60.0, 5)
    integ.compute_auxiliary()
    for name in ("temperature", "hydraulic_conductivity", "water_table"):
        f = getattr(integ.state, name)
        ring = f.to_ring(fill_value=-999.0)
        assert np.array_equal(ring, grid.to_ring(f.numpy(), fill_value=-999.0))
        assert np.all(ring[..., ~mask] == -999.0)
    assert np.all(np.isnan(integ.state.temperature.to_ring()[:, ~mask]))
    # ring -> masked columns
    new_sat = rng.uniform(0.2, 0.9, (9, mask.size))
    integ.state.saturation_water_ice.set_from_ring(new_sat)
    assert np.array_equal(integ.state.saturation_water_ice.numpy(), grid.from_ring(new_sat))
    with pytest.raises(trm.TerrariumError):
        integ._lib.check(integ._lib.get_field_ring(integ._h, 1, new_sat.ctypes.data, 5, 0.0), "get_field_ring")


@pytest.mark.parametrize("math", ["faithful", "fast"])
@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_land_model_immobile_water_1000_steps(math, stepper):
    """LandModel on its default soil (NoFlow hydrology): surface energy balance driving heat conduction with freeze-thaw."""
    n = 500
    gpu = synthetic_land_case("cuda", n, heun=stepper == "heun", math=math, windspeed=0.5, richards=False)
    cpu = synthetic_land_case("oracle", n, heun=stepper == "heun", windspeed=0.5, richards=False)
    fields = FIELDS[:2] + FIELDS[3:] + ("skin_temperature", "ground_heat_flux", "latent_heat_flux", "sensible_heat_flux", "infiltration", "surface_runoff")
    for _ in range(4):
        gpu.step(60.0, 250)
        cpu.step(60.0, 250)
        compare(gpu, cpu, fields, TOL)
    gpu.compute_auxiliary(); cpu.compute_auxiliary()
    compare(gpu, cpu, ("hydraulic_conductivity", "surface_net_radiation", "evaporation_ground"), TOL)


# ---------------------------------------------------------------------------------------------
# fast-math formulas across regimes: the one-root conductivity, the reciprocal-square-root retention curve, clamped seeds
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(6))
def test_fast_math_regimes_randomised(seed):
    """Random columns spanning frozen / thawing / thawed soil, nearly dry to exactly saturated layers, residual water
    content, van Genuchten alpha from 0.3 to 5 per metre and ice impedance 0 .. 9: 150 steps of the fast build against
    the oracle, per-step ulp-level formula differences must stay below the parity bar."""
    rng = np.random.default_rng(100 + seed)
    ncol, nz = 257, int(rng.integers(5, 40))
    alpha, theta_res, omega = float(rng.uniform(0.3, 5.0)), float(rng.choice([0.0, 0.03, 0.08])), float(rng.uniform(0.0, 9.0))
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_min=0.05, dz_max=10.0, N=nz), ncol)
    hp = trm.ConstantSoilHydraulics(swrc=trm.VanGenuchten(alpha=alpha, n=2.0, theta_res=theta_res),
                                    unsat_hydraulic_cond=trm.UnsatKVanGenuchten(impedance=omega), sat_hydraulic_cond=float(rng.uniform(1e-7, 5e-6)))
    soil = trm.SoilEnergyWaterCarbon(hydrology=trm.SoilHydrology(trm.RichardsEq(), hydraulic_properties=hp))
    T0 = rng.uniform(-12.0, 12.0, ncol)
    zc = grid.znodes_center().astype(np.float64)
    sat0 = rng.uniform(0.02, 1.3, (nz, ncol)).clip(max=1.0)           # about a quarter of the cells exactly saturated
    T_init = T0[None, :] + rng.uniform(-2.0, 2.0, (nz, ncol))
    T_ub = T0 + (4.0 if seed % 2 else -4.0)
    pair = []
    for engine in ("cuda", "oracle"):
        model = trm.SoilModel(grid, soil=soil)
        integ = make(engine, model, trm.ForwardEuler(dt=20.0), boundary_conditions=trm.PrescribedSurfaceTemperature("T_ub", T_ub),
                     initializers={"temperature": T_init, "saturation_water_ice": sat0}, math="fast")
        pair.append(integ)
    for integ in pair:
        integ.step(20.0, 150)
    compare(pair[0], pair[1], FIELDS + ("water_table",), TOL)
    # the matric head psi_m = -(1/alpha) sqrt(se^-2 - 1) loses digits to the cancellation in 1 - se^2 as se -> 1 in ANY
    # implementation: with se = 1 - d one ulp of se moves psi_m by 1.1e-16 / (2 d) of its value (5e-7 at d = 1e-10, which
    # these random profiles contain). Against the field's scale that is invisible; pointwise it is bounded at 1e-6 here.
    compare(pair[0], pair[1], ("pressure_head",), TOL, pw_tol=1.0e-6)
    # (the surface excess is zero or rounding noise of a top layer sitting at saturation here: absolute comparison)
    assert np.max(np.abs(pair[0].state.surface_excess_water.numpy() - pair[1].state.surface_excess_water.numpy())) <= 1.0e-13
    pair[0].compute_auxiliary(); pair[1].compute_auxiliary()
    # K(x) = K_sat sqrt(x) (1 - sqrt(1 - x^(2/3)))^2 has an unbounded derivative at saturation: one ulp in x^(2/3) next to
    # x = 1 moves K by 1.5e-8 K_sat in the reference formula itself (1 - sqrt(1.1e-16) vs 1 - 0), whatever computes the
    # power. Cells whose liquid saturation sits within a few ulp of 1 are therefore compared at 1e-7; all others at the bar.
    Kc, Ko = pair[0].state.hydraulic_conductivity.numpy(), pair[1].state.hydraulic_conductivity.numpy()
    assert np.max(np.abs(Kc - Ko)) <= 1.0e-7 * np.max(np.abs(Ko))
    x = (pair[1].state.saturation_water_ice.numpy() * pair[1].state.liquid_water_fraction.numpy())
    near = (x > 1.0 - 1.0e-12) & (x < 1.0)
    face_near = np.zeros(Ko.shape, dtype=bool)
    face_near[:-1] |= near; face_near[1:] |= near          # a face takes the smaller of the two adjacent cell values
    ok = ~face_near
    assert np.max(np.abs(Kc - Ko)[ok]) <= TOL * np.max(np.abs(Ko))


def test_debug_nancheck(monkeypatch):
    """TERRARIUM_DEBUG=true (src/diagnostics/debugging.jl): a NaN in the state makes trm_step fail instead of continuing."""
    monkeypatch.setenv("TERRARIUM_DEBUG", "true")
    integ = synthetic_soil_case("cuda", 40, math="fast")
    integ.step(60.0, 2)
    U = integ.state.internal_energy.numpy()
    U[3, 5] = np.nan
    integ.state.internal_energy.set(U)
    with pytest.raises(trm.TerrariumError, match="TERRARIUM_DEBUG"):
        integ.step(60.0, 1)


@pytest.mark.parametrize("math", ["faithful", "fast"])
def test_heun_negative_saturation_stage_state(math):
    """Heun with a strong sink: the STAGE state of two thirds of the columns goes negative and needs the downward sweep of
    adjust_saturation_profile! (soil_hydrology.jl:201-216). Under the recompute protocol of the Float64 kernels stage 1 then
    stores the stage state of exactly those columns and flags them; every other column is rebuilt by stage 2 from the base
    state and k1. The sweep leaves the deficient layer at exactly zero saturation, whose matric head is -Inf in the reference
    formulation, so the stage-2 Darcy fluxes -- and with them the new saturation -- of such a column are NaN in the
    reference (and in the oracle); its internal energy stays finite and depends on the swept stage saturation through the
    thermal conductivity and the heat capacity. Asserted: the same columns break, everything finite agrees to 1e-12."""
    n = 96

    def build(engine):
        rng = np.random.default_rng(7)
        grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.1, N=20), n)
        model = trm.SoilModel(grid, soil=richards_soil(vwc_forcing=-1.0e-4))
        sat0 = rng.uniform(0.004, 0.02, (20, n))
        sat0[:, ::3] = 0.9   # every third column stays on the fast path
        return make(engine, model, trm.Heun(dt=60.0), initializers={"temperature": 5.0, "saturation_water_ice": sat0}, math=math)

    cpu, gpu = build("oracle"), build("cuda")
    cpu.step(60.0, 1); gpu.step(60.0, 1)
    sat_c = cpu.state.saturation_water_ice.numpy()
    assert (~np.isfinite(sat_c)).any(axis=0).sum() == 64 and np.isfinite(sat_c[:, ::3]).all()
    broken = (~np.isfinite(sat_c)).any(axis=0)            # columns whose new saturation is NaN in the reference formulation
    sat_g = gpu.state.saturation_water_ice.numpy()
    assert np.array_equal((~np.isfinite(sat_g)).any(axis=0), broken)
    assert max_scaled_err(sat_g[:, ~broken], sat_c[:, ~broken]) <= 1e-12
    Ug, Uc = gpu.state.internal_energy.numpy(), cpu.state.internal_energy.numpy()     # finite in every column
    assert np.isfinite(Uc).all() and max_scaled_err(Ug, Uc) <= 1e-12
    # (water table and surface excess water of a broken column are outcomes of arithmetic on NaN: compared elsewhere)
    for name in ("water_table", "surface_excess_water"):
        assert np.array_equal(getattr(gpu.state, name).numpy()[~broken], getattr(cpu.state, name).numpy()[~broken]), name
