"""Vegetated LandModel: PALADYN vegetation + canopy hydrology.

Known-answer / property tests of the reference (test/vegetation/*.jl, test/surface_hydrology/canopy_*_tests.jl,
test/coupled_models/land_model_tests.jl:42-76) re-encoded as whole-model evaluations on BOTH engines, an independent
numpy restatement of the per-column formulas for the oracle, and CUDA-vs-oracle parity (``-m gpu``, through the C ABI).
"""
import numpy as np
import pytest

from common import CUDA_MATH, ENGINES, make, max_scaled_err, richards_soil, synthetic_columns, trm

VEG_2D = ["carbon_vegetation", "vegetation_area_fraction", "canopy_water", "balanced_leaf_area_index", "leaf_area_index",
          "phenology_factor", "canopy_water_conductance", "leaf_to_air_co2_ratio", "net_assimilation", "leaf_respiration",
          "gross_primary_production", "autotrophic_respiration", "net_primary_production", "soil_moisture_limiting_factor",
          "canopy_water_interception", "canopy_water_removal", "saturation_canopy_water", "rainfall_ground",
          "evaporation_canopy", "transpiration", "evaporation_ground", "latent_heat_flux", "ground_heat_flux", "skin_temperature",
          "infiltration", "surface_runoff"]


def veg_land(engine, ncol=4, nf=np.float64, N=50, ts=None, inputs=None, inits=None, vegetation=None, math="faithful", soil=None, grid=None):
    grid = grid or trm.ColumnGrid(trm.B200(), nf, trm.ExponentialSpacing(dz_max=1.0, N=N), ncol)
    land = trm.LandModel(grid, soil=soil or richards_soil(), vegetation=vegetation or trm.VegetationCarbon())
    base = {"temperature": lambda x, z: 5.0 - 0.02 * z + 0 * x, "saturation_water_ice": lambda x, z: np.minimum(1, 0.8 - 0.05 * z) + 0 * x,
            "carbon_vegetation": 0.1}
    base.update(inits or {})
    return make(engine, land, ts or trm.ForwardEuler(dt=60.0), inputs, initializers=base, math=math)


# ---------------------------------------------------------------------------------------------
# host mirror: defaults of LandModel(grid) / LandModel(grid; vegetation = nothing), land_model.jl:26,111-125
def test_land_model_defaults():
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_max=1.0, N=10), 1)
    veg = trm.LandModel(grid)
    assert isinstance(veg.vegetation, trm.VegetationCarbon)
    assert isinstance(veg.surface_hydrology.evapotranspiration, trm.PALADYNCanopyEvapotranspiration)
    assert isinstance(veg.surface_hydrology.canopy_interception, trm.PALADYNCanopyInterception)
    assert isinstance(veg.soil.hydrology.vertical_flow, trm.RichardsEq)
    bare = trm.LandModel(grid, vegetation=None)
    assert isinstance(bare.surface_hydrology.evapotranspiration, trm.BareGroundEvaporation)
    assert isinstance(bare.surface_hydrology.canopy_interception, trm.NoCanopyInterception)
    assert isinstance(bare.soil.hydrology.vertical_flow, trm.NoFlow)


# test/coupled_models/land_model_tests.jl:42-76 : one 60 s step of the coupled vegetation-soil model stays finite
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("stepper", ["euler", "heun"])
def test_coupled_vegetation_soil_step(engine, stepper):
    ts = trm.ForwardEuler(dt=60.0) if stepper == "euler" else trm.Heun(dt=60.0)
    integ = veg_land(engine, ts=ts)
    trm.timestep(integ, 60.0)
    for name in ("saturation_water_ice", "internal_energy", "ground_heat_flux", "carbon_vegetation"):
        assert np.all(np.isfinite(getattr(integ.state, name).numpy())), name
    # the latent heat flux follows the summed humidity flux of the canopy scheme (turbulent_fluxes.jl:137-150)
    st = integ.state
    Qh = st.evaporation_ground.numpy() + st.evaporation_canopy.numpy() + st.transpiration.numpy()
    np.testing.assert_allclose(st.latent_heat_flux.numpy(), 2.257e6 * 1.293 * Qh, rtol=1e-12)


# test/vegetation/root_distribution_tests.jl: the root fractions sum to one over the column; plant_available_water_tests.jl
@pytest.mark.parametrize("engine", ENGINES)
def test_root_fraction_and_plant_available_water(engine):
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.2, N=10), 3)
    strat = trm.HomogeneousStratigraphy(porosity=trm.ConstantSoilPorosity(mineral_porosity=0.5))
    for sat, T0, expect in ((1.0, 10.0, 1.0), (0.0, 10.0, 0.0), (1.0, -10.0, 0.0), (0.2, 10.0, 0.25)):
        soil = richards_soil(strat=strat)
        integ = veg_land(engine, grid=grid, soil=soil, inits={"temperature": T0, "saturation_water_ice": sat})
        integ.compute_auxiliary()
        rf = integ.state.root_fraction.numpy()
        np.testing.assert_allclose(rf.sum(axis=0), 1.0, rtol=1e-12)
        assert np.all(rf > 0) and np.all(np.diff(rf[:, 0]) > 0)   # uniform grid: density decays with depth (layer 0 = bottom)
        paw = integ.state.plant_available_water.numpy()
        np.testing.assert_allclose(paw, expect, atol=1e-12)
        np.testing.assert_allclose(integ.state.soil_moisture_limiting_factor.numpy(), (paw * rf).sum(axis=0), rtol=1e-12, atol=1e-15)


# independent numpy restatement of the per-column formulas (photosynthesis.jl, stomatal_conductance.jl,
# autotrophic_respiration.jl, carbon_dynamics.jl, vegetation_dynamics.jl, canopy_interception.jl, canopy_evapotranspiration.jl)
def _esat(T):
    return np.where(T <= 0, 611.0 * np.exp(22.46 * T / (T + 272.62)), 611.0 * np.exp(17.62 * T / (T + 243.12)))


def reference_column(Ta, sw, pres, q, V, rain, co2, SAI, Rdl, Cv, nu, w, An_prev, beta, Ts, Tg):
    eps = 0.622
    ea = q * pres / (eps + (1 - eps) * q)
    vpd = np.maximum(_esat(Ta) - ea, 0.1)
    LAIb = Cv / (2.0 / 10.0 + 2.0)
    LAI = LAIb
    gw = 0.5 / 1000 * (1 - np.exp(-0.5 * LAI)) * beta + 1.6 * (1 + 2.3 / np.sqrt(vpd)) * An_prev / co2 * 1e6
    lamc = 1 - 1 / (1 + 2.3 / np.sqrt(vpd * 1e-3))
    pO2, pa = 0.209 * pres, co2 * 1e-6 * pres
    ex = (Ta - 25.0) * 0.1
    tau, Kc, Ko = 2600.0 * 0.57 ** ex, 30.0 * 2.1 ** ex, 3.0e4 * 1.2 ** ex
    Gs = pO2 / (2 * tau)
    PAR = 0.5 * sw * (1 - 0.17) * 4.6e-6
    APAR = 0.5 * PAR * (1 - np.exp(-0.5 * LAI))
    pi = lamc * pa
    k1, k2, k3 = 2 * np.log(1 / 0.99 - 1) / (-4.0 - 15.0), 0.5 * (-4.0 + 15.0), np.log(0.99 / 0.01) / (42.0 - 30.0)
    Tst = np.where((Ta > -4.0) & (Ta < 42.0), 1 / (1 + np.exp(k1 * (k2 - Ta))) * (1 - 0.01 * np.exp(k3 * (Ta - 30.0))), 0.0)
    c1 = 0.08 * Tst * 12.0 * (pi - Gs) / (pi + 2 * Gs)
    c2 = (pi - Gs) / (pi + Kc * (1 + pO2 / Ko))
    Vc = c1 * APAR * (pi + Kc * (1 + pO2 / Ko)) / (pi - Gs)
    Rd = 0.08 * Vc * beta
    JE, JC = c1 * APAR, c2 * Vc
    Ag = (JE + JC - np.sqrt((JE + JC) ** 2 - 4 * 0.7 * JE * JC)) / (2 * 0.7) * beta
    on = (sw > 0) & (Ta > -3.0) & (LAI > 0)
    An = np.where(on, Ag - Rd, 0.0)
    Rd = np.where(on, Rd, 0.0)
    GPP = An * 1e-3
    ft = lambda T: np.exp(308.56 * (1 / 56.02 - 1 / (46.02 + T)))
    Rm = Rdl / 1000 + 0.066 * ft(Ta) * (2.0 * (0.2 + 2.0)) / (Cv * 10.0 * 330.0) + 0.066 * np.where(Tg > 7, ft(Tg), 0.0) * 1.0 * 0.2 / (10.0 * Cv * 29.0)
    Ra = Rm + 0.25 * (GPP - Rm)
    NPP = GPP - Ra
    wmax = 2.0e-4 * (LAI + SAI)
    fcan = np.where(wmax > 0, w / np.where(wmax > 0, wmax, 1.0), 0.0)
    Ican = 0.2 * rain * (1 - np.exp(-0.5 * (LAI + SAI)))
    Rcan = np.maximum(w, 0) / 86400.0
    Vc_ = np.maximum(V, 0.01)
    ra = 1 / (1.2e-3 * np.maximum(Vc_, 1e-6))
    dqs = eps * np.maximum(_esat(Ts) - ea, 0.1) / pres
    dqg = eps * np.maximum(_esat(Tg) - ea, 0.1) / pres
    re = (1 - np.exp(-LAI - SAI)) / (0.006 * Vc_)
    transp = dqs / (ra + 1 / np.maximum(gw, np.sqrt(np.finfo(np.float64).eps)))
    Egnd = 1.0 * dqg / (ra + re)
    Ecan = fcan * dqs / ra
    lam = np.where(LAIb < 1.0, 0.0, np.where(LAIb <= 6.0, (LAIb - 1.0) / 5.0, 1.0))
    dC = (1 - lam) * NPP - (0.3 / 10 + 0.3 / 10 + 0.05 * 2.0) * LAIb
    nus = np.maximum(nu, 0.001)
    dnu = (lam * NPP / Cv) * nus * (1 - nu) - 0.002 * nus
    dw = Ican - Ecan - Rcan
    return dict(balanced_leaf_area_index=LAIb, leaf_area_index=LAI, canopy_water_conductance=gw, leaf_to_air_co2_ratio=lamc,
                net_assimilation=An, leaf_respiration=Rd, gross_primary_production=GPP, autotrophic_respiration=Ra,
                net_primary_production=NPP, saturation_canopy_water=fcan, canopy_water_interception=Ican, canopy_water_removal=Rcan,
                rainfall_ground=rain - Ican + Rcan, transpiration=transp, evaporation_ground=Egnd, evaporation_canopy=Ecan,
                dC=dC, dnu=dnu, dw=dw)


def _varied_inputs(n):
    rng = np.random.default_rng(7)
    return dict(air_temperature=rng.uniform(-8.0, 45.0, n), surface_shortwave_down=np.where(np.arange(n) % 5 == 0, 0.0, rng.uniform(10.0, 800.0, n)),
                air_pressure=rng.uniform(9.0e4, 1.02e5, n), specific_humidity=rng.uniform(1e-3, 8e-3, n), windspeed=rng.uniform(0.0, 6.0, n),
                rainfall=rng.uniform(0.0, 5e-7, n), CO2=rng.uniform(280.0, 600.0, n), SAI=rng.uniform(0.0, 1.5, n),
                daily_leaf_respiration=rng.uniform(0.0, 1e-3, n), surface_longwave_down=300.0)


@pytest.mark.parametrize("engine", ENGINES)
def test_vegetation_auxiliaries_against_formulas(engine):
    n = 64
    rng = np.random.default_rng(11)
    inp = _varied_inputs(n)
    Cv, nu, w = rng.uniform(0.2, 14.0, n), rng.uniform(0.0, 0.9, n), rng.uniform(0.0, 3e-4, n)
    Cv[3] = 0.0; Cv[4] = 30.0   # LAI = 0 (no photosynthesis, Inf respiration as coded) and LAI_b > LAI_max
    An0 = rng.uniform(0.0, 2e-4, n)
    Ts0 = rng.uniform(-5.0, 25.0, n)
    integ = veg_land(engine, ncol=n, inputs=inp, inits={"carbon_vegetation": Cv, "vegetation_area_fraction": nu, "canopy_water": w,
                                                         "net_assimilation": An0, "skin_temperature": Ts0,
                                                         "temperature": lambda x, z: Ts0[None, :] * 0 + np.linspace(-2.0, 12.0, n)[None, :] - 0.02 * z})
    # one ForwardEuler step without finalize: the auxiliaries hold the evaluation on the state BEFORE the step
    integ.step(60.0, 1)
    st = integ.state
    Tg = np.linspace(-2.0, 12.0, n) - 0.02 * integ.grid.znodes_center()[-1]
    beta = st.soil_moisture_limiting_factor.numpy()
    ref = reference_column(inp["air_temperature"], inp["surface_shortwave_down"], inp["air_pressure"], inp["specific_humidity"], inp["windspeed"],
                           inp["rainfall"], inp["CO2"], inp["SAI"], inp["daily_leaf_respiration"], Cv, nu, w, An0, beta, Ts0, Tg)
    with np.errstate(all="ignore"):
        for name, val in ref.items():
            if name in ("dC", "dnu", "dw"):
                continue
            got = getattr(st, name).numpy()
            ok = np.isfinite(val)
            np.testing.assert_allclose(got[ok], val[ok], rtol=2e-10, atol=1e-300, err_msg=name)
            assert np.array_equal(np.isfinite(got), ok), name
        fin = np.isfinite(ref["dC"]) & np.isfinite(ref["dnu"])
        np.testing.assert_allclose(st.carbon_vegetation.numpy()[fin], (Cv + 60.0 * ref["dC"])[fin], rtol=1e-9)
        np.testing.assert_allclose(st.vegetation_area_fraction.numpy()[fin], (nu + 60.0 * ref["dnu"])[fin], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(st.canopy_water.numpy(), w + 60.0 * ref["dw"], rtol=1e-9, atol=1e-14)
    # properties the reference tests assert (photosynthesis_tests.jl, stomatal_conductance_tests.jl, canopy_interception_tests.jl)
    An, Rd = st.net_assimilation.numpy(), st.leaf_respiration.numpy()
    dark = (inp["surface_shortwave_down"] == 0) | (inp["air_temperature"] <= -3.0) | (Cv == 0)
    assert np.all(An[dark] == 0) and np.all(Rd[dark] == 0)
    hot = (inp["air_temperature"] >= 42.0) & ~dark   # T_stress = 0 outside (T_CO2_low, T_CO2_high)
    assert np.all(An[hot] == 0)
    lam = st.leaf_to_air_co2_ratio.numpy()
    assert np.all((lam > 0) & (lam < 1))
    assert np.all(st.phenology_factor.numpy() == 1.0)
    assert np.all(st.canopy_water_interception.numpy() <= 0.2 * inp["rainfall"] + 1e-30)


# ---------------------------------------------------------------------------------------------
# CUDA path against the oracle
def _stable_vegetation():
    """The as-coded turnover rates (1/year used per second) empty the carbon pool within a minute; the long parity run uses
    rates that keep the state in a physical range so that every branch keeps being exercised."""
    return trm.VegetationCarbon(carbon_dynamics=trm.PALADYNCarbonDynamics(gamma_L=1e-9, gamma_R=1e-9, gamma_S=1e-10),
                                vegetation_dynamics=trm.PALADYNVegetationDynamics(gamma_v_min=1e-8))


def synthetic_vegetated_case(engine, ncol, nf=np.float64, heun=False, nz=30, math="faithful", dt=60.0, vegetation=None):
    """BASELINE config 4 with vegetation: LandModel + VegetationCarbon under the synthetic atmosphere of BASELINE.md section 5."""
    lat, lon, T0 = synthetic_columns(ncol)
    rng = np.random.default_rng(5)
    grid = trm.ColumnGrid(trm.B200(), nf, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=nz), ncol)
    land = trm.LandModel(grid, soil=richards_soil(), vegetation=vegetation or _stable_vegetation())
    day = 86400.0
    hours = np.arange(0, 73, dtype=np.float64)
    rain = np.where((hours % 24) < 6, 2.0e-8, 0.0)
    inputs = {
        "air_temperature": trm.Sinusoid(mean=T0, amp=8.0, phase=lon, period=day),
        "surface_shortwave_down": trm.Sinusoid(mean=0.0, amp=600.0, phase=lon, period=day, lo=0.0),
        "surface_longwave_down": 300.0, "specific_humidity": 0.005, "air_pressure": 101325.0, "windspeed": 0.5,
        "rainfall": trm.TimeSeries(hours * 3600.0, np.repeat(rain[:, None], ncol, axis=1)),
        "SAI": rng.uniform(0.1, 1.0, ncol), "CO2": 400.0,
    }
    inits = {
        "temperature": lambda x, z: T0[None, :] - 0.05 * z,
        "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x,
        "skin_temperature": T0, "carbon_vegetation": rng.uniform(6.0, 14.0, ncol), "vegetation_area_fraction": rng.uniform(0.05, 0.8, ncol),
    }
    ts = (trm.Heun if heun else trm.ForwardEuler)(dt=dt)
    return make(engine, land, ts, inputs, initializers=inits, math=math)


def _compare(a, b, names3=("temperature", "internal_energy", "saturation_water_ice", "liquid_water_fraction"), tol=1e-9, names2=VEG_2D):
    for name in names3:
        assert max_scaled_err(getattr(a.state, name).numpy(), getattr(b.state, name).numpy()) <= tol, name
    for name in names2:
        x, y = getattr(a.state, name).numpy(), getattr(b.state, name).numpy()
        assert np.array_equal(np.isfinite(x), np.isfinite(y)), name
        ok = np.isfinite(y)
        if ok.any() and np.max(np.abs(y[ok])) > 0:
            assert max_scaled_err(x[ok], y[ok]) <= tol, (name, max_scaled_err(x[ok], y[ok]))
        else:
            assert np.all(x[ok] == 0), name


@pytest.mark.gpu
@pytest.mark.parametrize("math", CUDA_MATH)
@pytest.mark.parametrize("heun", [False, True], ids=["euler", "heun"])
def test_vegetated_land_parity_1000_steps(math, heun):
    """FP64, 1000 steps of 60 s (16.7 h: night, sunrise, rain on / off), max error relative to the field scale <= 1e-9."""
    ncol = 96
    cu = synthetic_vegetated_case("cuda", ncol, heun=heun, math=math)
    orc = synthetic_vegetated_case("oracle", ncol, heun=heun)
    for chunk in (1, 9, 290, 700):
        cu.step(60.0, chunk); orc.step(60.0, chunk)
        _compare(cu, orc)
    cu.compute_auxiliary(); orc.compute_auxiliary()
    _compare(cu, orc, names3=("plant_available_water", "hydraulic_conductivity", "pressure_head"))
    C = cu.state.carbon_vegetation.numpy()
    assert np.all(np.isfinite(C)) and np.all(C > 0)
    assert np.any(cu.state.net_assimilation.numpy() > 0) and np.any(cu.state.canopy_water.numpy() > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("heun", [False, True], ids=["euler", "heun"])
def test_vegetated_land_default_parameters_parity(heun):
    """Reference default parameters (fast as-coded carbon turnover): short steps, timestep! with finalize, both steppers."""
    ncol = 40
    cu = synthetic_vegetated_case("cuda", ncol, heun=heun, dt=1.0, vegetation=trm.VegetationCarbon())
    orc = synthetic_vegetated_case("oracle", ncol, heun=heun, dt=1.0, vegetation=trm.VegetationCarbon())
    for _ in range(5):
        trm.timestep(cu, 1.0); trm.timestep(orc, 1.0)   # finalize = true: compute_auxiliary! after every step
        _compare(cu, orc)
    cu.step(1.0, 40); orc.step(1.0, 40)
    _compare(cu, orc)


@pytest.mark.gpu
def test_vegetated_land_f32_and_noflow_soil():
    """Float32 state (tolerance of the number format) and a vegetated LandModel on immobile soil water."""
    ncol = 64
    cu = synthetic_vegetated_case("cuda", ncol, nf=np.float32)
    orc = synthetic_vegetated_case("oracle", ncol, nf=np.float32)
    cu.step(60.0, 50); orc.step(60.0, 50)
    _compare(cu, orc, tol=2e-4)
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_max=1.0, N=20), 8)
    pair = [veg_land(e, grid=grid, soil=trm.SoilEnergyWaterCarbon(), vegetation=_stable_vegetation(), inits={"carbon_vegetation": 4.0, "saturation_water_ice": 0.7})
            for e in ("cuda", "oracle")]
    for it in pair:
        it.step(60.0, 100)
    _compare(pair[0], pair[1], names3=("temperature", "internal_energy"))


@pytest.mark.gpu
@pytest.mark.parametrize("nz,ncol", [(2, 1), (7, 129), (64, 333), (128, 40)])
def test_vegetated_land_layer_counts_and_ragged_columns(nz, ncol):
    """Column counts that do not fill a block / a warp, 2 .. 128 layers (metric rows in the large shared-memory layout),
    Heun + finalize in the middle of the run, the generic streaming kernel (TRM_KERNEL is read at handle creation)."""
    import os
    for kernel in ("smem", "stream"):
        os.environ["TRM_KERNEL"] = kernel
        try:
            cu = synthetic_vegetated_case("cuda", ncol, heun=True, nz=nz, math="fast")
        finally:
            os.environ.pop("TRM_KERNEL", None)
        orc = synthetic_vegetated_case("oracle", ncol, heun=True, nz=nz)
        for it in (cu, orc):
            it.step(60.0, 30)
            it.compute_auxiliary()
            it.step(60.0, 30)
        _compare(cu, orc)


@pytest.mark.gpu
def test_user_write_between_steps_refreshes_the_soil_moisture_factor():
    """set!(saturation_water_ice, ...) between steps: the carried soil moisture limiting factor is recomputed from the
    stored fields before the next surface launch (beta_kernel), like the oracle's full re-evaluation."""
    cu = synthetic_vegetated_case("cuda", 50, math="fast")
    orc = synthetic_vegetated_case("oracle", 50)
    for it in (cu, orc):
        it.step(60.0, 10)
        it.state.saturation_water_ice.set(lambda x, z: np.minimum(1.0, 0.3 - 0.02 * z) + 0 * x)
        it.step(60.0, 10)
    _compare(cu, orc)
    assert np.all(cu.state.soil_moisture_limiting_factor.numpy() < 1.0)


# ground_resistance_factor.jl:32-57 (Lee & Pielke 1992): (1 - cos(pi theta_w / theta_fc))^2 / 4 below field capacity
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("vegetated", [False, True], ids=["bare", "vegetated"])
def test_soil_moisture_resistance_factor(engine, vegetated):
    n = 6
    sat_top = np.array([0.05, 0.2, 0.4, 0.5, 0.6, 1.0])   # theta_w = 0.49 sat: below field capacity (0.25) up to sat = 0.51
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_max=1.0, N=12), n)

    def run(factor):
        if vegetated:
            sh = trm.SurfaceHydrology(evapotranspiration=trm.PALADYNCanopyEvapotranspiration(ground_resistance_factor=factor))
            land = trm.LandModel(grid, soil=richards_soil(), surface_hydrology=sh)
        else:
            sh = trm.SurfaceHydrology(evapotranspiration=trm.BareGroundEvaporation(ground_resistance_factor=factor))
            land = trm.LandModel(grid, soil=richards_soil(), vegetation=None, surface_hydrology=sh)
        inits = {"temperature": 8.0, "saturation_water_ice": lambda x, z: sat_top[None, :] + 0 * z, "carbon_vegetation": 5.0, "skin_temperature": 9.0}
        integ = make(engine, land, trm.ForwardEuler(dt=60.0), {"windspeed": 2.0}, initializers={k: v for k, v in inits.items() if vegetated or k != "carbon_vegetation"})
        integ.step(60.0, 1)
        return integ.state.evaporation_ground.numpy()

    E1, Esm, Ehalf = run(1.0), run(trm.SoilMoistureResistanceFactor()), run(trm.ConstantEvaporationResistanceFactor(0.5))
    thw = 0.49 * sat_top
    beta = np.where(thw < 0.25, (1 - np.cos(np.pi * thw / 0.25)) ** 2 / 4, 1.0)
    assert np.all(E1 > 0)
    np.testing.assert_allclose(Esm, beta * E1, rtol=1e-12)
    np.testing.assert_allclose(Ehalf, 0.5 * E1, rtol=1e-12)
    assert beta[0] < 0.1 and 0.9 < beta[3] < 1.0 and np.all(beta[4:] == 1.0)


# SoilHydraulicsSURFEX (soil_hydraulic_properties.jl:112-156): field capacity / wilting point from the clay content
@pytest.mark.parametrize("engine", ENGINES)
def test_surfex_field_capacity_and_wilting_point(engine):
    clay = 0.3
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.2, N=10), 2)
    hp = trm.SoilHydraulicsSURFEX(swrc=trm.VanGenuchten(alpha=2.0, n=2.0), unsat_hydraulic_cond=trm.UnsatKVanGenuchten())
    soil = trm.SoilEnergyWaterCarbon(hydrology=trm.SoilHydrology(trm.RichardsEq(), hydraulic_properties=hp),
                                     strat=trm.HomogeneousStratigraphy(texture=trm.SoilTexture(sand=0.5, clay=clay)))
    fc, wp = 89.0e-3 * (clay * 100) ** 0.35, 37.13e-3 * np.sqrt(clay * 100)
    assert trm.build_params(trm.LandModel(grid, soil=soil)).field_capacity == pytest.approx(fc, rel=1e-15)
    integ = veg_land(engine, grid=grid, soil=soil, inits={"temperature": 10.0, "saturation_water_ice": 0.5, "carbon_vegetation": 5.0})
    integ.compute_auxiliary()
    want = min(1.0, max(0.0, (0.49 * 0.5 - wp) / (fc - wp)))
    np.testing.assert_allclose(integ.state.plant_available_water.numpy(), want, rtol=1e-12)
    with pytest.raises(ValueError):
        trm.SoilTexture(sand=0.8, clay=0.5)


# ---------------------------------------------------------------------------------------------
# column budgets of the coupled model: what enters through the surface is what the column gains
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("vegetated", [False, True], ids=["bare", "vegetated"])
def test_land_model_energy_and_water_budgets_close(engine, vegetated):
    """ForwardEuler: the column energy changes by -G dt per step (G: ground heat flux, positive upward, Flux BC on the
    internal energy, land_model.jl:56-62) and the column water (soil + surface excess) by porosity * infiltration dt -- as
    coded the infiltration is a Flux BC on the SATURATION (land_model.jl:59-61, soil_model_bcs.jl:29), so a layer's water
    content gains porosity times the flux; the bottom boundary is closed. Evaluated step by step from the fields the
    surface block leaves behind."""
    ncol = 24
    if vegetated:
        integ = synthetic_vegetated_case(engine, ncol, dt=60.0)
    else:
        from common import synthetic_land_case
        integ = synthetic_land_case(engine, ncol, windspeed=0.5)
    dz = np.diff(integ.grid.znodes_face().astype(np.float64))[:, None]

    def budgets():
        U, s = integ.state.internal_energy.numpy(), integ.state.saturation_water_ice.numpy()
        return (U * dz).sum(axis=0), (s * 0.49 * dz).sum(axis=0) + integ.state.surface_excess_water.numpy()

    E0, W0 = budgets()
    dE, dW = np.zeros(ncol), np.zeros(ncol)
    for _ in range(120):   # two hours: rain on, infiltration active
        integ.step(60.0, 1)
        dE -= 60.0 * integ.state.ground_heat_flux.numpy()     # fluxes of the evaluation at the start of the step
        dW += 0.49 * 60.0 * integ.state.infiltration.numpy()
    E1, W1 = budgets()
    assert np.any(np.abs(dW) > 0) and np.all(np.abs(dE) > 0)
    np.testing.assert_allclose(E1 - E0, dE, rtol=1e-9, atol=1e-9 * np.max(np.abs(dE)))
    np.testing.assert_allclose(W1 - W0, dW, rtol=1e-9, atol=5e-12)   # 5e-12 m: rounding of the ~2e2 m column totals


@pytest.mark.gpu
def test_vegetated_land_large_domain_properties():
    """2 M columns (every block and warp shape of the surface / stage launches at scale): finite state, bounded factors,
    water budget closed against the accumulated infiltration, and the first 512 columns equal to an oracle run of the same
    columns (columns are independent)."""
    ncol, sub = 2_000_000, 512
    cu = synthetic_vegetated_case("cuda", ncol, math="fast")
    dz = np.diff(cu.grid.znodes_face().astype(np.float64))[:, None]
    W0 = (cu.state.saturation_water_ice.numpy() * 0.49 * dz).sum(axis=0) + cu.state.surface_excess_water.numpy()
    dW = np.zeros(ncol)
    for _ in range(20):
        cu.step(60.0, 1)
        dW += 0.49 * 60.0 * cu.state.infiltration.numpy()
    W1 = (cu.state.saturation_water_ice.numpy() * 0.49 * dz).sum(axis=0) + cu.state.surface_excess_water.numpy()
    # (the column totals are ~2e2 m of water: their own rounding, ~1e-13 m, bounds how well 1e-5 m of change can be resolved)
    np.testing.assert_allclose(W1 - W0, dW, rtol=1e-9, atol=5e-12)
    st = cu.state
    for name in ("temperature", "internal_energy", "saturation_water_ice", "carbon_vegetation", "canopy_water", "ground_heat_flux", "transpiration"):
        assert np.all(np.isfinite(getattr(st, name).numpy())), name
    b = st.soil_moisture_limiting_factor.numpy()
    assert np.all((b >= 0) & (b <= 1 + 1e-12))
    s = st.saturation_water_ice.numpy()
    assert np.all((s >= 0) & (s <= 1))
    # the same first columns on the oracle, built from slices of the large case's per-column data
    lat, lon, T0 = synthetic_columns(ncol)
    rng = np.random.default_rng(5)
    SAI, C0, nu0 = rng.uniform(0.1, 1.0, ncol), rng.uniform(6.0, 14.0, ncol), rng.uniform(0.05, 0.8, ncol)
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=30), sub)
    land = trm.LandModel(grid, soil=richards_soil(), vegetation=_stable_vegetation())
    hours = np.arange(0, 73, dtype=np.float64)
    rain = np.where((hours % 24) < 6, 2.0e-8, 0.0)
    inputs = {"air_temperature": trm.Sinusoid(mean=T0[:sub], amp=8.0, phase=lon[:sub], period=86400.0),
              "surface_shortwave_down": trm.Sinusoid(mean=0.0, amp=600.0, phase=lon[:sub], period=86400.0, lo=0.0),
              "surface_longwave_down": 300.0, "specific_humidity": 0.005, "air_pressure": 101325.0, "windspeed": 0.5,
              "rainfall": trm.TimeSeries(hours * 3600.0, np.repeat(rain[:, None], sub, axis=1)), "SAI": SAI[:sub], "CO2": 400.0}
    inits = {"temperature": lambda x, z: T0[None, :sub] - 0.05 * z, "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x,
             "skin_temperature": T0[:sub], "carbon_vegetation": C0[:sub], "vegetation_area_fraction": nu0[:sub]}
    orc = make("oracle", land, trm.ForwardEuler(dt=60.0), inputs, initializers=inits)
    orc.step(60.0, 20)
    for name in ("temperature", "saturation_water_ice", "carbon_vegetation", "ground_heat_flux", "transpiration", "net_assimilation"):
        a, b_ = getattr(cu.state, name).numpy()[..., :sub], getattr(orc.state, name).numpy()
        assert max_scaled_err(a, b_) <= 1e-9, name


# test/surface_hydrology/canopy_interception_tests.jl, canopy_evapotranspiration_tests.jl: edge properties of the scalar rules
@pytest.mark.parametrize("engine", ENGINES)
def test_canopy_edge_properties(engine):
    n = 6
    #            no rain   no leaves   wet canopy   negative store   dense canopy   open stomata
    rain = np.array([0.0,   1.0e-8,     1.0e-8,      1.0e-8,          1.0e-8,        1.0e-8])
    Cv   = np.array([2.2,   0.0,        2.2,         2.2,             8.8,           2.2])      # LAI = C_veg / 2.2
    SAI  = np.array([0.5,   0.0,        0.5,         0.5,             1.0,           0.5])
    w    = np.array([0.0,   1.0e-4,     1.0e-4,      -1.0,            1.0e-4,        0.0])
    An0  = np.array([0.0,   0.0,        0.0,         0.0,             0.0,           5.0e-4])
    integ = veg_land(engine, ncol=n, inputs={"rainfall": rain, "SAI": SAI, "windspeed": 1.0, "surface_shortwave_down": 0.0},
                     inits={"carbon_vegetation": Cv, "canopy_water": w, "net_assimilation": An0, "temperature": 8.0, "skin_temperature": 9.0})
    integ.step(60.0, 1)
    st = integ.state
    I, R, f, rg = (st.canopy_water_interception.numpy(), st.canopy_water_removal.numpy(), st.saturation_canopy_water.numpy(),
                   st.rainfall_ground.numpy())
    assert I[0] == 0 and I[1] == 0 and np.all((I[2:] > 0) & (I[2:] < rain[2:]))          # no rain / no canopy -> no interception
    assert f[0] == 0 and f[1] == 0 and 0 < f[2] < 1 and f[4] < f[2]                      # a denser canopy is less saturated
    assert R[0] == 0 and R[3] == 0 and R[2] > 0                                          # removal of a non-positive store is zero
    np.testing.assert_array_equal(rg, rain - I + R)
    tr = st.transpiration.numpy()
    assert np.all(np.isfinite(tr)) and np.all(tr > 0) and tr[5] > 5 * tr[0]
    assert 0 < tr[1] < 1.0e-8   # no conductance at all (no leaves): r_s = 1 / sqrt(eps), tiny but positive (canopy_evapotranspiration.jl:51-56)
    Ec, Eg = st.evaporation_canopy.numpy(), st.evaporation_ground.numpy()
    assert Ec[0] == 0 and Ec[2] > 0 and np.all(Eg > 0) and Eg[4] < Eg[2]                 # more leaves -> larger ground-canopy resistance
    # the canopy store tendency: interception - evaporation - removal
    np.testing.assert_allclose(st.canopy_water.numpy(), w + 60.0 * (I - Ec - R), rtol=1e-12, atol=1e-20)
