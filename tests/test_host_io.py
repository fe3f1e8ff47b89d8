"""Per-step exchange with a host-side coupler through mapped host memory (trm_bind_host_io): the coupling flow of
examples/simulations/speedy_dry_land.jl:45-68 (forcing in, surface state out, every step) without a copy per step.

The stage kernel reads the input from / writes the result to page-locked host memory; the results must be identical,
bit for bit, to the same run driven with `trm_set_input_field` + `trm_get_field`."""
import numpy as np
import pytest

from common import make, richards_soil, synthetic_columns, synthetic_land_case, trm

pytestmark = pytest.mark.gpu


def soil_case(n, heun, math, nf=np.float64):
    lat, lon, T0 = synthetic_columns(n)
    grid = trm.ColumnGrid(trm.B200(), nf, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=30), n)
    model = trm.SoilModel(grid, soil=richards_soil())
    bcs = trm.PrescribedSurfaceTemperature("T_ub", T0)
    inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z,
             "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
    integ = make("cuda", model, (trm.Heun if heun else trm.ForwardEuler)(dt=60.0), boundary_conditions=bcs, initializers=inits, math=math)
    return integ, lon, T0


@pytest.mark.parametrize("nf", [np.float64, np.float32])
@pytest.mark.parametrize("math", ["faithful", "fast"])
@pytest.mark.parametrize("heun", [False, True])
def test_mapped_host_exchange_equals_copies(heun, math, nf):
    n, nsteps, nslots = 1000 + 13, 11, 3
    a, lon, T0 = soil_case(n, heun, math, nf)
    b, _, _ = soil_case(n, heun, math, nf)

    def forcing(i):
        return (T0 + 10.0 * np.sin(2 * np.pi * (i * 60.0) / 86400.0 - lon)).astype(nf)

    # reference leg: blocking calls
    want_gt = []
    for i in range(nsteps):
        a.state.T_ub.set(forcing(i))
        a.step(60.0, 1)
        want_gt.append(a.state.ground_temperature.numpy())
    # mapped leg: one trm_step_async per step, the host refills a slot after the step that used it has completed
    ring_in, ring_out = b.bind_host_io("T_ub", "ground_temperature", nslots=nslots)
    it0 = b.clock.iteration
    got_gt = [None] * nsteps
    for i in range(nsteps):
        if i >= nslots:
            b.host_io_wait(it0 + i - nslots + 1)
            got_gt[i - nslots] = np.array(ring_out[(it0 + i) % nslots])
        ring_in[(it0 + i) % nslots, :] = forcing(i)
        b.step_async(60.0, 1)
    b.synchronize()
    for i in range(nsteps - nslots, nsteps):
        got_gt[i] = np.array(ring_out[(it0 + i) % nslots])
    for i in range(nsteps):
        assert np.array_equal(got_gt[i], want_gt[i]), i
    for name in ("internal_energy", "temperature", "saturation_water_ice", "pressure_head", "liquid_water_fraction"):
        assert np.array_equal(getattr(a.state, name).numpy(), getattr(b.state, name).numpy()), name
    # unbinding restores the stored source
    b._lib.check(b._lib.bind_host_io(b._h, -1, None, -1, None, 0), "unbind")
    b.step(60.0, 1)
    assert np.isfinite(b.state.temperature.numpy()).all()


def test_mapped_host_exchange_land_model():
    """LandModel: an atmospheric input read by the surface kernel from host memory, a surface field (skin temperature)
    written back through the small copy kernel."""
    n, nsteps, nslots = 700, 8, 4
    a = synthetic_land_case("cuda", n, math="fast", windspeed=0.5)
    b = synthetic_land_case("cuda", n, math="fast", windspeed=0.5)
    lat, lon, T0 = synthetic_columns(n)

    def forcing(i):
        return T0 + 8.0 * np.sin(2 * np.pi * (i * 60.0) / 86400.0 - lon)

    want = []
    for i in range(nsteps):
        a.state.inputs.air_temperature.set(forcing(i))
        a.step(60.0, 1)
        want.append(a.state.skin_temperature.numpy())
    ring_in, ring_out = b.bind_host_io("air_temperature", "skin_temperature", nslots=nslots)
    it0 = b.clock.iteration
    got = [None] * nsteps
    for i in range(nsteps):
        if i >= nslots:
            b.host_io_wait(it0 + i - nslots + 1)
            got[i - nslots] = np.array(ring_out[(it0 + i) % nslots])
        ring_in[(it0 + i) % nslots, :] = forcing(i)
        b.step_async(60.0, 1)
    b.synchronize()
    for i in range(nsteps - nslots, nsteps):
        got[i] = np.array(ring_out[(it0 + i) % nslots])
    for i in range(nsteps):
        assert np.array_equal(got[i], want[i]), i
    assert np.array_equal(a.state.temperature.numpy(), b.state.temperature.numpy())


def test_bind_host_io_rejects_pageable_memory():
    integ, _, _ = soil_case(64, False, "fast")
    import ctypes as C
    buf = np.zeros((2, 64))
    rc = integ._lib.bind_host_io(integ._h, integ._bc_inputs["T_ub"], buf.ctypes.data_as(C.c_void_p), -1, None, 2)
    assert rc == trm.abi.TRM_ERR_INVALID
    assert b"page-locked" in integ._lib.last_error()
