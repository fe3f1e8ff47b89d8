"""Simulation / output writers / callbacks (docs/src/running/time_stepping.md:86-180 of the reference):
scheduled NetCDF output with asynchronous snapshots, callbacks, time-step alignment, FieldTimeSeries read-back."""
import numpy as np
import pytest

from common import ENGINES, make, synthetic_soil_case, trm
from test_vegetation import synthetic_vegetated_case


@pytest.mark.parametrize("engine", ENGINES)
def test_simulation_with_output_writer_and_callbacks(engine, tmp_path):
    integ = synthetic_soil_case(engine, 37, dt=300.0)
    ref = synthetic_soil_case(engine, 37, dt=300.0)
    sim = trm.Simulation(integ, stop_time=6 * 3600.0, dt=300.0)
    path = str(tmp_path / "soil.nc")
    sim.output_writers["soil"] = trm.NetCDFWriter(integ, {"temperature": integ.state.temperature, "saturation": integ.state.saturation_water_ice,
                                                           "water_table": "water_table", "K": "hydraulic_conductivity"},
                                                  filename=path, schedule=trm.TimeInterval(2 * 3600.0), overwrite_existing=True)
    seen = []
    sim.callbacks["progress"] = trm.Callback(lambda s: seen.append((s.model.clock.iteration, s.model.clock.time)), trm.IterationInterval(30))
    trm.run_simulation(sim)
    sim.close()
    assert integ.clock.time == 6 * 3600.0 and integ.clock.iteration == 72 and sim.steps_taken == 72
    assert seen == [(0, 0.0), (30, 9000.0), (60, 18000.0)]
    T = trm.FieldTimeSeries(path, "temperature")
    S = trm.FieldTimeSeries(path, "saturation")
    K = trm.FieldTimeSeries(path, "K")
    assert list(T.times) == [0.0, 7200.0, 14400.0, 21600.0] and T.data.shape == (4, 30, 37) and K.data.shape == (4, 31, 37)
    np.testing.assert_array_equal(T.z, integ.grid.znodes_center())
    # every record equals the state of an undisturbed run at that time (run! semantics: auxiliaries current at output)
    ref.compute_auxiliary()
    for i in range(4):
        assert np.array_equal(T[i], ref.state.temperature.numpy()), i
        assert np.array_equal(S[i], ref.state.saturation_water_ice.numpy()), i
        assert np.array_equal(K[i], ref.state.hydraulic_conductivity.numpy()), i
        if i < 3:
            trm.run(ref, steps=24, dt=300.0)
    assert np.array_equal(T[-1], integ.state.temperature.numpy())
    # a finished simulation does not run again until the integrator is re-initialised (time_stepping.md:131-133)
    trm.run_simulation(sim)
    assert sim.steps_taken == 72
    with pytest.raises(FileExistsError):
        trm.NetCDFWriter(integ, ["temperature"], filename=path, schedule=trm.IterationInterval(1))


@pytest.mark.parametrize("engine", ENGINES)
def test_time_step_alignment_and_stop_iteration(engine, tmp_path):
    integ = synthetic_soil_case(engine, 5, dt=300.0)
    sim = trm.Simulation(integ, stop_time=2500.0, dt=300.0)
    sim.output_writers["o"] = trm.NetCDFWriter(integ, ["temperature"], filename=str(tmp_path / "a.nc"), schedule=trm.TimeInterval(1000.0))
    sim.run(); sim.close()
    T = trm.FieldTimeSeries(str(tmp_path / "a.nc"), "temperature")
    # 3 x 300 + 100 to land on 1000, 3 x 300 + 100 to land on 2000, 300 + 200 to land on the stop time
    assert list(T.times) == [0.0, 1000.0, 2000.0] and integ.clock.time == 2500.0 and sim.steps_taken == 10
    ref = synthetic_soil_case(engine, 5, dt=300.0)
    for dt in (300.0, 300.0, 300.0, 100.0):
        ref.step(dt, 1)
    assert np.array_equal(T[1], ref.state.temperature.numpy())
    integ2 = synthetic_soil_case(engine, 5, dt=300.0)
    sim2 = trm.Simulation(integ2, stop_iteration=7, dt=300.0)
    sim2.run()
    assert integ2.clock.iteration == 7 and integ2.clock.time == 2100.0


@pytest.mark.parametrize("engine", ENGINES)
def test_vegetated_simulation_equals_timestep_loop(engine, tmp_path):
    """The vegetated LandModel finalizes every step (timestep! semantics): identical to a timestep! loop that starts, like
    the simulation, with an evaluation of the auxiliaries."""
    a = synthetic_vegetated_case(engine, 11)
    b = synthetic_vegetated_case(engine, 11)
    sim = trm.Simulation(a, stop_iteration=12, dt=60.0)
    assert sim.finalize_every_step
    sim.output_writers["veg"] = trm.NetCDFWriter(a, ["carbon_vegetation", "canopy_water_conductance", "temperature"], filename=str(tmp_path / "v.nc"),
                                                 schedule=trm.IterationInterval(4))
    sim.run(); sim.close()
    b.compute_auxiliary()   # run!(sim) starts with update_state! (Oceananigans initialises the simulation): one more evaluation
    for _ in range(12):
        trm.timestep(b, 60.0)
    for name in ("temperature", "carbon_vegetation", "canopy_water_conductance", "net_assimilation", "ground_heat_flux"):
        assert np.array_equal(getattr(a.state, name).numpy(), getattr(b.state, name).numpy()), name
    g = trm.FieldTimeSeries(str(tmp_path / "v.nc"), "canopy_water_conductance")
    assert len(g) == 4 and np.array_equal(g[-1], b.state.canopy_water_conductance.numpy())


@pytest.mark.parametrize("engine", ENGINES)
def test_averaged_time_interval(engine, tmp_path):
    """AveragedTimeInterval: every record is the time average of the fields after each step of the window that ends at
    the output time (device-side accumulator), the first record the instantaneous initial field."""
    integ = synthetic_soil_case(engine, 9, dt=300.0)
    ref = synthetic_soil_case(engine, 9, dt=300.0)
    sim = trm.Simulation(integ, stop_time=3600.0, dt=300.0)
    path = str(tmp_path / "avg.nc")
    sim.output_writers["avg"] = trm.NetCDFWriter(integ, ["temperature", "ground_temperature", "hydraulic_conductivity"], filename=path,
                                                 schedule=trm.AveragedTimeInterval(1800.0))
    sim.run(); sim.close()
    T, Tg, K = (trm.FieldTimeSeries(path, n) for n in ("temperature", "ground_temperature", "hydraulic_conductivity"))
    assert list(T.times) == [0.0, 1800.0, 3600.0]
    ref.compute_auxiliary()
    assert np.array_equal(T[0], ref.state.temperature.numpy()) and np.array_equal(K[0], ref.state.hydraulic_conductivity.numpy())
    for rec in (1, 2):
        accT, accG, accK = 0.0, 0.0, 0.0
        for _ in range(6):
            trm.timestep(ref, 300.0)
            accT = accT + 300.0 * ref.state.temperature.numpy()
            accG = accG + 300.0 * ref.state.ground_temperature.numpy()
            accK = accK + 300.0 * ref.state.hydraulic_conductivity.numpy()
        np.testing.assert_allclose(T[rec], accT / 1800.0, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(Tg[rec], accG / 1800.0, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(K[rec], accK / 1800.0, rtol=1e-13)
    assert np.array_equal(integ.state.temperature.numpy(), ref.state.temperature.numpy())


@pytest.mark.parametrize("engine", ENGINES)
def test_bare_ground_land_trajectory_does_not_depend_on_the_output_interval(engine, tmp_path):
    """compute_auxiliary! is not idempotent for the bare-ground LandModel either (the surface block reads the stored skin
    temperature and writes a new one), and the reference's Simulation finalizes every step (model_integrator.jl:64,125-131):
    the state after 12 steps must not depend on how often something observes it, and must equal a timestep! loop."""
    from common import synthetic_land_case
    runs = []
    for every in (1, 4, 12):
        integ = synthetic_land_case(engine, 9, windspeed=0.5)
        sim = trm.Simulation(integ, stop_iteration=12, dt=60.0)
        assert sim.finalize_every_step
        sim.output_writers["o"] = trm.NetCDFWriter(integ, ["skin_temperature", "temperature"], filename=str(tmp_path / f"o{every}.nc"),
                                                   schedule=trm.IterationInterval(every))
        sim.run(); sim.close()
        runs.append(integ)
    loop = synthetic_land_case(engine, 9, windspeed=0.5)
    loop.compute_auxiliary()
    for _ in range(12):
        trm.timestep(loop, 60.0)
    for name in ("skin_temperature", "temperature", "ground_heat_flux", "internal_energy"):
        want = getattr(loop.state, name).numpy()
        for integ in runs:
            assert np.array_equal(getattr(integ.state, name).numpy(), want), name
    # a SoilModel has no auxiliary that is read back: it may batch steps
    from common import synthetic_soil_case
    assert not trm.Simulation(synthetic_soil_case(engine, 3), stop_iteration=1, dt=60.0).finalize_every_step


@pytest.mark.parametrize("engine", ENGINES)
def test_partitioned_run_writes_one_file_per_rank(engine, tmp_path):
    """Every rank of a partitioned run holds its own column range: the writer must not share one file between ranks."""
    from common import make, synthetic_columns
    n = 10
    lat, lon, T0 = synthetic_columns(n)
    files = []
    for rank in range(2):
        grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(N=8), n)
        integ = make(engine, trm.SoilModel(grid), trm.ForwardEuler(dt=300.0), boundary_conditions=trm.PrescribedSurfaceTemperature("T_ub", T0),
                     initializers={"temperature": lambda x, z: T0[None, :] - 0.05 * z, "saturation_water_ice": 1.0}, partition=(rank, 2))
        sim = trm.Simulation(integ, stop_iteration=3, dt=300.0)
        w = trm.NetCDFWriter(integ, ["temperature"], filename=str(tmp_path / "part.nc"), schedule=trm.IterationInterval(1))
        sim.output_writers["t"] = w
        sim.run(); sim.close()
        files.append((w.filename, integ.col0, integ.col1, integ.state.temperature.numpy()))
    assert files[0][0].endswith("part_rank0.nc") and files[1][0].endswith("part_rank1.nc")
    from scipy.io import netcdf_file
    for name, c0, c1, T in files:
        with netcdf_file(name, "r", mmap=False) as f:
            assert int(f.column_range_start) == c0 and int(f.column_range_stop) == c1 and int(f.columns_global) == n
            assert np.array_equal(f.variables["temperature"][-1], T)
    assert files[0][2] == files[1][1] and files[1][2] == n
