"""Multi-rank host logic on CPU: world size 2, gloo backend, one integrator per rank over its column range.
The engine is the CPU oracle (no GPU here); partitioning, reductions and gathers are the code under test."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NCOL, STEPS = 101, 20   # odd: ranks get 51 and 50 columns


def _case(partition):
    from common import make, richards_soil, synthetic_columns, trm
    lat, lon, T0 = synthetic_columns(NCOL)
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=12), NCOL)
    model = trm.SoilModel(grid, soil=richards_soil())
    bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0))
    inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z,
             "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
    return make("oracle", model, trm.ForwardEuler(dt=60.0), boundary_conditions=bcs, initializers=inits, partition=partition)


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from terrarium_jl_b200 import distributed as td
    integ = _case((rank, world))
    assert integ.ncol == (51 if rank == 0 else 50) and integ.col0 == (0 if rank == 0 else 51)
    integ.step(60.0, STEPS)
    diag = td.reduce_diagnostics(integ.diagnostics())
    T = td.gather_field(integ, "temperature")
    slowest = td.max_over_ranks(float(rank + 1))
    if rank == 0:
        np.savez(out, T=T, **{k: np.float64(v) for k, v in diag.items()}, slowest=slowest)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_match_one(tmp_path):
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, out), nprocs=2, join=True)
    got = np.load(out)
    single = _case(None)
    single.step(60.0, STEPS)
    d = single.diagnostics()
    assert np.array_equal(got["T"], single.state.temperature.numpy())       # columns are independent: bit-identical
    assert got["ncol"] == NCOL and got["nan_count"] == 0
    assert got["water"] == pytest.approx(d["water"], rel=1e-13) and got["energy"] == pytest.approx(d["energy"], rel=1e-13)
    assert got["t_min"] == d["t_min"] and got["t_max"] == d["t_max"]
    assert got["slowest"] == 2.0


def _ring_case(partition):
    """ColumnRingGrid over the N72 land mask (every 7th land point) with a time-varying raster as surface temperature:
    each rank gathers its own columns of the ring raster (idxmap = findall(mask)[col0:col1])."""
    from common import make, trm
    from test_global_configs import FIXTURE
    d = np.load(FIXTURE)
    land = np.flatnonzero(np.unpackbits(d["N72_bits"])[:41472])
    mask = np.zeros(41472, dtype=bool)
    mask[land[::7]] = True
    grid = trm.ColumnRingGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_max=1.0, N=8), mask)
    rng = np.random.default_rng(11)
    times = np.arange(5) * 600.0
    raster = trm.RasterInputSource(values=5.0 + rng.normal(0.0, 3.0, (5, 41472)), times=times)
    bcs = trm.PrescribedSurfaceTemperature("T_ub", raster)
    integ = make("oracle", trm.SoilModel(grid), trm.ForwardEuler(dt=100.0), boundary_conditions=bcs, partition=partition,
                 initializers={"temperature": 1.0, "saturation_water_ice": 0.5})
    return grid, integ


def _ring_worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from terrarium_jl_b200 import distributed as td
    grid, integ = _ring_case((rank, world))
    assert integ.ncol == (1002 if rank == 0 else 1001) and integ.col0 == (0 if rank == 0 else 1002)   # 2003 = ceil(14017 / 7)
    integ.step(100.0, 15)
    T = td.gather_field(integ, "temperature")
    if rank == 0:
        np.savez(out, T=T, ring=grid.to_ring(T[-1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_on_a_ring_grid_with_raster_forcing(tmp_path):
    out = str(tmp_path / "ring0.npz")
    mp.spawn(_ring_worker, args=(2, 31500 + os.getpid() % 2000, out), nprocs=2, join=True)
    got = np.load(out)
    grid, single = _ring_case(None)
    assert grid.Nc == 2003
    single.step(100.0, 15)
    T = single.state.temperature.numpy()
    assert np.array_equal(got["T"], T)
    ring = got["ring"]
    assert ring.shape == (41472,) and np.isnan(ring).sum() == 41472 - 2003 and np.array_equal(ring[grid.mask], T[-1].reshape(-1))


# ---------------------------------------------------------------------------------------------
# device side: zero-copy views of the library's buffers and the NCCL output gather
def _cuda_case(partition, ncol=NCOL):
    from common import make, richards_soil, synthetic_columns, trm
    lat, lon, T0 = synthetic_columns(ncol)
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=12), ncol)
    model = trm.SoilModel(grid, soil=richards_soil())
    bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0))
    inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z,
             "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
    return make("cuda", model, trm.ForwardEuler(dt=60.0), boundary_conditions=bcs, initializers=inits, partition=partition)


@pytest.mark.gpu
def test_field_tensor_is_a_zero_copy_view():
    import torch
    from terrarium_jl_b200 import distributed as td
    integ = _cuda_case(None)
    integ.step(60.0, 5)
    for name in ("temperature", "saturation_water_ice", "hydraulic_conductivity", "water_table"):
        t = td.field_tensor(integ, name)
        want = getattr(integ.state, name).numpy()
        assert t.is_cuda and tuple(t.shape) == (want.shape if want.ndim == 2 else (1,) + want.shape)
        assert np.array_equal(t.cpu().numpy().reshape(want.shape), want)
    # a read-only view does not invalidate the stored closure fields: the next step launches the recompute variant
    # (one launch), and the result equals an undisturbed run
    ref = _cuda_case(None)
    ref.step(60.0, 10)
    integ.step(60.0, 5)
    assert np.array_equal(integ.state.temperature.numpy(), ref.state.temperature.numpy())
    full = td.gather_field_device(integ, "temperature")   # world size 1: a copy of the local field
    assert np.array_equal(full.cpu().numpy(), ref.state.temperature.numpy())


def nccl_gather_worker():
    """Run under torchrun on >= 2 GPUs (tests/test_distributed.py is the entry point): every rank gathers the
    temperature over NCCL from the library's device buffers and compares it with a single-rank run of the whole domain."""
    import torch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from terrarium_jl_b200 import distributed as td
    import terrarium_jl_b200 as trm_
    ncol = 1003
    from common import make, richards_soil, synthetic_columns
    lat, lon, T0 = synthetic_columns(ncol)
    grid = trm_.ColumnGrid(trm_.B200(local), np.float64, trm_.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=12), ncol)

    def build(partition):
        model = trm_.SoilModel(grid, soil=richards_soil())
        bcs = trm_.PrescribedSurfaceTemperature("T_ub", trm_.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0))
        inits = {"temperature": lambda x, z: T0[None, :] - 0.05 * z, "saturation_water_ice": lambda x, z: np.minimum(1.0, 0.5 - 0.1 * z) + 0 * x}
        return make("cuda", model, trm_.ForwardEuler(dt=60.0), boundary_conditions=bcs, initializers=inits, partition=partition)

    part, whole = build((rank, world)), build(None)
    part.step(60.0, STEPS); whole.step(60.0, STEPS)
    for name in ("temperature", "water_table"):
        got = td.gather_field_device(part, name).cpu().numpy()
        want = getattr(whole.state, name).numpy()
        assert np.array_equal(got.reshape(want.shape), want), name
    d = td.reduce_diagnostics(part.diagnostics())
    assert d["ncol"] == ncol and d["nan_count"] == 0
    # the same reduction inside the library (trm_diagnostics_allreduce): rank 0 creates the NCCL unique id through the C ABI,
    # the bytes travel over the process group the host already has, every rank initialises its handle's communicator
    import ctypes as C
    from terrarium_jl_b200 import _abi as abi
    lib, h = part._lib, part._h
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_char * 128)()
        lib.check(lib.nccl_get_unique_id(buf), "nccl_get_unique_id")
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda()
    dist.broadcast(uid, src=0)
    raw = bytes(uid.cpu().numpy().tobytes())
    lib.check(lib.nccl_comm_init(h, world, rank, C.c_char_p(raw)), "nccl_comm_init")
    g = abi.trm_diag()
    lib.check(lib.diagnostics_allreduce(h, C.byref(g)), "diagnostics_allreduce")
    w = whole.diagnostics()
    assert g.ncol == ncol and g.nan_count == 0
    assert g.t_min == w["t_min"] and g.t_max == w["t_max"] and g.sat_min == w["sat_min"] and g.sat_max == w["sat_max"]
    assert abs(g.energy - d["energy"]) <= 1e-12 * abs(d["energy"]) and abs(g.water - d["water"]) <= 1e-12 * abs(d["water"])
    assert abs(g.water - w["water"]) <= 1e-11 * abs(w["water"])
    dist.barrier()
    if rank == 0:
        print(f"nccl gather over {world} ranks: OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    nccl_gather_worker()
