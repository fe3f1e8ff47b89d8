#!/usr/bin/env python
"""bench.py -- throughput of the fused per-column land time-step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # restated reference CPU path on the host cores

Workload (BASELINE.json configs[4]): synthetic 10 M-column x 30-layer domain, coupled soil
energy + Richards hydrology, ForwardEuler, dt = 60 s, Float64, sinusoidal surface temperature forcing
evaluated on the device.  A "step" is ONE fused stage-kernel launch over every column of the rank's column
range.  For N > 1 the 10 M columns are split by contiguous column range (strong scaling, no halo, no
data-path collective); NCCL only reduces the timing and the global budgets.

Prints ONE JSON line (rank 0).  `value` = column-layer-steps/s with the state resident in HBM, timed with
CUDA events on the library's stream; `e2e` = the same metric when every step also uploads that step's
surface forcing from pinned host memory and downloads the ground temperature (coupled-model usage, cf.
examples/simulations/speedy_dry_land.jl) through the public C ABI.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "column-layer-steps/sec"
UNIT = "column-layer-steps/s"
NZ = 30
DT = 60.0
SEED = 20260101
CPU_BUILD = ""


def algorithmic_bytes_per_cell(itemsize: int, nz: int, model: str = "soil", heun: bool = False, moved: bool = False) -> float:
    """read U, sat; write U, sat, T, liq, psi (7 values per cell) + per column: surface_excess_water
    R/W, water_table R/W, sinusoid forcing parameters (mean, amp, phase) R = 7 values per column. The same figure is the
    algorithmic one for Heun ("Euler or fused Heun: 7 s"): a step reads the state once and writes the new state once.

    `moved = True`: what the TWO stage launches of a Heun step actually move (recompute protocol, both number
    formats): stage 1 reads U, sat and writes k1 of both; stage 2 reads U, sat, both k1 and writes U, sat, T, liq, psi = 13
    values per cell. Per column, the bare-ground LandModel moves 22 values (skin
    temperature and surface excess water, 8 forcing parameters / table rows in; 10 surface fields out; G and infiltration read
    back by the stage kernel), the vegetated one 49 (adds 3 prognostic variables R/W, the previous net assimilation, the soil
    moisture factor R/W, SAI, 17 auxiliaries out); a Heun step re-reads G and the infiltration in stage 2 (+2) and, vegetated,
    evaluates the vegetation block again on the stage state (+27)."""
    per_cell = 7.0
    per_col = {"soil": 7.0, "land": 22.0, "land-veg": 49.0}[model]
    if heun and moved:
        per_cell = 13.0
        per_col += {"soil": 4.0, "land": 2.0, "land-veg": 27.0}[model]
    return per_cell * itemsize + per_col * itemsize / nz


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(nf_name: str):
    """dram bytes per launch of the stage kernel from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(nf_name)


def synthetic(ncol_global: int, c0: int, c1: int):
    rng = np.random.default_rng(SEED)
    lat = rng.uniform(-np.pi / 2, np.pi / 2, ncol_global)[c0:c1]
    lon = rng.uniform(0.0, 2 * np.pi, ncol_global)[c0:c1]
    T0 = 20.0 - np.abs(40.0 * np.sin(lat))
    return lon, T0


def build_case(trm, make_integrator, ncol_global, rank, world, device, nf, math, forcing="sinusoid", model_kind="soil", heun=False):
    from common import richards_soil
    grid = trm.ColumnGrid(trm.B200(device), nf, trm.ExponentialSpacing(dz_min=0.05, dz_max=100.0, N=NZ), ncol_global)
    c0, c1 = grid.partition(rank, world)
    lon, T0 = synthetic(ncol_global, c0, c1)
    if model_kind in ("land", "land-veg"):
        # secondary workloads (BASELINE configs[3]): bare-ground LandModel, or the full LandModel with the PALADYN vegetation
        # and canopy hydrology, under the synthetic atmosphere of BASELINE.md section 5 with a calm wind (see
        # tests/test_parity.py on the stability of the as-coded skin / soil coupling at 3 m/s)
        veg = None
        if model_kind == "land-veg":   # turnover rates per second that keep the carbon pool in range (tests/test_vegetation.py)
            veg = trm.VegetationCarbon(carbon_dynamics=trm.PALADYNCarbonDynamics(gamma_L=1e-9, gamma_R=1e-9, gamma_S=1e-10),
                                       vegetation_dynamics=trm.PALADYNVegetationDynamics(gamma_v_min=1e-8))
        model = trm.LandModel(grid, soil=richards_soil(), vegetation=veg)
        day = 86400.0
        hours = np.arange(0, 73, dtype=np.float64)
        rain = np.where((hours % 24) < 6, 2.0e-8, 0.0)
        inputs = {"air_temperature": trm.Sinusoid(mean=T0, amp=8.0, phase=lon, period=day),
                  "surface_shortwave_down": trm.Sinusoid(mean=0.0, amp=600.0, phase=lon, period=day, lo=0.0),
                  "surface_longwave_down": 300.0, "specific_humidity": 0.005, "air_pressure": 101325.0, "windspeed": 0.5,
                  "rainfall": trm.TimeSeries(hours * 3600.0, np.repeat(rain[:, None], c1 - c0, axis=1))}
        zc = grid.znodes_center().astype(np.float64)
        temp = (T0[None, :] - 0.05 * zc[:, None]).astype(nf)
        sat = np.ascontiguousarray(np.broadcast_to(np.minimum(1.0, 0.5 - 0.1 * zc)[:, None], temp.shape), dtype=nf)
        ts = (trm.Heun if heun else trm.ForwardEuler)(dt=DT)
        inits = {"temperature": temp, "saturation_water_ice": sat, "skin_temperature": T0}
        if veg is not None:
            inputs.update(SAI=0.5, CO2=400.0)
            inits.update(carbon_vegetation=10.0, vegetation_area_fraction=0.5)
        integ = make_integrator(model, ts, inputs, initializers=inits, partition=(rank, world), math=math)
        return integ, lon, T0
    model = trm.SoilModel(grid, soil=richards_soil())
    if forcing == "sinusoid":
        value = trm.Sinusoid(mean=T0, amp=10.0, phase=lon, period=86400.0)
    else:
        value = T0 + 10.0 * np.sin(-lon)
    bcs = trm.PrescribedSurfaceTemperature("T_ub", value)
    zc = grid.znodes_center().astype(np.float64)
    temp = (T0[None, :] - 0.05 * zc[:, None]).astype(nf)
    sat = np.ascontiguousarray(np.broadcast_to(np.minimum(1.0, 0.5 - 0.1 * zc)[:, None], temp.shape), dtype=nf)
    # temp / sat already hold only this rank's columns
    integ = make_integrator(model, (trm.Heun if heun else trm.ForwardEuler)(dt=DT), None, boundary_conditions=bcs,
                            initializers={"temperature": temp, "saturation_water_ice": sat},
                            partition=(rank, world), math=math)
    return integ, lon, T0


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        power = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def cpu_reference(ncol_sample, steps, warmup, nf, budget_s=None):
    """Restated reference CPU path (C++/OpenMP oracle in reference-structured mode) on the host cores.
    With `budget_s` the step count is chosen from a 2-step probe so that the timed sample takes about that long."""
    import oracle_integrator as oi
    import terrarium_jl_b200 as trm
    if "TERRARIUM_ORACLE_LIB" not in os.environ:   # the CPU arm uses the instruction set of the host it is timed on
        os.environ["TERRARIUM_ORACLE_LIB"] = oi.build_native_oracle()
    integ, _, _ = build_case(trm, oi.oracle_initialize, ncol_sample, 0, 1, 0, nf, "faithful")
    cdll = oi.oracle_library().cdll
    if os.environ.get("TORCHELASTIC_RUN_ID") and "TERRARIUM_CPU_THREADS" not in os.environ:
        # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and may use the whole host
        cdll.orc_set_num_threads(len(os.sched_getaffinity(0)))
    elif "TERRARIUM_CPU_THREADS" in os.environ:
        cdll.orc_set_num_threads(int(os.environ["TERRARIUM_CPU_THREADS"]))
    cores = int(cdll.orc_num_threads())
    integ.step(DT, max(warmup, 1))
    if budget_s is not None:
        t0 = time.perf_counter()
        integ.step(DT, 2)
        per_step = (time.perf_counter() - t0) / 2
        steps = int(min(max(budget_s / max(per_step, 1e-6), 3), 400))
    t0 = time.perf_counter()
    integ.step(DT, steps)
    el = time.perf_counter() - t0
    global CPU_BUILD
    CPU_BUILD = "-O3 -march=native" if os.environ.get("TERRARIUM_ORACLE_LIB", "").endswith("_native.so") else "-O3 -march=x86-64-v3 (AVX2)"
    return ncol_sample * NZ * steps / el, el, cores, steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--columns", type=int, default=10_000_000, help="total columns of the synthetic domain")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--math", default="fast", choices=["fast", "faithful"])
    ap.add_argument("--block", type=int, default=0, help="threads per block of the stage kernel (0 = library default)")
    ap.add_argument("--cpu-columns", type=int, default=1048576, help="columns of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-mode", default="mapped", choices=["mapped", "copy"],
                    help="end-to-end leg: `mapped` = trm_bind_host_io (the stage kernel reads the forcing from / writes the ground "
                         "temperature to page-locked host memory, one trm_step_async per step); `copy` = trm_set_input_field_async + "
                         "trm_step_async + trm_get_field_async (copy engines, three calls per step)")
    ap.add_argument("--model", default="soil", choices=["soil", "land", "land-veg"],
                    help="secondary workloads (not the headline): bare-ground LandModel, LandModel with PALADYN vegetation")
    ap.add_argument("--timestepper", default="euler", choices=["euler", "heun"], help="secondary workloads: Heun (two stage launches per step)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the `secondary` array (Float32, Heun, LandModel workloads, 20 steps each)")
    ap.add_argument("--secondary-steps", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    nf = np.float64 if args.dtype == "f64" else np.float32
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        ncs = args.cpu_columns
        v, el, cores, _ = cpu_reference(ncs, args.steps, args.warmup, nf)
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "soil energy + Richards hydrology, ForwardEuler, dt=60 s, Nz=30 exponential grid, sinusoidal surface "
                                   "temperature; restated reference CPU path (C++/OpenMP oracle, one loop nest per reference kernel); the "
                                   "Julia reference cannot run here (no julia in the image)",
                       "columns_per_step": ncs, "nz": NZ},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{ncs} columns x {NZ} layers x {args.steps} steps of the same workload; g++ {CPU_BUILD}, OpenMP"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import terrarium_jl_b200 as trm
    from terrarium_jl_b200 import distributed as td

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    # one process per GPU, on the cores of the GPU's NUMA node (pinned forcing / result buffers live next to its PCIe root)
    numa_cpus = td.bind_to_gpu_numa(local_rank) if world > 1 else []
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    secondary = args.model != "soil" or args.timestepper != "euler"
    if secondary:
        args.no_e2e = True   # the end-to-end leg drives the soil model's surface temperature input
    integ, lon, T0 = build_case(trm, trm.initialize, args.columns, rank, world, local_rank, nf, args.math,
                                model_kind=args.model, heun=args.timestepper == "heun")
    lib, h = integ._lib, integ._h
    if args.block:
        lib.check(lib.set_block_size(h, args.block), "set_block_size")
    ncol_local = integ.ncol
    itemsize = np.dtype(nf).itemsize

    d0 = integ.diagnostics()
    integ.step(DT, args.warmup)
    ms = C.c_float()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = lib.launch_count(h)
    w0 = time.perf_counter()
    integ.step(DT, args.steps)   # K fused stage-kernel launches; device time bracketed by CUDA events on the library's stream
    barrier()
    wall = time.perf_counter() - w0
    launches = lib.launch_count(h) - l0
    lib.check(lib.last_step_ms(h, C.byref(ms)), "last_step_ms")
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = td.max_over_ranks(ms.value)   # device time of the K launches, max over ranks
    total_cells = args.columns * NZ
    value = total_cells * args.steps / (dev_ms * 1e-3)

    # global budgets (NCCL all-reduce of the per-rank diagnostics): the run must conserve water
    g0, g1 = td.reduce_diagnostics(d0), td.reduce_diagnostics(integ.diagnostics())
    bud = [g0["water"], g1["water"], g1["nan_count"], g1["energy"]]

    # ---- end-to-end through the public C ABI with HOST buffers: every step the host hands over that step's surface
    #      forcing (pinned memory -> trm_set_input_field_async), advances one step (trm_step_async) and receives the
    #      ground temperature (trm_get_field_async -> pinned memory).  Uploads / downloads run on copy streams and
    #      overlap the stage kernel of the neighbouring steps; the timed region ends with trm_sync.
    def run_e2e(mode):
        k2 = max(3, args.steps)   # same step count as the kernel-only measurement
        NBUF = 4                  # host buffers are reused round robin (fresh forcing in / ground temperature out every step)
        in_id = integ._bc_inputs["T_ub"]
        gt_id = trm.abi.FIELD_IDS["ground_temperature"]
        t = integ.clock.time

        def forcing(i):   # the host-side "atmosphere": this step's surface temperature per column
            return (T0 + 10.0 * np.sin(2 * np.pi * (t + i * DT) / 86400.0 - lon)).astype(nf)

        if mode == "mapped":
            ring_in, ring_out = integ.bind_host_io("T_ub", "ground_temperature", nslots=NBUF)
            it0 = integ.clock.iteration
            for i in range(NBUF):
                ring_in[(it0 + i) % NBUF, :] = forcing(i)

            def e2e_step(i):
                # the coupler owns slot (it0 + i) % NBUF again once the step that used it last has completed: that is
                # where it would read that step's ground temperature and write this step's forcing
                if i >= NBUF:
                    integ.host_io_wait(it0 + i - NBUF + 1)
                lib.check(lib.step_async(h, DT, 1), "step_async")
            what = ("per step: ONE call, trm_step_async; the stage kernel reads that step's surface temperature forcing from, and "
                    "writes the ground temperature to, page-locked mapped host memory (trm_bind_host_io, %d-slot rings; the host "
                    "waits for the step that last used a slot before reusing it); wall clock around the loop incl. final trm_sync, "
                    "max over ranks" % NBUF)
            outs = ring_out
        else:
            tdtype = torch.float64 if nf == np.float64 else torch.float32
            forc = [torch.empty(ncol_local, dtype=tdtype).pin_memory() for _ in range(NBUF)]
            outs = [torch.empty(ncol_local, dtype=tdtype).pin_memory() for _ in range(NBUF)]
            for i, f in enumerate(forc):
                f.copy_(torch.from_numpy(forcing(i)))

            def e2e_step(i):
                lib.check(lib.set_input_field_async(h, in_id, C.c_void_p(forc[i % NBUF].data_ptr())), "set_input_field_async")
                lib.check(lib.step_async(h, DT, 1), "step_async")
                lib.check(lib.get_field_async(h, gt_id, C.c_void_p(outs[i % NBUF].data_ptr()), ncol_local), "get_field_async")
            what = ("per step: trm_set_input_field_async(surface temperature forcing, pinned host) + trm_step_async + "
                    "trm_get_field_async(ground_temperature -> pinned host); copies overlap the stage kernel on copy "
                    "streams; wall clock around the loop incl. final trm_sync, max over ranks")

        for i in range(3):
            e2e_step(i)
        lib.check(lib.sync(h), "sync")
        barrier()
        w0 = time.perf_counter()
        for i in range(3, 3 + k2):
            e2e_step(i)
        lib.check(lib.sync(h), "sync")
        barrier()
        el = td.max_over_ranks(time.perf_counter() - w0)
        res = {"value": total_cells * k2 / el, "unit": UNIT, "h2d_bytes_per_step": args.columns * itemsize,
               "d2h_bytes_per_step": args.columns * itemsize, "steps": k2, "ms_per_step": 1e3 * el / k2,
               "mode": mode, "what": what,
               "host_cpus_of_rank0": len(numa_cpus) if numa_cpus else None}
        if mode == "mapped":
            # the last step's slot holds the ground temperature the library reports for the final state
            last = (it0 + 3 + k2 - 1) % NBUF
            gt = integ.state.ground_temperature.numpy()
            assert np.array_equal(np.asarray(ring_out[last]), gt), "mapped host output differs from ground_temperature"
            assert bool(np.isfinite(np.asarray(ring_out)).all())
            lib.check(lib.bind_host_io(h, -1, None, -1, None, 0), "unbind_host_io")
        else:
            assert all(bool(torch.isfinite(o).all()) for o in outs)
        return res

    e2e = None
    if not args.no_e2e:
        # both public end-to-end paths are timed (the host side of these boxes decides which one is faster: the mapped path
        # depends on the latency of PCIe reads issued by the kernel, the copy path on the copy engines); `--e2e-mode` names
        # the preferred one, the line reports the faster as `e2e` and the other one next to it
        first = run_e2e(args.e2e_mode)
        second = run_e2e("copy" if args.e2e_mode == "mapped" else "mapped")
        e2e, other = (first, second) if first["ms_per_step"] <= second["ms_per_step"] else (second, first)
        e2e["other_mode"] = {k: other[k] for k in ("mode", "value", "ms_per_step", "steps")}

    peak, peak_src = peaks()
    bpc = algorithmic_bytes_per_cell(itemsize, NZ, args.model, args.timestepper == "heun")
    per_launch_ms = dev_ms / args.steps
    # ncu dram bytes of one launch over 10 M columns (profiles/traffic.json), scaled to this rank's column count
    traffic = profiled_traffic(args.dtype)
    if traffic is not None:
        traffic = traffic * ncol_local / 10_000_000 if not secondary else None
    achieved = bpc * (ncol_local * NZ) / (per_launch_ms * 1e-3) / 1e9   # this rank's kernel (ranks are symmetric)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_column_layer_step": bpc,
                "kernel": ("trm::stage_kernel" if os.environ.get("TRM_KERNEL") == "stream" else
                           ("trm::euler2_kernel (two columns per thread, packed f32x2)" if args.dtype == "f32" and args.math == "fast"
                            and os.environ.get("TRM_F32X2", "1") != "0" else "trm::euler_kernel"))
                          + f"<{args.dtype}, {'RICHARDS' if args.model == 'soil' else 'LAND'}, recompute, {args.math}>"
                          + (" (two stage launches per step)" if args.timestepper == "heun" else "")
                          + (" + trm::surface_kernel" if args.model != "soil" else ""),
                "launch_ms": per_launch_ms,
                "moved_bytes_per_column_layer_step": algorithmic_bytes_per_cell(itemsize, NZ, args.model, args.timestepper == "heun", moved=True)}

    # ---- secondary workloads on the same domain (not the headline; each with its own roofline and clock record) ----
    secondary_lines = []
    if not secondary and not args.no_secondary:
        integ.close()
        del integ
        cases = [("f32", "soil", "euler"), ("f64", "soil", "heun"), ("f32", "soil", "heun"), ("f64", "land", "euler"),
                 ("f64", "land-veg", "euler"), ("f64", "land-veg", "heun"), ("f32", "land-veg", "euler")]
        for dt_name, model_kind, stepper in cases:
            nf2 = np.float64 if dt_name == "f64" else np.float32
            it2, _, _ = build_case(trm, trm.initialize, args.columns, rank, world, local_rank, nf2, args.math,
                                   model_kind=model_kind, heun=stepper == "heun")
            it2.step(DT, max(args.warmup, 3))
            smp = ClockSampler(local_rank)
            if rank == 0:
                smp.start()
            barrier()
            la = it2._lib.launch_count(it2._h)
            it2.step(DT, args.secondary_steps)
            barrier()
            nl = it2._lib.launch_count(it2._h) - la
            m2 = C.c_float()
            it2._lib.check(it2._lib.last_step_ms(it2._h, C.byref(m2)), "last_step_ms")
            ck = smp.stop() if rank == 0 else None
            ms2 = td.max_over_ranks(m2.value) / args.secondary_steps
            d2 = it2.diagnostics()
            nan2 = td.reduce_diagnostics(d2)["nan_count"]
            isz = np.dtype(nf2).itemsize
            bpc2 = algorithmic_bytes_per_cell(isz, NZ, model_kind, stepper == "heun")
            ach2 = bpc2 * (it2.ncol * NZ) / (ms2 * 1e-3) / 1e9
            moved2 = algorithmic_bytes_per_cell(isz, NZ, model_kind, stepper == "heun", moved=True)
            secondary_lines.append({
                "workload": {"soil": "soil energy + Richards", "land": "bare-ground LandModel", "land-veg": "vegetated LandModel"}[model_kind]
                            + f", {stepper}, {dt_name}", "dtype": dt_name, "timestepper": stepper, "steps": args.secondary_steps,
                "ms_per_step": ms2, "value": total_cells / (ms2 * 1e-3), "unit": UNIT, "gpu_launches": int(nl), "nan_count": nan2,
                "roofline": {"bound": "hbm", "achieved": ach2, "peak": peak, "unit": "GB/s", "frac": ach2 / peak,
                             "algorithmic_bytes_per_column_layer_step": bpc2,
                             # what the launches of one step move (Heun: two stage launches) and the same fraction on that basis
                             "moved_bytes_per_column_layer_step": moved2, "frac_of_moved_bytes": ach2 / peak * moved2 / bpc2},
                "clocks": ck})
            it2.close()
            del it2

    # ---- small domains (one GPU): the same soil workload at the column counts of the reference's own configurations, 600
    #      steps per trm_step call. Below 114 688 (Float32) / 65 536 (Float64) columns the library runs them on the
    #      warp-per-column kernel (csrc/warp_kernel.cuh: all steps of a call in one launch, state in registers -- no HBM
    #      roofline applies, the figure is time per step); TRM_WARP=0 gives the streaming kernels' time next to it. ----
    small_lines = []
    if not secondary and not args.no_secondary and world == 1:
        for label, ncs, dt_name, stepper in (("single column", 1, "f64", "euler"), ("N72 land mask", 14017, "f32", "euler"),
                                             ("N72 land mask", 14017, "f64", "heun"), ("N145 land mask", 56951, "f32", "euler")):
            nf2 = np.float64 if dt_name == "f64" else np.float32
            row = {"workload": f"soil energy + Richards, {label}, {stepper}, {dt_name}", "columns": ncs, "steps_per_call": 600}
            for warp in ("1", "0"):
                os.environ["TRM_WARP"] = warp
                it2, _, _ = build_case(trm, trm.initialize, ncs, 0, 1, local_rank, nf2, args.math, heun=stepper == "heun")
                it2.step(DT, 20)
                la = it2._lib.launch_count(it2._h)
                it2.step(DT, 600)
                nl = it2._lib.launch_count(it2._h) - la
                m2 = C.c_float()
                it2._lib.check(it2._lib.last_step_ms(it2._h, C.byref(m2)), "last_step_ms")
                key = "warp_per_column" if warp == "1" else "streaming"
                row[key] = {"us_per_step": 1e3 * m2.value / 600, "gpu_launches": int(nl),
                            "value": ncs * NZ * 600 / (m2.value * 1e-3), "unit": UNIT, "nan_count": it2.diagnostics()["nan_count"]}
                it2.close()
                del it2
            os.environ.pop("TRM_WARP", None)
            small_lines.append(row)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        try:
            v, el, cores, nst = cpu_reference(args.cpu_columns, 10, 1, nf, budget_s=12.0)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{args.cpu_columns} columns x {NZ} layers x {nst} steps of the same workload ({el:.1f} s); restated "
                                      f"reference CPU path (C++/OpenMP oracle, one loop nest per reference kernel, g++ {CPU_BUILD}), all host cores"}
        except Exception as exc:  # the oracle is test infrastructure; its absence must not hide the GPU number
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {exc}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_launch_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "configs[4]: synthetic 10M-column soil energy + Richards hydrology (VanGenuchten alpha=2 n=2, "
                                   "UnsatKVanGenuchten, K_sat=1e-5), ForwardEuler dt=60 s, Nz=30 exponential grid, sinusoidal surface "
                                   "temperature evaluated on the device",
                       "columns": args.columns, "nz": NZ, "columns_per_gpu": ncol_local, "math": args.math,
                       "processes": {"soil": "soil energy + Richards", "land": "bare-ground LandModel", "land-veg": "vegetated LandModel"}[args.model], "timestepper": args.timestepper,
                       "partition": "contiguous column ranges, no halo, no data-path collective",
                       "cache": "inputs larger than L2 (state read per step = %.1f GB per GPU)" % (2 * ncol_local * NZ * itemsize / 1e9)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "secondary": secondary_lines, "small_domains": small_lines,
            "wall_s_timed_region": wall,
            "budgets": {"water_before": bud[0], "water_after": bud[1], "nan_count": bud[2], "energy_after": bud[3]},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
