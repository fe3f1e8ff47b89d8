"""The reference's own benchmark design (test/benchmarks/gpu/soil_heat_hydrology_global.jl:39-103) on this implementation:
SoilModel with soil energy + Richards hydrology on a ColumnRingGrid over FullGaussianGrid(2^i), i = 1..10 (8 * 4^i
columns, all land), Float32, ExponentialSpacing(N = 30), default SoilInitializer, surface temperature 30 sin(2 pi t / 365 d),
`run!(state, period = Hour(1), dt = 60)` = 60 steps per sample, 10 samples, minimum / median wall time in ms and the
authors' derived metric simulated-years-per-day = 1000 * 24 * 3600 / (24 * median_ms).

    python profiles/reference_benchmark_design.py --engine cuda                 # on a B200
    python profiles/reference_benchmark_design.py --engine oracle --max-i 7     # restated reference CPU path (OpenMP)

Wall time is host time around `run!` (what BenchmarkTools measures in the reference), which includes the final
`compute_auxiliary!` and the synchronisation. The reference publishes no results for this design."""
import argparse
import csv
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import terrarium_jl_b200 as trm  # noqa: E402


def set_up_model(engine, nf, nrings, math):
    npoints = (4 * nrings) * (2 * nrings)              # FullGaussianGrid(nrings): 2 nrings latitude rings of 4 nrings points
    grid = trm.ColumnRingGrid(trm.B200(), nf, trm.ExponentialSpacing(N=30), np.ones(npoints, dtype=bool))
    soil = trm.SoilEnergyWaterCarbon(hydrology=trm.SoilHydrology(trm.RichardsEq()))
    model = trm.SoilModel(grid, soil=soil, initializer=trm.SoilInitializer())
    bcs = trm.PrescribedSurfaceTemperature("T_ub", trm.Sinusoid(mean=0.0, amp=30.0, phase=0.0, period=24 * 3600 * 365.0))
    if engine == "oracle":
        import oracle_integrator as oi
        return npoints, oi.oracle_initialize(model, trm.ForwardEuler(), boundary_conditions=bcs)
    return npoints, trm.initialize(model, trm.ForwardEuler(), boundary_conditions=bcs, math=math)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="cuda", choices=["cuda", "oracle"])
    ap.add_argument("--samples", type=int, default=10)
    ap.add_argument("--max-i", type=int, default=10)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--math", default="fast", choices=["fast", "faithful"])
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    nf = np.float32 if args.dtype == "f32" else np.float64
    _, integ = set_up_model(args.engine, nf, 8, args.math)   # quick test, as in the reference script
    trm.timestep(integ, 60.0)
    rows = []
    for i in range(1, args.max_i + 1):
        nrings = 2 ** i
        npoints, integ = set_up_model(args.engine, nf, nrings, args.math)
        trm.run(integ, period=3600.0, dt=60.0)              # BenchmarkTools' warm-up evaluation
        times = []
        for _ in range(args.samples):
            t0 = time.perf_counter()
            trm.run(integ, period=3600.0, dt=60.0)
            integ.synchronize()
            times.append(1e3 * (time.perf_counter() - t0))
        mid, lo = statistics.median(times), min(times)
        assert np.isfinite(integ.state.temperature.numpy()).all()
        rows.append({"nrings": nrings, "npoints": npoints, "min_time_ms": lo, "mid_time_ms": mid,
                     "sypd": 1000 * 24 * 3600 / (24 * mid), "column_layer_steps_per_s": npoints * 30 * 60 / (mid * 1e-3)})
        print(f"nrings {nrings:5d}  columns {npoints:9d}  min {lo:10.3f} ms  median {mid:10.3f} ms  SYPD {rows[-1]['sypd']:12.1f}  "
              f"{rows[-1]['column_layer_steps_per_s'] / 1e9:8.3f} G column-layer-steps/s", flush=True)
    if args.out:
        with open(args.out, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(rows[0]))
            w.writeheader()
            w.writerows(rows)


if __name__ == "__main__":
    main()
