"""Per-step device time of the reference's real (small) configurations on one B200: BASELINE configs 1-4 at their own column
counts (single column ; N72 land mask = 14 017 columns ; N145 = 56 951), one `trm_step(h, dt, nsteps)` call, CUDA events of the
library (`trm_last_step_ms`). Each case runs on the one-thread-per-column streaming kernels (TRM_WARP=0) and on the library's
default for small SoilModel domains, the warp-per-column kernel (csrc/warp_kernel.cuh).

    python profiles/small_domains.py [--steps 600] [--out profiles/r02_small_domains.csv]"""
import argparse
import csv
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from common import synthetic_land_case, synthetic_soil_case  # noqa: E402
from test_vegetation import synthetic_vegetated_case  # noqa: E402

CASES = [
    ("config 1: quick-start column, heat + freeze-thaw", 1, dict(kind="soil", richards=False, heun=False, nf=np.float64, dt=300.0)),
    ("config 2: soil heat, N72 (14 017 columns), Float32", 14017, dict(kind="soil", richards=False, heun=False, nf=np.float32, dt=300.0)),
    ("config 3: soil energy + Richards, N145 (56 951 columns), Float32, ForwardEuler", 56951, dict(kind="soil", richards=True, heun=False, nf=np.float32, dt=60.0)),
    ("config 3: soil energy + Richards, N145, Float32, Heun", 56951, dict(kind="soil", richards=True, heun=True, nf=np.float32, dt=60.0)),
    ("config 3: soil energy + Richards, N145, Float64, Heun", 56951, dict(kind="soil", richards=True, heun=True, nf=np.float64, dt=60.0)),
    ("soil energy + Richards, 8 192 columns, Float32, ForwardEuler", 8192, dict(kind="soil", richards=True, heun=False, nf=np.float32, dt=60.0)),
    ("config 4: vegetated LandModel, N145 (56 951 columns), Float32, Heun", 56951, dict(kind="veg", heun=True, nf=np.float32, dt=60.0)),
    ("config 4: vegetated LandModel, N145, Float32, ForwardEuler", 56951, dict(kind="veg", heun=False, nf=np.float32, dt=60.0)),
    ("vegetated LandModel, N72 (14 017 columns), Float32, Heun", 14017, dict(kind="veg", heun=True, nf=np.float32, dt=60.0)),
    ("vegetated LandModel, N72, Float64, Heun", 14017, dict(kind="veg", heun=True, nf=np.float64, dt=60.0)),
    ("bare-ground LandModel, N72, Float32, Heun", 14017, dict(kind="land", heun=True, nf=np.float32, dt=60.0)),
]


def build(ncol, kind, nf, heun, dt, richards=True):
    if kind == "soil":
        return synthetic_soil_case("cuda", ncol, nf=nf, richards=richards, heun=heun, math="fast", dt=dt)
    if kind == "land":
        return synthetic_land_case("cuda", ncol, nf=nf, heun=heun, math="fast", windspeed=0.5)
    return synthetic_vegetated_case("cuda", ncol, nf=nf, heun=heun, math="fast")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rows = []
    for name, ncol, kw in CASES:
        for warp in ("0", "1"):
            os.environ["TRM_WARP"] = warp
            integ = build(ncol, **kw)
            integ.step(kw["dt"], 20)
            best = None
            l0 = integ._lib.launch_count(integ._h)
            for _ in range(3):
                integ.step(kw["dt"], args.steps)
                ms = C.c_float()
                integ._lib.check(integ._lib.last_step_ms(integ._h, C.byref(ms)), "last_step_ms")
                best = ms.value if best is None else min(best, ms.value)
            nl = (integ._lib.launch_count(integ._h) - l0) // 3
            us = 1e3 * best / args.steps
            rows.append({"case": name, "columns": ncol, "kernel": "warp-per-column" if warp == "1" else "streaming",
                         "steps_per_call": args.steps, "launches_per_call": nl, "us_per_step": round(us, 3),
                         "column_layer_steps_per_s": round(ncol * 30 * args.steps / (best * 1e-3), 1)})
            print(f"{name:90s} TRM_WARP={warp}  launches {nl:5d}  {us:9.3f} us/step", flush=True)
            integ.close()
    if args.out:
        with open(args.out, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(rows[0]))
            w.writeheader()
            w.writerows(rows)


if __name__ == "__main__":
    main()
