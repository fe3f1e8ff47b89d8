"""Where the warp-per-column kernel stops paying: per-step device time of the soil energy + Richards step (600 steps per call)
on both kernel families over a range of column counts. usage: python profiles/warp_crossover.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from common import synthetic_soil_case  # noqa: E402

os.environ["TRM_WARP_COLS"] = "1000000000"
for nf, name in ((np.float32, "f32"), (np.float64, "f64")):
    for heun in (False, True):
        for ncol in (16384, 32768, 57344, 81920, 114688, 163840):
            row = []
            for warp in ("0", "1"):
                os.environ["TRM_WARP"] = warp
                g = synthetic_soil_case("cuda", ncol, nf=nf, heun=heun, math="fast")
                g.step(60.0, 20)
                best = 1e30
                for _ in range(2):
                    g.step(60.0, 300)
                    ms = C.c_float()
                    g._lib.check(g._lib.last_step_ms(g._h, C.byref(ms)), "last_step_ms")
                    best = min(best, ms.value)
                row.append(1e3 * best / 300)
                g.close()
            print(f"{name} {'heun ' if heun else 'euler'} {ncol:7d} columns: streaming {row[0]:8.2f} us/step   warp-per-column {row[1]:8.2f} us/step", flush=True)

# ---- one step per call (per-step callers, host-evaluated boundary conditions): the tile copies are paid every step ----
print("one step per trm_step call:")
for nf, name in ((np.float32, "f32"), (np.float64, "f64")):
    for heun in (False, True):
        for ncol in (8192, 16384, 32768, 57344):
            row = []
            for warp in ("0", "1"):
                os.environ["TRM_WARP"] = warp
                g = synthetic_soil_case("cuda", ncol, nf=nf, heun=heun, math="fast")
                g.step(60.0, 5)
                tot = 0.0
                for _ in range(100):
                    g.step(60.0, 1)
                    ms = C.c_float()
                    g._lib.check(g._lib.last_step_ms(g._h, C.byref(ms)), "last_step_ms")
                    tot += ms.value
                row.append(1e3 * tot / 100)
                g.close()
            print(f"{name} {'heun ' if heun else 'euler'} {ncol:7d} columns: streaming {row[0]:8.2f} us/step   warp-per-column {row[1]:8.2f} us/step", flush=True)
