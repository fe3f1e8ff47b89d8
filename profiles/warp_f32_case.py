import ctypes as C, os, sys
sys.path.insert(0, 'tests')
import numpy as np
from common import synthetic_soil_case
for heun in (False, True):
    for ncol in (1, 14017, 56951, 98304):
        g = synthetic_soil_case("cuda", ncol, nf=np.float32, heun=heun, math="fast")
        g.step(60.0, 20)
        best = 1e30
        for _ in range(2):
            g.step(60.0, 300)
            ms = C.c_float(); g._lib.check(g._lib.last_step_ms(g._h, C.byref(ms)), "x"); best = min(best, ms.value)
        print(f"f32 {'heun ' if heun else 'euler'} {ncol:6d} columns  {1e3*best/300:8.2f} us/step", flush=True)
        g.close()
