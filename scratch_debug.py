import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np
from common import synthetic_land_case
for math in ("faithful", "fast"):
    gpu = synthetic_land_case("cuda", 600, math=math)
    cpu = synthetic_land_case("oracle", 600)
    for step in range(1000):
        prev = {n: getattr(gpu.state, n).numpy() for n in ("temperature", "internal_energy", "saturation_water_ice", "pressure_head", "liquid_water_fraction", "surface_excess_water", "skin_temperature", "ground_heat_flux", "infiltration")}
        gpu.step(60.0, 1); cpu.step(60.0, 1)
        bad = None
        for n in ("temperature", "internal_energy", "saturation_water_ice", "pressure_head", "skin_temperature", "ground_heat_flux"):
            a = getattr(gpu.state, n).numpy()
            if not np.all(np.isfinite(a)):
                idx = np.argwhere(~np.isfinite(a))
                print(math, "step", step, "field", n, "first bad idx", idx[:5].tolist(), "count", len(idx))
                bad = idx[0]
                break
        if bad is not None:
            c = bad[-1]
            np.set_printoptions(precision=17, linewidth=200)
            for n, v in prev.items():
                print("prev", n, v[..., c] if v.ndim > 1 else v[c])
            for n in prev:
                a = getattr(gpu.state, n).numpy(); b = getattr(cpu.state, n).numpy()
                print("now gpu", n, a[..., c] if a.ndim > 1 else a[c])
                print("now cpu", n, b[..., c] if b.ndim > 1 else b[c])
            break
    else:
        print(math, "no NaN in 1000 steps")
