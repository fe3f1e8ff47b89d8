import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np
from common import *
n = 96
def build(engine, math):
    rng = np.random.default_rng(7)
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.1, N=20), n)
    model = trm.SoilModel(grid, soil=richards_soil(vwc_forcing=-2.0e-4))
    sat0 = rng.uniform(0.0, 0.05, (20, n)); sat0[:, ::3] = 0.9
    return make(engine, model, trm.ForwardEuler(dt=60.0), initializers={"temperature": 5.0, "saturation_water_ice": sat0}, math=math)
np.set_printoptions(linewidth=220, precision=8)
cpu = build("oracle", "faithful"); gpu = build("cuda", "faithful")
print("init sat", gpu.state.saturation_water_ice.numpy()[:, 2])
print("init psi gpu", gpu.state.pressure_head.numpy()[:, 2])
print("init psi cpu", cpu.state.pressure_head.numpy()[:, 2])
cpu.step(60.0, 1); gpu.step(60.0, 1)
for nme in ("saturation_water_ice", "pressure_head", "temperature"):
    print(nme, "gpu", getattr(gpu.state, nme).numpy()[:, 2])
    print(nme, "cpu", getattr(cpu.state, nme).numpy()[:, 2])
print("wt", gpu.state.water_table.numpy()[2], cpu.state.water_table.numpy()[2])
