import sys
sys.path.insert(0, 'tests')
import numpy as np
from common import synthetic_soil_case
g = synthetic_soil_case('cuda', 14017, nf=np.float32, richards=True, math='fast')
g.step(60.0, 2)
g.step(60.0, 200)
print("ok", float(g.state.temperature.numpy().mean()))
