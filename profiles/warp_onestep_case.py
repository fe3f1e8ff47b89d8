import os, sys
sys.path.insert(0, 'tests')
import numpy as np
from common import synthetic_soil_case
os.environ["TRM_WARP_COLS"] = "1000000"
g = synthetic_soil_case('cuda', 8192, nf=np.float32, richards=True, math='fast')
for _ in range(8):
    g.step(60.0, 1)
print("ok")
