"""Launch list of a vegetated LandModel Heun step on the N72 column count (14 017 columns, Float32): which launch costs what on
a small domain. usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv python profiles/land_small_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from test_vegetation import synthetic_vegetated_case  # noqa: E402

g = synthetic_vegetated_case("cuda", 14017, nf=np.float32, heun=True, math="fast")
g.step(60.0, 6)
print("ok", float(g.state.temperature.numpy().mean()))
