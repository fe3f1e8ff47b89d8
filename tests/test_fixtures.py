"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU oracle):
the oracle must still reproduce them (CPU), and the CUDA path must hit them through the C ABI (-m gpu)."""
import os
import sys

import numpy as np
import pytest

from common import ENGINES, max_scaled_err

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_golden_fixture(engine, name):
    want = np.load(os.path.join(HERE, "golden", name + ".npz"))
    got = make_golden.run(name, engine)
    # the oracle on another libm may differ in the last bits of pow / exp / sin; the CUDA path within the 1e-9 parity bar
    tol = 1.0e-12 if engine == "oracle" else 1.0e-9
    for field in want.files:
        assert np.all(np.isfinite(got[field]))
        assert max_scaled_err(got[field], want[field]) <= tol, (name, field)
