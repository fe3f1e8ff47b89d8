import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


# The GPU suite was written against the one-thread-per-column streaming kernels at small column counts, and those kernels
# serve the 10 M-column benchmark: it keeps running on them. The warp-per-column kernel that the library picks for small
# domains by default (csrc/warp_kernel.cuh) is covered by tests/test_warp_kernel.py, which switches it on around its runs
# and replays the parity tests of the other modules through it. TRM_TEST_WARP=1 runs the WHOLE suite on the warp path.
os.environ["TRM_WARP"] = "1" if os.environ.get("TRM_TEST_WARP") == "1" else "0"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def trm():
    import terrarium_jl_b200
    return terrarium_jl_b200


@pytest.fixture(scope="session")
def oracle():
    import oracle_integrator
    oracle_integrator.oracle_library()
    return oracle_integrator
