import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    """Build the device library and the CPU oracle in-tree (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as G
    G.build()
    return True


def golden_cases():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    out = []
    for fn in sorted(os.listdir(d)):
        if fn.endswith(".npz") and not fn.startswith("pick"):
            z = np.load(os.path.join(d, fn))
            out.append((fn[:-4], z))
    return out


def pick_golden():
    """tables through the reference's own qselect_eval / fetch metrics (tests/golden/make_golden.py pick_goldens)"""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "pick.npz"))
