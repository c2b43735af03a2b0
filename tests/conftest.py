import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """The reference's own saved trajectory (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "cmaes_plot_trajectory.npz"))
    return {k.replace("_", " "): g[k] for k in g.files}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / den
