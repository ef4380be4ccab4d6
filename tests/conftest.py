import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def rt():
    """The product binding; the session-scoped context is shared by all GPU tests."""
    from dxrexperiments_b200 import rtcore
    return rtcore


@pytest.fixture(scope="session")
def ctx(rt):
    c = rt.Context(0)
    yield c
    c.close()


def rel_rmse(a: np.ndarray, b: np.ndarray) -> float:
    """Relative RMSE of image a against reference b (the north-star criterion: <= 1e-3)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-30))
