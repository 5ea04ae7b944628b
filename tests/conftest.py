import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def c1():
    return np.load(os.path.join(GOLDEN, "c1_st_bb.npz"))


@pytest.fixture(scope="session")
def m2():
    return np.load(os.path.join(GOLDEN, "m2_stu_nsx.npz"))


def rel_err(a, b, floor=1.0e-300):
    """max |a-b| / max|b| per row-scale: elementwise error relative to the
    largest entry of the reference array (entries far below it carry no weight
    in any downstream sum)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), floor))
