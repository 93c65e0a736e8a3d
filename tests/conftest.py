import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


def chain_from_golden(g):
    M = int(g["nslices"])
    return ([g[f"strikes_{m}"] for m in range(M)], [g[f"types_{m}"] for m in range(M)])


@pytest.fixture(scope="session")
def cuda_lib():
    """the CUDA library must be present -- GPU tests fail loudly otherwise (there is no fallback to skip to)."""
    from stochvolmodels_b200 import _capi
    return _capi.load_library()
