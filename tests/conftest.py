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


def _cuda_usable() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain ``pytest`` on a box without a usable CUDA device skips the gpu-marked tests instead of failing 65 of them on
    'CUDA driver version is insufficient'.  On the B200 box (or with B200SV_REQUIRE_GPU=1) nothing is skipped: a missing GPU or a
    missing libb200sv.so must fail loudly there -- there is no CPU fallback to hide behind."""
    if os.environ.get("B200SV_REQUIRE_GPU") == "1" or _cuda_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (set B200SV_REQUIRE_GPU=1 to make this an error)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
