import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing at the first
    one (the product path itself still fails loudly there: tests/test_capi_cpu.py checks that)."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (gpu-marked test)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def world():
    from yag_slam_b200 import synth
    return synth.make_world()
