import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] >= 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need an sm_100 device: without one they are skipped (never silently passed on a fallback)."""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no sm_100 GPU in this process")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def gpu_ctx():
    import vbmc_b200
    return vbmc_b200.default_context()
