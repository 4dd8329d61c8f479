import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box")


def pytest_collection_modifyitems(config, items):
    """GPU tests are only collected as runnable when a device exists; they never silently pass on CPU."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    return load


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure libgillb200.so exists (build() is cheap when it is up to date)."""
    lib = os.path.join(ROOT, "gill_b200", "libgillb200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g

        g.build()
