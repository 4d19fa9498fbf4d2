import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` need a CUDA device (the product has no CPU path): skip them where there is none."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA (sm_100) device; boxfusion_b200 has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _restore_oracle_backend():
    """Tests pick the oracle's IoU backend (scipy/Qhull like the reference, or its C restatement) by assigning the module
    global; put it back so that no test depends on the order they run in."""
    from oracle import port
    saved = port.IOU_BACKEND
    yield
    port.IOU_BACKEND = saved
