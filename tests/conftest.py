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
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_prepare():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "prepare_ref.npz"))


@pytest.fixture(scope="session")
def golden_radar():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "radar_ref.npz"))


@pytest.fixture(scope="session")
def golden_temporal():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "temporal_ref.npz"))
