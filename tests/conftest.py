import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="session")
def golden_images(golden_dir):
    import numpy as np
    z = np.load(os.path.join(golden_dir, "images.npz"))
    return {"00000": z["img00000"], "00042": z["img00042"]}


@pytest.fixture(scope="session")
def golden_prims(golden_dir):
    import numpy as np
    return np.load(os.path.join(golden_dir, "primitives.npz"))


@pytest.fixture(scope="session")
def golden_drivers(golden_dir):
    import json
    with open(os.path.join(golden_dir, "drivers.json")) as f:
        return json.load(f)
