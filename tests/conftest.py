import os
import sys
import numpy as np
import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tables():
    from roreg_b200 import group
    return group.load()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    meta = z["meta"]
    return z, int(meta[0]), int(meta[1]), int(meta[2]), [int(s) for s in meta[3:]]
