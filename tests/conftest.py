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


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device: skip them on a machine without one (plain `pytest` here).  On a GPU box they always
    run - a missing libroreg_b200.so then FAILS them loudly (ops.Context raises), it never skips."""
    try:
        import torch
        ok = torch.cuda.is_available()
    except Exception:
        ok = False
    if ok:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def tables():
    from roreg_b200 import group
    return group.load()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    meta = z["meta"]
    return z, int(meta[0]), int(meta[1]), int(meta[2]), [int(s) for s in meta[3:]]
