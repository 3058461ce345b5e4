import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native libraries exist (nvcc cross-compiles without a GPU)."""
    from pyro_b200 import _build
    _build.build()


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(autouse=True)
def _clean_state():
    import pyro_b200
    pyro_b200.clear_param_store()
    torch.set_default_dtype(torch.float32)
    yield
    pyro_b200.clear_param_store()
    from pyro_b200.poutine import runtime
    del runtime._STACK[:]
    torch.set_default_dtype(torch.float32)
