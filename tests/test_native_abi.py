"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/pyro_b200.h declares; entry points validate their arguments without launching."""
import ctypes
import os
import re

import pytest

from pyro_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "pyro_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(N.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert set(names) == set(N.SIGNATURES), set(names) ^ set(N.SIGNATURES)


def test_argument_validation_without_gpu():
    L = N.lib()
    assert L.b2_version() >= 100
    assert L.b2_last_error(-3).decode().startswith("unknown distribution family")
    assert L.b2_site_score_workspace() > 0
    t = N.b2_tensor()
    # bad family / null pointers are rejected before any CUDA call
    assert L.b2_site_score(999, ctypes.byref(t), ctypes.byref(t), 1, None, 1.0, None, 1.0, 1.0, 0,
                           None, None, None, None, None, 0, None) == -3
    assert L.b2_site_score(0, None, None, 2, None, 1.0, None, 1.0, 1.0, 0, None, None, None, None,
                           None, 0, None) == -4
    assert L.b2_glm_bernoulli_logits(None, None, None, None, 10, 32, 8, 1.0, 1.0, 1.0, 0, None, None,
                                     None, None, None, 0, None) == -4
    assert L.b2_nuts_small(None, None, None, None, None, None, 1, 1, 10, 1000.0, 0, None, None, None,
                           None, None, None, None) == -4


def test_cpu_tensors_are_refused_loudly():
    import torch
    import pyro_b200.distributions as dist
    with pytest.raises(RuntimeError, match="no CPU"):
        dist.Normal(torch.zeros(3), torch.ones(3)).log_prob(torch.zeros(3))
