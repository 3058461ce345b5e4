"""Small host utilities (RNG state, NaN warnings): pyro/util.py:48-63, pyro/infer/util.py:48-91."""
import random
import warnings

import numpy as np
import torch


def set_rng_seed(rng_seed):
    torch.manual_seed(rng_seed)
    random.seed(rng_seed)
    np.random.seed(rng_seed)
    from .distributions import _ops
    _ops.reseed(rng_seed)   # the in-kernel Philox streams of the fused draws


def get_rng_state():
    return {"torch": torch.get_rng_state(), "random": random.getstate(),
            "numpy": np.random.get_state()}


def set_rng_state(state):
    torch.set_rng_state(state["torch"])
    random.setstate(state["random"])
    np.random.set_state(state["numpy"])


def torch_item(x):
    """Python number of a 0-d tensor (ONE device sync) or the number itself."""
    return x if isinstance(x, (int, float)) else x.item()


def warn_if_nan(value, msg=""):
    if isinstance(value, float):
        if value != value:
            warnings.warn("Encountered NaN{}".format(": " + msg if msg else "."), stacklevel=2)
    return value
