"""Minimal stand-in for the `opt_einsum` package, which reference Pyro imports at module load
(pyro/poutine/trace_struct.py:21, pyro/infer/util.py:11-12) but which is not installed in this
offline image.  Only used by tests/golden/make_golden.py to import the UNMODIFIED reference from
/root/reference; the hot paths exercised there never enumerate, so nothing here is ever called
for real work."""
import contextlib

import torch

from . import sharing  # noqa: F401

_SYMBOLS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def get_symbol(i):
    return _SYMBOLS[i] if i < len(_SYMBOLS) else chr(i + 140)


@contextlib.contextmanager
def shared_intermediates(cache=None):
    yield {} if cache is None else cache


def contract(equation, *operands, **kwargs):
    return torch.einsum(equation, *operands)


def contract_expression(*args, **kwargs):
    raise NotImplementedError("opt_einsum stand-in: contract_expression is not available")


def contract_path(*args, **kwargs):
    raise NotImplementedError("opt_einsum stand-in: contract_path is not available")
