"""Global parameter store (mirror of pyro/params/param_store.py:125-156,276-336).

Parameters live as UNCONSTRAINED leaf tensors (``requires_grad``); the constrained value is
recomputed with ``torch.distributions.transform_to(constraint)`` at every ``pyro.param`` call, so
the fused optimiser updates the unconstrained storage in place.
"""
import weakref

import torch
from torch.distributions import constraints, transform_to


class ParamStoreDict:
    def __init__(self):
        self._params = {}        # name -> unconstrained leaf
        self._constraints = {}   # name -> constraint
        self._param_to_name = {}

    def clear(self):
        self._params = {}
        self._constraints = {}
        self._param_to_name = {}

    def __contains__(self, name):
        return name in self._params

    def __len__(self):
        return len(self._params)

    def keys(self):
        return self._params.keys()

    def items(self):
        for name in self._params:
            yield name, self[name]

    def named_parameters(self):
        return self._params.items()

    def get_all_param_names(self):
        return set(self._params.keys())

    def setdefault(self, name, init_constrained_value, constraint=constraints.real):
        if name not in self._params:
            if callable(init_constrained_value):
                init_constrained_value = init_constrained_value()
            self.__setitem__(name, init_constrained_value, constraint)
        return self[name]

    def __setitem__(self, name, value, constraint=None):
        if constraint is None:
            constraint = self._constraints.get(name, constraints.real)
        with torch.no_grad():
            unconstrained = transform_to(constraint).inv(value.detach()).contiguous().clone()
        unconstrained.requires_grad_(True)
        self._params[name] = unconstrained
        self._constraints[name] = constraint
        self._param_to_name[unconstrained] = name

    def __getitem__(self, name):
        unconstrained = self._params[name]
        constraint = self._constraints[name]
        if constraint is constraints.positive:
            # transform_to(positive) = Affine(0, 1) o Exp: the affine part is the identity, skip its
            # two launches (and two more in backward) per parameter per step
            from . import _native as N
            if N.LAZY_PARAM and (unconstrained.is_cuda or N.EMULATE_RSAMPLE):
                # exp deferred: a Normal guide site's draw kernel takes log(scale) as it is stored
                from ._lazyparam import LazyExpParam
                constrained = LazyExpParam(unconstrained)
                constrained.unconstrained = weakref.ref(unconstrained)
                constrained._pyro_unconstrained_param = unconstrained
                return constrained
            constrained = unconstrained.exp()
        else:
            constrained = transform_to(constraint)(unconstrained)
        constrained.unconstrained = weakref.ref(unconstrained)
        constrained._pyro_unconstrained_param = unconstrained
        return constrained

    def get_param(self, name, init_tensor=None, constraint=constraints.real, event_dim=None):
        if init_tensor is None:
            return self[name]
        return self.setdefault(name, init_tensor, constraint)

    def param_name(self, p):
        return self._param_to_name.get(p)

    def match(self, name):
        import re
        pattern = re.compile(name)
        return {n: self[n] for n in self._params if pattern.match(n)}

    # -- checkpointing (pyro/params/param_store.py:276-336): same on-disk schema ------------------
    def get_state(self):
        return {"params": {k: v.detach() for k, v in self._params.items()},
                "constraints": dict(self._constraints)}

    def set_state(self, state):
        assert set(state.keys()) == {"params", "constraints"}
        for name, p in state["params"].items():
            c = state["constraints"][name]
            u = p.detach().clone().requires_grad_(True)
            self._params[name] = u
            self._constraints[name] = c
            self._param_to_name[u] = name

    def save(self, filename):
        with open(filename, "wb") as f:
            torch.save(self.get_state(), f)

    def load(self, filename, map_location=None):
        with open(filename, "rb") as f:
            state = torch.load(f, map_location=map_location, weights_only=False)
        self.set_state(state)


_PARAM_STORE = ParamStoreDict()


def get_param_store():
    return _PARAM_STORE


def clear_param_store():
    _PARAM_STORE.clear()
