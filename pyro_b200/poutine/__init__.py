"""Effect handlers used by the two hot paths: trace, replay, condition, block, scale, mask, plate
(+ broadcast), seed, lift-free param capture.

Mirror of the reference handlers' documented behaviour:
  trace      pyro/poutine/trace_messenger.py:66-200
  replay     pyro/poutine/replay_messenger.py:60-81
  condition  pyro/poutine/condition_messenger.py
  block      pyro/poutine/block_messenger.py
  scale/mask pyro/poutine/scale_messenger.py:52-53, mask_messenger.py
  plate      pyro/poutine/indep_messenger.py:14-120, subsample_messenger.py:90-200,
             broadcast_messenger.py:46-93, plate_messenger.py
"""
from collections import namedtuple
from numbers import Number

import torch

from .runtime import _STACK, Messenger, apply_stack, new_message
from .trace_struct import Trace


# ---------------------------------------------------------------------------------------------
class TraceMessenger(Messenger):
    def __init__(self, graph_type="flat", param_only=False):
        self.graph_type = graph_type
        self.param_only = param_only
        self.trace = Trace()

    def __enter__(self):
        self.trace = Trace()
        return super().__enter__()

    def get_trace(self):
        return self.trace.copy()

    def _pyro_post_sample(self, msg):
        if self.param_only:
            return
        if msg["infer"].get("_do_not_trace"):
            return
        self.trace.add_node(msg["name"], **msg.copy())

    def _pyro_post_param(self, msg):
        self.trace.add_node(msg["name"], **msg.copy())

    def __call__(self, fn):
        return TraceHandler(self, fn)


class TraceHandler:
    def __init__(self, msngr, fn):
        self.fn = fn
        self.msngr = msngr

    def __call__(self, *args, **kwargs):
        with self.msngr:
            self.msngr.trace.add_node("_INPUT", name="_INPUT", type="args", args=args, kwargs=kwargs)
            ret = self.fn(*args, **kwargs)
            self.msngr.trace.add_node("_RETURN", name="_RETURN", type="return", value=ret)
        return ret

    @property
    def trace(self):
        return self.msngr.trace

    def get_trace(self, *args, **kwargs):
        self(*args, **kwargs)
        return self.msngr.get_trace()


class ReplayMessenger(Messenger):
    def __init__(self, trace=None, params=None):
        if trace is None and params is None:
            raise ValueError("must provide trace or params to replay against")
        self.trace = trace
        self.params = params

    def _pyro_sample(self, msg):
        name = msg["name"]
        if self.trace is not None and name in self.trace:
            guide_msg = self.trace.nodes[name]
            if msg["is_observed"]:
                return
            if guide_msg["type"] != "sample" or guide_msg["is_observed"]:
                raise RuntimeError("site {} must be sampled in trace".format(name))
            msg["done"] = True
            msg["value"] = guide_msg["value"]
            msg["infer"] = guide_msg["infer"]

    def _pyro_param(self, msg):
        if self.params is not None and msg["name"] in self.params:
            msg["done"] = True
            msg["value"] = self.params[msg["name"]]


class ConditionMessenger(Messenger):
    def __init__(self, data):
        self.data = data

    def _pyro_sample(self, msg):
        name = msg["name"]
        if isinstance(self.data, Trace):
            if name in self.data.nodes:
                msg["value"] = self.data.nodes[name]["value"]
                msg["is_observed"] = msg["value"] is not None
        elif name in self.data:
            if msg["is_observed"]:
                raise RuntimeError("Cannot condition an already observed site: {}".format(name))
            msg["value"] = self.data[name]
            msg["is_observed"] = msg["value"] is not None


class SubstituteMessenger(Messenger):
    """Replace param values (used to run guides/models on given parameter tensors)."""

    def __init__(self, data):
        self.data = data

    def _pyro_param(self, msg):
        if msg["name"] in self.data:
            msg["value"] = self.data[msg["name"]]
            msg["done"] = True


class BlockMessenger(Messenger):
    def __init__(self, hide_fn=None, expose_fn=None, hide_all=True, expose_all=False, hide=None,
                 expose=None, hide_types=None, expose_types=None):
        if hide_fn is not None:
            self.hide_fn = hide_fn
        elif expose_fn is not None:
            self.hide_fn = lambda msg: not expose_fn(msg)
        else:
            if hide is not None or hide_types is not None:
                hide = hide or []
                hide_types = hide_types or []
                self.hide_fn = lambda msg: msg["name"] in hide or msg["type"] in hide_types
            elif expose is not None or expose_types is not None:
                expose = expose or []
                expose_types = expose_types or []
                self.hide_fn = lambda msg: not (msg["name"] in expose or msg["type"] in expose_types)
            elif expose_all:
                self.hide_fn = lambda msg: False
            else:
                self.hide_fn = lambda msg: True

    def _process_message(self, msg):
        msg["stop"] = bool(self.hide_fn(msg))


class ScaleMessenger(Messenger):
    def __init__(self, scale):
        if isinstance(scale, torch.Tensor):
            if torch._C._get_tracing_state() is None and scale.numel() == 1 and not (scale > 0).all():
                raise ValueError("Expected scale > 0")
        elif not (scale > 0):
            raise ValueError("Expected scale > 0 but got {}".format(scale))
        self.scale = scale

    def _process_message(self, msg):
        msg["scale"] = self.scale * msg["scale"]


class MaskMessenger(Messenger):
    def __init__(self, mask):
        if isinstance(mask, torch.Tensor):
            if mask.dtype != torch.bool:
                raise ValueError("Expected mask to be a BoolTensor but got {}".format(type(mask)))
        elif mask not in (True, False):
            raise ValueError("Expected mask to be a boolean but got {}".format(type(mask)))
        self.mask = mask

    def _process_message(self, msg):
        msg["mask"] = self.mask if msg["mask"] is None else msg["mask"] & self.mask


class SeedMessenger(Messenger):
    def __init__(self, rng_seed):
        self.rng_seed = rng_seed

    def __enter__(self):
        from ..util import get_rng_state, set_rng_seed
        self.old_state = get_rng_state()
        set_rng_seed(self.rng_seed)
        return super().__enter__()

    def __exit__(self, *a):
        from ..util import set_rng_state
        set_rng_state(self.old_state)
        return super().__exit__(*a)


# ---------------------------------------------------------------------------------------------
# plates
# ---------------------------------------------------------------------------------------------
class CondIndepStackFrame(namedtuple("CondIndepStackFrame", ["name", "dim", "size", "counter", "full_size"])):
    @property
    def vectorized(self):
        return self.dim is not None


class _DimAllocator:
    """Tracks which (negative) batch dims are taken by active vectorised plates."""

    def __init__(self):
        self._stack = []  # index i <-> dim -(i+1); entry = plate name or None

    def allocate(self, name, dim):
        if name in self._stack:
            raise ValueError('duplicate plate "{}"'.format(name))
        if dim is None:
            dim = -1
            while -dim <= len(self._stack) and self._stack[-1 - dim] is not None:
                dim -= 1
        elif dim >= 0:
            raise ValueError("Expected dim < 0 to index from the right, actual {}".format(dim))
        while dim < -len(self._stack):
            self._stack.append(None)
        if self._stack[-1 - dim] is not None:
            raise ValueError('at plates "{}" and "{}", collide at dim={}'.format(
                name, self._stack[-1 - dim], dim))
        self._stack[-1 - dim] = name
        return dim

    def free(self, name, dim):
        free_idx = -1 - dim
        assert self._stack[free_idx] == name
        self._stack[free_idx] = None
        while self._stack and self._stack[-1] is None:
            self._stack.pop()


_DIM_ALLOCATOR = _DimAllocator()


class _Subsample:
    """Random subsample indices of a plate (pyro/poutine/subsample_messenger.py:22-87)."""
    has_rsample = False

    def __init__(self, size, subsample_size, device=None):
        self.size = size
        self.subsample_size = subsample_size
        self.device = device

    def __call__(self, sample_shape=torch.Size()):
        n = self.subsample_size
        if n is None or n >= self.size:
            return torch.arange(self.size, device=self.device)
        return torch.randperm(self.size, device=self.device)[:n].clone()

    def log_prob(self, x):
        return torch.zeros((), device=x.device if isinstance(x, torch.Tensor) else None)


class PlateMessenger(Messenger):
    """``pyro.plate``: declares a conditionally independent batch dim, rescales log-probs by
    ``size / subsample_size`` and broadcasts site distributions to the plate size."""

    def __init__(self, name, size=None, subsample_size=None, subsample=None, dim=None,
                 use_cuda=None, device=None):
        self.name = name
        self.dim = dim
        self.device = device
        if size is None:
            assert subsample_size is None and subsample is None
            size = -1
            subsample_size = -1
        self.size = size
        self._requested_subsample_size = subsample_size
        self._given_subsample = subsample
        self.subsample_size = subsample_size if subsample_size is not None else size
        self._indices = subsample
        self.counter = 0
        self._vectorized = None

    # -- subsample draw (a sample site of type "sample" hidden from traces by default pruning) ----
    def _draw_subsample(self):
        if self.size == -1:
            return
        if self._given_subsample is not None:
            self._indices = self._given_subsample
            self.subsample_size = len(self._given_subsample)
            return
        msg = new_message(type="sample", name=self.name,
                          fn=_Subsample(self.size, self._requested_subsample_size, self.device),
                          infer={"_subsample": True})
        apply_stack(msg)
        self._indices = msg["value"]
        self.subsample_size = self._indices.shape[0] if isinstance(self._indices, torch.Tensor) \
            else len(self._indices)

    @property
    def indices(self):
        if self._indices is None and self.size != -1:
            self._indices = torch.arange(self.size, device=self.device)
        return self._indices

    def __enter__(self):
        if self._vectorized is not False:
            self._vectorized = True
        if self._vectorized:
            self.dim = _DIM_ALLOCATOR.allocate(self.name, self.dim)
            self._draw_subsample()
        super().__enter__()
        return self.indices if self.size != -1 else None

    def __exit__(self, *args):
        if self._vectorized:
            _DIM_ALLOCATOR.free(self.name, self.dim)
        return super().__exit__(*args)

    def __iter__(self):
        # sequential plate
        self._vectorized = False
        self.dim = None
        self._draw_subsample()
        idx = self.indices
        n = self.subsample_size if self.size != -1 else 0
        for i in (idx.tolist() if isinstance(idx, torch.Tensor) else range(n)):
            self.counter += 1
            with self:
                yield i if isinstance(i, Number) else i

    def _process_message(self, msg):
        frame = CondIndepStackFrame(self.name, self.dim, self.subsample_size, self.counter, self.size)
        msg["cond_indep_stack"] = (frame,) + msg["cond_indep_stack"]
        if self.size != -1 and self.subsample_size != self.size:
            msg["scale"] = msg["scale"] * self.size / self.subsample_size
        if msg["type"] == "sample":
            _broadcast_site(msg)

    def _postprocess_message(self, msg):
        if msg["type"] in ("param", "subsample") and self.dim is not None:
            event_dim = msg["kwargs"].get("event_dim")
            if event_dim is not None and self.size != -1:
                dim = self.dim - event_dim
                shape = msg["value"].shape
                if len(shape) >= -dim and shape[dim] != 1:
                    if self.subsample_size < self.size:
                        value = msg["value"]
                        new_value = value.index_select(dim, self.indices.to(value.device))
                        if msg["type"] == "param":
                            if hasattr(value, "_pyro_unconstrained_param"):
                                param = value._pyro_unconstrained_param
                            else:
                                param = value.unconstrained()
                            if not hasattr(param, "_pyro_subsample"):
                                param._pyro_subsample = {}
                            param._pyro_subsample[dim] = self.indices
                            new_value._pyro_unconstrained_param = param
                        msg["value"] = new_value


def _broadcast_site(msg):
    """Expand the site's distribution to the sizes of the enclosing vectorised plates
    (pyro/poutine/broadcast_messenger.py:46-93)."""
    if msg["done"] or msg["type"] != "sample" or not hasattr(msg["fn"], "expand"):
        return
    if msg["infer"].get("_subsample"):
        return
    dist = msg["fn"]
    actual = dist.batch_shape
    target = [None if s == 1 else s for s in actual]
    for f in msg["cond_indep_stack"]:
        if f.dim is None or f.size == -1:
            continue
        assert f.dim < 0
        target = [None] * (-f.dim - len(target)) + target
        if target[f.dim] is not None and target[f.dim] != f.size:
            raise ValueError("Shape mismatch inside plate('{}') at site {} dim {}, {} vs {}".format(
                f.name, msg["name"], f.dim, f.size, target[f.dim]))
        target[f.dim] = f.size
    for i in range(-len(target) + 1, 1):
        if target[i] is None:
            target[i] = actual[i] if len(actual) >= -i else 1
    if tuple(target) != tuple(actual):
        had = dist.has_rsample
        msg["fn"] = dist.expand(target)
        if msg["fn"].has_rsample != had:
            msg["fn"].has_rsample = had


# ---------------------------------------------------------------------------------------------
# functional forms
# ---------------------------------------------------------------------------------------------
def _make_handler(cls):
    def handler(fn=None, *args, **kwargs):
        if fn is not None and not (callable(fn) or isinstance(fn, Trace)) and cls is not TraceMessenger:
            raise ValueError("{} is not callable, did you mean to pass it as a keyword arg?".format(fn))
        msngr = cls(*args, **kwargs)
        return msngr(fn) if fn is not None else msngr
    handler.__name__ = cls.__name__
    return handler


def trace(fn=None, graph_type="flat", param_only=False):
    msngr = TraceMessenger(graph_type=graph_type, param_only=param_only)
    return msngr(fn) if fn is not None else msngr


def replay(fn=None, trace=None, params=None):
    msngr = ReplayMessenger(trace=trace, params=params)
    return msngr(fn) if fn is not None else msngr


def condition(fn=None, data=None):
    msngr = ConditionMessenger(data=data)
    return msngr(fn) if fn is not None else msngr


def substitute(fn=None, data=None):
    msngr = SubstituteMessenger(data=data)
    return msngr(fn) if fn is not None else msngr


def block(fn=None, **kwargs):
    msngr = BlockMessenger(**kwargs)
    return msngr(fn) if fn is not None else msngr


def scale(fn=None, scale=None):
    msngr = ScaleMessenger(scale=scale)
    return msngr(fn) if fn is not None else msngr


def mask(fn=None, mask=None):
    msngr = MaskMessenger(mask=mask)
    return msngr(fn) if fn is not None else msngr


def seed(fn=None, rng_seed=None):
    msngr = SeedMessenger(rng_seed)
    return msngr(fn) if fn is not None else msngr


def prune_subsample_sites(trace):
    """Copy of ``trace`` without the plates' subsample-index sites (pyro/poutine/util.py:40-48)."""
    trace = trace.copy()
    for name, site in list(trace.nodes.items()):
        if site["type"] == "sample" and site["infer"].get("_subsample"):
            trace.remove_node(name)
    return trace


__all__ = ["trace", "replay", "condition", "substitute", "block", "scale", "mask", "seed",
           "Trace", "Messenger", "TraceMessenger", "ReplayMessenger", "ConditionMessenger",
           "BlockMessenger", "ScaleMessenger", "MaskMessenger", "PlateMessenger",
           "CondIndepStackFrame", "prune_subsample_sites", "apply_stack", "new_message"]
