"""``sample`` / ``param`` / ``plate`` / ``deterministic`` primitives
(mirror of pyro/primitives.py:125-192,45-110,~400)."""
import torch
from torch.distributions import constraints

from . import params as _params
from .poutine import PlateMessenger
from .poutine.runtime import am_i_wrapped, apply_stack, new_message


def sample(name, fn, *args, obs=None, obs_mask=None, infer=None, **kwargs):
    infer = {} if infer is None else infer.copy()
    if obs_mask is not None:
        raise NotImplementedError("obs_mask is outside the hot-path scope of pyro_b200")
    if not am_i_wrapped():
        if obs is not None:
            return obs
        return fn(*args, **kwargs)
    msg = new_message(type="sample", name=name, fn=fn, is_observed=obs is not None, args=args,
                      kwargs=kwargs, value=obs, infer=infer)
    apply_stack(msg)
    return msg["value"]


def param(name, init_tensor=None, constraint=constraints.real, event_dim=None):
    store = _params.get_param_store()
    args = (name,) if init_tensor is None else (name, init_tensor)
    if not am_i_wrapped():
        return store.get_param(name, init_tensor, constraint, event_dim)
    msg = new_message(type="param", name=name, fn=store.get_param, args=args,
                      kwargs={"constraint": constraint, "event_dim": event_dim})
    apply_stack(msg)
    return msg["value"]


def deterministic(name, value, event_dim=None):
    from .distributions import Delta
    event_dim = value.dim() if event_dim is None else event_dim
    return sample(name, Delta(value, event_dim=event_dim), obs=value,
                  infer={"_deterministic": True})


def factor(name, log_factor, *, has_rsample=None):
    from .distributions import Delta
    unit = Delta(torch.zeros((), device=log_factor.device), log_density=log_factor, event_dim=0)
    unit = unit.expand(log_factor.shape) if log_factor.dim() else unit
    sample(name, unit, obs=torch.zeros((), device=log_factor.device).expand(log_factor.shape),
           infer={"is_auxiliary": True})


class plate(PlateMessenger):
    """Vectorised (``with``) or sequential (``for``) conditional-independence context."""
    pass


def get_param_store():
    return _params.get_param_store()


def clear_param_store():
    _params.clear_param_store()
