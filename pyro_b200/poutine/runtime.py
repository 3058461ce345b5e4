"""Effect-handler runtime: the messenger stack and the sample/param primitives' dispatch.

Host-side mirror of pyro/poutine/runtime.py:334-390 (``default_process_message``, ``apply_stack``)
and pyro/poutine/messenger.py.  north_star keeps this machinery in Python; it only has to exist
on the GPU box (where reference Pyro is not installed) so that unchanged model/guide code can
drive the fused kernels.  Written from the documented semantics, not copied.
"""
import functools

_STACK = []          # active messengers, outermost first
_PARAM_STORE = None  # set by pyro_b200.params


def am_i_wrapped():
    return len(_STACK) > 0


def new_message(**kw):
    msg = dict(type=None, name=None, fn=None, is_observed=False, args=(), kwargs={}, value=None,
               infer={}, scale=1.0, mask=None, cond_indep_stack=(), done=False, stop=False,
               continuation=None)
    msg.update(kw)
    return msg


def default_process_message(msg):
    """Run the site function unless some handler already produced a value."""
    if msg["done"] or msg["is_observed"] or msg["value"] is not None:
        msg["done"] = True
        return msg
    msg["value"] = msg["fn"](*msg["args"], **msg["kwargs"])
    msg["done"] = True
    return msg


def apply_stack(msg):
    """Send ``msg`` down the stack (innermost handler first), apply the default behaviour, then
    let the visited handlers post-process it on the way back up."""
    visited = 0
    for handler in reversed(_STACK):
        visited += 1
        handler._process_message(msg)
        if msg["stop"]:
            break
    default_process_message(msg)
    for handler in _STACK[len(_STACK) - visited:]:
        handler._postprocess_message(msg)
    cont = msg["continuation"]
    if cont is not None:
        cont(msg)
    return msg


class Messenger:
    """Context manager that sees every message issued while it is active."""

    def __call__(self, fn):
        if not callable(fn):
            raise ValueError("{} is not callable, did you mean to pass it as a keyword arg?".format(fn))
        handler = self

        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            with handler:
                return fn(*args, **kwargs)

        wrapped.msngr = handler
        wrapped.fn = fn
        return wrapped

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, exc_type, exc_value, tb):
        if exc_type is None:
            if _STACK and _STACK[-1] is self:
                _STACK.pop()
            else:
                raise ValueError("This Messenger is not on the bottom of the stack")
        else:
            # unwind through (and including) this handler
            if self in _STACK:
                i = _STACK.index(self)
                del _STACK[i:]
        return None

    def _process_message(self, msg):
        fn = getattr(self, "_pyro_" + msg["type"], None)
        if fn is not None:
            fn(msg)

    def _postprocess_message(self, msg):
        fn = getattr(self, "_pyro_post_" + msg["type"], None)
        if fn is not None:
            fn(msg)

    def _reset(self):
        pass


class _BoundMessengerFn:
    """``handler(fn)`` object exposing ``get_trace`` style helpers (used by TraceHandler)."""
    pass
