"""Optimisers: ``ClippedAdam`` and ``AdagradRMSProp`` as ONE fused multi-tensor launch.

Interface mirrors pyro/optim/optim.py:72-198 (``PyroOptim``: constructed from an args dict or a
per-parameter callable, called as ``optim(params)``, ``get_state/set_state/save/load``) and the
update rules of pyro/optim/clipped_adam.py:52-100 and pyro/optim/adagrad_rmsprop.py:54-87.

Differences in mechanism, not in results: the reference keeps one ``torch.optim`` object per
parameter and steps them in a Python loop (optim.py:125-155); here all parameters of one dtype
are updated by one kernel pair, per-parameter scalars (step count, decayed lr) live in device
arrays so the captured CUDA graph of an SVI step can be replayed, and gradients are zeroed in the
same pass (the reference reallocates zeros, pyro/infer/util.py:85-91).
"""
import ctypes

import torch

from .. import _native as N
from ..params import get_param_store


class PyroOptim:
    """Base: bookkeeping shared by the fused optimisers."""

    _state_names = ()
    _store = staticmethod(get_param_store)   # pyro_b200/bind.py points this at the reference's param store

    def __init__(self, optim_args, clip_args=None):
        if not (callable(optim_args) or isinstance(optim_args, dict)):
            raise ValueError("optim_args must be function that returns defaults or a defaults dictionary")
        if clip_args is not None and not (callable(clip_args) or isinstance(clip_args, dict)):
            raise ValueError("clip_args must be function that returns defaults or a defaults dictionary")
        self.pt_optim_args = optim_args
        self.pt_clip_args = clip_args
        self._host = {}      # param -> {"args": dict, "step": int, "lr": float, tensors...}
        self._tables = {}    # dtype -> table dict
        self._retired = []   # replaced tables: a captured CUDA graph may still read their device arrays
        self.graph_epoch = 0  # bumped when state is replaced; SVI re-captures its graph when it changes
        self._state_waiting_to_be_consumed = {}

    # ---- per-parameter hyper-parameters ---------------------------------------------------------
    def _args_for(self, param):
        if callable(self.pt_optim_args):
            name = self._store().param_name(param)
            return dict(self.pt_optim_args(name))
        return dict(self.pt_optim_args)

    def _clip_for(self, param):
        if self.pt_clip_args is None:
            return None
        if callable(self.pt_clip_args):
            return self.pt_clip_args(self._store().param_name(param))
        return self.pt_clip_args

    def _init_param(self, p):
        raise NotImplementedError

    def _ensure(self, p):
        if p not in self._host:
            self._host[p] = self._init_param(p)
            name = self._store().param_name(p)
            waiting = self._state_waiting_to_be_consumed.pop(name, None)
            if waiting is not None:
                self._load_one(p, waiting)
        return self._host[p]

    # ---- stepping ----------------------------------------------------------------------------------
    def __call__(self, params, *args, **kwargs):
        params = [p for p in params if p.grad is not None]
        if not params:
            return
        for p in params:
            N.require_cuda(p, type(self).__name__)
            self._ensure(p)
            clip = self._clip_for(p)
            if clip:
                # pyro/optim/optim.py:227-267 gradient clipping hooks (torch utilities)
                if "clip_norm" in clip:
                    torch.nn.utils.clip_grad_norm_(p, clip["clip_norm"])
                if "clip_value" in clip:
                    torch.nn.utils.clip_grad_value_(p, clip["clip_value"])
        by_dtype = {}
        for p in params:
            by_dtype.setdefault(p.dtype, []).append(p)
        for dtype, ps in by_dtype.items():
            table = self._table_for(dtype, ps)
            self._launch(table)

    def prepare(self, params):
        """Build the device tables for ``params`` (with their CURRENT gradient tensors) without launching
        anything -- used before a CUDA-graph capture so that the capture itself allocates nothing."""
        params = [p for p in params if p.grad is not None]
        by_dtype = {}
        for p in params:
            N.require_cuda(p, type(self).__name__)
            self._ensure(p)
            by_dtype.setdefault(p.dtype, []).append(p)
        for dtype, ps in by_dtype.items():
            self._table_for(dtype, ps)

    def _table_key(self, ps):
        return tuple((id(p), p.data_ptr(), p.grad.data_ptr()) for p in ps)

    def _table_for(self, dtype, ps):
        key = self._table_key(ps)
        table = self._tables.get(dtype)
        if table is not None and table["key"] == key:
            return table
        if N.capturing():
            # Same parameters, new gradient tensors: the captured step cleared ``.grad`` so that
            # autograd ASSIGNS fresh gradients (graph-pool tensors with replay-stable addresses)
            # instead of launching one accumulate kernel per parameter.  The kernels read the gradient
            # pointers from the device table at replay time, so the table's contents are re-pointed
            # right after capture (``flush_pending``); nothing executes during capture itself.
            if table is not None and [k[:2] for k in table["key"]] == [k[:2] for k in key]:
                table["pending"] = list(ps)
                return table
            raise RuntimeError("pyro_b200.optim: parameter set changed during CUDA graph capture")
        if table is not None:
            self._sync_to_host(table)
            self._retired.append(table)   # keep its device arrays alive (baked into captured graphs)
            del self._retired[:-8]
        dev = ps[0].device
        n = len(ps)

        def ptrs(ts):
            return torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)

        table = {"key": key, "params": list(ps), "n": n, "dtype": dtype, "device": dev,
                 "p": ptrs(ps), "g": ptrs([p.grad for p in ps]),
                 "numel": torch.tensor([p.numel() for p in ps], dtype=torch.int64, device=dev),
                 "max_numel": max(p.numel() for p in ps)}
        for sname in self._state_names:
            table[sname] = ptrs([self._host[p][sname] for p in ps])
        self._fill_scalars(table)
        self._tables[dtype] = table
        return table

    def flush_pending(self):
        """After a CUDA-graph capture: point the device tables at the gradient tensors the captured
        backward pass produced."""
        for table in self._tables.values():
            ps = table.pop("pending", None)
            if ps is None:
                continue
            ptrs = torch.tensor([p.grad.data_ptr() for p in ps], dtype=torch.int64)
            table["g"].copy_(ptrs.to(table["device"]))
            table["key"] = self._table_key(ps)

    def _fill_scalars(self, table):
        raise NotImplementedError

    def _sync_to_host(self, table):
        raise NotImplementedError

    def _launch(self, table):
        raise NotImplementedError

    # ---- state (pyro/optim/optim.py:157-198; torch.optim state_dict schema per parameter name) -----
    def _state_dict_one(self, p):
        raise NotImplementedError

    def _load_one(self, p, state):
        raise NotImplementedError

    def get_state(self):
        for table in self._tables.values():
            self._sync_to_host(table)
        out = {}
        for p in self._host:
            out[self._store().param_name(p)] = self._state_dict_one(p)
        return out

    def set_state(self, state_dict):
        self._state_waiting_to_be_consumed.update(state_dict)
        # parameters already being optimised take their state immediately
        for p in list(self._host):
            name = self._store().param_name(p)
            if name in self._state_waiting_to_be_consumed:
                self._load_one(p, self._state_waiting_to_be_consumed.pop(name))
        # The device tables are refreshed IN PLACE (step counts, learning rates, hyper-parameters): a
        # captured CUDA graph reads them through baked-in addresses, so they must neither move nor be
        # freed.  ``graph_epoch`` additionally tells SVI to re-capture.
        for table in self._tables.values():
            old = {k: table[k] for k in ("hyper", "lrs", "steps") if k in table}
            self._fill_scalars(table)
            for k, t_old in old.items():
                t_old.copy_(table[k])
                table[k] = t_old
        self.graph_epoch += 1

    def save(self, filename):
        with open(filename, "wb") as f:
            torch.save(self.get_state(), f)

    def load(self, filename, map_location=None):
        with open(filename, "rb") as f:
            state = torch.load(f, map_location=map_location, weights_only=False)
        self.set_state(state)


def _dt(dtype):
    return N.B2_F32 if dtype == torch.float32 else N.B2_F64


class ClippedAdam(PyroOptim):
    """pyro/optim/clipped_adam.py: Adam with element-wise gradient clamp (``clip_norm``) and
    multiplicative lr decay (``lrd``) applied before every update."""

    _state_names = ("exp_avg", "exp_avg_sq")
    _defaults = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, clip_norm=10.0, lrd=1.0)

    def _init_param(self, p):
        args = dict(self._defaults)
        args.update(self._args_for(p))
        return {"args": args, "step": 0, "lr": float(args["lr"]),
                "exp_avg": torch.zeros_like(p, memory_format=torch.contiguous_format),
                "exp_avg_sq": torch.zeros_like(p, memory_format=torch.contiguous_format)}

    def _fill_scalars(self, table):
        ps, dev = table["params"], table["device"]
        hyper = torch.zeros(len(ps), 8, dtype=torch.float64)
        for i, p in enumerate(ps):
            a = self._host[p]["args"]
            hyper[i, 0], hyper[i, 1] = a["betas"]
            hyper[i, 2] = a["eps"]
            hyper[i, 3] = a["weight_decay"]
            hyper[i, 4] = a["clip_norm"]
            hyper[i, 5] = a["lrd"]
        table["hyper"] = hyper.to(dev)
        table["lrs"] = torch.tensor([self._host[p]["lr"] for p in ps], dtype=torch.float64, device=dev)
        table["steps"] = torch.tensor([self._host[p]["step"] for p in ps], dtype=torch.int32, device=dev)

    def _sync_to_host(self, table):
        lrs = table["lrs"].tolist()
        steps = table["steps"].tolist()
        for p, lr, st in zip(table["params"], lrs, steps):
            self._host[p]["lr"] = lr
            self._host[p]["step"] = st

    def _launch(self, t):
        N.check(N.lib().b2_clipped_adam(
            t["n"], t["p"].data_ptr(), t["g"].data_ptr(), t["exp_avg"].data_ptr(),
            t["exp_avg_sq"].data_ptr(), t["numel"].data_ptr(), t["hyper"].data_ptr(),
            t["lrs"].data_ptr(), t["steps"].data_ptr(), _dt(t["dtype"]), 1, t["max_numel"],
            N.stream_ptr(t["device"])), "b2_clipped_adam")

    def _state_dict_one(self, p):
        h = self._host[p]
        a = h["args"]
        return {"state": {0: {"step": h["step"], "exp_avg": h["exp_avg"], "exp_avg_sq": h["exp_avg_sq"]}},
                "param_groups": [{"lr": h["lr"], "betas": tuple(a["betas"]), "eps": a["eps"],
                                  "weight_decay": a["weight_decay"], "clip_norm": a["clip_norm"],
                                  "lrd": a["lrd"], "params": [0]}]}

    def _load_one(self, p, state):
        h = self._host[p]
        st = state["state"].get(0, {}) if state.get("state") else {}
        if st:
            h["step"] = int(st["step"])
            h["exp_avg"].copy_(st["exp_avg"])
            h["exp_avg_sq"].copy_(st["exp_avg_sq"])
        g = state["param_groups"][0]
        h["lr"] = float(g["lr"])
        for k in ("betas", "eps", "weight_decay", "clip_norm", "lrd"):
            if k in g:
                h["args"][k] = g[k]


class AdagradRMSProp(PyroOptim):
    """pyro/optim/adagrad_rmsprop.py (``eta``, ``delta``, ``t``)."""

    _state_names = ("sum",)
    _defaults = dict(eta=1.0, delta=1.0e-16, t=0.1)

    def _init_param(self, p):
        args = dict(self._defaults)
        args.update(self._args_for(p))
        return {"args": args, "step": 0,
                "sum": torch.zeros_like(p, memory_format=torch.contiguous_format)}

    def _fill_scalars(self, table):
        ps, dev = table["params"], table["device"]
        hyper = torch.zeros(len(ps), 4, dtype=torch.float64)
        for i, p in enumerate(ps):
            a = self._host[p]["args"]
            hyper[i, 0], hyper[i, 1], hyper[i, 2] = a["eta"], a["delta"], a["t"]
        table["hyper"] = hyper.to(dev)
        table["steps"] = torch.tensor([self._host[p]["step"] for p in ps], dtype=torch.int32, device=dev)

    def _sync_to_host(self, table):
        for p, st in zip(table["params"], table["steps"].tolist()):
            self._host[p]["step"] = st

    def _launch(self, t):
        N.check(N.lib().b2_adagrad_rmsprop(
            t["n"], t["p"].data_ptr(), t["g"].data_ptr(), t["sum"].data_ptr(),
            t["numel"].data_ptr(), t["hyper"].data_ptr(), t["steps"].data_ptr(), _dt(t["dtype"]), 1,
            t["max_numel"], N.stream_ptr(t["device"])), "b2_adagrad_rmsprop")

    def _state_dict_one(self, p):
        h = self._host[p]
        a = h["args"]
        return {"state": {0: {"step": h["step"], "sum": h["sum"]}},
                "param_groups": [{"eta": a["eta"], "delta": a["delta"], "t": a["t"], "params": [0]}]}

    def _load_one(self, p, state):
        h = self._host[p]
        st = state["state"].get(0, {}) if state.get("state") else {}
        if st:
            h["step"] = int(st["step"])
            h["sum"].copy_(st["sum"])
        g = state["param_groups"][0]
        for k in ("eta", "delta", "t"):
            if k in g:
                h["args"][k] = g[k]


__all__ = ["PyroOptim", "ClippedAdam", "AdagradRMSProp"]
