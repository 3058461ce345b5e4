"""SVI driver (mirror of pyro/infer/svi.py:38-162).

``step`` = capture the parameters touched by the loss, ``loss_and_grads``, one fused optimiser
call; gradients are zeroed inside the optimiser kernel.  With a loss whose ``capture_graph`` is
true (``JitTrace_ELBO``) the whole step is captured into a CUDA graph on the second call and
replayed afterwards, so the per-step host cost is one graph launch plus one 4-byte read-back.
"""
import warnings

import torch

from .. import poutine
from ..util import torch_item, warn_if_nan
from .elbo import ELBO


class SVI:
    _poutine = poutine   # pyro_b200/bind.py substitutes the reference's poutine for reference models

    def __init__(self, model, guide, optim, loss, loss_and_grads=None, num_samples=0, num_steps=0,
                 **kwargs):
        if num_steps:
            warnings.warn("The `num_steps` argument to SVI is deprecated", FutureWarning)
        self.model = model
        self.guide = guide
        self.optim = optim
        self.num_steps = num_steps
        self.num_samples = num_samples
        self._loss_obj = loss
        if isinstance(loss, ELBO):
            self.loss = loss.loss
            self.loss_and_grads = loss.loss_and_grads
            self._loss_and_grads_tensor = getattr(loss, "loss_and_grads_tensor", None)
        else:
            if loss_and_grads is None:
                def _loss_and_grads(model, guide, *args, **kwargs):
                    loss_val = loss(model, guide, *args, **kwargs)
                    if getattr(loss_val, "requires_grad", False):
                        loss_val.backward(retain_graph=True)
                    return loss_val
                loss_and_grads = _loss_and_grads
            self.loss = loss
            self.loss_and_grads = loss_and_grads
            self._loss_and_grads_tensor = None
        self._capture = bool(getattr(loss, "capture_graph", False))
        self._graph = None
        self._graph_state = None
        self._steps_done = 0

    # ---- evaluation ---------------------------------------------------------------------------------
    def evaluate_loss(self, *args, **kwargs):
        with torch.no_grad():
            loss = self.loss(self.model, self.guide, *args, **kwargs)
            return loss if isinstance(loss, float) else torch_item(loss)

    # ---- one eager step -----------------------------------------------------------------------------
    def _grads(self, args, kwargs, want_tensor=True):
        """Phase 1: loss + gradients.  Returns (loss, [unconstrained parameters touched])."""
        with self._poutine.trace(param_only=True) as param_capture:
            if want_tensor and self._loss_and_grads_tensor is not None:
                loss = self._loss_and_grads_tensor(self.model, self.guide, *args, **kwargs)
            else:
                loss = self.loss_and_grads(self.model, self.guide, *args, **kwargs)
        params = []
        seen = set()
        for site in param_capture.trace.nodes.values():
            if site["type"] != "param":
                continue
            v = site["value"]
            u = getattr(v, "_pyro_unconstrained_param", None)
            if u is None:
                u = v.unconstrained() if hasattr(v, "unconstrained") else v
            if id(u) not in seen:
                seen.add(id(u))
                params.append(u)
        return loss, params

    @staticmethod
    def _world():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size()
        return 1

    def _pack(self, params, loss):
        grads = [p.grad for p in params if p.grad is not None]
        return torch.cat([loss.detach().reshape(1).to(grads[0].dtype)] + [g.reshape(-1) for g in grads])

    def _unpack(self, params, flat):
        off = 1
        for p in params:
            if p.grad is None:
                continue
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].reshape(p.grad.shape))
            off += n
        return flat[0]

    def _eager_step(self, args, kwargs, want_tensor=False):
        loss, params = self._grads(args, kwargs, want_tensor)
        self._last_params = params
        loss = self._allreduce(params, loss)
        self.optim(params)  # fused update; zeroes the gradients in the same pass
        return loss

    def _allreduce(self, params, loss):
        """Data parallelism: every rank scores a shard (of the particle plate, or of a data plate
        whose ``size/subsample_size`` rescaling keeps each rank's ELBO unbiased), so the ELBO
        estimate and its gradient are the MEAN over ranks.  ONE all-reduce per step over a packed
        buffer [loss, grad_1 .. grad_n] (the reference's only analogue is the
        per-parameter Horovod all-reduce of pyro/optim/horovod.py:41-45 +
        examples/svi_horovod.py:134); the replicated fused optimiser then keeps the parameters
        identical on every rank."""
        import torch.distributed as dist
        if self._world() == 1:
            return loss
        if not isinstance(loss, torch.Tensor) or not any(p.grad is not None for p in params):
            return loss
        flat = self._pack(params, loss)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= dist.get_world_size()
        return self._unpack(params, flat)

    def step(self, *args, **kwargs):
        """One gradient step; returns the loss estimate as a python float."""
        if self._capture and self._steps_done >= 1:
            loss = self._captured_step(args, kwargs)
        else:
            loss = self._eager_step(args, kwargs, want_tensor=True)
        self._steps_done += 1
        if isinstance(loss, torch.Tensor):
            loss = torch_item(loss)
        warn_if_nan(loss, "loss")
        return loss

    def step_async(self, *args, **kwargs):
        """Like ``step`` but returns the loss as a 0-d device tensor without synchronising."""
        if self._capture and self._steps_done >= 1:
            loss = self._captured_step(args, kwargs)
        else:
            loss = self._eager_step(args, kwargs, want_tensor=True)
        self._steps_done += 1
        return loss

    # ---- CUDA-graph captured step ------------------------------------------------------------------
    def _captured_step(self, args, kwargs):
        if kwargs:
            raise ValueError("graph-captured SVI steps take tensor positional arguments only")
        if self._graph is None:
            # the capturing call: ONE real (eager) update whose loss is returned, then the capture itself,
            # which records kernels without running them -- exactly one update per step() call, as in
            # pyro/infer/svi.py:134-162
            return self._capture_graph(args)
        st = self._graph_state
        if len(args) != len(st["static_args"]):
            raise ValueError("graph-captured SVI step called with a different number of arguments")
        for i, (a, s) in enumerate(zip(args, st["static_args"])):
            if isinstance(a, torch.Tensor):
                if a.shape != s.shape or a.dtype != s.dtype:
                    raise ValueError("graph-captured SVI step called with different argument shapes")
                if a.data_ptr() != s.data_ptr():
                    if not st["owned"][i]:
                        # the graph reads the caller's FIRST tensor in place (zero-copy for a resident data
                        # set).  A different tensor now arrives: never overwrite the caller's memory --
                        # re-capture once with private input buffers, then copy into those every step.
                        self._graph = None
                        self._graph_state = None
                        return self._capture_graph(args, private=True)
                    s.copy_(a, non_blocking=True)
            elif a != s:
                raise ValueError("non-tensor argument of a graph-captured SVI step changed")
        if getattr(self.optim, "graph_epoch", 0) != st["optim_epoch"]:
            # optimiser state was replaced (set_state / load): its device tables moved
            self._graph = None
            self._graph_state = None
            return self._capture_graph(args, private=any(st["owned"]))
        if st.get("graph_b") is None:
            self._graph.replay()
            return st["loss"]
        # multi-rank: [graph A: loss + grads into the flat payload] -> NCCL all-reduce -> [graph B: optimiser]
        import torch.distributed as dist
        self._graph.replay()
        dist.all_reduce(st["flat"], op=dist.ReduceOp.SUM)
        st["graph_b"].replay()
        return st["loss"]

    def _capture_graph(self, args, private=False):
        """Run ONE eager step on ``args`` (its loss is the return value of this call), then capture the
        step into a CUDA graph.  Capture records kernels without executing them, so parameters and
        optimiser state advance exactly once."""
        dev = None
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
        if dev is None:
            raise RuntimeError("graph capture needs CUDA tensor arguments")
        if self._loss_and_grads_tensor is None:
            raise RuntimeError("this loss cannot be captured (no loss_and_grads_tensor)")
        static_args, owned = [], []
        for a in args:
            if isinstance(a, torch.Tensor):
                if private or not a.is_cuda:
                    static_args.append(a.to(dev, copy=True))
                    owned.append(True)
                else:
                    static_args.append(a)           # zero-copy: read in place, never written
                    owned.append(False)
            else:
                static_args.append(a)
                owned.append(True)
        # the real step of this call, on a side stream so that every lazily created buffer exists
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        from ..distributions import _ops as _dops
        _dops.SLOT_LEAVES.clear()
        with torch.cuda.stream(side):
            eager_loss = self._eager_step(tuple(static_args), {}, want_tensor=True)
        slot_leaves = set(_dops.SLOT_LEAVES)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if isinstance(eager_loss, torch.Tensor):
            eager_loss = eager_loss.detach().clone()
        graph = torch.cuda.CUDAGraph()
        flush = getattr(self.optim, "flush_pending", None)
        state = {"static_args": static_args, "owned": owned,
                 "optim_epoch": getattr(self.optim, "graph_epoch", 0)}
        if self._world() == 1:
            with torch.cuda.graph(graph):
                if flush is not None:
                    # no stored gradients during the captured backward: autograd hands each parameter
                    # its fresh gradient tensor (no accumulate launch per parameter); the optimiser's
                    # pointer table is re-pointed below
                    # (parameters whose gradient a kernel adds straight into .grad -- the latent-sites backward --
                    # keep their buffer: the optimiser zeroed it, nothing is accumulated by the engine)
                    for p in self._last_params:
                        if id(p) not in slot_leaves:
                            p.grad = None
                loss = self._eager_step(tuple(static_args), {}, want_tensor=True)
                loss = loss.reshape(()) if isinstance(loss, torch.Tensor) else torch.as_tensor(loss, device=dev)
            if flush is not None:
                flush()
            state["loss"] = loss
        else:
            # The collective stays outside the graphs: capture the two halves around it.  The
            # gradients are re-pointed to views of ONE flat buffer [loss, grad_1 .. grad_n], so the
            # backward pass accumulates straight into the all-reduce payload and the optimiser reads
            # it back in place -- no pack / unpack copies on either side of the collective.
            world = self._world()
            live = [p for p in self._last_params if p.grad is not None]   # from the eager step above
            total = 1 + sum(p.grad.numel() for p in live)
            flat = torch.zeros(total, dtype=live[0].grad.dtype, device=dev)
            off = 1
            for p in live:
                n = p.grad.numel()
                p.grad = flat[off:off + n].view(p.grad.shape)
                off += n
            # the optimiser builds its pointer tables for the new gradient storage now (no launch),
            # not during capture
            prepare = getattr(self.optim, "prepare", None)
            if prepare is not None:
                prepare(live)
            import os
            import torch.distributed as dist
            single = None
            if os.environ.get("B2_NCCL_IN_GRAPH", "0") == "1" and dist.get_backend() == "nccl":
                # OPT-IN (B2_NCCL_IN_GRAPH=1): ONE graph for the whole step -- NCCL collectives are capturable,
                # so the all-reduce of the [loss, grads] payload sits between the backward and the optimiser
                # inside the graph and a replay is a single launch.  Two precautions (the first 2-GPU run of
                # round 2 hung without them): the capture is thread-local, so the CUDA calls of NCCL's watchdog
                # thread cannot invalidate it on one rank only, and the ranks AGREE (eager all-reduce of a flag)
                # on whether every capture succeeded before any of them replays a graph with a collective in it.
                ok = 1.0
                try:
                    g1 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g1, capture_error_mode="thread_local"):
                        loss, params2 = self._grads(tuple(static_args), {}, True)
                        flat[0:1].copy_(loss.detach().reshape(1).to(flat.dtype))
                        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                        flat /= world
                        out = flat[0].clone()
                        self.optim(params2)
                    single = (g1, out)
                except Exception:  # noqa: BLE001 -- fall back to the two-graph form below
                    ok = 0.0
                    single = None
                    torch.cuda.synchronize(dev)
                agree = torch.tensor([ok], device=dev)
                dist.all_reduce(agree, op=dist.ReduceOp.MIN)
                if float(agree) < 1.0:
                    single = None
            if single is not None:
                graph, out = single
                state.update({"loss": out, "flat": flat, "graph_b": None, "nccl_in_graph": True})
            else:
                with torch.cuda.graph(graph):
                    loss, params2 = self._grads(tuple(static_args), {}, True)
                    flat[0:1].copy_(loss.detach().reshape(1).to(flat.dtype))
                graph_b = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph_b, pool=graph.pool()):
                    flat /= world
                    out = flat[0].clone()
                    self.optim(params2)
                state.update({"loss": out, "flat": flat, "graph_b": graph_b, "nccl_in_graph": False})
        self._graph = graph
        self._graph_state = state
        return eager_loss
