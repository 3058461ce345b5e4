"""Plate-aware bookkeeping for the score-function (non-reparameterised) ELBO terms.

Behavioural mirror of pyro/infer/util.py:94-105 (``get_plate_stacks``) and :122-165
(``MultiFrameTensor``): a dictionary from sets of vectorised plate frames to the running sum of
tensors that live in exactly those plates, with ``sum_to`` reducing every entry down to the plates
a target site lives in (the Rao-Blackwellisation used by Trace_ELBO).
"""
import torch


def zero_grads(tensors):
    for p in tensors:
        if p.grad is not None:
            p.grad = torch.zeros_like(p.grad)


def get_plate_stacks(trace):
    out = {}
    for name, node in trace.nodes.items():
        if node["type"] == "sample" and not node["infer"].get("_subsample"):
            out[name] = [f for f in node["cond_indep_stack"] if f.vectorized]
    return out


class MultiFrameTensor(dict):
    def __init__(self, *items):
        super().__init__()
        self.add(*items)

    def add(self, *items):
        for stack, value in items:
            key = frozenset(f for f in stack if f.vectorized)
            for f in key:
                if not (f.dim < 0 and -value.dim() <= f.dim):
                    raise ValueError("plate dim {} out of range for tensor of shape {}".format(
                        f.dim, tuple(value.shape)))
            self[key] = self[key] + value if key in self else value

    def sum_to(self, target_frames):
        total = None
        for frames, value in self.items():
            for f in frames:
                if f not in target_frames and value.shape[f.dim] != 1:
                    value = value.sum(f.dim, keepdim=True)
            while value.dim() > 0 and value.shape[0] == 1:
                value = value.squeeze(0)
            total = value if total is None else total + value
        return 0.0 if total is None else total
