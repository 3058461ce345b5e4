"""ORACLE (test infrastructure).  Optimiser update rules, restated with CPU tensor ops in the
reference's operation order."""
import math

import torch


class ClippedAdam:
    """pyro/optim/clipped_adam.py:52-100 for one parameter tensor (PyroOptim keeps one optimizer per
    parameter, pyro/optim/optim.py:125-155)."""

    def __init__(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_norm=10.0, lrd=1.0):
        self.lr, self.betas, self.eps = lr, betas, eps
        self.weight_decay, self.clip_norm, self.lrd = weight_decay, clip_norm, lrd
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None

    def step(self, p, grad):
        """In-place update of ``p`` (a leaf's .data) given ``grad``."""
        self.lr *= self.lrd                                             # :62
        grad = grad.clamp(-self.clip_norm, self.clip_norm)              # :68
        if self.exp_avg is None:                                        # :72-77
            self.exp_avg = torch.zeros_like(grad)
            self.exp_avg_sq = torch.zeros_like(grad)
        beta1, beta2 = self.betas
        self.step_count += 1                                            # :82
        if self.weight_decay != 0:                                      # :84-85
            grad = grad.add(p, alpha=self.weight_decay)
        self.exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)            # :88
        self.exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)  # :89
        denom = self.exp_avg_sq.sqrt().add_(self.eps)                   # :91
        bc1 = 1 - beta1 ** self.step_count                              # :93-95
        bc2 = 1 - beta2 ** self.step_count
        step_size = self.lr * math.sqrt(bc2) / bc1
        p.addcdiv_(self.exp_avg, denom, value=-step_size)               # :97
        return p


class AdagradRMSProp:
    """pyro/optim/adagrad_rmsprop.py:54-87 for one parameter tensor."""

    def __init__(self, eta=1.0, delta=1.0e-16, t=0.1):
        self.eta, self.delta, self.t = eta, delta, t
        self.step_count = 0
        self.sum = None

    def step(self, p, grad):
        self.step_count += 1
        if self.step_count == 1:
            self.sum = grad * grad
        else:
            self.sum *= 1.0 - self.t
            self.sum += self.t * grad * grad
        lr = self.eta * (self.step_count ** (-0.5 + self.delta))
        std = self.sum.sqrt()
        p.addcdiv_(grad, 1.0 + std, value=-lr)
        return p
