"""ORACLE (test infrastructure).  The Trace_ELBO SVI step, restated for CPU tensors.

``trace_elbo_from_sites`` follows pyro/infer/trace_elbo.py:82-112 for an explicit list of scored
sites (reparameterised guides): loss = -(sum_model lp - sum_guide lp)/P and
surrogate = -(sum_model lp - sum_guide entropy_term)/P with entropy_term = lp.

``LogisticSVI`` is BASELINE config 2 end to end as reference Pyro executes it on CPU
(model tests/infer/mcmc/test_hmc.py:189-198 scaled as SURVEY.md 8d; vectorised particles
pyro/infer/elbo.py:186-216; per-site log_prob + scale_and_mask + sum
pyro/poutine/trace_struct.py:248-328; backward pyro/infer/trace_elbo.py:153-157; one ClippedAdam
per parameter pyro/optim/optim.py:125-155; zero_grads pyro/infer/util.py:85-91).  The guide noise
can be injected so the GPU path and the reference can be compared on identical draws
(the replay technique of tests/infer/test_gradient.py:77-91).
"""
import torch

from . import dists
from .optim import ClippedAdam


def trace_elbo_from_sites(model_lps, guide_lps, num_particles):
    """model_lps / guide_lps: lists of already scaled+masked log_prob tensors.
    Returns (loss float tensor, surrogate_loss tensor)."""
    elbo = 0.0
    surrogate = 0.0
    for lp in model_lps:
        s = lp.sum()
        elbo = elbo + s.detach()
        surrogate = surrogate + s
    for lp in guide_lps:
        s = lp.sum()
        elbo = elbo - s.detach()
        surrogate = surrogate - s
    return -elbo / num_particles, -surrogate / num_particles


class LogisticSVI:
    """Bayesian logistic regression, mean-field Normal guide, vectorised Trace_ELBO + ClippedAdam.

    model:  w ~ Normal(0, 1).to_event(1) [D];  b ~ Normal(0, 10);
            with plate("data", N): y ~ Bernoulli(logits = X w + b)
    guide:  w ~ Normal(w_loc, w_scale).to_event(1);  b ~ Normal(b_loc, b_scale)
            scales are pyro.param(..., constraint=positive): stored as log, constrained by exp
            (pyro/params/param_store.py:125-156, torch transform_to(positive) = ExpTransform).
    """

    def __init__(self, D, num_particles, lr=0.01, dtype=torch.float32, init_scale=0.1, **adam):
        self.P = num_particles
        self.D = D
        self.dtype = dtype
        # unconstrained parameters, in the order the guide declares them
        self.params = {
            "w_loc": torch.zeros(D, dtype=dtype, requires_grad=True),
            "w_scale": torch.full((D,), float(init_scale), dtype=dtype).log().requires_grad_(True),
            "b_loc": torch.zeros((), dtype=dtype, requires_grad=True),
            "b_scale": torch.tensor(float(init_scale), dtype=dtype).log().requires_grad_(True),
        }
        self.optims = {k: ClippedAdam(lr=lr, **adam) for k in self.params}

    def loss_and_grads(self, X, y, eps_w=None, eps_b=None):
        P, D = self.P, self.D
        w_loc, b_loc = self.params["w_loc"], self.params["b_loc"]
        w_scale, b_scale = self.params["w_scale"].exp(), self.params["b_scale"].exp()
        if eps_w is None:
            eps_w = torch.randn(P, 1, D, dtype=self.dtype)
        if eps_b is None:
            eps_b = torch.randn(P, 1, dtype=self.dtype)
        # guide rsample: loc + eps*scale (torch/distributions/normal.py:82-85), particle plate
        # at dim -2 (max_plate_nesting = 1 + 1)
        w = w_loc + eps_w * w_scale            # [P, 1, D]
        b = b_loc + eps_b * b_scale            # [P, 1]
        guide_lps = [dists.normal(w, w_loc, w_scale), dists.normal(b, b_loc, b_scale)]
        zero = torch.zeros((), dtype=self.dtype)
        logits = (w * X).sum(-1) + b           # [P, N]  (X: [N, D] broadcast against [P, 1, D])
        model_lps = [dists.normal(w, zero, torch.ones((), dtype=self.dtype)),
                     dists.normal(b, zero, torch.full((), 10.0, dtype=self.dtype)),
                     dists.bernoulli_logits(y, logits)]
        loss, surrogate = trace_elbo_from_sites(model_lps, guide_lps, P)
        surrogate.backward()
        return loss

    def step(self, X, y, eps_w=None, eps_b=None):
        loss = self.loss_and_grads(X, y, eps_w, eps_b)
        with torch.no_grad():
            for k, p in self.params.items():
                self.optims[k].step(p.data, p.grad)
                p.grad = torch.zeros_like(p.grad)  # zero_grads
        return float(loss)

    def constrained(self):
        return {"w_loc": self.params["w_loc"].detach().clone(),
                "w_scale": self.params["w_scale"].detach().exp(),
                "b_loc": self.params["b_loc"].detach().clone(),
                "b_scale": self.params["b_scale"].detach().exp()}


class LogisticSVIMatmul(LogisticSVI):
    """Same step with the linear predictor as a matmul (``X @ w.T``), the formulation a user would
    write for large N; used as the CPU baseline timing because broadcasting ``w * X`` materialises
    a [P, N, D] tensor."""

    def loss_and_grads(self, X, y, eps_w=None, eps_b=None):
        P, D = self.P, self.D
        w_loc, b_loc = self.params["w_loc"], self.params["b_loc"]
        w_scale, b_scale = self.params["w_scale"].exp(), self.params["b_scale"].exp()
        if eps_w is None:
            eps_w = torch.randn(P, 1, D, dtype=self.dtype)
        if eps_b is None:
            eps_b = torch.randn(P, 1, dtype=self.dtype)
        w = w_loc + eps_w * w_scale
        b = b_loc + eps_b * b_scale
        guide_lps = [dists.normal(w, w_loc, w_scale), dists.normal(b, b_loc, b_scale)]
        zero = torch.zeros((), dtype=self.dtype)
        logits = w.squeeze(-2) @ X.T + b       # [P, N]
        model_lps = [dists.normal(w, zero, torch.ones((), dtype=self.dtype)),
                     dists.normal(b, zero, torch.full((), 10.0, dtype=self.dtype)),
                     dists.bernoulli_logits(y, logits)]
        loss, surrogate = trace_elbo_from_sites(model_lps, guide_lps, P)
        surrogate.backward()
        return loss
