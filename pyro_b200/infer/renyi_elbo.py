"""RenyiELBO (another estimator on the same kernels).

Restates pyro/infer/renyi_elbo.py:97-137 (``loss``) and :139-236 (``loss_and_grads``): per-particle ELBOs
``elbo_p = sum_sites log p - sum_sites log q`` summed over the dependent plate dims
(pyro/infer/util.py:108-119), ``elbo = log mean_p exp((1 - alpha) elbo_p) / (1 - alpha)`` and the surrogate
``sum_p w_p surrogate_p / P`` with the self-normalised weights ``w_p``.  Unlike Trace_ELBO this estimator
needs PER-PARTICLE sums with a per-particle upstream weight, so every site goes through the materialised
``log_prob`` of the fused kernels (one launch forward, one launch backward with the upstream gradient folded
in) instead of the sum-only fast path.
"""
import math

import torch

from ..distributions import is_identically_zero
from ..util import torch_item, warn_if_nan
from .elbo import ELBO, get_importance_trace


def get_dependent_plate_dims(sites):
    """pyro/infer/util.py:108-119: dims of the plates that are not common to all sample sites."""
    plate_sets = [site["cond_indep_stack"] for site in sites if site["type"] == "sample"]
    all_plates = set().union(*plate_sets)
    common_plates = all_plates.intersection(*plate_sets)
    return sorted({f.dim for f in (all_plates - common_plates) if f.dim is not None})


def torch_sum(tensor, dims):
    """pyro/util.py torch_sum: sum out ``dims`` (negative, counted from the right) that exist."""
    assert all(d < 0 for d in dims)
    leftmost = -tensor.dim()
    dims = [d for d in dims if leftmost <= d]
    return tensor.sum(dims) if dims else tensor


class RenyiELBO(ELBO):
    def __init__(self, alpha=0, num_particles=2, max_plate_nesting=float("inf"), max_iarange_nesting=None,
                 vectorize_particles=False, strict_enumeration_warning=True):
        if alpha == 1:
            raise ValueError("The order alpha should not be equal to 1. Please use Trace_ELBO class"
                             "for the case alpha = 1.")
        self.alpha = alpha
        super().__init__(num_particles=num_particles, max_plate_nesting=max_plate_nesting,
                         max_iarange_nesting=max_iarange_nesting, vectorize_particles=vectorize_particles,
                         strict_enumeration_warning=strict_enumeration_warning)

    def _get_trace(self, model, guide, args, kwargs):
        return get_importance_trace("flat", self.max_plate_nesting, model, guide, args, kwargs)

    def _particles(self, model, guide, args, kwargs, grads):
        elbo_particles, surrogate_particles = [], []
        traces = None
        for model_trace, guide_trace in self._get_traces(model, guide, args, kwargs):
            traces = (model_trace, guide_trace)
            elbo_particle, surrogate = 0.0, 0.0
            sum_dims = get_dependent_plate_dims(model_trace.nodes.values())
            for name, site in model_trace.nodes.items():
                if site["type"] == "sample":
                    lps = torch_sum(site["log_prob"], sum_dims)
                    elbo_particle = elbo_particle + lps.detach()
                    surrogate = surrogate + lps
            for name, site in guide_trace.nodes.items():
                if site["type"] == "sample":
                    log_prob, score_function_term, entropy_term = site["score_parts"]
                    lps = torch_sum(site["log_prob"], sum_dims)
                    elbo_particle = elbo_particle - lps.detach()
                    if not is_identically_zero(entropy_term):
                        surrogate = surrogate - lps
                        if not is_identically_zero(score_function_term):
                            raise NotImplementedError   # pyro issue 1222, as in the reference
                    if not is_identically_zero(score_function_term):
                        surrogate = surrogate + (self.alpha / (1.0 - self.alpha)) * lps
            elbo_particles.append(elbo_particle)
            surrogate_particles.append(surrogate)
        if traces is None or not isinstance(elbo_particles[0], torch.Tensor):
            return None, None, traces
        if self.vectorize_particles and self.num_particles > 1:
            return elbo_particles[0], surrogate_particles[0], traces
        return torch.stack(elbo_particles), torch.stack(surrogate_particles), traces

    def loss(self, model, guide, *args, **kwargs):
        with torch.no_grad():
            elbo_particles, _, _ = self._particles(model, guide, args, kwargs, False)
            if elbo_particles is None:
                return 0.0
            log_weights = (1.0 - self.alpha) * elbo_particles
            log_mean_weight = torch.logsumexp(log_weights, dim=0) - math.log(self.num_particles)
            loss = -torch_item(log_mean_weight.sum()) / (1.0 - self.alpha)
        warn_if_nan(loss, "loss")
        return loss

    def loss_and_grads(self, model, guide, *args, **kwargs):
        elbo_particles, surrogate_particles, traces = self._particles(model, guide, args, kwargs, True)
        if elbo_particles is None:
            return 0.0
        log_weights = (1.0 - self.alpha) * elbo_particles
        log_mean_weight = torch.logsumexp(log_weights, dim=0, keepdim=True) - math.log(self.num_particles)
        elbo = torch_item(log_mean_weight.sum()) / (1.0 - self.alpha)
        trainable = any(site["type"] == "param" for trace in traces for site in trace.nodes.values())
        if trainable and getattr(surrogate_particles, "requires_grad", False):
            normalized_weights = (log_weights - log_mean_weight).exp()
            surrogate_elbo = (normalized_weights * surrogate_particles).sum() / self.num_particles
            (-surrogate_elbo).backward()
        loss = -elbo
        warn_if_nan(loss, "loss")
        return loss
