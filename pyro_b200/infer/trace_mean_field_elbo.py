"""TraceMeanField_ELBO with analytic KL terms on the fused kernels.

Restates pyro/infer/trace_mean_field_elbo.py:104-156: observed sites contribute their
``log_prob_sum``; each latent site contributes ``-KL(guide_fn || model_fn)`` (scaled and masked
like the site), falling back to ``log_prob_sum - entropy_term.sum()`` when no analytic KL is
registered.  Normal||Normal and Gamma||Gamma KLs (torch/distributions/kl.py:301-306,468-471) are
evaluated, summed and differentiated by ONE fused kernel per site.
"""
import torch

from ..distributions import fused_kl_sum, kl_divergence, scale_and_mask
from ..poutine.trace_struct import _fused
from ..util import torch_item, warn_if_nan
from .trace_elbo import Trace_ELBO


def _kl_site_args(model_site, guide_site):
    scale = model_site["scale"]
    mask = model_site["mask"]
    if isinstance(scale, torch.Tensor):
        # tensor-valued scale: reference path (keeps it differentiable; no device sync, capture-safe)
        from .. import _native as N
        if scale.numel() != 1 or scale.requires_grad or (scale.is_cuda and N.capturing()):
            scale = None
        else:
            scale = float(scale)
    if mask is True:
        mask = None
    return scale, mask


class TraceMeanField_ELBO(Trace_ELBO):
    def loss(self, model, guide, *args, **kwargs):
        with torch.no_grad():
            loss = self.loss_and_grads_tensor(model, guide, *args, _no_backward=True, **kwargs)
        loss = torch_item(loss)
        warn_if_nan(loss, "loss")
        return loss if loss is not None else 0.0

    def _mean_field_particle(self, model_trace, guide_trace):
        """(loss contribution ``-elbo/P`` as a 0-d tensor, [unit-upstream terms])"""
        P = self.num_particles
        terms, elbo_terms = [], []

        def push(t, coeff_already_applied=True):
            elbo_terms.append(t.detach())
            if t.requires_grad:
                terms.append(t)

        for name, model_site in model_trace.nodes.items():
            if model_site["type"] != "sample":
                continue
            if model_site["is_observed"]:
                t = _fused(model_site, weight=-1.0 / P, sum_coeff=1.0, unit=True)
                if t is None:
                    lp = scale_and_mask(model_site["fn"].log_prob(model_site["value"]),
                                        model_site["scale"], model_site["mask"]).sum()
                    elbo_terms.append(lp.detach())
                    if lp.requires_grad:
                        terms.append((-1.0 / P) * lp)
                else:
                    push(t)
                continue
            guide_site = guide_trace.nodes[name]
            scale, mask = _kl_site_args(model_site, guide_site)
            t = None
            if scale is not None and mask is not False:
                # elbo -= KL ; surrogate_loss += KL / P
                t = fused_kl_sum(guide_site["fn"], model_site["fn"], mask, scale, weight=1.0 / P,
                                 sum_coeff=-1.0, unit=True)
            if t is not None:
                push(t)
                continue
            try:
                kl_qp = kl_divergence(guide_site["fn"], model_site["fn"])
                kl_qp = scale_and_mask(kl_qp, scale=guide_site["scale"], mask=guide_site["mask"]).sum()
                elbo_terms.append(-kl_qp.detach())
                if kl_qp.requires_grad:
                    terms.append((1.0 / P) * kl_qp)
            except NotImplementedError:
                # entropy-term fallback (trace_mean_field_elbo.py:133-139)
                tm = _fused(model_site, weight=-1.0 / P, sum_coeff=1.0, unit=True)
                tg = _fused(guide_site, weight=1.0 / P, sum_coeff=-1.0, unit=True)
                for t2, site, c in ((tm, model_site, 1.0), (tg, guide_site, -1.0)):
                    if t2 is not None:
                        push(t2)
                    else:
                        lp = scale_and_mask(site["fn"].log_prob(site["value"]), site["scale"], site["mask"]).sum()
                        elbo_terms.append(c * lp.detach())
                        if lp.requires_grad:
                            terms.append((-c / P) * lp)
        if not elbo_terms:
            return torch.zeros(()), terms
        # loss contribution -elbo/P assembled on the device in one launch (b2_elbo_combine)
        from ..distributions import _ops
        ref = elbo_terms[0]
        parts = [e.reshape(()).to(ref.dtype) if e.dtype != ref.dtype else e.reshape(()) for e in elbo_terms]
        return _ops.elbo_combine(parts, [-1.0 / P] * len(parts)), terms

    def _score_and_backward(self, model_trace, guide_trace, allow_general=False, _no_backward=False):
        for name, site in guide_trace.nodes.items():
            if site["type"] == "sample" and not getattr(site["fn"], "has_rsample", False):
                raise ValueError("TraceMeanField_ELBO requires fully reparameterised guides; "
                                 "site '{}' is not".format(name))
        loss_particle, terms = self._mean_field_particle(model_trace, guide_trace)
        trainable = any(site["type"] == "param" for trace in (model_trace, guide_trace)
                        for site in trace.nodes.values())
        if trainable and terms and not _no_backward:
            from .trace_elbo import _one_like
            torch.autograd.backward(terms, [_one_like(t) for t in terms], retain_graph=self.retain_graph)
        return loss_particle

    def loss_and_grads_tensor(self, model, guide, *args, _no_backward=False, **kwargs):
        loss = None
        for model_trace, guide_trace in self._get_traces(model, guide, args, kwargs):
            loss_particle = self._score_and_backward(model_trace, guide_trace, _no_backward=_no_backward)
            loss = loss_particle if loss is None else loss + loss_particle
        return loss if loss is not None else 0.0
