"""Posterior predictive sampling (interface of pyro/infer/predictive.py:173-330,
the subset that does not need enumeration): draw latent sites from ``posterior_samples`` (e.g.
``MCMC.get_samples()``) or from ``guide``, replay the model against them and collect ``return_sites``.

``parallel=True`` wraps the model in an outermost plate of size ``num_samples`` (the model must broadcast
over a leading batch dim, as ``vectorize_particles`` requires) so the whole predictive set is ONE model
execution whose sites are drawn by the kernels' samplers; otherwise the model runs ``num_samples`` times.
"""
import torch

from .. import poutine
from ..primitives import plate


def _guess_max_plate_nesting(model, args, kwargs):
    with poutine.block():
        trace = poutine.trace(model).get_trace(*args, **kwargs)
    dims = [f.dim for s in trace.nodes.values() if s["type"] == "sample"
            for f in s["cond_indep_stack"] if f.vectorized]
    return -min(dims) if dims else 0


class Predictive(torch.nn.Module):
    def __init__(self, model, posterior_samples=None, guide=None, num_samples=None, return_sites=(),
                 parallel=False):
        super().__init__()
        if posterior_samples is None:
            if num_samples is None:
                raise ValueError("Either posterior_samples or num_samples must be specified.")
            posterior_samples = {}
        for name, v in posterior_samples.items():
            batch = v.shape[0]
            if num_samples is None:
                num_samples = batch
            elif num_samples != batch:
                num_samples = batch
        if num_samples is None:
            raise ValueError("No sample sites in posterior samples to infer `num_samples`.")
        if guide is not None and posterior_samples:
            raise ValueError("`posterior_samples` cannot be provided with the `guide` argument.")
        self.model, self.guide = model, guide
        self.posterior_samples = posterior_samples
        self.num_samples = num_samples
        self.return_sites = tuple(return_sites) if return_sites else None
        self.parallel = parallel

    def forward(self, *args, **kwargs):
        with torch.no_grad():
            if self.parallel:
                nest = _guess_max_plate_nesting(self.model, args, kwargs)
                dim = -(nest + 1)
                model = plate("_num_predictive_samples", self.num_samples, dim=dim)(self.model)
                if self.guide is not None:
                    guide = plate("_num_predictive_samples", self.num_samples, dim=dim)(self.guide)
                    guide_trace = poutine.trace(guide).get_trace(*args, **kwargs)
                    model_trace = poutine.trace(poutine.replay(model, trace=guide_trace)).get_trace(*args, **kwargs)
                    fed = set(guide_trace.nodes)
                else:
                    data = {}
                    for name, v in self.posterior_samples.items():
                        # [S, *site_shape] -> sample dim at plate position `dim` of the batch shape
                        pad = nest - (v.dim() - 1)
                        data[name] = v.reshape((v.shape[0],) + (1,) * max(pad, 0) + tuple(v.shape[1:])) if pad > 0 else v
                    model_trace = poutine.trace(poutine.condition(model, data=data)).get_trace(*args, **kwargs)
                    fed = set(data)
                out = {}
                for name, site in model_trace.nodes.items():
                    if site["type"] != "sample" or site["infer"].get("_subsample"):
                        continue
                    if (self.return_sites is not None and name not in self.return_sites) or \
                            (self.return_sites is None and name in fed):
                        continue
                    out[name] = site["value"].detach()
                return out
            out = {}
            for i in range(self.num_samples):
                if self.guide is not None:
                    guide_trace = poutine.trace(self.guide).get_trace(*args, **kwargs)
                    model_trace = poutine.trace(poutine.replay(self.model, trace=guide_trace)).get_trace(*args, **kwargs)
                    fed = set(guide_trace.nodes)
                else:
                    data = {k: v[i] for k, v in self.posterior_samples.items()}
                    model_trace = poutine.trace(poutine.condition(self.model, data=data)).get_trace(*args, **kwargs)
                    fed = set(data)
                for name, site in model_trace.nodes.items():
                    if site["type"] != "sample" or site["infer"].get("_subsample"):
                        continue
                    if (self.return_sites is not None and name not in self.return_sites) or \
                            (self.return_sites is None and name in fed):
                        continue
                    out.setdefault(name, []).append(site["value"].detach())
            return {k: torch.stack(v) for k, v in out.items()}
