"""ELBO base class: particle loop / vectorised-particle plate and trace generation.

Mirror of pyro/infer/elbo.py:108-237 (constructor arguments, ``_guess_max_plate_nesting``,
``_vectorized_num_particles``, ``_get_traces``) and of the importance-trace builder
pyro/infer/enum.py:45-85.
"""
import warnings
from abc import ABCMeta, abstractmethod

from .. import poutine
from ..primitives import plate


LAZY_LINEAR = True   # hand latent values to the model as lazy-aware tensors (pyro_b200/lazy.py)


def get_importance_trace(graph_type, max_plate_nesting, model, guide, args, kwargs, detach=False,
                         score=True):
    """Run the guide, replay the model against it, prune subsample sites.  With ``score`` the
    traces are scored the reference way (``compute_log_prob`` / ``compute_score_parts``); the fused
    ELBO path passes ``score=False`` and scores each site with one fused kernel instead."""
    guide_trace = poutine.trace(guide, graph_type=graph_type).get_trace(*args, **kwargs)
    if detach:
        guide_trace.detach_()
    if not score and LAZY_LINEAR:
        # the model sees latent values as SiteValue tensors, so an unchanged `w @ X.T + b` likelihood
        # reaches the fused GLM kernel (pyro_b200/lazy.py); plain tensors again before scoring
        from ..lazy import unwrap_site_values, wrap_site_values
        wrap_site_values(guide_trace)
        try:
            model_trace = poutine.trace(poutine.replay(model, trace=guide_trace),
                                        graph_type=graph_type).get_trace(*args, **kwargs)
        finally:
            unwrap_site_values(guide_trace)
        unwrap_site_values(model_trace)
    else:
        model_trace = poutine.trace(poutine.replay(model, trace=guide_trace),
                                    graph_type=graph_type).get_trace(*args, **kwargs)
    check_model_guide_match(model_trace, guide_trace, max_plate_nesting)
    guide_trace = poutine.prune_subsample_sites(guide_trace)
    model_trace = poutine.prune_subsample_sites(model_trace)
    if score:
        model_trace.compute_log_prob()
        guide_trace.compute_score_parts()
    return model_trace, guide_trace


def check_model_guide_match(model_trace, guide_trace, max_plate_nesting=float("inf")):
    """Host validation (pyro/util.py check_model_guide_match): every unobserved guide site must
    appear in the model, and vice versa."""
    guide_vars = set(n for n, s in guide_trace.nodes.items()
                     if s["type"] == "sample" and not s["infer"].get("_subsample"))
    aux_vars = set(n for n, s in guide_trace.nodes.items()
                   if s["type"] == "sample" and s["infer"].get("is_auxiliary"))
    model_vars = set(n for n, s in model_trace.nodes.items()
                     if s["type"] == "sample" and not s["is_observed"]
                     and not s["infer"].get("_subsample"))
    if not (guide_vars - aux_vars <= set(n for n, s in model_trace.nodes.items() if s["type"] == "sample")):
        warnings.warn("Found vars in guide but not model: {}".format(guide_vars - model_vars - aux_vars))
    if not (model_vars <= guide_vars):
        warnings.warn("Found vars in model but not guide: {}".format(model_vars - guide_vars))


class ELBO(object, metaclass=ABCMeta):
    def __init__(self, num_particles=1, max_plate_nesting=float("inf"), max_iarange_nesting=None,
                 vectorize_particles=False, strict_enumeration_warning=True,
                 ignore_jit_warnings=False, jit_options=None, retain_graph=None,
                 tail_adaptive_beta=-1.0):
        if max_iarange_nesting is not None:
            warnings.warn("max_iarange_nesting is deprecated; use max_plate_nesting instead",
                          DeprecationWarning)
            max_plate_nesting = max_iarange_nesting
        self.max_plate_nesting = max_plate_nesting
        self.num_particles = num_particles
        self.vectorize_particles = vectorize_particles
        self.retain_graph = retain_graph
        if self.vectorize_particles and self.num_particles > 1:
            self.max_plate_nesting += 1
        self.strict_enumeration_warning = strict_enumeration_warning
        self.ignore_jit_warnings = ignore_jit_warnings
        self.jit_options = jit_options
        self.tail_adaptive_beta = tail_adaptive_beta

    def _guess_max_plate_nesting(self, model, guide, args, kwargs):
        with poutine.block():
            guide_trace = poutine.trace(guide).get_trace(*args, **kwargs)
            model_trace = poutine.trace(poutine.replay(model, trace=guide_trace)).get_trace(*args, **kwargs)
        guide_trace = poutine.prune_subsample_sites(guide_trace)
        model_trace = poutine.prune_subsample_sites(model_trace)
        sites = [site for trace in (model_trace, guide_trace) for site in trace.nodes.values()
                 if site["type"] == "sample"]
        dims = [frame.dim for site in sites for frame in site["cond_indep_stack"] if frame.vectorized]
        self.max_plate_nesting = -min(dims) if dims else 0
        if self.vectorize_particles and self.num_particles > 1:
            self.max_plate_nesting += 1

    def _vectorized_num_particles(self, fn):
        if self.num_particles == 1:
            return fn
        return plate("num_particles_vectorized", self.num_particles, dim=-self.max_plate_nesting)(fn)

    def _get_vectorized_trace(self, model, guide, args, kwargs):
        return self._get_trace(self._vectorized_num_particles(model),
                               self._vectorized_num_particles(guide), args, kwargs)

    @abstractmethod
    def _get_trace(self, model, guide, args, kwargs):
        raise NotImplementedError

    def _get_traces(self, model, guide, args, kwargs):
        if self.vectorize_particles:
            if self.max_plate_nesting == float("inf"):
                self._guess_max_plate_nesting(model, guide, args, kwargs)
            yield self._get_vectorized_trace(model, guide, args, kwargs)
        else:
            for i in range(self.num_particles):
                yield self._get_trace(model, guide, args, kwargs)
