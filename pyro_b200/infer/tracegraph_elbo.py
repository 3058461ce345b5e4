"""TraceGraph_ELBO: Rao-Blackwellised score-function estimator with baselines.

Restates pyro/infer/tracegraph_elbo.py:29-102 (baseline options and ``_construct_baseline``), :186-245
(``_compute_elbo``), :248-292 (``TrackNonReparam``) and :295-376 (the estimator class).  The reference finds
the cost terms downstream of a non-reparameterisable site by PROVENANCE TRACKING
(pyro/ops/provenance.py): the site's sampled value is tagged with its name and the tag follows the value
through every torch operation, so a site's ``log_prob`` carries the names of all non-reparameterisable sites it
depends on.  :class:`ProvenanceTensor` below is that mechanism -- a tensor subclass whose
``__torch_function__`` unions the tags of the inputs onto the outputs; where a value enters one of this
package's native scoring kernels (an autograd.Function, opaque to torch functions) the tag is carried across
explicitly from the operands (``_site_provenance``).

The scoring itself is the materialised ``log_prob`` / ``score_parts`` path of the fused kernels
(``Trace.compute_log_prob`` / ``compute_score_parts``), as for the general path of Trace_ELBO.
"""
from collections import defaultdict

import torch

from ..distributions import Distribution, is_identically_zero
from ..params import get_param_store
from ..poutine import Messenger, _Subsample
from ..primitives import param as pyro_param
from ..util import torch_item, warn_if_nan
from .elbo import ELBO, get_importance_trace
from .util import MultiFrameTensor


# ---- provenance (pyro/ops/provenance.py:12-130) ------------------------------------------------------
class ProvenanceTensor(torch.Tensor):
    """A tensor that remembers which non-reparameterisable sample sites its value depends on."""

    @staticmethod
    def wrap(t, provenance):
        if not provenance or not isinstance(t, torch.Tensor):
            return t
        if isinstance(t, ProvenanceTensor):
            provenance = provenance | t._provenance
            t = t._t
        out = t.as_subclass(ProvenanceTensor)
        out._t = t          # the ORIGINAL tensor object: every computation runs on it (autograd leaves stay leaves)
        out._provenance = frozenset(provenance)
        return out

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        prov = set()

        def strip(x):
            if isinstance(x, ProvenanceTensor):
                prov.update(getattr(x, "_provenance", ()))
                return x._t
            if isinstance(x, (list, tuple)):
                return type(x)(strip(v) for v in x)
            if isinstance(x, dict):
                return {k: strip(v) for k, v in x.items()}
            return x

        out = func(*strip(args), **strip(kwargs))
        if not prov:
            return out
        prov = frozenset(prov)

        def tag(x):
            if isinstance(x, torch.Tensor):
                return ProvenanceTensor.wrap(x, prov)
            if isinstance(x, (list, tuple)):
                return type(x)(tag(v) for v in x)
            return x

        return tag(out)


def get_provenance(x):
    """Names attached to ``x`` (a tensor, or a distribution: the union over its tensor attributes)."""
    if isinstance(x, ProvenanceTensor):
        return frozenset(getattr(x, "_provenance", ()))
    if isinstance(x, Distribution):
        out = set()
        for v in vars(x).values():
            if isinstance(v, (torch.Tensor, Distribution)):
                out |= get_provenance(v)
            elif isinstance(v, (list, tuple)):
                for e in v:
                    if isinstance(e, (torch.Tensor, Distribution)):
                        out |= get_provenance(e)
        return frozenset(out)
    return frozenset()


def detach_provenance(x):
    return x._t if isinstance(x, ProvenanceTensor) else x


def _site_provenance(site):
    """Provenance of a site's log_prob: its value's and its distribution's parameters'."""
    return get_provenance(site["value"]) | get_provenance(site["fn"])


class TrackNonReparam(Messenger):
    """Tag the value of every non-reparameterisable, unobserved sample site with the site's name
    (pyro/infer/tracegraph_elbo.py:248-292)."""

    def _pyro_post_sample(self, msg):
        if (msg["type"] == "sample" and not isinstance(msg["fn"], _Subsample) and not msg["is_observed"]
                and not getattr(msg["fn"], "has_rsample", False) and isinstance(msg["value"], torch.Tensor)):
            msg["value"] = ProvenanceTensor.wrap(msg["value"], frozenset({msg["name"]}))


# ---- baselines (pyro/infer/tracegraph_elbo.py:29-102) ------------------------------------------------
def _get_baseline_options(site):
    options = site["infer"].get("baseline", {}).copy()
    out = (options.pop("nn_baseline", None), options.pop("nn_baseline_input", None),
           options.pop("use_decaying_avg_baseline", False), options.pop("baseline_beta", 0.90),
           options.pop("baseline_value", None))
    if options:
        raise ValueError("Unrecognized baseline options: {}".format(options.keys()))
    return out


def _construct_baseline(node, guide_site, downstream_cost):
    baseline = 0.0
    baseline_loss = 0.0
    nn_baseline, nn_baseline_input, use_decaying_avg, beta, baseline_value = _get_baseline_options(guide_site)
    use_nn = nn_baseline is not None
    use_value = baseline_value is not None
    use_baseline = use_nn or use_decaying_avg or use_value
    assert not (use_nn and use_value), "cannot use baseline_value and nn_baseline simultaneously"
    if use_decaying_avg:
        name = "__baseline_avg_downstream_cost_" + node
        with torch.no_grad():
            old = pyro_param(name, torch.zeros(downstream_cost.shape, dtype=downstream_cost.dtype,
                                               device=guide_site["value"].device))
            old = detach_provenance(old)
            new = (1 - beta) * detach_provenance(downstream_cost) + beta * old
        get_param_store()[name] = new
        baseline = baseline + old
    if use_nn:
        baseline = baseline + nn_baseline(detach_provenance(nn_baseline_input).detach())
    elif use_value:
        baseline = baseline + baseline_value
    if use_nn or use_value:
        baseline_loss = baseline_loss + torch.pow(detach_provenance(downstream_cost).detach() - baseline, 2.0).sum()
    if use_baseline and downstream_cost.shape != baseline.shape:
        raise ValueError("Expected baseline at site {} to be {} instead got {}".format(
            node, downstream_cost.shape, baseline.shape))
    return use_baseline, baseline_loss, baseline


# ---- the estimator (pyro/infer/tracegraph_elbo.py:186-245) ---------------------------------------------
def _compute_elbo(model_trace, guide_trace):
    elbo = 0.0
    surrogate_elbo = 0.0
    baseline_loss = 0.0
    # non-reparameterisable guide site -> cost terms that depend on its value
    downstream_costs = defaultdict(MultiFrameTensor)

    for name, site in model_trace.nodes.items():
        if site["type"] == "sample":
            elbo = elbo + site["log_prob_sum"]
            surrogate_elbo = surrogate_elbo + site["log_prob_sum"]
            for key in _site_provenance(site):
                downstream_costs[key].add((site["cond_indep_stack"], detach_provenance(site["log_prob"])))

    for name, site in guide_trace.nodes.items():
        if site["type"] == "sample":
            elbo = elbo - site["log_prob_sum"]
            entropy_term = site["score_parts"].entropy_term
            if not is_identically_zero(entropy_term):
                surrogate_elbo = surrogate_elbo - entropy_term.sum()
            for key in _site_provenance(site):
                downstream_costs[key].add((site["cond_indep_stack"], -detach_provenance(site["log_prob"])))

    for node, downstream_cost in downstream_costs.items():
        guide_site = guide_trace.nodes[node]
        downstream_cost = downstream_cost.sum_to(guide_site["cond_indep_stack"])
        score_function = guide_site["score_parts"].score_function
        use_baseline, baseline_loss_term, baseline = _construct_baseline(node, guide_site, downstream_cost)
        if use_baseline:
            downstream_cost = downstream_cost - baseline
            baseline_loss = baseline_loss + baseline_loss_term
        surrogate_elbo = surrogate_elbo + (detach_provenance(score_function) *
                                           detach_provenance(downstream_cost).detach()).sum()

    surrogate_loss = -surrogate_elbo + baseline_loss
    return detach_provenance(elbo), detach_provenance(surrogate_loss)


class TraceGraph_ELBO(ELBO):
    """Drop-in for ``pyro.infer.TraceGraph_ELBO`` (same constructor, ``loss``, ``loss_and_grads``, baseline
    options under ``infer={"baseline": {...}}``)."""

    def _get_trace(self, model, guide, args, kwargs):
        with TrackNonReparam():
            # guide values reach the model by replay, tags attached
            return get_importance_trace("dense", self.max_plate_nesting, model, guide, args, kwargs)

    def loss(self, model, guide, *args, **kwargs):
        elbo = 0.0
        for model_trace, guide_trace in self._get_traces(model, guide, args, kwargs):
            elbo_particle = torch_item(detach_provenance(model_trace.log_prob_sum())) - \
                torch_item(detach_provenance(guide_trace.log_prob_sum()))
            elbo += elbo_particle / float(self.num_particles)
        loss = -elbo
        warn_if_nan(loss, "loss")
        return loss

    def _loss_and_surrogate_loss(self, model, guide, args, kwargs):
        loss = 0.0
        surrogate_loss = 0.0
        for model_trace, guide_trace in self._get_traces(model, guide, args, kwargs):
            lp, slp = _compute_elbo(model_trace, guide_trace)
            loss = loss + lp
            surrogate_loss = surrogate_loss + slp
        return loss / self.num_particles, surrogate_loss / self.num_particles

    def loss_and_grads(self, model, guide, *args, **kwargs):
        elbo, surrogate_loss = self._loss_and_surrogate_loss(model, guide, args, kwargs)
        if getattr(surrogate_loss, "requires_grad", False):
            surrogate_loss.backward(retain_graph=self.retain_graph)
        loss = -torch_item(elbo)
        warn_if_nan(loss, "loss")
        return loss


__all__ = ["TraceGraph_ELBO", "TrackNonReparam", "ProvenanceTensor", "get_provenance"]
