"""Execution trace: an ordered record of the sites one run of a model/guide visited, plus the
scoring entry points the ELBOs call.

Mirror of pyro/poutine/trace_struct.py:203-328 (``log_prob_sum``, ``compute_log_prob``,
``compute_score_parts``).  These three methods are where the reference calls every
``fn.log_prob``; here they reach the fused kernels through ``pyro_b200.distributions``.
"""
from collections import OrderedDict

import torch

from ..distributions import scale_and_mask


class Trace:
    def __init__(self):
        self.nodes = OrderedDict()
        self._edges = []

    # -- graph bookkeeping ----------------------------------------------------------------------
    def add_node(self, site_name, **kwargs):
        if site_name in self.nodes:
            site = self.nodes[site_name]
            if site["type"] != kwargs["type"]:
                raise RuntimeError("{} is already in the trace as a {}".format(site_name, site["type"]))
            elif kwargs["type"] != "param":
                raise RuntimeError("Multiple {} sites named '{}'".format(kwargs["type"], site_name))
        self.nodes[site_name] = kwargs

    def add_edge(self, a, b):
        self._edges.append((a, b))

    def remove_node(self, site_name):
        self.nodes.pop(site_name)

    def __contains__(self, name):
        return name in self.nodes

    def __iter__(self):
        return iter(self.nodes.keys())

    def __len__(self):
        return len(self.nodes)

    def copy(self):
        new = Trace()
        for name, site in self.nodes.items():
            new.nodes[name] = site.copy()
        new._edges = list(self._edges)
        return new

    def detach_(self):
        for site in self.nodes.values():
            if site["type"] == "sample" and isinstance(site["value"], torch.Tensor):
                site["value"] = site["value"].detach()

    # -- views ----------------------------------------------------------------------------------
    @property
    def stochastic_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "sample" and not s["is_observed"]]

    @property
    def observation_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "sample" and s["is_observed"]]

    @property
    def param_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "param"]

    @property
    def reparameterized_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "sample" and not s["is_observed"]
                and getattr(s["fn"], "has_rsample", False)]

    @property
    def nonreparam_stochastic_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "sample" and not s["is_observed"]
                and not getattr(s["fn"], "has_rsample", False)]

    def iter_stochastic_nodes(self):
        for name, site in self.nodes.items():
            if site["type"] == "sample" and not site["is_observed"]:
                yield name, site

    def format_shapes(self, title="Trace Shapes:", last_site=None):
        rows = [title]
        for name, site in self.nodes.items():
            if site["type"] == "sample":
                fn = site["fn"]
                bs = tuple(getattr(fn, "batch_shape", ()))
                es = tuple(getattr(fn, "event_shape", ()))
                vs = tuple(getattr(site["value"], "shape", ()))
                rows.append("  {:>20} dist {} | {}   value {}".format(name, bs, es, vs))
            elif site["type"] == "param":
                rows.append("  {:>20} param {}".format(name, tuple(site["value"].shape)))
            if name == last_site:
                break
        return "\n".join(rows)

    # -- scoring --------------------------------------------------------------------------------
    def _score(self, name, site, what):
        try:
            if what == "log_prob":
                return site["fn"].log_prob(site["value"], *site["args"], **site["kwargs"])
            return site["fn"].score_parts(site["value"], *site["args"], **site["kwargs"])
        except ValueError as e:
            raise ValueError("Error while computing {} at site '{}':\n{}\n{}".format(
                what, name, e, self.format_shapes(last_site=name))) from e

    def log_prob_sum(self, site_filter=lambda name, site: True):
        result = 0.0
        for name, site in self.nodes.items():
            if site["type"] == "sample" and site_filter(name, site):
                if "log_prob_sum" in site:
                    log_p = site["log_prob_sum"]
                else:
                    fused = _fused(site, weight=1.0, sum_coeff=1.0, unit=False)
                    if fused is not None:
                        log_p = fused
                    else:
                        log_p = self._score(name, site, "log_prob")
                        log_p = scale_and_mask(log_p, site["scale"], site["mask"]).sum()
                    site["log_prob_sum"] = log_p
                result = result + log_p
        return result

    def compute_log_prob(self, site_filter=lambda name, site: True):
        for name, site in self.nodes.items():
            if site["type"] == "sample" and site_filter(name, site):
                if "log_prob" not in site:
                    log_p = self._score(name, site, "log_prob")
                    site["unscaled_log_prob"] = log_p
                    log_p = scale_and_mask(log_p, site["scale"], site["mask"])
                    site["log_prob"] = log_p
                    site["log_prob_sum"] = log_p.sum()

    def compute_score_parts(self):
        for name, site in self.nodes.items():
            if site["type"] == "sample" and "score_parts" not in site:
                value = self._score(name, site, "score_parts")
                site["unscaled_log_prob"] = value.log_prob
                value = value.scale_and_mask(site["scale"], site["mask"])
                site["score_parts"] = value
                site["log_prob"] = value.log_prob
                site["log_prob_sum"] = value.log_prob.sum()


class ScaledTerm:
    """A precomputed 0-d ``sum log_prob`` (fused draw) and the coefficient it enters the ELBO with."""
    __slots__ = ("tensor", "coeff")

    def __init__(self, tensor, coeff):
        self.tensor, self.coeff = tensor, coeff


def _fused(site, weight, sum_coeff, unit=True, claim=False):
    """Fused ``sum_coeff * sum(scale*mask*log_prob)`` of a site, or None if it has no fused path
    (tensor-valued scale, python-bool mask False, exotic distribution).  With ``claim`` (and the
    unit-upstream contract) a site whose value is a fused reparameterised draw from its own
    distribution returns a :class:`ScaledTerm` around the sum computed at sampling time."""
    fn = site["fn"]
    scale = site["scale"]
    mask = site["mask"]
    if isinstance(scale, torch.Tensor):
        # a tensor-valued scale stays on the reference path (materialised log_prob * scale keeps the
        # scale differentiable, pyro/distributions/util.py:311-328); float() would also synchronise
        # the device and is illegal during graph capture
        if scale.numel() != 1 or scale.requires_grad:
            return None
        from .. import _native as N
        if scale.is_cuda and N.capturing():
            return None
        scale = float(scale)
    if mask is False:
        return None
    if mask is True:
        mask = None
    if site["args"] or site["kwargs"]:
        return None
    if claim and unit and mask is None:
        from ..distributions import claim_rsample_score
        lq = claim_rsample_score(fn, site["value"], weight * scale)
        if lq is not None:
            return ScaledTerm(lq, sum_coeff * scale)
    f = getattr(fn, "_fused_sum", None)
    if f is None:
        return None
    return f(site["value"], mask, scale, weight, sum_coeff, unit)
