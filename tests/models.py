"""Model / guide definitions shared by the CPU (emulated) and GPU (real kernel) test tiers and by
bench.py.  Written against the ``pyro`` interface; ``import pyro_b200 as pyro`` is the only change
from reference user code."""
import torch
from torch.distributions import constraints

import pyro_b200 as pyro
import pyro_b200.distributions as dist
from pyro_b200 import poutine


def logistic_model(X, y):
    """tests/infer/mcmc/test_hmc.py:189-198 scaled as SURVEY.md 8d (BASELINE config 2)."""
    D = X.shape[-1]
    w = pyro.sample("w", dist.Normal(X.new_zeros(D), X.new_ones(D)).to_event(1))
    b = pyro.sample("b", dist.Normal(X.new_zeros(()), X.new_full((), 10.0)))
    with pyro.plate("data", X.shape[0]):
        # w: [D] or [P, 1, D] (vectorised particles); b: [] or [P, 1] -> logits [N] or [P, N]
        if w.dim() > 1:
            logits = w.squeeze(-2) @ X.T + b
        else:
            logits = X @ w + b
        pyro.sample("y", dist.Bernoulli(logits=logits), obs=y)


def logistic_model_fused(X, y):
    """Same model with the linear predictor handed over lazily, so the likelihood site is scored by
    the fused GLM kernel (X and y read once for value and gradient)."""
    D = X.shape[-1]
    zero = dist.constant(0.0, X)   # cached constants: no fill kernels inside the step
    w = pyro.sample("w", dist.Normal(zero, 1.0).expand([D]).to_event(1))
    b = pyro.sample("b", dist.Normal(zero, 10.0))
    with pyro.plate("data", X.shape[0]):
        pyro.sample("y", dist.Bernoulli(logits=dist.linear_predictor(X, w, b)), obs=y)


def logistic_model_sharded(X, y, idx, n_total):
    """Data-parallel form: this rank holds rows ``idx`` of an ``n_total``-row data set.  The data
    plate's ``size / subsample_size`` rescaling (pyro/poutine/subsample_messenger.py:159-174) makes
    each rank's ELBO an unbiased estimate; SVI averages loss and gradients over ranks (the scheme of
    examples/svi_horovod.py:94-134)."""
    D = X.shape[-1]
    w = pyro.sample("w", dist.Normal(X.new_zeros(D), X.new_ones(D)).to_event(1))
    b = pyro.sample("b", dist.Normal(X.new_zeros(()), X.new_full((), 10.0)))
    with pyro.plate("data", n_total, subsample=idx):
        logits = w.squeeze(-2) @ X.T + b if w.dim() > 1 else X @ w + b
        pyro.sample("y", dist.Bernoulli(logits=logits), obs=y)


def logistic_guide_sharded(X, y, idx, n_total):
    logistic_guide(X, y)


def logistic_guide(X, y):
    D = X.shape[-1]
    w_loc = pyro.param("w_loc", lambda: X.new_zeros(D))
    w_scale = pyro.param("w_scale", lambda: X.new_full((D,), 0.1), constraint=constraints.positive)
    b_loc = pyro.param("b_loc", lambda: X.new_zeros(()))
    b_scale = pyro.param("b_scale", lambda: X.new_full((), 0.1), constraint=constraints.positive)
    pyro.sample("w", dist.Normal(w_loc, w_scale).to_event(1))
    pyro.sample("b", dist.Normal(b_loc, b_scale))


class InjectNoise(poutine.Messenger):
    """Replace the guide's Normal draws by loc + eps*scale for recorded eps (the replay technique of
    tests/infer/test_gradient.py:77-91), so runs are comparable across devices and with the
    reference goldens."""

    def __init__(self, eps, fused_draw=False):
        self.eps = eps
        self.fused_draw = fused_draw   # route the draw through Normal.rsample_with_noise (b2 family 14)

    def _pyro_sample(self, msg):
        if msg["name"] in self.eps and not msg["is_observed"]:
            fn = msg["fn"]
            base = fn
            while hasattr(base, "base_dist"):
                base = base.base_dist
            e = self.eps[msg["name"]]
            if self.fused_draw:
                shape = torch.broadcast_shapes(e.shape, base.batch_shape)
                msg["value"] = base.rsample_with_noise(e.to(base.loc.dtype).expand(shape).contiguous())
            else:
                msg["value"] = base.loc + e.to(base.loc.dtype) * base.scale
            msg["done"] = True


def eight_schools(sigma, y=None):
    """examples/eight_schools/mcmc.py:27-34"""
    J = sigma.shape[0]
    eta = pyro.sample("eta", dist.Normal(sigma.new_zeros(J), sigma.new_ones(J)))
    mu = pyro.sample("mu", dist.Normal(sigma.new_zeros(1), 10 * sigma.new_ones(1)))
    tau = pyro.sample("tau", dist.HalfCauchy(scale=25 * sigma.new_ones(1)))
    theta = mu + tau * eta
    return pyro.sample("obs", dist.Normal(theta, sigma), obs=y)


def logreg_mcmc_model(X, y):
    D = X.shape[-1]
    beta = pyro.sample("beta", dist.Normal(X.new_zeros(D), X.new_ones(D)))
    logits = (X * beta.unsqueeze(-2)).sum(-1) if beta.dim() > 1 else (X * beta).sum(-1)
    return pyro.sample("y", dist.Bernoulli(logits=logits), obs=y)


class SparseGammaDEF:
    """Sparse gamma deep exponential family (BASELINE config 5; structure of
    examples/sparse_gamma_def.py:43-165): three layers of Gamma weights and per-datapoint Gamma
    latents, Poisson likelihood, mean-field Gamma guide with softplus-parameterised shape/mean.
    ``inject`` (optional) maps site name -> callable(alpha, rate) -> value, used by tests to make
    the latent draws a deterministic function of the parameters on every device."""

    def __init__(self, image_size=4096, widths=(100, 40, 15), device="cpu", dtype=torch.float32, inject=None,
                 particles=1):
        self.top, self.mid, self.bottom = widths
        self.image_size = image_size
        t = lambda v: torch.tensor(v, device=device, dtype=dtype)  # noqa: E731
        self.alpha_z, self.beta_z, self.alpha_w, self.beta_w = t(0.1), t(0.1), t(0.1), t(0.3)
        self.device, self.dtype = device, dtype
        self.inject = inject
        self.particles = particles  # only used to give injected values the particle dim

    def model(self, x):
        n = x.size(0)
        shapes = {"top": (self.top, self.mid), "mid": (self.mid, self.bottom), "bottom": (self.bottom, self.image_size)}
        w = {}
        for name, (a, b) in shapes.items():
            with pyro.plate("w_%s_plate" % name, a * b):
                v = pyro.sample("w_%s" % name, dist.Gamma(self.alpha_w, self.beta_w))
            w[name] = v.reshape(a, b) if v.dim() == 1 else v.reshape(-1, a, b)
        with pyro.plate("data", n):
            z = pyro.sample("z_top", dist.Gamma(self.alpha_z, self.beta_z).expand([self.top]).to_event(1))
            mean = torch.matmul(z, w["top"])
            z = pyro.sample("z_mid", dist.Gamma(self.alpha_z, self.beta_z / mean).to_event(1))
            mean = torch.matmul(z, w["mid"])
            z = pyro.sample("z_bottom", dist.Gamma(self.alpha_z, self.beta_z / mean).to_event(1))
            mean = torch.matmul(z, w["bottom"])
            pyro.sample("obs", dist.Poisson(mean).to_event(1), obs=x)

    def guide(self, x):
        n = x.size(0)
        sp = torch.nn.functional.softplus
        gen = torch.Generator().manual_seed(0)

        def init(shape, mean):
            return lambda: (mean + 0.1 * torch.randn(shape, generator=gen)).to(self.device, self.dtype)

        def gamma_site(name, shape, event):
            alpha = sp(pyro.param("alpha_%s" % name, init(shape, 0.5)))
            mean = sp(pyro.param("mean_%s" % name, init(shape, 0.0)))
            d = dist.Gamma(alpha, alpha / mean)
            d = d.to_event(1) if event else d
            site = name.replace("_q", "")
            if self.inject is not None:
                v = self.inject[site](alpha, alpha / mean)
                if self.particles > 1:
                    # batch dims: [particles, plate] (+ event dim for the z sites)
                    v = v.expand((self.particles,) + tuple(v.shape))
                with InjectValue({site: v}):
                    pyro.sample(site, d)
            else:
                pyro.sample(site, d)

        for name, width in (("w_q_top", self.top * self.mid), ("w_q_mid", self.mid * self.bottom),
                            ("w_q_bottom", self.bottom * self.image_size)):
            with pyro.plate(name.replace("_q", "") + "_plate", width):
                gamma_site(name, (width,), False)
        with pyro.plate("data", n):
            gamma_site("z_q_top", (n, self.top), True)
            gamma_site("z_q_mid", (n, self.mid), True)
            gamma_site("z_q_bottom", (n, self.bottom), True)


class InjectValue(poutine.Messenger):
    def __init__(self, vals):
        self.vals = vals

    def _pyro_sample(self, msg):
        if msg["name"] in self.vals and not msg["is_observed"]:
            # only the value is fixed; the site still goes through plate broadcasting
            msg["value"] = self.vals[msg["name"]]
