"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    PYTHONPATH=tests/golden/opt_einsum_standin:/root/reference python tests/golden/make_golden.py

It imports reference Pyro 1.9.1 from /root/reference (plus a ~30-line stand-in for the absent
`opt_einsum` package, tests/golden/opt_einsum_standin/) and records inputs + outputs of the hot
paths.  The GPU box has no /root/reference, so these files are what travels.  Everything is
float64 unless a key says otherwise; files are small (< 1 MB total).

Files written
  dist_fixtures.json   the reference's own known-answer fixtures (tests/distributions/conftest.py:
                       params, test data, scipy log-pdf) for the families on the hot path, plus
                       the reference log_prob on them
  dist_random.npz      seeded random batches per family: inputs, reference log_prob, autograd grads
  kl.npz               KL(Normal||Normal), KL(Gamma||Gamma) from torch.distributions.kl
  elbo_grad.npz        Trace_ELBO loss + grads on small models with recorded guide noise
  svi_logistic.npz     5 SVI steps (Trace_ELBO, 8 vectorised particles, ClippedAdam) of Bayesian
                       logistic regression with recorded noise: losses and parameters per step
  optim.npz            ClippedAdam / AdagradRMSProp trajectories on given gradients
  def_meanfield.npz    config-5 structure (sparse gamma DEF, TraceMeanField_ELBO, AdagradRMSProp): 6 losses
                       and the final unconstrained parameters
  hmm.npz              GaussianHMM.log_prob and parameter gradients (homogeneous, heterogeneous, wide)
  mcmc.npz             potentials + gradients, velocity_verlet trajectories, integrator KATs,
                       adaptation schedules, dual averaging / Welford sequences,
                       eight_schools NUTS posterior moments (4 chains, 200+200), stats (r_hat, ESS)
"""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "opt_einsum_standin"))
sys.path.insert(0, "/root/reference")

import pyro  # noqa: E402
import pyro.distributions as dist  # noqa: E402
import pyro.poutine as poutine  # noqa: E402
from pyro.infer import SVI, Trace_ELBO, TraceMeanField_ELBO  # noqa: E402
from pyro.infer.mcmc import MCMC, NUTS  # noqa: E402
from pyro.ops.integrator import potential_grad, velocity_verlet  # noqa: E402

assert pyro.__version__ == "1.9.1", pyro.__version__
torch.set_default_dtype(torch.float64)


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


# ------------------------------------------------------------------------------------------------
def dist_fixtures():
    sys.path.insert(0, "/root/reference")
    from tests.distributions.conftest import continuous_dists, discrete_dists
    wanted = {"Normal", "Gamma", "Beta", "Dirichlet", "MultivariateNormal", "HalfCauchy", "Cauchy",
              "Bernoulli", "Poisson", "Categorical", "Exponential", "LogNormal", "HalfNormal",
              "Uniform"}
    out = []
    for fx in continuous_dists + discrete_dists:
        name = fx.pyro_dist.__name__
        if name not in wanted:
            continue
        for idx in range(fx.get_num_test_data()):
            params = fx.get_dist_params(idx)
            data = fx.get_test_data(idx)
            try:
                d = fx.pyro_dist(**params)
                ref_lp = d.log_prob(data)
            except Exception as e:  # pragma: no cover
                print("skip", name, idx, e)
                continue
            scipy_lp = None
            if fx.scipy_arg_fn is not None:
                try:
                    scipy_lp = np.asarray(fx.get_scipy_batch_logpdf(idx), dtype=np.float64).tolist()
                except Exception:
                    try:
                        scipy_lp = np.asarray(fx.get_scipy_logpdf(idx), dtype=np.float64).tolist()
                    except Exception:
                        scipy_lp = None
            out.append({"dist": name, "idx": idx,
                        "params": {k: npy(v).tolist() for k, v in params.items()},
                        "test_data": npy(data).tolist(),
                        "reference_log_prob": npy(ref_lp).tolist(),
                        "scipy_log_prob": scipy_lp})
    with open(os.path.join(HERE, "dist_fixtures.json"), "w") as f:
        json.dump(out, f)
    print("dist_fixtures:", len(out), "cases")


def dist_random():
    g = torch.Generator().manual_seed(20260922)
    n = 257  # odd on purpose (vector-width tails)

    def rn(*s):
        return torch.randn(*s, generator=g)

    def ru(*s):
        return torch.rand(*s, generator=g)

    cases = {}

    def record(key, make, value, params, discrete_value=False):
        ps = [p.clone().requires_grad_(True) for p in params]
        v = value.clone()
        if not discrete_value:
            v.requires_grad_(True)
        lp = make(*ps).log_prob(v)
        grads = torch.autograd.grad(lp.sum(), ([v] if not discrete_value else []) + ps, allow_unused=True)
        cases[key + ".value"] = npy(value)
        for i, p in enumerate(params):
            cases[key + ".p%d" % i] = npy(p)
        cases[key + ".lp"] = npy(lp)
        gi = 0
        if not discrete_value:
            cases[key + ".dvalue"] = npy(grads[0]) if grads[0] is not None else np.zeros(value.shape)
            gi = 1
        for i in range(len(params)):
            cases[key + ".dp%d" % i] = npy(grads[gi + i]) if grads[gi + i] is not None else np.zeros(())

    loc, scale = rn(n), 0.3 + 2 * ru(n)
    record("normal", dist.Normal, rn(n) * 2, [loc, scale])
    record("cauchy", dist.Cauchy, rn(n) * 2, [loc, scale])
    record("lognormal", dist.LogNormal, (rn(n) * 0.7).exp(), [loc * 0.3, scale])
    record("halfcauchy", dist.HalfCauchy, rn(n).abs() + 0.01, [scale])
    record("halfnormal", dist.HalfNormal, rn(n).abs() + 0.01, [scale])
    record("exponential", dist.Exponential, rn(n).abs() + 0.01, [scale])
    a, b = 0.2 + 3 * ru(n), 0.2 + 3 * ru(n)
    record("gamma", dist.Gamma, dist.Gamma(a, b).sample().clamp(min=1e-6), [a, b])
    record("beta", dist.Beta, dist.Beta(a + 0.3, b + 0.3).sample().clamp(1e-4, 1 - 1e-4), [a + 0.3, b + 0.3])
    record("uniform", dist.Uniform, loc + 0.3 * scale, [loc, loc + scale])
    lg = rn(n) * 3
    record("bernoulli_logits", lambda l: dist.Bernoulli(logits=l), (ru(n) < 0.5).double(), [lg], True)
    record("bernoulli_probs", lambda p: dist.Bernoulli(probs=p), (ru(n) < 0.5).double(),
           [ru(n).clamp(0.02, 0.98)], True)
    rate = 0.2 + 6 * ru(n)
    record("poisson", dist.Poisson, dist.Poisson(rate).sample(), [rate], True)
    # broadcast case: params [D], value [P, N, D]-like
    record("normal_bcast", dist.Normal, rn(5, 7, 4), [rn(4), 0.5 + ru(4)])
    record("normal_bcast2", dist.Normal, rn(6, 1, 4).expand(6, 3, 4).contiguous(), [rn(6, 1, 1), 0.5 + ru(1, 3, 1)])
    # event families
    conc = 0.3 + 3 * ru(33, 5)
    record("dirichlet", dist.Dirichlet, dist.Dirichlet(conc).sample().clamp(min=1e-6), [conc])
    conc_b = 0.3 + 3 * ru(5)
    record("dirichlet_bcast", dist.Dirichlet, dist.Dirichlet(conc_b).sample((21,)).clamp(min=1e-6), [conc_b])
    for K in (3, 40):
        logits = rn(29, K)
        record("categorical%d" % K, lambda l: dist.Categorical(logits=l),
               torch.randint(0, K, (29,), generator=g), [logits], True)
    logits_b = rn(6)
    record("categorical_bcast", lambda l: dist.Categorical(logits=l),
           torch.randint(0, 6, (4, 31), generator=g), [logits_b], True)
    for nn in (2, 5, 37):
        A = rn(13, nn, nn)
        L = torch.linalg.cholesky(A @ A.transpose(-1, -2) + nn * torch.eye(nn))
        record("mvn%d" % nn, lambda m, t: dist.MultivariateNormal(m, scale_tril=t), rn(13, nn), [rn(13, nn), L])
    A = rn(4, 4)
    Lb = torch.linalg.cholesky(A @ A.T + 4 * torch.eye(4))
    record("mvn_bcast", lambda m, t: dist.MultivariateNormal(m, scale_tril=t), rn(19, 4), [rn(4), Lb])
    np.savez_compressed(os.path.join(HERE, "dist_random.npz"), **cases)
    print("dist_random:", len(cases), "arrays")


def kl_cases():
    g = torch.Generator().manual_seed(7)
    n = 129
    out = {}
    for name, mk, params in [
        ("normal", dist.Normal, [torch.randn(n, generator=g), 0.3 + torch.rand(n, generator=g),
                                 torch.randn(n, generator=g), 0.3 + torch.rand(n, generator=g)]),
        ("gamma", dist.Gamma, [0.2 + 3 * torch.rand(n, generator=g), 0.2 + 3 * torch.rand(n, generator=g),
                               0.2 + 3 * torch.rand(n, generator=g), 0.2 + 3 * torch.rand(n, generator=g)]),
    ]:
        ps = [p.clone().requires_grad_(True) for p in params]
        kl = torch.distributions.kl_divergence(mk(ps[0], ps[1]), mk(ps[2], ps[3]))
        grads = torch.autograd.grad(kl.sum(), ps)
        for i, p in enumerate(params):
            out["%s.p%d" % (name, i)] = npy(p)
            # torch's polygamma(1) in float64 is only ~1e-9 accurate; keep as-is, tests allow for it
            out["%s.dp%d" % (name, i)] = npy(grads[i])
        out["%s.kl" % name] = npy(kl)
    np.savez_compressed(os.path.join(HERE, "kl.npz"), **out)
    print("kl ok")


# ------------------------------------------------------------------------------------------------
class _Noise:
    """Deterministic replacement for guide noise: records what torch.randn-like draws produced by
    replaying a fixed list of eps tensors into Normal.rsample (loc + eps*scale)."""


def logistic_model(X, y):
    D = X.shape[-1]
    w = pyro.sample("w", dist.Normal(torch.zeros(D), torch.ones(D)).to_event(1))
    b = pyro.sample("b", dist.Normal(torch.tensor(0.0), torch.tensor(10.0)))
    with pyro.plate("data", X.shape[0]):
        # w: [D] or [P, 1, D] (vectorised particles); b: [] or [P, 1]  ->  logits [N] or [P, N]
        logits = (w * X).sum(-1) + b
        pyro.sample("y", dist.Bernoulli(logits=logits), obs=y)


def logistic_guide(X, y):
    D = X.shape[-1]
    w_loc = pyro.param("w_loc", torch.zeros(D))
    w_scale = pyro.param("w_scale", torch.full((D,), 0.1), constraint=dist.constraints.positive)
    b_loc = pyro.param("b_loc", torch.tensor(0.0))
    b_scale = pyro.param("b_scale", torch.tensor(0.1), constraint=dist.constraints.positive)
    pyro.sample("w", dist.Normal(w_loc, w_scale).to_event(1))
    pyro.sample("b", dist.Normal(b_loc, b_scale))


def svi_logistic():
    """Reference SVI on logistic regression with injected guide noise.

    The noise is injected by conditioning the guide's sample sites on loc + eps*scale computed
    from the CURRENT parameters (poutine.condition keeps the rsample-style gradient path because
    the conditioned value is a differentiable function of the params) -- equivalent to the
    reference drawing eps itself, but reproducible across devices (the replay technique of
    tests/infer/test_gradient.py:77-91)."""
    torch.manual_seed(0)
    N, D, P, steps = 192, 4, 8, 5
    X = torch.randn(N, D)
    w_true = torch.randn(D) / math.sqrt(D)
    y = torch.bernoulli(torch.sigmoid(X @ w_true + 0.5))
    eps_w = torch.randn(steps, P, 1, D)
    eps_b = torch.randn(steps, P, 1)
    out = {"X": npy(X), "y": npy(y), "eps_w": npy(eps_w), "eps_b": npy(eps_b), "P": P}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        torch.set_default_dtype(dtype)
        pyro.clear_param_store()
        Xd, yd = X.to(dtype), y.to(dtype)
        step_box = {"i": 0}

        def guide(X_, y_):
            D_ = X_.shape[-1]
            w_loc = pyro.param("w_loc", torch.zeros(D_))
            w_scale = pyro.param("w_scale", torch.full((D_,), 0.1), constraint=dist.constraints.positive)
            b_loc = pyro.param("b_loc", torch.tensor(0.0))
            b_scale = pyro.param("b_scale", torch.tensor(0.1), constraint=dist.constraints.positive)
            i = step_box["i"]
            wv = w_loc + eps_w[i].to(dtype) * w_scale
            bv = b_loc + eps_b[i].to(dtype) * b_scale
            with poutine.condition(data={"w": wv, "b": bv}):
                w = pyro.sample("w", dist.Normal(w_loc, w_scale).to_event(1))
                b = pyro.sample("b", dist.Normal(b_loc, b_scale))
            return w, b

        # conditioning marks the sites observed; Trace_ELBO needs them latent in the guide, so
        # un-observe them with a tiny messenger
        class Unobserve(poutine.messenger.Messenger):
            # visited right after `condition` on the way down the stack: keep the injected value
            # but make the site latent again
            def _pyro_sample(self, msg):
                if msg["name"] in ("w", "b"):
                    msg["is_observed"] = False

        def guide_latent(X_, y_):
            with Unobserve():
                return guide(X_, y_)

        elbo = Trace_ELBO(num_particles=P, vectorize_particles=True, max_plate_nesting=1)
        svi = SVI(logistic_model, guide_latent, pyro.optim.ClippedAdam({"lr": 0.01}), elbo)
        losses, traj = [], []
        for i in range(steps):
            step_box["i"] = i
            losses.append(svi.step(Xd, yd))
            store = pyro.get_param_store()
            traj.append(np.concatenate([npy(store[k]).reshape(-1).astype(np.float64)
                                        for k in ("w_loc", "w_scale", "b_loc", "b_scale")]))
        out["losses_" + tag] = np.asarray(losses)
        out["params_" + tag] = np.stack(traj)
        # gradient at the initial point (separate run, no optimiser step)
        pyro.clear_param_store()
        step_box["i"] = 0
        with poutine.trace(param_only=True) as cap:
            loss0 = elbo.loss_and_grads(logistic_model, guide_latent, Xd, yd)
        grads = {}
        for name, site in cap.trace.nodes.items():
            grads[name] = npy(site["value"].unconstrained().grad).astype(np.float64)
        out["loss0_" + tag] = np.asarray(loss0)
        for k, v in grads.items():
            out["grad0_%s_%s" % (k, tag)] = v
    torch.set_default_dtype(torch.float64)
    np.savez_compressed(os.path.join(HERE, "svi_logistic.npz"), **out)
    print("svi_logistic ok; losses f64:", out["losses_f64"])


def elbo_grad():
    """Trace_ELBO / TraceMeanField_ELBO on a small Gamma-Poisson + Normal model incl. a plate with
    subsampling scale and a mask, noise injected as above; plus the closed-form KAT of
    tests/infer/test_gradient.py (Normal-Normal)."""
    pyro.clear_param_store()
    torch.manual_seed(1)
    N = 12
    data = torch.randn(N) + 1.5
    counts = torch.poisson(torch.full((N,), 3.0))
    mask = torch.rand(N) < 0.7
    eps = torch.randn(6, 1)       # 6 particles
    ueps = torch.rand(6, 1).clamp(0.05, 0.95)

    def model():
        z = pyro.sample("z", dist.Normal(0.0, 2.0))
        rate = pyro.sample("rate", dist.Gamma(2.0, 0.5))
        with pyro.plate("data", 2 * N, subsample_size=N, dim=-1):
            with poutine.mask(mask=mask):
                pyro.sample("x", dist.Normal(z, 1.3), obs=data)  # z: [] or [P, 1]
            pyro.sample("c", dist.Poisson(rate), obs=counts)

    def guide():
        loc = pyro.param("loc", torch.tensor(0.3))
        scale = pyro.param("scale", torch.tensor(0.7), constraint=dist.constraints.positive)
        conc = pyro.param("conc", torch.tensor(3.0), constraint=dist.constraints.positive)
        grate = pyro.param("grate", torch.tensor(1.2), constraint=dist.constraints.positive)
        zv = loc + eps * scale
        # a reparameterised Gamma draw through the inverse CDF is not available; use a value that
        # is a differentiable function of the params so the pathwise term is exercised
        rv = conc / grate * (0.5 + ueps)

        class Unobserve(poutine.messenger.Messenger):
            def _pyro_sample(self, msg):
                if msg["name"] in ("z", "rate"):
                    msg["is_observed"] = False

        with Unobserve(), poutine.condition(data={"z": zv, "rate": rv}):
            pyro.sample("z", dist.Normal(loc, scale))
            pyro.sample("rate", dist.Gamma(conc, grate))
        # subsample indices must match the model's plate: fix them
    out = {"data": npy(data), "counts": npy(counts), "mask": npy(mask), "eps": npy(eps), "ueps": npy(ueps)}
    for cls, tag in ((Trace_ELBO, "trace"), (TraceMeanField_ELBO, "meanfield")):
        pyro.clear_param_store()
        elbo = cls(num_particles=6, vectorize_particles=True, max_plate_nesting=1)
        with poutine.trace(param_only=True) as cap:
            loss = elbo.loss_and_grads(model, guide)
        out[tag + ".loss"] = np.asarray(loss)
        for name, site in cap.trace.nodes.items():
            out["%s.grad.%s" % (tag, name)] = npy(site["value"].unconstrained().grad)
    np.savez_compressed(os.path.join(HERE, "elbo_grad.npz"), **out)
    print("elbo_grad ok", out["trace.loss"], out["meanfield.loss"])


def optim_cases():
    torch.manual_seed(3)
    out = {}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        p0 = torch.randn(37, dtype=dtype)
        grads = torch.randn(12, 37, dtype=dtype) * torch.logspace(-2, 1.5, 12, dtype=dtype)[:, None]
        out["p0_" + tag] = npy(p0)
        out["grads_" + tag] = npy(grads)
        for name, cls, kw in [
            ("clipped_adam", pyro.optim.clipped_adam.ClippedAdam,
             dict(lr=0.05, betas=(0.9, 0.99), clip_norm=2.0, lrd=0.97, weight_decay=0.01)),
            ("clipped_adam_default", pyro.optim.clipped_adam.ClippedAdam, dict(lr=0.01)),
            ("adagrad_rmsprop", pyro.optim.adagrad_rmsprop.AdagradRMSProp, dict(eta=4.5, t=0.1)),
        ]:
            p = p0.clone().requires_grad_(True)
            opt = cls([p], **kw)
            traj = []
            for g in grads:
                p.grad = g.clone()
                opt.step()
                traj.append(npy(p).copy())
            out["%s_%s" % (name, tag)] = np.stack(traj)
    np.savez_compressed(os.path.join(HERE, "optim.npz"), **out)
    print("optim ok")


def mcmc_cases():
    from pyro.infer.mcmc.adaptation import WarmupAdapter
    from pyro.infer.mcmc.util import initialize_model
    from pyro.ops.dual_averaging import DualAveraging
    from pyro.ops.welford import WelfordCovariance
    from pyro.ops import stats
    out = {}
    # ---- eight schools potential / grad / leapfrog ---------------------------------------------------
    J = 8
    y = torch.tensor([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0])
    sigma = torch.tensor([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0])

    def eight_schools(sigma, y=None):
        eta = pyro.sample("eta", dist.Normal(torch.zeros(J), torch.ones(J)))
        mu = pyro.sample("mu", dist.Normal(torch.zeros(1), 10 * torch.ones(1)))
        tau = pyro.sample("tau", dist.HalfCauchy(scale=25 * torch.ones(1)))
        theta = mu + tau * eta
        return pyro.sample("obs", dist.Normal(theta, sigma), obs=y)

    pyro.set_rng_seed(0)
    init, potential_fn, transforms, _ = initialize_model(eight_schools, model_args=(sigma, y))
    g = torch.Generator().manual_seed(11)
    Z = torch.randn(9, J + 2, generator=g)          # layout [mu, log_tau, eta]
    Us, Gs = [], []
    for z in Z:
        zd = {"mu": z[0:1].clone(), "tau": z[1:2].clone(), "eta": z[2:].clone()}
        grads, pe = potential_grad(potential_fn, zd)
        Us.append(float(pe))
        Gs.append(torch.cat([grads["mu"], grads["tau"], grads["eta"]]))
    out["es.y"], out["es.sigma"], out["es.Z"] = npy(y), npy(sigma), npy(Z)
    out["es.U"], out["es.G"] = np.asarray(Us), npy(torch.stack(Gs))
    # fixed-length trajectory from identical (z, r)
    z0 = {"mu": Z[0, 0:1].clone(), "tau": Z[0, 1:2].clone(), "eta": Z[0, 2:].clone()}
    r0v = torch.randn(J + 2, generator=g)
    r0 = {"mu": r0v[0:1].clone(), "tau": r0v[1:2].clone(), "eta": r0v[2:].clone()}
    minv = 0.5 + torch.rand(J + 2, generator=g)

    def kinetic_grad(r):
        return {"mu": minv[0:1] * r["mu"], "tau": minv[1:2] * r["tau"], "eta": minv[2:] * r["eta"]}

    zt, rt, gt, pet = velocity_verlet(z0, r0, potential_fn, kinetic_grad, 0.05, num_steps=7)
    out["es.vv.r0"], out["es.vv.minv"] = npy(r0v), npy(minv)
    out["es.vv.z"] = npy(torch.cat([zt["mu"], zt["tau"], zt["eta"]]))
    out["es.vv.r"] = npy(torch.cat([rt["mu"], rt["tau"], rt["eta"]]))
    out["es.vv.U"] = np.asarray(float(pet))
    # ---- logistic regression potential -------------------------------------------------------------------
    torch.manual_seed(5)
    N, D = 200, 3
    X = torch.randn(N, D)
    beta_true = torch.tensor([1.0, 2.0, 3.0])
    yb = torch.bernoulli(torch.sigmoid(X @ beta_true))

    def logreg(X, yb):
        beta = pyro.sample("beta", dist.Normal(torch.zeros(D), torch.ones(D)))
        return pyro.sample("y", dist.Bernoulli(logits=(X * beta).sum(-1)), obs=yb)

    _, pfn, _, _ = initialize_model(logreg, model_args=(X, yb))
    B = torch.randn(6, D, generator=g)
    Ul, Gl = [], []
    for b in B:
        grads, pe = potential_grad(pfn, {"beta": b.clone()})
        Ul.append(float(pe))
        Gl.append(grads["beta"])
    out["lr.X"], out["lr.y"], out["lr.B"] = npy(X), npy(yb), npy(B)
    out["lr.U"], out["lr.G"] = np.asarray(Ul), npy(torch.stack(Gl))
    # ---- integrator KATs (closed form; tests/ops/test_integrator.py harmonic oscillator) ---------------
    def harmonic(z):
        return 0.5 * (z["x"] ** 2).sum()
    zf, rf, _, _ = velocity_verlet({"x": torch.tensor([1.0])}, {"x": torch.tensor([0.0])}, harmonic,
                                   lambda r: r, 0.01, num_steps=628)
    out["ho.z"], out["ho.r"] = npy(zf["x"]), npy(rf["x"])
    # ---- adaptation ----------------------------------------------------------------------------------------------
    for w in (5, 19, 100, 150, 200, 500, 1000):
        ad = WarmupAdapter(adapt_step_size=True, adapt_mass_matrix=True)
        ad._warmup_steps = w
        sched = ad._build_adaptation_schedule()
        out["sched.%d" % w] = np.asarray([[s.start, s.end] for s in sched])
    da = DualAveraging(prox_center=math.log(10 * 0.3))
    gs = torch.rand(25, generator=g) - 0.4
    xs = []
    for gg in gs:
        da.step(float(gg))
        xs.append(list(da.get_state()))
    out["da.g"], out["da.x"] = npy(gs), np.asarray(xs)
    wf = WelfordCovariance(diagonal=True)
    S = torch.randn(17, 5, generator=g) * torch.tensor([0.1, 1.0, 3.0, 0.5, 2.0])
    for srow in S:
        wf.update(srow)
    out["wf.samples"] = npy(S)
    out["wf.cov_reg"] = npy(wf.get_covariance(regularize=True))
    out["wf.cov"] = npy(wf.get_covariance(regularize=False))
    # ---- stats -----------------------------------------------------------------------------------------------------
    x = torch.randn(4, 300, 3, generator=g).cumsum(1) * 0.05 + torch.randn(4, 300, 3, generator=g)
    out["stats.x"] = npy(x)
    out["stats.rhat"] = npy(stats.split_gelman_rubin(x, chain_dim=0, sample_dim=1))
    out["stats.neff"] = npy(stats.effective_sample_size(x, chain_dim=0, sample_dim=1))
    # ---- eight_schools posterior (BASELINE config 1: 4 chains x (200 + 200)), sequential chains ------------
    pyro.set_rng_seed(0)
    post = {"mu": [], "tau": [], "eta": []}
    nleap = 0
    import pyro.ops.integrator as integ
    import pyro.infer.mcmc.hmc as hmc_mod
    import pyro.infer.mcmc.nuts as nuts_mod
    calls = {"n": 0}
    orig = integ.potential_grad

    def counted(fn, z):
        calls["n"] += 1
        return orig(fn, z)
    integ.potential_grad = counted
    hmc_mod.potential_grad = counted
    nuts_mod.potential_grad = counted
    import time
    t0 = time.time()
    for chain in range(4):
        pyro.set_rng_seed(chain)
        mc = MCMC(NUTS(eight_schools), num_samples=200, warmup_steps=200, num_chains=1,
                  disable_progbar=True)
        mc.run(sigma, y)
        s = mc.get_samples()
        for k in post:
            post[k].append(s[k])
    dt = time.time() - t0
    integ.potential_grad = hmc_mod.potential_grad = nuts_mod.potential_grad = orig
    for k in post:
        v = torch.stack(post[k])  # [4, 200, ...]
        out["es.post.%s.mean" % k] = npy(v.reshape(-1, v.shape[-1]).mean(0))
        out["es.post.%s.std" % k] = npy(v.reshape(-1, v.shape[-1]).std(0))
    out["es.post.leapfrogs"] = np.asarray(calls["n"])
    out["es.post.seconds"] = np.asarray(dt)
    print("eight_schools reference: %d leapfrogs in %.1fs -> %.0f leapfrog/s (sequential chains)"
          % (calls["n"], dt, calls["n"] / dt))
    # long run for tight posterior moments (statistical parity target)
    pyro.set_rng_seed(123)
    mc = MCMC(NUTS(eight_schools), num_samples=4000, warmup_steps=500, num_chains=1, disable_progbar=True)
    mc.run(sigma, y)
    s = mc.get_samples()
    for k in ("mu", "tau", "eta"):
        out["es.long.%s.mean" % k] = npy(s[k].mean(0))
        out["es.long.%s.std" % k] = npy(s[k].std(0))
    np.savez_compressed(os.path.join(HERE, "mcmc.npz"), **out)
    print("mcmc ok; es long mu/tau mean:", out["es.long.mu.mean"], out["es.long.tau.mean"])


def hmm_cases():
    """GaussianHMM.log_prob + parameter gradients from the reference (parallel-scan formulation,
    pyro/distributions/hmm.py:565-582) for time-homogeneous and time-heterogeneous models."""
    g = torch.Generator().manual_seed(42)
    torch.set_default_dtype(torch.float64)
    out = {}

    def rn(*s):
        return torch.randn(*s, generator=g)

    def spd(*lead, n):
        A = rn(*lead, n, n)
        return A @ A.transpose(-1, -2) / n + 0.5 * torch.eye(n)

    for tag, H, O, T, B, hetero in (("homog", 3, 2, 7, (4,), False), ("hetero", 2, 3, 5, (), True),
                                    ("wide", 6, 4, 19, (2,), False)):
        tl = (T,) if hetero else ()
        params = {"init_loc": rn(H), "init_cov": spd(n=H),
                  "F": 0.5 * rn(*tl, H, H), "trans_loc": 0.1 * rn(*tl, H), "trans_cov": spd(*tl, n=H),
                  "Hm": rn(*tl, H, O), "obs_loc": 0.1 * rn(*tl, O), "obs_scale": 0.5 + torch.rand(*tl, O, generator=g)}
        value = rn(*B, T, O)
        ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        d = dist.GaussianHMM(
            dist.MultivariateNormal(ps["init_loc"], covariance_matrix=ps["init_cov"]), ps["F"],
            dist.MultivariateNormal(ps["trans_loc"], covariance_matrix=ps["trans_cov"]), ps["Hm"],
            dist.Normal(ps["obs_loc"], ps["obs_scale"]).to_event(1), duration=T)
        lp = d.log_prob(value)
        grads = torch.autograd.grad(lp.sum(), list(ps.values()))
        out[tag + ".value"] = npy(value)
        out[tag + ".lp"] = npy(lp)
        for (k, v), gr in zip(params.items(), grads):
            out["%s.%s" % (tag, k)] = npy(v)
            out["%s.grad.%s" % (tag, k)] = npy(gr)
    np.savez_compressed(os.path.join(HERE, "hmm.npz"), **out)
    print("hmm ok", out["homog.lp"])


def def_meanfield():
    """BASELINE config 5 structure at reduced size through the reference: sparse gamma DEF model of
    tests/models.py (the SAME source, with `pyro_b200` imports rewritten to `pyro`),
    TraceMeanField_ELBO with 3 vectorised particles, AdagradRMSProp, latent values injected as a
    deterministic function of the variational parameters."""
    import re
    import types
    src = open(os.path.join(os.path.dirname(HERE), "models.py")).read()
    src = src.replace("import pyro_b200 as pyro", "import pyro") \
             .replace("import pyro_b200.distributions as dist", "import pyro.distributions as dist") \
             .replace("from pyro_b200 import poutine", "from pyro import poutine") \
             .replace("poutine.Messenger", "poutine.messenger.Messenger")
    src = re.sub(r"def logistic_model_fused.*?\n\n\ndef ", "def ", src, flags=re.S)
    src = re.sub(r"def logistic_model_sharded.*?\n\n\ndef logistic_guide_sharded", "def logistic_guide_sharded", src, flags=re.S)
    mod = types.ModuleType("refmodels")
    exec(compile(src, "refmodels", "exec"), mod.__dict__)
    torch.manual_seed(0)
    torch.set_default_dtype(torch.float64)
    n, img, widths, P = 6, 32, (10, 6, 4), 3
    x = torch.poisson(torch.full((n, img), 2.0))
    inj = {s_: (lambda a, r, k=k: (a / r) * (0.6 + 0.1 * k)) for k, s_ in
           enumerate(["w_top", "w_mid", "w_bottom", "z_top", "z_mid", "z_bottom"])}
    pyro.clear_param_store()
    m = mod.SparseGammaDEF(img, widths, dtype=torch.float64, inject=inj, particles=P)
    svi = SVI(m.model, m.guide, pyro.optim.AdagradRMSProp({"eta": 0.5, "t": 0.1}),
              TraceMeanField_ELBO(num_particles=P, vectorize_particles=True, max_plate_nesting=1))
    losses = [svi.step(x) for _ in range(6)]
    out = {"x": npy(x), "losses": np.asarray(losses), "P": P, "widths": np.asarray(widths)}
    for k, v in pyro.get_param_store().named_parameters():
        out["param." + k] = npy(v)
    np.savez_compressed(os.path.join(HERE, "def_meanfield.npz"), **out)
    print("def_meanfield ok", losses[:3])


if __name__ == "__main__":
    which = sys.argv[1:] or ["dist_fixtures", "dist_random", "kl", "optim", "svi_logistic",
                             "elbo_grad", "mcmc", "def_meanfield", "hmm"]
    fns = {"dist_fixtures": dist_fixtures, "dist_random": dist_random, "kl": kl_cases,
           "optim": optim_cases, "svi_logistic": svi_logistic, "elbo_grad": elbo_grad,
           "mcmc": mcmc_cases, "def_meanfield": def_meanfield, "hmm": hmm_cases}
    for w in which:
        fns[w]()
