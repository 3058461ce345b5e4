"""The oracle (oracle/) is pinned against the UNMODIFIED reference: its own known-answer fixtures
(tests/distributions/conftest.py, scipy log-pdfs, atol 1e-5 as tests/common.py:246-248) and
reference outputs recorded by tests/golden/make_golden.py.  CPU only."""
import math

import numpy as np
import pytest
import torch

from conftest import load_json, load_npz
from oracle import dists, mcmc as omcmc, optim as ooptim, svi as osvi

T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731

_FIXTURE_FN = {
    "Normal": lambda x, p: dists.normal(x, T(p["loc"]), T(p["scale"])),
    "Cauchy": lambda x, p: dists.cauchy(x, T(p["loc"]), T(p["scale"])),
    "HalfCauchy": lambda x, p: dists.half_cauchy(x, T(p["scale"])),
    "LogNormal": lambda x, p: dists.log_normal(x, T(p["loc"]), T(p["scale"])),
    "Exponential": lambda x, p: dists.exponential(x, T(p["rate"])),
    "Uniform": lambda x, p: dists.uniform(x, T(p["low"]), T(p["high"])),
    "Gamma": lambda x, p: dists.gamma(x, T(p["concentration"]), T(p["rate"])),
    "Beta": lambda x, p: dists.beta(x, T(p["concentration1"]), T(p["concentration0"])),
    "Dirichlet": lambda x, p: dists.dirichlet(x, T(p["concentration"])),
    "Poisson": lambda x, p: dists.poisson(x, T(p["rate"])),
    "Bernoulli": lambda x, p: (dists.bernoulli_logits(x, T(p["logits"])) if "logits" in p
                               else dists.bernoulli_probs(x, T(p["probs"]))),
    "Categorical": lambda x, p: dists.categorical(x, T(p["logits"]) if "logits" in p else T(p["probs"]).log()),
    "MultivariateNormal": lambda x, p: dists.mvn_tril(
        x, T(p["loc"]),
        T(p["scale_tril"]) if "scale_tril" in p else torch.linalg.cholesky(T(p["covariance_matrix"]))),
}


@pytest.mark.parametrize("case", load_json("dist_fixtures.json"),
                         ids=lambda c: "%s-%d" % (c["dist"], c["idx"]))
def test_reference_fixtures(case):
    fn = _FIXTURE_FN.get(case["dist"])
    if fn is None:
        pytest.skip("family not on the hot path")
    if case["dist"] == "MultivariateNormal" and not ({"scale_tril", "covariance_matrix"} & set(case["params"])):
        pytest.skip("precision parametrisation")
    x = T(case["test_data"])
    lp = fn(x, case["params"])
    ref = T(case["reference_log_prob"])
    assert torch.allclose(lp.reshape(ref.shape) if lp.numel() == ref.numel() else lp, ref, atol=1e-9, rtol=1e-9)
    if case["scipy_log_prob"] is not None:
        sc = T(case["scipy_log_prob"])
        # the reference's own bar: sum of log_prob vs scipy at atol 1e-5
        # (tests/distributions/test_distributions.py:60-71)
        if sc.numel() == lp.numel():
            assert torch.allclose(lp.reshape(-1), sc.reshape(-1), atol=1e-5)
        else:
            assert abs(float(lp.sum()) - float(sc.sum())) < 1e-5


_RANDOM = {
    "normal": lambda v, p: dists.normal(v, *p), "cauchy": lambda v, p: dists.cauchy(v, *p),
    "lognormal": lambda v, p: dists.log_normal(v, *p), "halfcauchy": lambda v, p: dists.half_cauchy(v, *p),
    "halfnormal": lambda v, p: dists.half_normal(v, *p), "exponential": lambda v, p: dists.exponential(v, *p),
    "gamma": lambda v, p: dists.gamma(v, *p), "beta": lambda v, p: dists.beta(v, *p),
    "uniform": lambda v, p: dists.uniform(v, *p),
    "bernoulli_logits": lambda v, p: dists.bernoulli_logits(v, *p),
    "bernoulli_probs": lambda v, p: dists.bernoulli_probs(v, *p),
    "poisson": lambda v, p: dists.poisson(v, *p),
    "normal_bcast": lambda v, p: dists.normal(v, *p), "normal_bcast2": lambda v, p: dists.normal(v, *p),
    "dirichlet": lambda v, p: dists.dirichlet(v, *p), "dirichlet_bcast": lambda v, p: dists.dirichlet(v, *p),
    "categorical3": lambda v, p: dists.categorical(v, *p), "categorical40": lambda v, p: dists.categorical(v, *p),
    "categorical_bcast": lambda v, p: dists.categorical(v, *p),
    "mvn2": lambda v, p: dists.mvn_tril(v, *p), "mvn5": lambda v, p: dists.mvn_tril(v, *p),
    "mvn37": lambda v, p: dists.mvn_tril(v, *p), "mvn_bcast": lambda v, p: dists.mvn_tril(v, *p),
}


@pytest.mark.parametrize("key", sorted(_RANDOM))
def test_random_batches_and_grads(key):
    g = load_npz("dist_random.npz")
    v = torch.as_tensor(g[key + ".value"])
    ps = []
    i = 0
    while key + ".p%d" % i in g:
        ps.append(torch.as_tensor(g[key + ".p%d" % i]).requires_grad_(True))
        i += 1
    has_dv = key + ".dvalue" in g
    if has_dv:
        v = v.requires_grad_(True)
    lp = _RANDOM[key](v, ps)
    assert torch.allclose(lp, torch.as_tensor(g[key + ".lp"]), atol=1e-10, rtol=1e-10)
    grads = torch.autograd.grad(lp.sum(), ([v] if has_dv else []) + ps, allow_unused=True)
    grads = [torch.zeros_like(t) if gr is None else gr for gr, t in zip(grads, ([v] if has_dv else []) + ps)]
    gi = 0
    if has_dv:
        assert torch.allclose(grads[0], torch.as_tensor(g[key + ".dvalue"]), atol=1e-9, rtol=1e-8)
        gi = 1
    for k in range(len(ps)):
        assert torch.allclose(grads[gi + k], torch.as_tensor(g[key + ".dp%d" % k]), atol=1e-8, rtol=1e-7)


def test_kl():
    g = load_npz("kl.npz")
    for name, fn in (("normal", dists.kl_normal_normal), ("gamma", dists.kl_gamma_gamma)):
        ps = [torch.as_tensor(g["%s.p%d" % (name, i)]) for i in range(4)]
        assert torch.allclose(fn(*ps), torch.as_tensor(g[name + ".kl"]), atol=1e-12, rtol=1e-12)


@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-12), ("f32", torch.float32, 2e-6)])
def test_optimisers(tag, dtype, tol):
    g = load_npz("optim.npz")
    p0 = torch.as_tensor(g["p0_" + tag])
    grads = torch.as_tensor(g["grads_" + tag])
    for name, mk in (
        ("clipped_adam", lambda: ooptim.ClippedAdam(lr=0.05, betas=(0.9, 0.99), clip_norm=2.0, lrd=0.97, weight_decay=0.01)),
        ("clipped_adam_default", lambda: ooptim.ClippedAdam(lr=0.01)),
        ("adagrad_rmsprop", lambda: ooptim.AdagradRMSProp(eta=4.5, t=0.1)),
    ):
        opt = mk()
        p = p0.clone()
        for i, gr in enumerate(grads):
            opt.step(p, gr.clone())
            ref = torch.as_tensor(g["%s_%s" % (name, tag)][i])
            assert torch.allclose(p, ref, atol=tol, rtol=tol), (name, i)


@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-9), ("f32", torch.float32, 3e-4)])
def test_logistic_svi_matches_reference(tag, dtype, tol):
    g = load_npz("svi_logistic.npz")
    X, y = torch.as_tensor(g["X"]).to(dtype), torch.as_tensor(g["y"]).to(dtype)
    eps_w, eps_b = torch.as_tensor(g["eps_w"]).to(dtype), torch.as_tensor(g["eps_b"]).to(dtype)
    P = int(g["P"])
    for cls in (osvi.LogisticSVI, osvi.LogisticSVIMatmul):
        m = cls(X.shape[1], P, lr=0.01, dtype=dtype)
        # gradient at the initial point
        loss0 = m.loss_and_grads(X, y, eps_w[0], eps_b[0])
        assert abs(float(loss0) - float(g["loss0_" + tag])) <= tol * max(1.0, abs(float(g["loss0_" + tag])))
        for k in ("w_loc", "w_scale", "b_loc", "b_scale"):
            ref = torch.as_tensor(g["grad0_%s_%s" % (k, tag)]).to(dtype)
            assert torch.allclose(m.params[k].grad.reshape(ref.shape), ref, atol=tol * 50, rtol=tol * 50), k
        m = cls(X.shape[1], P, lr=0.01, dtype=dtype)
        for i in range(eps_w.shape[0]):
            loss = m.step(X, y, eps_w[i], eps_b[i])
            assert abs(loss - g["losses_" + tag][i]) <= tol * abs(g["losses_" + tag][i]) * 10
            c = m.constrained()
            flat = torch.cat([c[k].reshape(-1).double() for k in ("w_loc", "w_scale", "b_loc", "b_scale")])
            assert torch.allclose(flat, torch.as_tensor(g["params_" + tag][i]), atol=tol * 10, rtol=tol * 10)


def test_potentials_and_leapfrog():
    g = load_npz("mcmc.npz")
    y, sigma = torch.as_tensor(g["es.y"]), torch.as_tensor(g["es.sigma"])
    U = omcmc.eight_schools_potential(y, sigma)
    for z, u_ref, g_ref in zip(torch.as_tensor(g["es.Z"]), g["es.U"], torch.as_tensor(g["es.G"])):
        gr, u = omcmc.potential_grad(U, z)
        assert abs(float(u) - u_ref) < 1e-9 * max(1, abs(u_ref))
        assert torch.allclose(gr, g_ref, atol=1e-9, rtol=1e-9)
    z0 = torch.as_tensor(g["es.Z"])[0]
    z, r, _, u = omcmc.velocity_verlet(z0, torch.as_tensor(g["es.vv.r0"]), U,
                                       torch.as_tensor(g["es.vv.minv"]), 0.05, num_steps=7)
    assert torch.allclose(z, torch.as_tensor(g["es.vv.z"]), atol=1e-10)
    assert torch.allclose(r, torch.as_tensor(g["es.vv.r"]), atol=1e-10)
    assert abs(float(u) - float(g["es.vv.U"])) < 1e-9
    UL = omcmc.logistic_potential(torch.as_tensor(g["lr.X"]), torch.as_tensor(g["lr.y"]))
    for b, u_ref, g_ref in zip(torch.as_tensor(g["lr.B"]), g["lr.U"], torch.as_tensor(g["lr.G"])):
        gr, u = omcmc.potential_grad(UL, b)
        assert abs(float(u) - u_ref) < 1e-9 * max(1, abs(u_ref))
        assert torch.allclose(gr, g_ref, atol=1e-9, rtol=1e-9)
    # harmonic oscillator closed form (tests/ops/test_integrator.py): after t = 6.28, (q, p) ~ (1, 0)
    zf, rf, _, _ = omcmc.velocity_verlet(torch.tensor([1.0], dtype=torch.float64), torch.tensor([0.0], dtype=torch.float64),
                                         lambda z: 0.5 * (z ** 2).sum(), torch.ones(1, dtype=torch.float64), 0.01, 628)
    assert torch.allclose(zf, torch.as_tensor(g["ho.z"]), atol=1e-12)
    assert abs(float(zf) - math.cos(6.28)) < 1e-4


def test_adaptation_pieces():
    g = load_npz("mcmc.npz")
    for w in (5, 19, 100, 150, 200, 500, 1000):
        s = omcmc.adaptation_schedule(w)
        assert [[a.start, a.end] for a in s] == g["sched.%d" % w].tolist()
    da = omcmc.DualAveraging(prox_center=math.log(10 * 0.3))
    for gg, ref in zip(g["da.g"], g["da.x"]):
        da.step(float(gg))
        assert np.allclose(da.get_state(), ref, atol=1e-12)
    wf = omcmc.Welford()
    for s in torch.as_tensor(g["wf.samples"]):
        wf.update(s)
    assert torch.allclose(wf.get_covariance(True), torch.as_tensor(g["wf.cov_reg"]), atol=1e-12)
    assert torch.allclose(wf.get_covariance(False), torch.as_tensor(g["wf.cov"]), atol=1e-12)


def test_oracle_nuts_recovers_reference_posterior():
    """Statistical parity of the oracle's recursive NUTS with the reference run (tolerances in the
    spirit of tests/infer/mcmc/test_nuts.py: posterior means within a few MC standard errors)."""
    g = load_npz("mcmc.npz")
    y, sigma = torch.as_tensor(g["es.y"]), torch.as_tensor(g["es.sigma"])
    U = omcmc.eight_schools_potential(y, sigma)
    torch.manual_seed(0)
    chain = omcmc.NUTSChain(U, 10, seed=4)
    z0 = (torch.rand(10, dtype=torch.float64) * 4 - 2)
    samples, accs = chain.run(z0, 300, 1500)
    mu = samples[:, 0]
    tau = samples[:, 1].exp()
    assert abs(float(mu.mean()) - float(g["es.long.mu.mean"])) < 0.6
    assert abs(float(tau.mean()) - float(g["es.long.tau.mean"])) < 0.9
    assert 0.6 < np.mean(accs) < 0.99


@pytest.mark.parametrize("tag", ["homog", "hetero", "wide"])
def test_gaussian_hmm_first_principles_oracle(tag):
    """oracle/hmm.py (dense joint Gaussian of the stacked observations -- neither the reference's
    parallel scan nor the product's Kalman filter) against the reference's recorded log_prob
    (tests/golden/hmm.npz; shapes of tests/distributions/test_hmm.py:424-555)."""
    from oracle import hmm as ohmm
    g = load_npz("hmm.npz")
    P = {k: torch.as_tensor(g["%s.%s" % (tag, k)]) for k in
         ("init_loc", "init_cov", "F", "trans_loc", "trans_cov", "Hm", "obs_loc", "obs_scale")}
    value, ref = torch.as_tensor(g[tag + ".value"]), torch.as_tensor(g[tag + ".lp"])
    R = torch.diag_embed(P["obs_scale"] ** 2)
    seqs = value if value.dim() == 3 else value[None]
    out = torch.stack([ohmm.gaussian_hmm_log_prob(P["init_loc"], P["init_cov"], P["F"], P["trans_loc"],
                                                  P["trans_cov"], P["Hm"], P["obs_loc"], R, x) for x in seqs])
    assert torch.allclose(out.reshape(ref.shape), ref, rtol=1e-10, atol=1e-9)
