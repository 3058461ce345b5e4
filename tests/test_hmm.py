"""GaussianHMM (row c1): the Kalman-filter formulation against the reference's parallel-scan
log_prob and its parameter gradients (tests/golden/hmm.npz, recorded from the unmodified reference;
shapes follow tests/distributions/test_hmm.py:424-555).  The recursion is dense linear algebra
through torch (cuBLAS / cuSOLVER on the GPU), so the same test runs on both tiers."""
import numpy as np
import pytest
import torch

import pyro_b200.distributions as dist
from conftest import load_npz


def _run(tag, device, dtype, tol):
    g = load_npz("hmm.npz")
    P = {k: torch.as_tensor(g["%s.%s" % (tag, k)]).to(device, dtype).requires_grad_(True)
         for k in ("init_loc", "init_cov", "F", "trans_loc", "trans_cov", "Hm", "obs_loc", "obs_scale")}
    value = torch.as_tensor(g[tag + ".value"]).to(device, dtype)
    d = dist.GaussianHMM(
        dist.MultivariateNormal(P["init_loc"], covariance_matrix=P["init_cov"]), P["F"],
        dist.MultivariateNormal(P["trans_loc"], covariance_matrix=P["trans_cov"]), P["Hm"],
        dist.Normal(P["obs_loc"], P["obs_scale"]).to_event(1), duration=value.shape[-2])
    lp = d.log_prob(value)
    ref = torch.as_tensor(g[tag + ".lp"])
    assert torch.allclose(lp.detach().cpu().double(), ref, atol=tol * 10, rtol=tol)
    grads = torch.autograd.grad(lp.sum(), list(P.values()))
    for k, gr in zip(P, grads):
        r = torch.as_tensor(g["%s.grad.%s" % (tag, k)])
        if k in ("init_cov", "trans_cov"):
            # the reference differentiates through cholesky(cov) (lower triangle only); compare the
            # symmetrised gradient, which is what both formulations agree on
            gr = 0.5 * (gr + gr.transpose(-1, -2))
            r = 0.5 * (r + r.transpose(-1, -2))
        assert torch.allclose(gr.detach().cpu().double(), r, atol=tol * 100, rtol=tol * 100), (tag, k)


@pytest.mark.parametrize("tag", ["homog", "hetero", "wide"])
def test_gaussian_hmm_matches_reference_cpu(tag):
    _run(tag, "cpu", torch.float64, 1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["homog", "hetero", "wide"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
def test_gaussian_hmm_matches_reference_gpu(tag, dtype, tol):
    from conftest import device
    _run(tag, device(), dtype, tol)


def _steady_vs_sequential(device, dtype, H, O, T, tol, gtol):
    torch.manual_seed(H)
    F = (0.5 * torch.randn(H, H, dtype=dtype) / H ** 0.5).to(device).requires_grad_(True)
    Hm = torch.randn(H, O, dtype=dtype).to(device).requires_grad_(True)
    tsc = (torch.randn(H, dtype=dtype) * 0.1).exp().to(device).requires_grad_(True)
    osc = (torch.randn(O, dtype=dtype) * 0.1).exp().to(device).requires_grad_(True)
    isc = torch.ones(H, dtype=dtype, device=device).requires_grad_(True)
    bw = (0.1 * torch.randn(H, dtype=dtype)).to(device).requires_grad_(True)
    data = torch.randn(T, O, dtype=dtype).to(device)
    z = lambda n: torch.zeros(n, dtype=dtype, device=device)  # noqa: E731
    out = []
    for steady in (True, False):
        d = dist.GaussianHMM(dist.Normal(z(H), isc).to_event(1), F, dist.Normal(bw, tsc).to_event(1), Hm,
                             dist.Normal(z(O), osc).to_event(1), duration=T, steady_state=steady)
        lp = d.log_prob(data)
        out.append((float(lp), torch.autograd.grad(lp, [F, Hm, tsc, osc, isc, bw])))
    (a, ga), (b, gb) = out
    assert abs(a - b) <= tol * abs(b), (a, b)
    for x, y in zip(ga, gb):
        assert float((x - y).abs().max()) <= gtol * float(y.abs().max().clamp(min=1.0))


@pytest.mark.parametrize("H,O,T", [(6, 2, 700), (12, 3, 1000), (5, 1, 64)])
def test_steady_state_scan_equals_sequential_filter_cpu(H, O, T):
    """Time-invariant parameters (BASELINE config 3's case): once the predicted covariance has converged
    the filter switches to a blocked linear scan of the means; value and all parameter gradients equal
    the step-by-step recursion (itself pinned against the reference by the goldens above)."""
    _steady_vs_sequential("cpu", torch.float64, H, O, T, 1e-11, 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float64, 1e-11, 1e-8), (torch.float32, 1e-4, 2e-3)])
def test_steady_state_scan_equals_sequential_filter_gpu(dtype, tol, gtol):
    from conftest import device
    _steady_vs_sequential(device(), dtype, 16, 3, 1200, tol, gtol)


def test_kalman_filter_against_first_principles_oracle_on_seeded_inputs():
    """Beyond the goldens: random time-varying parameters, product filter vs oracle/hmm.py."""
    from oracle import hmm as ohmm
    torch.manual_seed(3)
    T, Hd, O = 11, 4, 3
    dt = torch.float64
    F = 0.6 * torch.randn(T, Hd, Hd, dtype=dt)
    Hm = torch.randn(T, Hd, O, dtype=dt)
    bw, bv = torch.randn(T, Hd, dtype=dt), torch.randn(T, O, dtype=dt)
    A = torch.randn(T, Hd, Hd, dtype=dt)
    Q = A @ A.transpose(-1, -2) + 0.5 * torch.eye(Hd, dtype=dt)
    osc = 0.5 + torch.rand(T, O, dtype=dt)
    m0 = torch.randn(Hd, dtype=dt)
    B = torch.randn(Hd, Hd, dtype=dt)
    P0 = B @ B.T + torch.eye(Hd, dtype=dt)
    x = torch.randn(T, O, dtype=dt)
    d = dist.GaussianHMM(dist.MultivariateNormal(m0, covariance_matrix=P0), F,
                         dist.MultivariateNormal(bw, covariance_matrix=Q), Hm,
                         dist.Normal(bv, osc).to_event(1), duration=T)
    ref = ohmm.gaussian_hmm_log_prob(m0, P0, F, bw, Q, Hm, bv, torch.diag_embed(osc ** 2), x)
    assert abs(float(d.log_prob(x)) - float(ref)) <= 1e-9 * abs(float(ref))


def _filter_case(tag, device, dtype):
    g = load_npz("hmm_filter.npz")
    t = lambda k: torch.as_tensor(g["%s.%s" % (tag, k)]).to(device, dtype)  # noqa: E731
    d = dist.GaussianHMM(dist.Normal(t("i_loc"), t("i_scale")).to_event(1), t("F"),
                         dist.Normal(t("t_loc"), t("t_scale")).to_event(1), t("H"),
                         dist.Normal(t("o_loc"), t("o_scale")).to_event(1), duration=t("x").shape[0])
    post = d.filter(t("x"))
    L = post.scale_tril
    return post.loc.detach().cpu().double(), (L @ L.transpose(-1, -2)).detach().cpu().double(), \
        d.log_prob(t("x")).detach().cpu().double(), g


@pytest.mark.parametrize("tag", ["inv", "var"])
def test_gaussian_hmm_filter_matches_reference_and_oracle(tag):
    """``GaussianHMM.filter`` (pyro/distributions/hmm.py:604-633): posterior of the final hidden state against
    the unmodified reference (tests/golden/hmm_filter.npz) and against the first-principles oracle (joint
    Gaussian of (z_T, x_1..x_T), conditioned on x); the oracle itself is pinned by the same golden."""
    from oracle import hmm as ohmm
    mean, cov, lp, g = _filter_case(tag, "cpu", torch.float64)
    rm, rc = torch.as_tensor(g[tag + ".mean"]), torch.as_tensor(g[tag + ".cov"])
    assert torch.allclose(mean, rm, atol=1e-9, rtol=1e-9) and torch.allclose(cov, rc, atol=1e-9, rtol=1e-9)
    assert abs(float(lp) - float(g[tag + ".logp"])) <= 1e-9 * abs(float(g[tag + ".logp"]))
    t = lambda k: torch.as_tensor(g["%s.%s" % (tag, k)])  # noqa: E731
    om, oc = ohmm.gaussian_hmm_filter(t("i_loc"), torch.diag_embed(t("i_scale") ** 2), t("F"), t("t_loc"),
                                      torch.diag_embed(t("t_scale") ** 2), t("H"), t("o_loc"),
                                      torch.diag_embed(t("o_scale") ** 2), t("x"))
    assert torch.allclose(om, rm, atol=1e-9, rtol=1e-9) and torch.allclose(oc, rc, atol=1e-9, rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["inv", "var"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
def test_gaussian_hmm_filter_matches_reference_gpu(tag, dtype, tol):
    from conftest import device
    mean, cov, _, g = _filter_case(tag, device(), dtype)
    assert torch.allclose(mean, torch.as_tensor(g[tag + ".mean"]), atol=tol, rtol=tol)
    assert torch.allclose(cov, torch.as_tensor(g[tag + ".cov"]), atol=tol, rtol=tol)
