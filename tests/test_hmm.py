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
