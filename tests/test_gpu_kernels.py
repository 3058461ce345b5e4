"""Parity tests proper: the CUDA path, called through the C ABI, against goldens recorded from the
unmodified reference and against the oracle on seeded inputs.  Tolerances (stated per test):
  * index / integer work (Categorical gather) ............ bit-exact selection
  * fp64 kernels vs reference ............................. 1e-12 relative (1e-10 on gradients)
  * fp32 kernels vs fp64 reference ........................ |d| <= 1e-5 * max(1, |lp|)
    (the reference's own atol, tests/common.py:246-248); gradients 2e-4
  * sums of N terms ....................................... 1e-5 relative in fp32 (fixed-order tree
    summation, more accurate than a sequential sum)
"""
import numpy as np
import pytest
import torch

import pyro_b200.distributions as dist
from conftest import EMULATE, device, load_npz
from oracle import dists as odists
from pyro_b200.distributions import _ops

pytestmark = pytest.mark.gpu
DEV = device()

MAKE = {
    "normal": lambda p: dist.Normal(*p), "cauchy": lambda p: dist.Cauchy(*p),
    "lognormal": lambda p: dist.LogNormal(*p), "halfcauchy": lambda p: dist.HalfCauchy(*p),
    "halfnormal": lambda p: dist.HalfNormal(*p), "exponential": lambda p: dist.Exponential(*p),
    "gamma": lambda p: dist.Gamma(*p), "beta": lambda p: dist.Beta(*p),
    "uniform": lambda p: dist.Uniform(*p),
    "bernoulli_logits": lambda p: dist.Bernoulli(logits=p[0]),
    "bernoulli_probs": lambda p: dist.Bernoulli(probs=p[0]),
    "poisson": lambda p: dist.Poisson(*p),
    "normal_bcast": lambda p: dist.Normal(*p), "normal_bcast2": lambda p: dist.Normal(*p),
    "dirichlet": lambda p: dist.Dirichlet(*p), "dirichlet_bcast": lambda p: dist.Dirichlet(*p),
    "categorical3": lambda p: dist.Categorical(logits=p[0]),
    "categorical40": lambda p: dist.Categorical(logits=p[0]),
    "categorical_bcast": lambda p: dist.Categorical(logits=p[0]),
    "mvn2": lambda p: dist.MultivariateNormal(p[0], scale_tril=p[1]),
    "mvn5": lambda p: dist.MultivariateNormal(p[0], scale_tril=p[1]),
    "mvn37": lambda p: dist.MultivariateNormal(p[0], scale_tril=p[1]),
    "mvn_bcast": lambda p: dist.MultivariateNormal(p[0], scale_tril=p[1]),
}


@pytest.fixture(params=["small", "large"])
def site_kernels(request):
    """The reference's fixtures are a few hundred elements: run them through the one-CTA kernel
    (default for sites <= B2_SITE_SMALL_N) AND through the multi-CTA vector / generic kernels."""
    from pyro_b200 import _native as N
    N.FORCE_LARGE_SITE_KERNELS = request.param == "large"
    yield request.param
    N.FORCE_LARGE_SITE_KERNELS = False


def _load(key, dtype):
    g = load_npz("dist_random.npz")
    v = torch.as_tensor(g[key + ".value"])
    v = v.to(DEV) if not v.is_floating_point() else v.to(DEV, dtype)
    ps, i = [], 0
    while key + ".p%d" % i in g:
        ps.append(torch.as_tensor(g[key + ".p%d" % i]).to(DEV, dtype))
        i += 1
    return g, v, ps


def _close(a, ref, tol):
    a = a.detach().double().cpu()
    ref = torch.as_tensor(ref).double().reshape(a.shape)
    bad = (a - ref).abs() > tol * ref.abs().clamp(min=1)
    assert not bool(bad.any()), "max err %.3e" % float(((a - ref).abs() / ref.abs().clamp(min=1)).max())


@pytest.mark.parametrize("key", sorted(MAKE))
@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float64, 1e-12, 1e-9), (torch.float32, 1e-5, 2e-4)])
def test_log_prob_and_backward_match_reference(site_kernels, key, dtype, tol, gtol):
    g, v, ps = _load(key, dtype)
    ps = [p.requires_grad_(True) for p in ps]
    has_dv = key + ".dvalue" in g
    if has_dv:
        v = v.requires_grad_(True)
    lp = MAKE[key](ps).log_prob(v)
    _close(lp, g[key + ".lp"], tol)
    if key.startswith("categorical"):
        # the gathered entry must be the exact normalised logit (index work is exact)
        lgn = (ps[0] - ps[0].logsumexp(-1, keepdim=True)).detach()
        vv = v.unsqueeze(-1)
        vv, lgb = torch.broadcast_tensors(vv, lgn)
        ref = lgb.gather(-1, vv[..., :1]).squeeze(-1)
        assert torch.allclose(lp.detach(), ref, atol=tol * 10, rtol=0)
    up = torch.ones_like(lp)
    grads = torch.autograd.grad(lp, ([v] if has_dv else []) + ps, grad_outputs=up, allow_unused=True)
    gi = 0
    if has_dv:
        _close(grads[0], g[key + ".dvalue"], gtol)
        gi = 1
    for k in range(len(ps)):
        ref = g[key + ".dp%d" % k]
        got = grads[gi + k]
        if got is None:
            assert np.all(ref == 0)
        else:
            _close(got, ref, gtol)


@pytest.mark.parametrize("key", ["normal", "gamma", "bernoulli_logits", "normal_bcast", "normal_bcast2",
                                 "dirichlet_bcast", "categorical_bcast", "mvn_bcast", "beta", "poisson"])
@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float64, 1e-11, 1e-9), (torch.float32, 2e-5, 3e-4)])
def test_fused_sum_path_scale_mask_weight(site_kernels, key, dtype, tol, gtol):
    """ONE kernel: sum(scale*mask*lp) and the final weighted gradients, vs autograd on the oracle."""
    g, v, ps = _load(key, dtype)
    torch.manual_seed(0)
    d = MAKE[key](ps)
    bshape = d.log_prob(v).shape
    mask = (torch.rand(bshape, device=DEV) < 0.7)
    scale, weight, coeff = 2.5, -0.125, -1.0
    diff_value = v.is_floating_point() and (key + ".dvalue") in g
    ps_r = [p.clone().requires_grad_(True) for p in ps]
    vr = v.clone().requires_grad_(True) if diff_value else v
    out = MAKE[key](ps_r)._fused_sum(vr, mask, scale, weight, coeff, True)
    grads = torch.autograd.grad(out, ([vr] if diff_value else []) + ps_r, allow_unused=True)
    # oracle in float64 on the CPU
    fam = d.family
    po = [p.detach().double().cpu().requires_grad_(True) for p in ps]
    vo = v.detach().cpu()
    if vo.is_floating_point():
        vo = vo.double()
    if diff_value:
        vo.requires_grad_(True)
    fn = odists.ELEMENTWISE[fam][0] if fam in odists.ELEMENTWISE else odists.EVENT[fam]
    lp = fn(vo, *po).expand(bshape)
    tot = torch.where(mask.cpu(), lp * scale, torch.zeros((), dtype=torch.float64)).sum()
    assert abs(float(out) - coeff * float(tot)) <= 10 * tol * max(1.0, abs(float(tot)))
    og = torch.autograd.grad(weight * tot, ([vo] if diff_value else []) + po, allow_unused=True)
    assert len(og) == len(grads)
    for a, b in zip(grads, og):
        if b is None:
            continue
        _close(a, b, gtol)


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float64, 1e-12, 1e-7), (torch.float32, 2e-5, 3e-4)])
def test_kl_kernels(site_kernels, dtype, tol, gtol):
    g = load_npz("kl.npz")
    for name, cls in (("normal", dist.Normal), ("gamma", dist.Gamma)):
        ps = [torch.as_tensor(g["%s.p%d" % (name, i)]).to(DEV, dtype).requires_grad_(True) for i in range(4)]
        kl = dist.kl_divergence(cls(ps[0], ps[1]), cls(ps[2], ps[3]))
        _close(kl, g[name + ".kl"], tol)
        grads = torch.autograd.grad(kl.sum(), ps)
        for k in range(4):
            _close(grads[k], g["%s.dp%d" % (name, k)], gtol)


@pytest.mark.parametrize("shape,dst_shape", [((7, 5, 3), (5, 3)), ((7, 5, 3), (7, 1, 3)), ((6, 4), (1, 4)),
                                             ((1000, 33), (33,)), ((3, 100000), (3, 1)), ((64, 1, 32), (32,)),
                                             ((300, 1000), (1000,)), ((64, 40, 50), (40, 50)), ((7, 5, 4096), (4096,)),
                                             ((256, 61440), (61440,))])
def test_reduce_to(shape, dst_shape):
    torch.manual_seed(0)
    src = torch.randn(shape, device=DEV, dtype=torch.float64)
    dst = torch.empty(dst_shape, device=DEV, dtype=torch.float64)
    _ops.reduce_to(src, dst)
    ref = src.sum_to_size(dst_shape) if len(dst_shape) == len(shape) else src.sum_to_size((1,) * (len(shape) - len(dst_shape)) + tuple(dst_shape)).reshape(dst_shape)
    assert torch.allclose(dst, ref, atol=1e-9, rtol=1e-12)


def test_large_sum_property_and_determinism():
    """BASELINE-size site: Bernoulli(logits [64, 1e6]) against obs [1e6] broadcast over particles.
    Properties: (i) fused sum == sum of materialised log_prob; (ii) bit-stable across launches;
    (iii) linearity in scale; (iv) gradient == y - sigmoid(l) on a sampled slab."""
    torch.manual_seed(0)
    P, N = 64, 1 << 20
    logits = torch.randn(P, N, device=DEV)
    y = (torch.rand(N, device=DEV) < 0.3).float()
    d = dist.Bernoulli(logits=logits)
    s1 = d._fused_sum(y, None, 1.0, 1.0, 1.0, True)
    s2 = d._fused_sum(y, None, 1.0, 1.0, 1.0, True)
    assert float(s1) == float(s2)
    lp = d.log_prob(y)
    ref = float(lp.double().sum())
    assert abs(float(s1) - ref) <= 1e-5 * abs(ref)
    s3 = d._fused_sum(y, None, 3.0, 1.0, 1.0, True)
    assert abs(float(s3) - 3 * float(s1)) <= 1e-5 * abs(ref) * 3
    lg = logits.clone().requires_grad_(True)
    out = dist.Bernoulli(logits=lg)._fused_sum(y, None, 1.0, -1.0 / P, 1.0, True)
    (gl,) = torch.autograd.grad(out, lg)
    expect = (-1.0 / P) * (y[None, :4096] - torch.sigmoid(logits[:, :4096]))
    assert torch.allclose(gl[:, :4096], expect, atol=1e-7, rtol=1e-5)
    # oracle on a slab
    o = odists.bernoulli_logits(y[:5000].cpu().double(), logits[:3, :5000].cpu().double())
    assert torch.allclose(lp[:3, :5000].cpu().double(), o, atol=1e-5)


def test_empty_and_ragged_sites(site_kernels):
    z = torch.zeros(0, 5, device=DEV)
    d = dist.Normal(torch.zeros(5, device=DEV), torch.ones(5, device=DEV))
    assert d.log_prob(z).shape == (0, 5)
    assert float(d._fused_sum(z, None, 1.0, 1.0, 1.0, True)) == 0.0
    for n in (1, 2, 3, 5, 31, 33, 1023, 1025):
        x = torch.randn(n, device=DEV, dtype=torch.float64)
        lp = dist.Normal(torch.zeros((), device=DEV, dtype=torch.float64), torch.ones((), device=DEV, dtype=torch.float64)).log_prob(x)
        ref = odists.normal(x.cpu(), torch.zeros((), dtype=torch.float64), torch.ones((), dtype=torch.float64))
        assert torch.allclose(lp.cpu(), ref, atol=1e-13)
    # unaligned view (forces the generic kernel)
    base = torch.randn(1001, device=DEV)
    x = base[1:]
    lp = dist.Normal(torch.zeros((), device=DEV), torch.ones((), device=DEV)).log_prob(x)
    assert torch.allclose(lp.cpu().double(), odists.normal(x.cpu().double(), torch.zeros((), dtype=torch.float64), torch.ones((), dtype=torch.float64)), atol=1e-5)
    # non-contiguous (transposed) operands
    a = torch.randn(37, 19, device=DEV, dtype=torch.float64)
    lpt = dist.Normal(a.t(), torch.ones((), device=DEV, dtype=torch.float64)).log_prob(torch.zeros(19, 37, device=DEV, dtype=torch.float64))
    assert torch.allclose(lpt.cpu(), odists.normal(torch.zeros(19, 37, dtype=torch.float64), a.t().cpu(), torch.ones((), dtype=torch.float64)), atol=1e-12)


def test_edge_values(site_kernels):
    """-inf outside the support, NaN propagation, masked NaNs do not leak."""
    hc = dist.HalfCauchy(torch.ones(3, device=DEV))
    lp = hc.log_prob(torch.tensor([-1.0, 0.0, 2.0], device=DEV))
    assert lp[0] == -float("inf") and torch.isfinite(lp[1:]).all()
    n = dist.Normal(torch.tensor([0.0, float("nan")], device=DEV), torch.ones(2, device=DEV))
    lp = n.log_prob(torch.zeros(2, device=DEV))
    assert torch.isfinite(lp[0]) and torch.isnan(lp[1])
    mask = torch.tensor([True, False], device=DEV)
    s = n._fused_sum(torch.zeros(2, device=DEV), mask, 1.0, 1.0, 1.0, True)
    assert torch.isfinite(s)
    p = dist.Poisson(torch.tensor([0.0, 2.0], device=DEV))
    lp = p.log_prob(torch.tensor([0.0, 3.0], device=DEV))
    assert abs(float(lp[0])) < 5e-6  # fp32 lgamma(1) via the shifted Stirling series: |err| ~ 1e-6


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-6)])
@pytest.mark.parametrize("shape,pshape", [((64, 1, 32), (32,)), ((64, 1), ()), ((3, 5, 7), (5, 1)),
                                          ((300, 1000), (1000,)), ((17,), (17,))])
def test_fused_normal_draw_and_backward(shape, pshape, dtype, tol):
    """b2 families 14 / 15 (draw + own score, and its backward) against torch autograd on the
    same noise: z, sum log q, and d/d(loc, scale) of  L = sum(gz * z) + c * sum log q(z)  reduced
    to the parameters' stored shapes (one-CTA kernel for the first three shapes, vector kernel +
    b2_reduce_to for [300, 1000])."""
    torch.manual_seed(0)
    loc = torch.randn(pshape, device=DEV, dtype=dtype)
    scale = (0.5 + torch.rand(pshape, device=DEV, dtype=dtype))
    eps = torch.randn(shape, device=DEV, dtype=dtype)
    gz = torch.randn(shape, device=DEV, dtype=dtype)
    c = 0.37
    z, lq = _ops.normal_rsample_score(loc, scale, eps)
    lo, so = loc.double().cpu().requires_grad_(True), scale.double().cpu().requires_grad_(True)
    zo = lo + eps.double().cpu() * so
    lqo = odists.normal(zo, lo, so).sum()
    _close(z, zo.detach(), tol)
    assert abs(float(lq) - float(lqo)) <= 50 * tol * max(1.0, abs(float(lqo)))
    gl, gs = _ops.normal_rsample_backward(gz, eps, loc, scale, c, True, True)
    L = (gz.double().cpu() * zo).sum() + c * lqo
    glo, gso = torch.autograd.grad(L, [lo, so])
    assert gl.shape == loc.shape and gs.shape == scale.shape
    n_red = eps.numel() / max(1, loc.numel())
    _close(gl, glo, 20 * tol * max(1.0, n_red ** 0.5))
    _close(gs, gso, 20 * tol * max(1.0, n_red ** 0.5))


def test_fused_draw_step_equals_sitewise_step():
    """Free-running guide (torch.randn noise, same seed): one Trace_ELBO loss_and_grads with the
    fused draw claimed by the ELBO equals the same step with every site scored by b2_site_score."""
    import models
    import pyro_b200 as pyro
    from pyro_b200 import _native as N
    from pyro_b200.infer import Trace_ELBO
    torch.manual_seed(1)
    X = torch.randn(500, 6, device=DEV)
    y = (torch.rand(500, device=DEV) < 0.4).float()
    res = []
    for fused in (True, False):
        N.FUSED_DRAW = fused
        N.PHILOX_DRAW = False      # both runs take their noise from torch.randn (same seed -> same draws)
        try:
            pyro.clear_param_store()
            torch.manual_seed(7)
            elbo = Trace_ELBO(num_particles=16, vectorize_particles=True, max_plate_nesting=1)
            loss = elbo.loss_and_grads(models.logistic_model, models.logistic_guide, X, y)
            store = pyro.get_param_store()
            grads = {k: store._params[k].grad.clone() for k in ("w_loc", "w_scale", "b_loc", "b_scale")}
            res.append((loss, grads))
        finally:
            N.FUSED_DRAW = True
            N.PHILOX_DRAW = True
    assert abs(res[0][0] - res[1][0]) <= 1e-5 * abs(res[1][0])
    for k in res[0][1]:
        assert torch.allclose(res[0][1][k], res[1][1][k], rtol=2e-4, atol=2e-4 * float(res[1][1][k].abs().max())), k


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_philox_draw_kernel(dtype):
    """b2_normal_rsample: z = loc + eps * scale for the eps it returns, the 0-d sum equals the oracle's Normal
    log density of z, the noise is standard normal (goodness of fit in the style of
    tests/distributions/test_distributions.py:138-164), consecutive launches and CUDA-graph replays draw
    fresh noise (the kernel advances its own counter), and reseeding reproduces the stream."""
    if EMULATE:
        pytest.skip("kernel test")
    import pyro_b200 as pyro
    from oracle import dists as od
    pyro.set_rng_seed(11)
    loc = torch.randn(7, 1, 33, device=DEV, dtype=dtype)
    scale = torch.rand(33, device=DEV, dtype=dtype) + 0.5
    shape = (5, 7, 4, 33)
    z, lq, eps = _ops.normal_rsample_philox(loc, scale, shape)
    tol = 2e-6 if dtype == torch.float32 else 1e-14     # the kernel contracts the multiply-add
    assert torch.allclose(z, loc + eps * scale, rtol=tol, atol=tol)
    ref = od.normal(z.double().cpu(), loc.double().cpu().expand(shape), scale.double().cpu().expand(shape)).sum()
    assert abs(float(lq) - float(ref)) <= (2e-5 if dtype == torch.float32 else 1e-10) * abs(float(ref))
    z2, _, eps2 = _ops.normal_rsample_philox(loc, scale, shape)
    assert not torch.equal(eps, eps2)
    zero, one = torch.zeros((), device=DEV, dtype=dtype), torch.ones((), device=DEV, dtype=dtype)
    big = torch.cat([_ops.normal_rsample_philox(zero, one, (64, 1024))[2].reshape(-1)
                     for _ in range(8)]).double().cpu()
    n = big.numel()
    assert abs(float(big.mean())) < 5 / n ** 0.5 and abs(float(big.var()) - 1) < 5 * (2 / n) ** 0.5
    assert abs(float((big ** 3).mean())) < 5 * (15 / n) ** 0.5 and abs(float((big ** 4).mean()) - 3) < 5 * (96 / n) ** 0.5
    assert abs(float((big.abs() < 1).double().mean()) - 0.682689) < 5 * (0.2171 / n) ** 0.5
    # graph replay: fresh noise every replay
    g = torch.cuda.CUDAGraph()
    _ops.normal_rsample_philox(loc, scale, shape)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        zg, _, eg = _ops.normal_rsample_philox(loc, scale, shape)
    g.replay()
    a = eg.clone()
    g.replay()
    assert not torch.equal(a, eg)
    # reseeding reproduces the stream
    pyro.set_rng_seed(11)
    _, _, again = _ops.normal_rsample_philox(loc, scale, shape)
    assert torch.equal(again, eps)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("log_scale", [False, True])
def test_latent_block_kernels(dtype, log_scale):
    """csrc/latent.cu through the C ABI: b2_latent_normal_draw (z = loc + eps*s with s = scale or exp(log_scale),
    the 0-d sum of log q(z) against the oracle, fresh noise per launch and per graph replay, reseeding),
    b2_latent_normal_prior (value-only prior sums of several sites in one launch) and
    b2_latent_normal_backward (gradient reaching z + prior's d log p/dz + the score's own total derivative,
    reduced to the stored shapes of loc / scale / log_scale) against autograd on the same expression."""
    if EMULATE:
        pytest.skip("kernel test")
    import pyro_b200 as pyro
    from oracle import dists as od
    pyro.set_rng_seed(5)
    tol = 3e-6 if dtype == torch.float32 else 1e-13
    loc = torch.randn(7, 1, 33, device=DEV, dtype=dtype)
    sc = torch.rand(33, device=DEV, dtype=dtype) + 0.5
    store = sc.log() if log_scale else sc            # what the kernel reads
    s = store.exp() if log_scale else store
    shape = (5, 7, 4, 33)
    z, lq, eps = _ops.latent_draw(loc, store, log_scale, shape)
    assert torch.allclose(z, loc + eps * s, rtol=tol, atol=tol)
    ref = od.normal(z.double().cpu(), loc.double().cpu().expand(shape), s.double().cpu().expand(shape)).sum()
    assert abs(float(lq) - float(ref)) <= (2e-5 if dtype == torch.float32 else 1e-10) * abs(float(ref))
    _, _, eps2 = _ops.latent_draw(loc, store, log_scale, shape)
    assert not torch.equal(eps, eps2)
    big = _ops.latent_draw(torch.zeros((), device=DEV, dtype=dtype), torch.ones((), device=DEV, dtype=dtype),
                           False, (64, 1024))[2].reshape(-1).double().cpu()
    n = big.numel()
    assert abs(float(big.mean())) < 5 / n ** 0.5 and abs(float(big.var()) - 1) < 5 * (2 / n) ** 0.5
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        _, _, eg = _ops.latent_draw(loc, store, log_scale, shape)
    g.replay()
    a = eg.clone()
    g.replay()
    assert not torch.equal(a, eg)
    pyro.set_rng_seed(5)
    assert torch.equal(_ops.latent_draw(loc, store, log_scale, shape)[2], eps)
    # value-only priors of two sites, one launch
    pl, ps = torch.zeros((), device=DEV, dtype=dtype), torch.rand(4, 1, device=DEV, dtype=dtype) + 0.7
    z2 = torch.randn(11, device=DEV, dtype=dtype)
    pl2, ps2 = torch.randn(11, device=DEV, dtype=dtype), torch.full((), 10.0, device=DEV, dtype=dtype)
    lps = _ops.latent_prior([(z, pl, ps), (z2, pl2, ps2)])
    r1 = od.normal(z.double().cpu(), pl.double().cpu().expand(shape), ps.double().cpu().expand(shape)).sum()
    r2 = od.normal(z2.double().cpu(), pl2.double().cpu(), ps2.double().cpu().expand(11)).sum()
    for got, want in zip(lps, (r1, r2)):
        assert abs(float(got) - float(want)) <= (2e-5 if dtype == torch.float32 else 1e-10) * abs(float(want))
    # backward
    gz = torch.randn(shape, device=DEV, dtype=dtype)
    c, pw = 0.3, -0.7
    for prior in (None, (pl, ps, pw)):
        gl, gs = _ops.latent_backward(gz, eps, z, loc, store, log_scale, c, prior, True, True)
        lo = loc.double().cpu().requires_grad_(True)
        so = store.double().cpu().requires_grad_(True)
        s_o = so.exp() if log_scale else so
        zo = lo + eps.double().cpu() * s_o
        L = (gz.double().cpu() * zo).sum() + c * od.normal(zo, lo.expand(shape), s_o.expand(shape)).sum()
        if prior is not None:
            L = L + pw * od.normal(zo, pl.double().cpu().expand(shape), ps.double().cpu().expand(shape)).sum()
        glo, gso = torch.autograd.grad(L, [lo, so])
        assert gl.shape == loc.shape and gs.shape == store.shape
        gt = 2e-4 if dtype == torch.float32 else 1e-9
        _close(gl, glo, gt * 20 ** 0.5)
        _close(gs, gso, gt * 140 ** 0.5)
    gl, gs = _ops.latent_backward(None, eps, z, loc, store, log_scale, c, None, True, False)
    assert gs is None and float(gl.abs().max()) == 0.0
    # leaves that already hold a .grad: the kernel adds into it and hands nothing back to autograd
    gl0, gs0 = _ops.latent_backward(gz, eps, z, loc, store, log_scale, c, (pl, ps, pw), True, True)
    lleaf, sleaf = loc.clone().requires_grad_(True), store.clone().requires_grad_(True)
    lleaf.grad, sleaf.grad = torch.ones_like(loc), torch.full_like(store, 2.0)
    r = _ops.latent_backward(gz, eps, z, lleaf, sleaf, log_scale, c, (pl, ps, pw), True, True, accumulate=True)
    assert r == (None, None)
    assert torch.allclose(lleaf.grad, 1.0 + gl0, rtol=1e-6, atol=1e-6)
    assert torch.allclose(sleaf.grad, 2.0 + gs0, rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gamma_rsample_kernel(dtype):
    """b2_gamma_rsample through ``Gamma.rsample``: goodness of fit of the draws (Kolmogorov-Smirnov and moments
    per concentration, in the style of tests/distributions/test_distributions.py:138-164), the
    implicit-reparameterisation derivative against ATen's ``_standard_gamma_grad`` (the reference's backward of
    torch/distributions/gamma.py:79-87) on the kernel's own draws, the autograd gradients w.r.t. broadcast
    concentration / rate, fresh noise per launch and per graph replay, reproducibility after reseeding."""
    if EMULATE:
        pytest.skip("kernel test")
    import pyro_b200 as pyro
    from scipy import stats
    pyro.set_rng_seed(3)
    alphas = torch.tensor([0.15, 0.8, 1.0, 3.5, 40.0], device=DEV, dtype=dtype)
    rates = torch.tensor([0.5, 2.0, 1.0, 4.0, 10.0], device=DEV, dtype=dtype)
    n = 40000
    conc = alphas[:, None].clone().requires_grad_(True)          # [5, 1] broadcast over the draws
    rate = rates[:, None].clone().requires_grad_(True)
    z = dist.Gamma(conc, rate).rsample((n,)).squeeze(-1)        # [n, 5]
    assert z.shape == (n, 5) and bool((z > 0).all())
    zc = z.detach().double().cpu().numpy()
    for k, (a, r) in enumerate(zip(alphas.tolist(), rates.tolist())):
        x = zc[:, k] * r
        assert stats.kstest(x, "gamma", args=(a,)).pvalue > 1e-3, a
        assert abs(x.mean() - a) < 5 * (a / n) ** 0.5 and abs(x.var() - a) < 6 * a * (2 / n + 6 / (a * n)) ** 0.5
    # gradients: d sum(w z) / d conc, d rate with the derivative of every draw from ATen on the same draws
    w = torch.randn(n, 5, device=DEV, dtype=dtype)
    gc, gr = torch.autograd.grad((w * z).sum(), [conc, rate])
    x = (z.detach() * rates).double()
    ref_dx = torch._standard_gamma_grad(alphas.double().expand(n, 5).contiguous(), x.contiguous())
    ref_gc = (w.double() * ref_dx / rates.double()).sum(0)
    ref_gr = (-(w.double() * z.detach().double()) / rates.double()).sum(0)
    scale_c = (w.double().abs() * ref_dx.abs() / rates.double()).sum(0)
    assert bool(((gc.squeeze(-1).double() - ref_gc).abs() <= 3e-3 * scale_c).all())   # ATen's own approximation error
    assert torch.allclose(gr.squeeze(-1).double(), ref_gr, rtol=1e-4 if dtype == torch.float32 else 1e-10)
    # per-draw derivative against ATen
    _, dz = _ops.gamma_rsample(alphas.expand(n, 5), rates.expand(n, 5), (n, 5))
    z2, dz2 = _ops.gamma_rsample(alphas.expand(n, 5), rates.expand(n, 5), (n, 5))
    ref2 = torch._standard_gamma_grad(alphas.double().expand(n, 5).contiguous(),
                                      (z2 * rates).double().contiguous()) / rates.double()
    assert float(((dz2.double() - ref2).abs() / ref2.abs().clamp(min=1e-30)).max()) < 5e-3
    assert not torch.equal(dz, dz2)
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        zg, _ = _ops.gamma_rsample(alphas, rates, (5,))
    g.replay()
    a0 = zg.clone()
    g.replay()
    assert not torch.equal(a0, zg)
    pyro.set_rng_seed(3)
    again = dist.Gamma(conc, rate).rsample((n,)).squeeze(-1)
    assert torch.equal(again.detach(), z.detach())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_elbo_combine(dtype):
    """b2_elbo_combine: weighted sum of 0-d device scalars in index order, one launch."""
    torch.manual_seed(0)
    terms = [torch.randn((), device=DEV, dtype=dtype) * 10 ** k for k in range(7)]
    coeffs = [1.0, -1.0, 0.5, -1.0 / 64, 2.0, 0.0, -3.0]
    out = _ops.elbo_combine(terms, coeffs)
    ref = sum(c * float(t.double()) for c, t in zip(coeffs, terms))
    assert out.shape == () and out.dtype == dtype
    assert abs(float(out) - ref) <= (1e-6 if dtype == torch.float32 else 1e-14) * abs(ref)


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float64, 1e-11, 1e-9), (torch.float32, 2e-5, 3e-4)])
@pytest.mark.parametrize("K,rows", [(3, 5000), (8, 5000), (17, 5000), (64, 700), (100, 300), (1000, 70)])
def test_event_families_row_per_thread_kernels(K, rows, dtype, tol, gtol):
    """Dirichlet / Categorical with many rows and a small event size take the one-thread-per-row
    kernels (the reference fixtures, a few dozen rows, exercise the sub-warp-group kernels): log_prob,
    fused sum and gradients against the oracle, including a concentration broadcast over the rows and
    a [P, rows] batch that the host merges into one dim.  K > 32 (a multiple of 4) takes the 16-byte
    vector kernels, two rows in flight per lane group."""
    torch.manual_seed(K)
    conc = (0.3 + 2 * torch.rand(rows, K)).to(DEV, dtype)
    x = torch.distributions.Dirichlet(torch.ones(K)).sample((rows,)).clamp(min=1e-4).to(DEV, dtype)
    x = x / x.sum(-1, keepdim=True)
    for c in (conc, conc[0]):
        cr = c.clone().requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        lp = dist.Dirichlet(cr).log_prob(xr)
        co = c.double().cpu().requires_grad_(True)
        xo = x.double().cpu().requires_grad_(True)
        ref = odists.dirichlet(xo, co)
        _close(lp, ref.detach(), tol)
        w = torch.randn(rows, dtype=torch.float64)
        g = torch.autograd.grad((lp * w.to(DEV, dtype)).sum(), [xr, cr])
        go = torch.autograd.grad((ref * w).sum(), [xo, co])
        _close(g[0], go[0], gtol)
        _close(g[1], go[1], gtol * (1 if c.dim() == 2 else rows ** 0.5))
    logits = torch.randn(2, rows, K).to(DEV, dtype)
    idx = torch.randint(0, K, (2, rows), device=DEV)
    lr = logits.clone().requires_grad_(True)
    lp = dist.Categorical(logits=lr).log_prob(idx)
    lo = logits.double().cpu().requires_grad_(True)
    ref = odists.categorical(idx.cpu(), lo)
    _close(lp, ref.detach(), tol)
    (g,) = torch.autograd.grad(lp.sum(), lr)
    (go,) = torch.autograd.grad(ref.sum(), lo)
    _close(g, go, gtol)
    s = dist.Categorical(logits=logits)._fused_sum(idx, None, 1.0, 1.0, 1.0, True)
    assert abs(float(s) - float(ref.sum())) <= 20 * tol * abs(float(ref.sum()))


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float64, 1e-11, 1e-9), (torch.float32, 2e-5, 5e-4)])
@pytest.mark.parametrize("n", [1, 2, 5, 8, 12, 32])
def test_mvn_row_per_thread_kernel(n, dtype, tol, gtol):
    """MultivariateNormal(scale_tril) with many rows: event size <= 8 takes one thread per row (forward
    and back substitution in registers), 8 < n <= 32 one warp per row with the factor's columns in
    registers; against the oracle, per-row and row-broadcast parameters."""
    torch.manual_seed(n)
    rows = 3000
    A = torch.randn(rows, n, n)
    L = torch.linalg.cholesky(A @ A.transpose(-1, -2) + n * torch.eye(n)).to(DEV, dtype)
    mu = torch.randn(rows, n).to(DEV, dtype)
    x = torch.randn(rows, n).to(DEV, dtype)
    for mu_, L_ in ((mu, L), (mu[0], L[0])):
        mr, Lr, xr = (t.clone().requires_grad_(True) for t in (mu_, L_, x))
        lp = dist.MultivariateNormal(mr, scale_tril=Lr).log_prob(xr)
        mo, Lo, xo = (t.double().cpu().requires_grad_(True) for t in (mu_, L_, x))
        ref = odists.mvn_tril(xo, mo, Lo)
        _close(lp, ref.detach(), tol)
        w = torch.randn(rows, dtype=torch.float64)
        g = torch.autograd.grad((lp * w.to(DEV, dtype)).sum(), [xr, mr, Lr])
        go = torch.autograd.grad((ref * w).sum(), [xo, mo, Lo])
        red = 1 if mu_.dim() == 2 else rows ** 0.5
        _close(g[0], go[0], gtol)
        _close(g[1], go[1], gtol * red)
        _close(g[2], torch.tril(go[2]), gtol * red)


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float64, 1e-10, 1e-8), (torch.float32, 5e-5, 2e-3)])
@pytest.mark.parametrize("n", [129, 200, 512, 1000])
def test_mvn_large_event_warp_kernel(n, dtype, tol, gtol):
    """MultivariateNormal(scale_tril) beyond n = 128 (round 2: up to 1024; config 3 has H = 512): the
    warp-per-row substitution kernel with 16 / 32 register slots per lane, against the oracle (fp64 on the
    CPU) -- log_prob and all three gradients, factor shared by the batch and one factor per row."""
    if EMULATE:
        pytest.skip("kernel test")
    torch.manual_seed(n)
    rows = 6
    A = torch.randn(rows, n, n, dtype=torch.float64)
    L = torch.linalg.cholesky(A @ A.transpose(-1, -2) / n + torch.eye(n, dtype=torch.float64)).to(DEV, dtype)
    mu = torch.randn(rows, n, dtype=torch.float64).to(DEV, dtype)
    x = torch.randn(rows, n, dtype=torch.float64).to(DEV, dtype)
    for mu_, L_ in ((mu, L), (mu[0], L[0])):
        mr, Lr, xr = (t.clone().requires_grad_(True) for t in (mu_, L_, x))
        lp = dist.MultivariateNormal(mr, scale_tril=Lr).log_prob(xr)
        mo, Lo, xo = (t.double().cpu().requires_grad_(True) for t in (mu_, L_, x))
        ref = odists.mvn_tril(xo, mo, Lo)
        _close(lp, ref.detach(), tol)
        w = torch.randn(rows, dtype=torch.float64)
        g = torch.autograd.grad((lp * w.to(DEV, dtype)).sum(), [xr, mr, Lr])
        go = torch.autograd.grad((ref * w).sum(), [xo, mo, Lo])
        _close(g[0], go[0], gtol)
        _close(g[1], go[1], gtol * (1 if mu_.dim() == 2 else rows ** 0.5))
        _close(g[2], torch.tril(go[2]), gtol * (1 if mu_.dim() == 2 else rows ** 0.5))
