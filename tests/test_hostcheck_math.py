"""The element functors and the NUTS core that the CUDA kernels execute, run on the CPU
(libpyro_b200_hostcheck.so) and checked against the reference goldens / the oracle.  These tests
do not need a GPU; the -m gpu tier repeats the comparison through the real kernels."""
import numpy as np
import pytest
import torch

import hostcheck as H
from conftest import load_npz
from oracle import mcmc as omcmc

FAMS = {"normal": 0, "bernoulli_logits": 1, "gamma": 2, "beta": 3, "poisson": 4, "cauchy": 5,
        "halfcauchy": 6, "exponential": 7, "lognormal": 8, "halfnormal": 9, "bernoulli_probs": 10,
        "uniform": 11}


@pytest.mark.parametrize("key", sorted(FAMS))
@pytest.mark.parametrize("dtype,tol,gtol", [(np.float64, 1e-12, 1e-10), (np.float32, 1e-5, 2e-4)])
def test_functors_match_reference(key, dtype, tol, gtol):
    g = load_npz("dist_random.npz")
    v = g[key + ".value"]
    ps = []
    i = 0
    while key + ".p%d" % i in g:
        ps.append(g[key + ".p%d" % i])
        i += 1
    lp, dx, dp = H.eval_family(FAMS[key], v, ps, dtype)
    ref = g[key + ".lp"]
    # tolerance: |d| <= tol * max(1, |lp|)  (reference bar: atol 1e-5, tests/common.py:246-248)
    assert np.all(np.abs(lp - ref) <= tol * np.maximum(1, np.abs(ref)))
    if key + ".dvalue" in g:
        r = g[key + ".dvalue"]
        assert np.all(np.abs(dx - r) <= gtol * np.maximum(1, np.abs(r)))
    for k in range(len(ps)):
        r = g[key + ".dp%d" % k]
        assert np.all(np.abs(dp[k] - r) <= gtol * np.maximum(1, np.abs(r))), k


def test_kl_functors():
    g = load_npz("kl.npz")
    for name, fam in (("normal", 12), ("gamma", 13)):
        ps = [g["%s.p%d" % (name, i)] for i in range(4)]
        lp, _, dp = H.eval_family(fam, None, ps)
        assert np.allclose(lp, g[name + ".kl"], atol=1e-12, rtol=1e-12)
        for k in range(4):
            # torch's float64 polygamma(1) (in the golden) is only ~1e-9 accurate
            assert np.allclose(dp[k], g["%s.dp%d" % (name, k)], atol=1e-7, rtol=1e-7)


def test_fast_gamma_functions_sweep_fp32():
    """The closed-form shift-by-4 + Stirling evaluation of lgamma / digamma (fp32 kernels) over six
    decades of concentration, through the Gamma functor: log density within 1e-5 * max(1, |lp|) and
    d/d concentration within 2e-4 * max(1, |g|) of the fp64 reference formulas (gamma.py:89-98)."""
    a = np.concatenate([np.logspace(-3, 3, 400), [0.5, 1.0, 2.0, 3.999999, 4.0, 4.000001]])
    x = np.full_like(a, 1.3)
    b = np.full_like(a, 0.7)
    lp, _, dp = H.eval_family(2, x, [a, b], np.float32)
    at = torch.tensor(a, requires_grad=True)
    ref = torch.distributions.Gamma(at, torch.tensor(b)).log_prob(torch.tensor(x))
    (ga,) = torch.autograd.grad(ref.sum(), at)
    ref, ga = ref.detach().numpy(), ga.numpy()
    assert np.all(np.abs(lp - ref) <= 1e-5 * np.maximum(1, np.abs(ref)))
    assert np.all(np.abs(dp[0] - ga) <= 2e-4 * np.maximum(1, np.abs(ga)))


def test_kl_functors_fp32():
    """fp32 KL kernels (fast lgamma / digamma / trigamma, SFU log and reciprocal on the device):
    value within 2e-5 * max(1, |kl|) of the fp64 reference, gradients within 3e-4."""
    g = load_npz("kl.npz")
    for name, fam in (("normal", 12), ("gamma", 13)):
        ps = [g["%s.p%d" % (name, i)] for i in range(4)]
        lp, _, dp = H.eval_family(fam, None, ps, np.float32)
        ref = g[name + ".kl"]
        assert np.all(np.abs(lp - ref) <= 2e-5 * np.maximum(1, np.abs(ref)))
        for k in range(4):
            r = g["%s.dp%d" % (name, k)]
            assert np.all(np.abs(dp[k] - r) <= 3e-4 * np.maximum(1, np.abs(r))), (name, k)


def test_fused_normal_draw_functors():
    """Families 14 / 15 (include/pyro_b200.h): the fused draw reproduces the reference's
    rsample (normal.py:82-85) and log_prob (normal.py:87-102) on the golden Normal fixture's
    parameters, and its backward is the chain rule of L = sum(gz*z) + c*sum log q(z)."""
    g = load_npz("dist_random.npz")
    loc, scale = g["normal.p0"], g["normal.p1"]
    rng = np.random.default_rng(0)
    eps, gz = rng.standard_normal(loc.shape), rng.standard_normal(loc.shape)
    lp, z, _ = H.eval_family(14, eps, [loc, scale])
    assert np.allclose(z, loc + eps * scale, rtol=1e-15, atol=0)
    lo, so = torch.tensor(loc, requires_grad=True), torch.tensor(scale, requires_grad=True)
    zt = lo + torch.tensor(eps) * so
    lq = torch.distributions.Normal(lo, so).log_prob(zt)
    assert np.allclose(lp, lq.detach().numpy(), rtol=1e-12, atol=1e-12)
    c = np.full(loc.shape, 0.37)
    _, _, dp = H.eval_family(15, gz, [eps, scale, c])
    L = (torch.tensor(gz) * zt).sum() + 0.37 * lq.sum()
    gl, gs = torch.autograd.grad(L, [lo, so])
    assert np.allclose(dp[0], gl.numpy(), rtol=1e-12, atol=1e-12)
    assert np.allclose(dp[1], gs.numpy(), rtol=1e-9, atol=1e-9)


def test_digamma_special_values():
    L = H.lib()
    for x in [-2.5, -0.3, 1e-8, 0.5, 1.0, 5.9, 6.0, 100.0, 1e6]:
        ref = float(torch.digamma(torch.tensor(x, dtype=torch.float64)))
        assert abs(L.b2h_digamma(x) - ref) <= 1e-12 * max(1, abs(ref))
    assert L.b2h_digamma(0.0) == -np.inf


def test_native_potentials_match_reference():
    g = load_npz("mcmc.npz")
    for z, u_ref, g_ref in zip(g["es.Z"], g["es.U"], g["es.G"]):
        U, gr = H.potential_hier_normal(g["es.y"], g["es.sigma"], 10.0, 25.0, z)
        assert abs(U - u_ref) < 1e-10 * max(1, abs(u_ref))
        assert np.allclose(gr, g_ref, atol=1e-10, rtol=1e-10)
    for b, u_ref, g_ref in zip(g["lr.B"], g["lr.U"], g["lr.G"]):
        U, gr = H.potential_logistic(g["lr.X"], g["lr.y"], 1.0, b)
        assert abs(U - u_ref) < 1e-10 * max(1, abs(u_ref))
        assert np.allclose(gr, g_ref, atol=1e-10, rtol=1e-10)


def _iterative_turning(rs):
    """U-turn decisions of the iterative checkpoint scheme (nuts_core.cuh) for a momentum sequence."""
    n = len(rs)
    rsub = np.zeros_like(rs[0])
    rck, sck = {}, {}
    for leaf in range(n):
        rsub = rsub + rs[leaf]
        idx_max = bin(leaf >> 1).count("1")
        if leaf % 2 == 0:
            rck[idx_max], sck[idx_max] = rs[leaf].copy(), rsub.copy()
        else:
            t = leaf
            nblk = 0
            while t & 1:
                nblk += 1
                t >>= 1
            for k in range(idx_max, idx_max - nblk, -1):
                blk = rsub - sck[k] + rck[k]
                rho = blk - 0.5 * (rck[k] + rs[leaf])
                if rck[k] @ rho <= 0 or rs[leaf] @ rho <= 0:
                    return leaf
    return None


def _recursive_turning(rs, lo, hi):
    """First leaf index at which the reference recursion (nuts.py:250-365) stops for leaves [lo,hi)."""
    if hi - lo == 1:
        return None
    mid = (lo + hi) // 2
    a = _recursive_turning(rs, lo, mid)
    if a is not None:
        return a
    b = _recursive_turning(rs, mid, hi)
    if b is not None:
        return b
    rsum = np.sum(rs[lo:hi], axis=0)
    rho = rsum - 0.5 * (rs[lo] + rs[hi - 1])
    if rs[lo] @ rho <= 0 or rs[hi - 1] @ rho <= 0:
        return hi - 1
    return None


def test_iterative_uturn_equals_recursive():
    rng = np.random.default_rng(0)
    hits = 0
    for trial in range(300):
        depth = rng.integers(1, 7)
        n = 1 << depth
        # momenta drifting in direction so that turns happen at varied places
        base = rng.standard_normal(3)
        rs = [base + 0.9 * rng.standard_normal(3) * (1 + 0.3 * i) for i in range(n)]
        rs = np.asarray(rs)
        a, b = _iterative_turning(rs), _recursive_turning(rs, 0, n)
        assert a == b
        hits += a is not None
    assert 30 < hits < 290


def test_device_nuts_core_recovers_reference_posterior():
    """nuts_core.cuh (the code nuts_small_kernel runs per chain), executed on the CPU for
    eight_schools: posterior moments must agree with the reference's long run
    (golden es.long.*), in the spirit of tests/infer/mcmc/test_nuts.py tolerances."""
    g = load_npz("mcmc.npz")
    y, sigma = g["es.y"], g["es.sigma"]
    C, D = 8, 10
    rng = np.random.default_rng(1)
    z = rng.uniform(-2, 2, (C, D))
    U = np.zeros(C)
    gr = np.zeros((C, D))
    for c in range(C):
        U[c], gr[c] = H.potential_hier_normal(y, sigma, 10.0, 25.0, z[c])
    eps = np.full(C, 0.1)
    minv = np.ones((C, D))
    # crude warm-up: a few hundred transitions at small step, then sample with a tuned step
    H.nuts_hier_normal(y, sigma, 10.0, 25.0, z, U, gr, eps, minv, 200, seed=3)
    eps = np.full(C, 0.35)
    samples, acc, depth, div, steps = H.nuts_hier_normal(y, sigma, 10.0, 25.0, z, U, gr, eps, minv, 1500, seed=4)
    mu = samples[..., 0].reshape(-1)
    tau = np.exp(samples[..., 1]).reshape(-1)
    eta = samples[..., 2:].reshape(-1, 8)
    assert abs(mu.mean() - float(g["es.long.mu.mean"][0])) < 0.5
    assert abs(tau.mean() - float(g["es.long.tau.mean"][0])) < 0.7
    assert np.max(np.abs(eta.mean(0) - g["es.long.eta.mean"])) < 0.12
    assert 0.5 < acc.mean() < 0.99
    assert steps.min() >= 1 and depth.max() <= 10
    # tree sizes: a transition of depth d that was not cut short has 2^d - 1 leapfrogs
    full = (div == 0)
    assert np.all(steps[full] <= (1 << depth[full].clip(max=10)) * 2)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors: zero key/counter, and the
    pi/e digits test)."""
    import ctypes
    out = np.zeros(4, np.uint32)
    H.lib().b2h_philox(ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0), 4,
                       out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]


def test_standard_gamma_grad_against_finite_differences_and_torch():
    """csrc/gamma_sample.cuh ``standard_gamma_grad`` (the code b2_gamma_rsample runs, built for the host): the
    implicit-reparameterisation derivative dx/da against a central finite difference of scipy's inverse
    incomplete gamma function (fp64, 1e-6), the fp32 evaluation against the fp64 one (3e-4), and ATen's
    ``_standard_gamma_grad`` -- the reference's backward of ``Gamma.rsample`` (torch/distributions/gamma.py:79-87)
    -- within the accuracy of ITS approximation (2e-3)."""
    import ctypes
    from scipy.special import gammainc, gammaincinv
    L = H.lib()
    L.b2h_standard_gamma_grad.restype = ctypes.c_double
    L.b2h_standard_gamma_grad.argtypes = [ctypes.c_double, ctypes.c_double]
    L.b2h_standard_gamma_gradf.restype = ctypes.c_float
    L.b2h_standard_gamma_gradf.argtypes = [ctypes.c_float, ctypes.c_float]
    rng = np.random.default_rng(0)
    worst = [0.0, 0.0, 0.0]
    for a in (0.05, 0.1, 0.3, 0.5, 0.9, 1.0, 1.5, 2.0, 3.0, 5.0, 8.0, 12.0, 20.0, 50.0, 100.0, 300.0, 1000.0):
        xs = np.concatenate([rng.gamma(a, size=30), [a * 0.1, a, a + 1, a + 1.0001, 3 * a + 3]])
        for x in xs[xs > 1e-30]:
            g64 = L.b2h_standard_gamma_grad(a, float(x))
            g32 = float(L.b2h_standard_gamma_gradf(a, float(x)))
            p = gammainc(a, x)
            eps = 1e-6 * max(1.0, a)
            fd = (gammaincinv(a + eps, p) - gammaincinv(a - eps, p)) / (2 * eps)
            if np.isfinite(fd) and 1e-12 < p < 1 - 1e-12:
                worst[0] = max(worst[0], abs(g64 - fd) / abs(fd))
            worst[1] = max(worst[1], abs(g32 - g64) / abs(g64))
            t = float(torch._standard_gamma_grad(torch.tensor([a], dtype=torch.float64),
                                                 torch.tensor([x], dtype=torch.float64)))
            worst[2] = max(worst[2], abs(t - g64) / abs(g64))
    assert worst[0] < 1e-6 and worst[1] < 3e-4 and worst[2] < 2e-3, worst


def test_standard_gamma_sampler_goodness_of_fit():
    """Marsaglia-Tsang on the Philox stream (csrc/gamma_sample.cuh, host build): Kolmogorov-Smirnov against the
    Gamma CDF and the first two moments, shapes below and above 1, fp32 and fp64 (in the style of
    tests/distributions/test_distributions.py:138-164)."""
    import ctypes
    from scipy import stats
    L = H.lib()
    n = 40000
    for a in (0.2, 0.7, 1.0, 2.5, 17.0):
        out = np.zeros(n, np.float64)
        L.b2h_standard_gamma_sample(ctypes.c_uint64(7), ctypes.c_int64(n), ctypes.c_double(a),
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        outf = np.zeros(n, np.float32)
        L.b2h_standard_gamma_samplef(ctypes.c_uint64(9), ctypes.c_int64(n), ctypes.c_float(a),
                                     outf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        for s in (out, outf.astype(np.float64)):
            assert (s > 0).all()
            assert stats.kstest(s, "gamma", args=(a,)).pvalue > 1e-3, a
            assert abs(s.mean() - a) < 5 * (a / n) ** 0.5
            assert abs(s.var() - a) < 6 * a * (2 / n + 6 / (a * n)) ** 0.5


def test_beta_shared_evaluation_sweep_fp32():
    """Round 2's Beta functor evaluates lgamma(c1+c0) - lgamma(c1) - lgamma(c0) and the three digammas together
    (one reciprocal, one shift logarithm; csrc/b2_math.cuh ``lbeta_terms_f32``): log density and both parameter
    gradients over five decades of both concentrations -- including the hand-over points of the joint route
    (1e-6, a sum of 1e9, the shift at 4) -- against the fp64 reference formulas (beta.py:87-91 ->
    dirichlet.py:90-97), same tolerances as the Gamma sweep."""
    g = np.concatenate([np.logspace(-2.5, 2.5, 40), [0.5, 1.0, 3.999999, 4.0, 4.000001]])
    c1, c0 = [v.ravel() for v in np.meshgrid(g, g)]
    extra1 = np.array([5e-7, 2e-6, 1e-6, 0.3, 6e8, 7e8, 2.0, 1e-3])
    extra0 = np.array([0.4, 5e-7, 2.0, 9e8, 6e8, 2e8, 3.999999, 1e3])
    c1, c0 = np.concatenate([c1, extra1]), np.concatenate([c0, extra0])
    x = np.full_like(c1, 0.37)
    lp, _, dp = H.eval_family(3, x, [c1, c0], np.float32)
    a1 = torch.tensor(c1, requires_grad=True)
    a0 = torch.tensor(c0, requires_grad=True)
    ref = torch.distributions.Beta(a1, a0).log_prob(torch.tensor(x))
    g1, g0 = torch.autograd.grad(ref.sum(), [a1, a0])
    ref, g1, g0 = ref.detach().numpy(), g1.numpy(), g0.numpy()
    # the density is a difference of terms that reach 1e3 .. 2e10 at these concentrations: the fp32 bound is
    # relative to the LARGEST term (torch's own fp32 evaluation shows the same cancellation), not to the result
    from scipy.special import gammaln
    terms = (np.abs(gammaln(c1 + c0)) + np.abs(gammaln(c1)) + np.abs(gammaln(c0)) +
             np.abs((c1 - 1) * np.log(0.37)) + np.abs((c0 - 1) * np.log(0.63)))
    tol = 1e-5 * np.maximum(1, np.abs(ref)) + 3e-7 * terms
    assert np.all(np.abs(lp - ref) <= tol), float(np.max(np.abs(lp - ref) / tol))
    ref32 = torch.distributions.Beta(torch.tensor(c1, dtype=torch.float32), torch.tensor(c0, dtype=torch.float32)
                                     ).log_prob(torch.tensor(x, dtype=torch.float32)).numpy()
    assert np.max(np.abs(lp - ref) / tol) <= 2 * max(1.0, np.max(np.abs(ref32 - ref) / tol))
    gt = 2e-4 + 3e-6 * np.log(c1 + c0 + 2)
    assert np.all(np.abs(dp[0] - g1) <= gt * np.maximum(1, np.abs(g1)))
    assert np.all(np.abs(dp[1] - g0) <= gt * np.maximum(1, np.abs(g0)))


def test_poisson_and_lognormal_fast_paths_fp32():
    """The fp32 device routes added in round 2 share their algebra with the host build: Poisson (integer counts
    below / above the 64-entry log-factorial table, non-integer values) and LogNormal / Exponential / HalfNormal
    (reciprocal + logarithm form) against torch's fp64 formulas."""
    k = np.concatenate([np.arange(0, 80, dtype=np.float64), [63.0, 64.0, 1000.0, 2.5, 0.3]])
    rate = np.linspace(0.05, 60.0, k.size)
    lp, _, dp = H.eval_family(4, k, [rate], np.float32)
    rt = torch.tensor(rate, requires_grad=True)
    ref = torch.tensor(k) * rt.log() - rt - torch.lgamma(torch.tensor(k) + 1)
    (gr,) = torch.autograd.grad(ref.sum(), rt)
    assert np.all(np.abs(lp - ref.detach().numpy()) <= 1e-5 * np.maximum(1, np.abs(ref.detach().numpy())))
    assert np.all(np.abs(dp[0] - gr.numpy()) <= 2e-4 * np.maximum(1, np.abs(gr.numpy())))
    x = np.logspace(-3, 3, 60)
    loc, sc = np.linspace(-2, 2, 60), np.logspace(-1, 1, 60)
    for fam, tdist, params in ((8, torch.distributions.LogNormal, [loc, sc]),
                               (7, torch.distributions.Exponential, [sc]),
                               (9, torch.distributions.HalfNormal, [sc])):
        lp, dx, dp = H.eval_family(fam, x, params, np.float32)
        ps = [torch.tensor(p, requires_grad=True) for p in params]
        xt = torch.tensor(x, requires_grad=True)
        ref = tdist(*ps).log_prob(xt)
        grads = torch.autograd.grad(ref.sum(), [xt] + ps)
        r = ref.detach().numpy()
        assert np.all(np.abs(lp - r) <= 1e-5 * np.maximum(1, np.abs(r))), fam
        assert np.all(np.abs(dx - grads[0].numpy()) <= 2e-4 * np.maximum(1, np.abs(grads[0].numpy()))), fam
        for got, want in zip(dp, grads[1:]):
            assert np.all(np.abs(got - want.numpy()) <= 2e-4 * np.maximum(1, np.abs(want.numpy()))), fam
