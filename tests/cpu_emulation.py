"""TEST INFRASTRUCTURE ONLY: stand-ins for the native entry points, built on the oracle, so the
HOST logic of pyro_b200 (effect handlers, plates/broadcasting, ELBO assembly, optimiser
bookkeeping, the lockstep NUTS driver, chain sharding) can be exercised by the CPU test tier
(`-m "not gpu"`), which has no GPU.  Activated explicitly by the ``cpu_emulation`` fixture; the
product never imports this module, and the `-m gpu` tier runs the same scenarios through the real
kernels.
"""
import contextlib

import torch

from oracle import dists as odists
from oracle import mcmc as omcmc
from oracle import optim as ooptim


def _site_score(family, value, params, shape, *, mask=None, scale=1.0, upstream=None, weight=1.0,
                sum_coeff=1.0, accumulate=False, want_logprob=False, out_sum=None,
                need_dvalue=False, need_dparams=None, event_size=None):
    need_dparams = need_dparams or [False] * len(params)
    with torch.enable_grad():
        ps = [p.detach().requires_grad_(True) for p in params]
        v = value
        if value is not None and value.is_floating_point():
            v = value.detach().requires_grad_(True)
        if event_size is None:
            fn, _ = odists.ELEMENTWISE[family]
            lp = fn(v, *ps)
            lp = lp.expand(shape) if tuple(lp.shape) != tuple(shape) else lp
        else:
            lp = odists.EVENT[family](v, *ps)
            lp = lp.expand(shape) if tuple(lp.shape) != tuple(shape) else lp
        slp = odists.scale_and_mask(lp, scale, mask)
        total = slp.sum()
        wanted = ([v] if need_dvalue else []) + [p for p, n in zip(ps, need_dparams) if n]
        grads = []
        if wanted:
            obj = (slp * upstream).sum() if upstream is not None else total
            grads = list(torch.autograd.grad(weight * obj, wanted, allow_unused=True))
            grads = [torch.zeros_like(w) if g is None else g for g, w in zip(grads, wanted)]
    if out_sum is not None:
        val = (sum_coeff * total.detach()).to(out_sum.dtype)
        if accumulate:
            out_sum.add_(val)
        else:
            out_sum.copy_(val)
    gv = grads.pop(0).detach() if need_dvalue else None
    gp = [(grads.pop(0).detach() if n else None) for n in need_dparams]
    return (slp.detach().clone() if want_logprob else None), gv, gp


def _adam_launch(self, t):
    for p in t["params"]:
        h = self._host[p]
        a = h["args"]
        o = ooptim.ClippedAdam(lr=h["lr"], betas=tuple(a["betas"]), eps=a["eps"],
                               weight_decay=a["weight_decay"], clip_norm=a["clip_norm"], lrd=a["lrd"])
        o.step_count, o.exp_avg, o.exp_avg_sq = h["step"], h["exp_avg"], h["exp_avg_sq"]
        with torch.no_grad():
            o.step(p.data, p.grad)
            p.grad.zero_()
        h["step"], h["lr"] = o.step_count, o.lr
    # keep the device-table mirror in sync with what the real kernel would have written
    t["lrs"] = torch.tensor([self._host[p]["lr"] for p in t["params"]], dtype=torch.float64)
    t["steps"] = torch.tensor([self._host[p]["step"] for p in t["params"]], dtype=torch.int32)


def _agr_launch(self, t):
    for p in t["params"]:
        h = self._host[p]
        a = h["args"]
        o = ooptim.AdagradRMSProp(eta=a["eta"], delta=a["delta"], t=a["t"])
        o.step_count = h["step"]
        o.sum = h["sum"] if h["step"] > 0 else None
        with torch.no_grad():
            o.step(p.data, p.grad)
            p.grad.zero_()
        h["sum"].copy_(o.sum)
        h["step"] = o.step_count
    t["steps"] = torch.tensor([self._host[p]["step"] for p in t["params"]], dtype=torch.int32)


def _leaf_vector(self, z, rcur, g, minv, active8, take8, rsub, zs, gs, rck, sck, leaf):
    """torch stand-in for b2_nuts_leaf_vector (same contract, in-place on rsub/zs/gs/rck/sck)."""
    active, take = active8.bool(), take8.bool()
    ru = rcur * minv.sqrt()
    rsub.add_(torch.where(active[:, None], ru, torch.zeros_like(ru)))
    zs.copy_(torch.where((take & active)[:, None], z, zs))
    gs.copy_(torch.where((take & active)[:, None], g, gs))
    idx_max = bin(leaf >> 1).count("1")
    turn = torch.zeros(z.shape[0], dtype=torch.bool)
    if leaf % 2 == 0:
        am = active[:, None]
        rck[idx_max] = torch.where(am, ru, rck[idx_max])
        sck[idx_max] = torch.where(am, rsub, sck[idx_max])
    else:
        t, nblk = leaf, 0
        while t & 1:
            nblk += 1
            t >>= 1
        for k in range(idx_max, idx_max - nblk, -1):
            blk = rsub - sck[k] + rck[k]
            rho = blk - 0.5 * (rck[k] + ru)
            turn = turn | ((rck[k] * rho).sum(-1) <= 0) | ((ru * rho).sum(-1) <= 0)
    return turn


def _leaf_hier(self, st, leaf):
    """torch stand-in for b2_nuts_leaf_hier (include/pyro_b200.h): one new leaf for every active
    chain, the end picked by ``dir`` advanced in place, scalar tree logic included; the uniform for
    the multinomial draw comes from the kernel's torch generator instead of Philox."""
    t = st["t"]
    done = t["done"].bool()
    act = ~done
    am = act[:, None]
    right = t["dir"].bool()
    rm = right[:, None]
    z = torch.where(rm, t["zR"], t["zL"])
    r = torch.where(rm, t["rR"], t["rL"])
    minv, eps = t["minv"], t["eps"]
    tk = (t["take"].bool() & act)[:, None]
    t["zs"].copy_(torch.where(tk, z, t["zs"]))
    _, g_old = self.potential.value_and_grad(z)
    e = eps[:, None]
    rh = r - 0.5 * e * g_old
    z2 = z + e * (minv * rh)
    U2, g2 = self.potential.value_and_grad(z2)
    r2 = rh - 0.5 * e * g2
    ke = 0.5 * (minv * r2 * r2).sum(-1)
    for side, sel in (("R", am & rm), ("L", am & ~rm)):
        t["z" + side].copy_(torch.where(sel, z2, t["z" + side]))
        t["r" + side].copy_(torch.where(sel, r2, t["r" + side]))
        t["gsc" + side].copy_(torch.where(sel, g2[:, :2], t["gsc" + side]))
    t["U"].copy_(torch.where(act, U2, t["U"]))
    ru = r2 * minv.sqrt()
    rsub = t["rsub"]
    if leaf == 0:
        rsub.copy_(torch.where(am, ru, rsub))
    else:
        rsub.add_(torch.where(am, ru, torch.zeros_like(ru)))
    idx_max = bin(leaf >> 1).count("1")
    turn = torch.zeros(z.shape[0], dtype=torch.bool)
    rck, sck = t["rck"], t["sck"]
    if leaf % 2 == 0:
        rck[idx_max] = torch.where(am, ru, rck[idx_max])
        sck[idx_max] = torch.where(am, rsub, sck[idx_max])
    else:
        q, nblk = leaf, 0
        while q & 1:
            nblk += 1
            q >>= 1
        for k in range(idx_max, idx_max - nblk, -1):
            rho = (rsub - sck[k] + rck[k]) - 0.5 * (rck[k] + ru)
            turn = turn | ((rck[k] * rho).sum(-1) <= 0) | ((ru * rho).sum(-1) <= 0)
    energy = U2 + ke
    energy = torch.where(torch.isnan(energy), torch.full_like(energy, float("inf")), energy)
    delta = energy - t["energy0"]
    div_now = act & (delta > 1000.0)
    accp = (-delta).exp().clamp(max=1.0)
    t["sum_accept"].add_(torch.where(act, accp, torch.zeros_like(accp)))
    t["num_prop"].add_(act.to(accp.dtype))
    t["num_leapfrogs"].add_(act.to(torch.int32))
    w_leaf = -delta
    lws = t["logw_sub"]
    if leaf == 0:
        nw, take = w_leaf, act.clone()
    else:
        m = torch.maximum(lws, w_leaf)
        ms = torch.where(torch.isinf(m) & (m < 0), torch.zeros_like(m), m)
        nw = ms + torch.log(torch.exp(lws - ms) + torch.exp(w_leaf - ms))
        take = act & (torch.rand(z.shape[0], generator=self._gen, dtype=z.dtype) < torch.exp(w_leaf - nw))
    lws.copy_(torch.where(act, nw, lws))
    t["Us"].copy_(torch.where(take, U2, t["Us"]))
    t["gsc_s"].copy_(torch.where(take[:, None], g2[:, :2], t["gsc_s"]))
    t["take"].copy_(torch.where(act, take, t["take"].bool()).to(torch.uint8))
    t["diverged"].copy_((t["diverged"].bool() | div_now).to(torch.uint8))
    t["done"].copy_((done | div_now | (act & turn & ~div_now)).to(torch.uint8))


def _tree_merge(self, t, rsum):
    act = ~t["done"].bool()
    am = act[:, None]
    sq = t["minv"].sqrt()
    ul, ur = t["rL"] * sq, t["rR"] * sq
    rsum.copy_(torch.where(am, rsum + t["rsub"], rsum))
    rho = rsum - 0.5 * (ul + ur)
    dots = torch.stack([(ul * rho).sum(-1), (ur * rho).sum(-1)], dim=-1)
    return torch.where(am, dots, torch.zeros_like(dots))


def _rows_copy(self, dst, src, mask):
    dst.copy_(torch.where(mask.bool()[:, None], src, dst))


def _native_value_and_grad(self, z, active=None, out_grad=None):
    from pyro_b200 import _native as N
    if self.model_id == N.MODEL_HIER_NORMAL:
        U = omcmc.eight_schools_potential(self._keep[0], self._keep[1], self._model.hyper[0], self._model.hyper[1])
    else:
        U = omcmc.logistic_potential(self._keep[0], self._keep[1], self._model.hyper[0])
    Us, gs = [], []
    for zc in z:
        g, u = omcmc.potential_grad(U, zc)
        Us.append(u)
        gs.append(g)
    G = torch.stack(gs)
    if out_grad is not None:
        a = torch.ones(z.shape[0], dtype=torch.bool) if active is None else active.bool()
        out_grad.copy_(torch.where(a[:, None], G, out_grad))
        G = out_grad
    return torch.stack(Us), G


def _leapfrog(self, z, r, g, eps, minv, active=None):
    a = torch.ones(z.shape[0], dtype=torch.bool) if active is None else active.bool()
    e = eps[:, None]
    am = a[:, None]
    r.copy_(torch.where(am, r + 0.5 * e * (-g), r))
    z.copy_(torch.where(am, z + e * (minv * r), z))
    g_old = g.clone()
    U, g_new = self.potential.value_and_grad(z, active)
    g_new = torch.where(am, g_new, g_old)
    r.copy_(torch.where(am, r + 0.5 * e * (-g_new), r))
    ke = 0.5 * (minv * r * r).sum(-1)
    self.num_leapfrogs += z.shape[0]
    return z, r, g_new, U, ke


def _reduce_to(src, dst):
    nd = src.dim()
    view = dst.reshape((1,) * (nd - dst.dim()) + tuple(dst.shape))
    view.copy_(src.sum_to_size(view.shape))


def _normal_rsample_score(loc, scale, eps):
    # families 14/15 of include/pyro_b200.h, restated with the oracle's Normal log density
    z = loc + eps * scale
    lq = odists.ELEMENTWISE[0][0](z, loc, scale).sum()
    return z.detach(), lq.detach()


def _normal_rsample_backward(gz, eps, loc, scale, c, need_loc, need_scale):
    gz = gz.expand(eps.shape)
    gloc = gz.sum_to_size(loc.shape) if need_loc else None
    gscale = (gz * eps - c / scale).sum_to_size(scale.shape) if need_scale else None
    return gloc, gscale


def _latent_prior(items):
    return [odists.ELEMENTWISE[0][0](z, pl, ps).sum().detach() for z, pl, ps in items]


def _latent_prior_combine(items, item_coeffs, terms, term_coeffs):
    tot = sum(c * odists.ELEMENTWISE[0][0](z, pl, ps).sum().detach() for c, (z, pl, ps) in zip(item_coeffs, items))
    return tot + sum(c * t.reshape(()) for c, t in zip(term_coeffs, terms))


class _no_defer:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _latent_backward(gz, eps, z, loc, scale, log_scale, c, prior, need_loc, need_scale, accumulate=False):
    # csrc/latent.cu latent_backward_kernel restated: g = gz + pw * dlogp/dz; d/dloc = g;
    # d/dscale = g*eps - c/s  or  d/dlog_scale = g*eps*s - c
    s = scale.exp() if log_scale else scale
    g = torch.zeros_like(eps) if gz is None else gz.expand(eps.shape)
    if prior is not None:
        pl, ps, pw = prior
        g = g + pw * (-(z - pl) / (ps * ps))
    gloc = g.sum_to_size(loc.shape) if need_loc else None
    gs = (g * eps * s - c) if log_scale else (g * eps - c / s)
    gscale = gs.sum_to_size(scale.shape) if need_scale else None
    if accumulate:
        import pyro_b200.distributions._ops as ops
        for t, g, which in ((loc, gloc, 0), (scale, gscale, 1)):
            slot = ops._grad_slot(t) if g is not None else None
            if slot is not None:
                slot.add_(g)
                if which == 0:
                    gloc = None
                else:
                    gscale = None
    return gloc, gscale


def _elbo_combine(terms, coeffs):
    return sum(c * t.reshape(()) for c, t in zip(coeffs, terms))


@contextlib.contextmanager
def enabled():
    """Patch the native seams with oracle-backed CPU stand-ins."""
    import pyro_b200._native as N
    import pyro_b200.distributions._ops as ops
    import pyro_b200.infer.mcmc.nuts as nuts
    import pyro_b200.infer.mcmc.potential as pot
    import pyro_b200.optim as optim
    import pyro_b200.distributions as pdist
    saved_glm = pdist._BernoulliLinear._fused_sum
    # the fused GLM kernel has no CPU stand-in: route the site through the dense Bernoulli path
    pdist._BernoulliLinear._fused_sum = lambda self, value, mask, scale, weight, sum_coeff, unit=True: \
        pdist.Bernoulli._fused_sum(self, value, mask, scale, weight, sum_coeff, unit)
    saved = (ops.site_score, N.require_cuda, optim.ClippedAdam._launch, optim.AdagradRMSProp._launch,
             pot.NativePotential.value_and_grad, nuts.HMC._leapfrog, ops.reduce_to)
    ops.site_score = _site_score
    N.require_cuda = lambda t, what: None
    optim.ClippedAdam._launch = _adam_launch
    optim.AdagradRMSProp._launch = _agr_launch
    pot.NativePotential.value_and_grad = _native_value_and_grad
    nuts.HMC._leapfrog = _leapfrog
    saved_leaf = nuts.NUTS._leaf_vector
    nuts.NUTS._leaf_vector = _leaf_vector
    saved_leaf_hier = (nuts.NUTS._leaf_hier, nuts.NUTS._tree_merge, nuts.NUTS._rows_copy)
    nuts.NUTS._leaf_hier, nuts.NUTS._tree_merge, nuts.NUTS._rows_copy = _leaf_hier, _tree_merge, _rows_copy
    ops.reduce_to = _reduce_to
    saved_rs = (ops.normal_rsample_score, ops.normal_rsample_backward, N.EMULATE_RSAMPLE)
    saved_latent = (ops.latent_prior, ops.latent_backward, ops.latent_prior_combine, ops.deferred_latent_backward)
    ops.latent_prior, ops.latent_backward = _latent_prior, _latent_backward
    ops.latent_prior_combine, ops.deferred_latent_backward = _latent_prior_combine, _no_defer
    saved_comb = ops.elbo_combine
    ops.elbo_combine = _elbo_combine
    ops.normal_rsample_score = _normal_rsample_score
    ops.normal_rsample_backward = _normal_rsample_backward
    N.EMULATE_RSAMPLE = True
    try:
        yield
    finally:
        ops.normal_rsample_score, ops.normal_rsample_backward, N.EMULATE_RSAMPLE = saved_rs
        (ops.latent_prior, ops.latent_backward, ops.latent_prior_combine,
         ops.deferred_latent_backward) = saved_latent
        ops.elbo_combine = saved_comb
        pdist._BernoulliLinear._fused_sum = saved_glm
        nuts.NUTS._leaf_vector = saved_leaf
        nuts.NUTS._leaf_hier, nuts.NUTS._tree_merge, nuts.NUTS._rows_copy = saved_leaf_hier
        (ops.site_score, N.require_cuda, optim.ClippedAdam._launch, optim.AdagradRMSProp._launch,
         pot.NativePotential.value_and_grad, nuts.HMC._leapfrog, ops.reduce_to) = saved
