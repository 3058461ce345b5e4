"""ORACLE (test infrastructure).  HMC/NUTS as reference Pyro runs it, restated for flat CPU tensors.

* ``velocity_verlet``        pyro/ops/integrator.py:14-65
* ``potential_grad``         pyro/ops/integrator.py:68-94 (NaN-energy convention)
* ``eight_schools_potential`` / ``logistic_potential``
                             pyro/infer/mcmc/util.py:275-286 applied to
                             examples/eight_schools/mcmc.py:27-34 and tests/infer/mcmc/test_hmc.py:189-198
* ``NUTSChain``              pyro/infer/mcmc/nuts.py:184-522 (RECURSIVE tree, multinomial sampling),
                             pyro/infer/mcmc/hmc.py:152-248 (kinetic energy, momentum draw,
                             find_reasonable_step_size)
* ``DualAveraging``, ``Welford``, ``adaptation_schedule``, ``WarmupAdapter``
                             pyro/ops/dual_averaging.py:43-79, pyro/ops/welford.py:7-51,
                             pyro/infer/mcmc/adaptation.py:65-202
State is one flat vector per chain (the reference keeps a dict of sites; with a diagonal mass
matrix over all sites the two are the same arithmetic).
"""
import math
from collections import namedtuple

import torch

from . import dists


# ---------------------------------------------------------------------------------------------
def potential_grad(potential_fn, z):
    z = z.detach().requires_grad_(True)
    U = potential_fn(z)
    (g,) = torch.autograd.grad(U, z)
    return g, U.detach()


def velocity_verlet(z, r, potential_fn, inv_mass, step_size, num_steps=1, z_grads=None):
    """Returns (z, r, z_grads, potential_energy).  kinetic_grad(r) = M^-1 r (diag)."""
    z, r = z.clone(), r.clone()
    U = None
    for _ in range(num_steps):
        if z_grads is None:
            z_grads, _ = potential_grad(potential_fn, z)
        r = r + 0.5 * step_size * (-z_grads)
        z = z + step_size * (inv_mass * r)
        z_grads, U = potential_grad(potential_fn, z)
        r = r + 0.5 * step_size * (-z_grads)
    return z, r, z_grads, U


def eight_schools_potential(y, sigma, s_mu=10.0, s_tau=25.0):
    """z = [mu, log tau, eta...]; tau's transform is ExpTransform (positive support)."""
    def U(z):
        mu, t, eta = z[0], z[1], z[2:]
        tau = t.exp()
        one = torch.ones((), dtype=z.dtype)
        lp = dists.normal(mu, torch.zeros((), dtype=z.dtype), one * s_mu)
        lp = lp + dists.half_cauchy(tau, one * s_tau)
        lp = lp + dists.normal(eta, torch.zeros((), dtype=z.dtype), one).sum()
        lp = lp + dists.normal(y, mu + tau * eta, sigma).sum()
        # log|d z / d tau| = -t  is subtracted from the log joint (mcmc/util.py:282-285)
        return -(lp + t)
    return U


def logistic_potential(X, y, prior_scale=1.0):
    def U(z):
        lp = dists.normal(z, torch.zeros((), dtype=z.dtype), torch.ones((), dtype=z.dtype) * prior_scale).sum()
        lp = lp + dists.bernoulli_logits(y, X @ z).sum()
        return -lp
    return U


# ---------------------------------------------------------------------------------------------
class DualAveraging:
    def __init__(self, prox_center=0.0, t0=10, kappa=0.75, gamma=0.05):
        self.prox_center, self.t0, self.kappa, self.gamma = prox_center, t0, kappa, gamma
        self.reset()

    def reset(self):
        self._x_avg = 0.0
        self._g_avg = 0.0
        self._t = 0

    def step(self, g):
        self._t += 1
        self._g_avg = (1 - 1 / (self._t + self.t0)) * self._g_avg + g / (self._t + self.t0)
        self._x_t = self.prox_center - (self._t ** 0.5) / self.gamma * self._g_avg
        weight_t = self._t ** (-self.kappa)
        self._x_avg = (1 - weight_t) * self._x_avg + weight_t * self._x_t

    def get_state(self):
        return self._x_t, self._x_avg


class Welford:
    def __init__(self):
        self.reset()

    def reset(self):
        self._mean, self._m2, self.n_samples = 0.0, 0.0, 0

    def update(self, sample):
        self.n_samples += 1
        delta_pre = sample - self._mean
        self._mean = self._mean + delta_pre / self.n_samples
        delta_post = sample - self._mean
        self._m2 = self._m2 + delta_pre * delta_post

    def get_covariance(self, regularize=True):
        cov = self._m2 / (self.n_samples - 1)
        if regularize:
            n = self.n_samples
            cov = (n / (n + 5.0)) * cov + 1e-3 * (5.0 / (n + 5.0))
        return cov


adapt_window = namedtuple("adapt_window", ["start", "end"])


def adaptation_schedule(warmup_steps):
    """pyro/infer/mcmc/adaptation.py:65-103"""
    sched = []
    if warmup_steps < 20:
        return [adapt_window(0, warmup_steps - 1)]
    start_buffer, end_buffer, init_window = 75, 50, 25
    if start_buffer + end_buffer + init_window > warmup_steps:
        start_buffer = int(0.15 * warmup_steps)
        end_buffer = int(0.1 * warmup_steps)
        init_window = warmup_steps - start_buffer - end_buffer
    sched.append(adapt_window(0, start_buffer - 1))
    end_window_start = warmup_steps - end_buffer
    next_size, next_start = init_window, start_buffer
    while next_start < end_window_start:
        cur_start, cur_size = next_start, next_size
        if 3 * cur_size <= end_window_start - cur_start:
            next_size = 2 * cur_size
        else:
            cur_size = end_window_start - cur_start
        next_start = cur_start + cur_size
        sched.append(adapt_window(cur_start, next_start - 1))
    sched.append(adapt_window(end_window_start, warmup_steps - 1))
    return sched


_Tree = namedtuple("_Tree", ["z_left", "r_left", "g_left", "z_right", "r_right", "g_right",
                             "z_prop", "U_prop", "g_prop", "r_sum", "weight", "turning",
                             "diverging", "sum_accept", "num_prop"])


def _logaddexp(x, y):
    m = max(x, y)
    if m == -math.inf:
        return m
    return m + math.log(math.exp(x - m) + math.exp(y - m))


class NUTSChain:
    """One chain of the reference sampler.  ``rng`` is a torch.Generator (CPU)."""

    def __init__(self, potential_fn, dim, dtype=torch.float64, step_size=1.0, adapt_step_size=True,
                 adapt_mass_matrix=True, target_accept_prob=0.8, max_tree_depth=10, seed=0):
        self.U = potential_fn
        self.D = dim
        self.dtype = dtype
        self.step_size = step_size
        self.adapt_step_size = adapt_step_size
        self.adapt_mass_matrix = adapt_mass_matrix
        self.target = target_accept_prob
        self.max_depth = max_tree_depth
        self.max_sliced_energy = 1000.0
        self.inv_mass = torch.ones(dim, dtype=dtype)
        self.rng = torch.Generator().manual_seed(seed)
        self.num_leapfrogs = 0
        self.divergences = 0
        self._direction_threshold = math.log(0.8)

    # -- helpers ------------------------------------------------------------------------------------
    def _randn(self, n):
        return torch.randn(n, generator=self.rng, dtype=self.dtype)

    def _rand(self):
        return float(torch.rand((), generator=self.rng, dtype=torch.float64))

    def _unscale(self, r):
        return r * self.inv_mass.sqrt()           # r_unscaled = M^{-1/2} r

    def _sample_r(self):
        r_u = self._randn(self.D)
        return r_u / self.inv_mass.sqrt(), r_u    # r = M^{1/2} r_unscaled

    def _leapfrog(self, z, r, g, eps):
        self.num_leapfrogs += 1
        return velocity_verlet(z, r, self.U, self.inv_mass, eps, z_grads=g)

    def find_reasonable_step_size(self, z):
        """hmc.py:170-229"""
        step_size = self.step_size
        U0 = float(self.U(z))
        r, r_u = self._sample_r()
        e0 = 0.5 * float(r_u @ r_u) + U0
        _, r_new, _, U1 = self._leapfrog(z, r, None, step_size)
        r_nu = self._unscale(r_new)
        delta = 0.5 * float(r_nu @ r_nu) + float(U1) - e0
        direction = 1 if self._direction_threshold < -delta else -1
        scale = 2.0 ** direction
        direction_new = direction
        while direction_new == direction and 1e-10 < step_size < 1e10:
            step_size = scale * step_size
            r, r_u = self._sample_r()
            e0 = 0.5 * float(r_u @ r_u) + U0
            _, r_new, _, U1 = self._leapfrog(z, r, None, step_size)
            r_nu = self._unscale(r_new)
            delta = 0.5 * float(r_nu @ r_nu) + float(U1) - e0
            direction_new = 1 if self._direction_threshold < -delta else -1
        return min(max(step_size, 1e-10), 1e10)

    # -- tree (nuts.py:184-365) ---------------------------------------------------------------------------
    @staticmethod
    def is_turning(r_left_u, r_right_u, r_sum):
        rho = r_sum - (r_left_u + r_right_u) / 2
        return float(r_left_u @ rho) <= 0 or float(r_right_u @ rho) <= 0

    def _build_basetree(self, z, r, g, log_slice, direction, energy_current):
        eps = self.step_size if direction == 1 else -self.step_size
        z_new, r_new, g_new, U_new = self._leapfrog(z, r, g, eps)
        r_nu = self._unscale(r_new)
        energy_new = float(U_new) + 0.5 * float(r_nu @ r_nu)
        if math.isnan(energy_new):
            energy_new = math.inf
        sliced = energy_new + log_slice
        diverging = sliced > self.max_sliced_energy
        delta = energy_new - energy_current
        accept_prob = min(1.0, math.exp(-delta)) if delta > -700 else 1.0
        return _Tree(z_new, r_new, g_new, z_new, r_new, g_new, z_new, float(U_new), g_new, r_nu,
                     -sliced, False, diverging, accept_prob, 1)

    def _build_tree(self, z, r, g, log_slice, direction, depth, energy_current):
        if depth == 0:
            return self._build_basetree(z, r, g, log_slice, direction, energy_current)
        half = self._build_tree(z, r, g, log_slice, direction, depth - 1, energy_current)
        if half.turning or half.diverging:
            return half
        if direction == 1:
            z, r, g = half.z_right, half.r_right, half.g_right
        else:
            z, r, g = half.z_left, half.r_left, half.g_left
        other = self._build_tree(z, r, g, log_slice, direction, depth - 1, energy_current)
        weight = _logaddexp(half.weight, other.weight)
        sum_accept = half.sum_accept + other.sum_accept
        num_prop = half.num_prop + other.num_prop
        r_sum = half.r_sum + other.r_sum
        other_prob = math.exp(other.weight - weight)
        z_prop, U_prop, g_prop = half.z_prop, half.U_prop, half.g_prop
        if self._rand() < other_prob:
            z_prop, U_prop, g_prop = other.z_prop, other.U_prop, other.g_prop
        if direction == 1:
            zl, rl, gl = half.z_left, half.r_left, half.g_left
            zr, rr, gr = other.z_right, other.r_right, other.g_right
        else:
            zl, rl, gl = other.z_left, other.r_left, other.g_left
            zr, rr, gr = half.z_right, half.r_right, half.g_right
        turning = other.turning or self.is_turning(self._unscale(rl), self._unscale(rr), r_sum)
        return _Tree(zl, rl, gl, zr, rr, gr, z_prop, U_prop, g_prop, r_sum, weight, turning,
                     other.diverging, sum_accept, num_prop)

    def sample(self, z, U, g):
        """nuts.py:367-522.  Returns (z, U, g, accept_prob, depth, diverging)."""
        r, r_u = self._sample_r()
        energy_current = 0.5 * float(r_u @ r_u) + float(U)
        log_slice = -energy_current
        zl = zr = z
        rl = rr = r
        gl = gr = g
        r_sum = r_u
        sum_accept, num_prop = 0.0, 0
        tree_weight = 0.0
        depth = 0
        diverged = False
        while depth < self.max_depth:
            direction = 1 if self._rand() < 0.5 else -1
            if direction == 1:
                new = self._build_tree(zr, rr, gr, log_slice, 1, depth, energy_current)
                zr, rr, gr = new.z_right, new.r_right, new.g_right
            else:
                new = self._build_tree(zl, rl, gl, log_slice, -1, depth, energy_current)
                zl, rl, gl = new.z_left, new.r_left, new.g_left
            sum_accept += new.sum_accept
            num_prop += new.num_prop
            if new.diverging:
                diverged = True
                break
            if new.turning:
                break
            depth += 1
            new_prob = math.exp(new.weight - tree_weight)
            if self._rand() < new_prob:
                z, U, g = new.z_prop, new.U_prop, new.g_prop
            r_sum = r_sum + new.r_sum
            if self.is_turning(self._unscale(rl), self._unscale(rr), r_sum):
                break
            tree_weight = _logaddexp(tree_weight, new.weight)
        return z, U, g, sum_accept / num_prop, depth, diverged

    # -- full run with warm-up (hmc.py:296-353, adaptation.py:166-202) ----------------------------------
    def run(self, z0, warmup_steps, num_samples):
        z = z0.clone()
        g, U = potential_grad(self.U, z)
        U = float(U)
        sched = adaptation_schedule(warmup_steps) if (self.adapt_step_size or self.adapt_mass_matrix) else []
        da = DualAveraging()
        wf = Welford()
        window = 0
        if self.adapt_step_size:
            self.step_size = self.find_reasonable_step_size(z)
            da.prox_center = math.log(10 * self.step_size)
            da.reset()
        samples, accs = [], []
        for t in range(1, warmup_steps + num_samples + 1):
            z, U, g, acc, depth, div = self.sample(z, U, g)
            if t > warmup_steps:
                samples.append(z.clone())
                accs.append(acc)
                self.divergences += int(div)
                continue
            # adapter.step(t, ...): adaptation.py:166-202
            if t >= warmup_steps or not sched:
                continue
            w = sched[window]
            nwin = len(sched)
            mass_phase = self.adapt_mass_matrix and (0 < window < nwin - 1)
            if self.adapt_step_size:
                da.step(self.target - acc)
                self.step_size = math.exp(da.get_state()[0])
            if mass_phase:
                wf.update(z)
            if t == w.end:
                if window == nwin - 1:
                    window += 1
                    if self.adapt_step_size:
                        self.step_size = math.exp(da.get_state()[1])
                    continue
                if window == 0:
                    window += 1
                    continue
                if mass_phase:
                    self.inv_mass = wf.get_covariance(regularize=True).to(self.dtype)
                    wf.reset()
                    if self.adapt_step_size:
                        self.step_size = self.find_reasonable_step_size(z)
                        da.prox_center = math.log(10 * self.step_size)
                        da.reset()
                window += 1
        return torch.stack(samples), accs
