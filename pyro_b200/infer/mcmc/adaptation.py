"""Warm-up adaptation for MANY chains at once, state kept on the device.

Restates, vectorised over the chain axis, pyro/infer/mcmc/adaptation.py:23-215 (Stan window
schedule 75 / 25*2^k / 50, per-window mass-matrix update, dual-averaging step size with
prox-centre log(10 eps0)), pyro/ops/dual_averaging.py:43-79 (t0=10, kappa=0.75, gamma=0.05) and
pyro/ops/welford.py:7-51 (Welford variance + Stan shrinkage n/(n+5)*cov + 1e-3*5/(n+5)).

Every chain has its own step size and (diagonal) inverse mass, exactly like the reference's
independent worker processes (pyro/infer/mcmc/api.py:88-142); the arithmetic is elementwise over
``[C]`` / ``[C, D]`` tensors so no host synchronisation is needed per transition
(the reference calls ``accept_prob.item()``, adaptation.py:183).
"""
import math
from collections import namedtuple

import torch

adapt_window = namedtuple("adapt_window", ["start", "end"])


def build_adaptation_schedule(warmup_steps, start_buffer=75, end_buffer=50, initial_window=25):
    """pyro/infer/mcmc/adaptation.py:65-103 (exact lists pinned by
    tests/infer/mcmc/test_adaptation.py:27-35)."""
    schedule = []
    if warmup_steps < 20:
        schedule.append(adapt_window(0, warmup_steps - 1))
        return schedule
    start_buffer_size, end_buffer_size, init_window_size = start_buffer, end_buffer, initial_window
    if start_buffer + end_buffer + initial_window > warmup_steps:
        start_buffer_size = int(0.15 * warmup_steps)
        end_buffer_size = int(0.1 * warmup_steps)
        init_window_size = warmup_steps - start_buffer_size - end_buffer_size
    schedule.append(adapt_window(0, start_buffer_size - 1))
    end_window_start = warmup_steps - end_buffer_size
    next_window_size = init_window_size
    next_window_start = start_buffer_size
    while next_window_start < end_window_start:
        cur_start, cur_size = next_window_start, next_window_size
        if 3 * cur_size <= end_window_start - cur_start:
            next_window_size = 2 * cur_size
        else:
            cur_size = end_window_start - cur_start
        next_window_start = cur_start + cur_size
        schedule.append(adapt_window(cur_start, next_window_start - 1))
    schedule.append(adapt_window(end_window_start, warmup_steps - 1))
    return schedule


class DualAveraging:
    """Nesterov dual averaging for ``C`` independent sequences (float64 tensors ``[C]``)."""

    def __init__(self, num_chains, device, prox_center=0.0, t0=10, kappa=0.75, gamma=0.05):
        self.C = num_chains
        self.device = device
        self.t0, self.kappa, self.gamma = t0, kappa, gamma
        self.prox_center = torch.full((num_chains,), float(prox_center), dtype=torch.float64, device=device)
        self.reset()

    def reset(self):
        self._x_avg = torch.zeros(self.C, dtype=torch.float64, device=self.device)
        self._g_avg = torch.zeros(self.C, dtype=torch.float64, device=self.device)
        self._x_t = torch.zeros(self.C, dtype=torch.float64, device=self.device)
        self._t = 0

    def step(self, g):
        self._t += 1
        t = self._t
        self._g_avg = (1 - 1 / (t + self.t0)) * self._g_avg + g / (t + self.t0)
        self._x_t = self.prox_center - (t ** 0.5) / self.gamma * self._g_avg
        weight_t = t ** (-self.kappa)
        self._x_avg = (1 - weight_t) * self._x_avg + weight_t * self._x_t

    def get_state(self):
        return self._x_t, self._x_avg


class WelfordDiag:
    """Per-chain running variance of ``[C, D]`` samples."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._mean = 0.0
        self._m2 = 0.0
        self.n_samples = 0

    def update(self, sample):
        self.n_samples += 1
        delta_pre = sample - self._mean
        self._mean = self._mean + delta_pre / self.n_samples
        delta_post = sample - self._mean
        self._m2 = self._m2 + delta_pre * delta_post

    def get_covariance(self, regularize=True):
        if self.n_samples < 2:
            raise RuntimeError("Insufficient samples to estimate covariance")
        cov = self._m2 / (self.n_samples - 1)
        if regularize:
            n = self.n_samples
            cov = (n / (n + 5.0)) * cov + 1e-3 * (5.0 / (n + 5.0))
        return cov


class WelfordFull:
    """Per-chain running covariance of ``[C, D]`` samples (pyro/ops/welford.py:7-51, ``diagonal=False``)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._mean = 0.0
        self._m2 = 0.0
        self.n_samples = 0

    def update(self, sample):
        self.n_samples += 1
        delta_pre = sample - self._mean
        self._mean = self._mean + delta_pre / self.n_samples
        delta_post = sample - self._mean
        self._m2 = self._m2 + delta_pre.unsqueeze(-1) * delta_post.unsqueeze(-2)

    def get_covariance(self, regularize=True):
        if self.n_samples < 2:
            raise RuntimeError("Insufficient samples to estimate covariance")
        cov = self._m2 / (self.n_samples - 1)
        if regularize:
            n = self.n_samples
            cov = (n / (n + 5.0)) * cov
            shrink = 1e-3 * (5.0 / (n + 5.0))
            cov = cov + shrink * torch.eye(cov.shape[-1], dtype=cov.dtype, device=cov.device)
        return cov


class WarmupAdapter:
    """Step-size and mass adaptation for ``C`` chains (diagonal mass; with ``full_mass`` the dense
    covariance of each window is handed to ``mass_update_fn`` as its Cholesky factor)."""

    def __init__(self, num_chains, dim, dtype, device, step_size=1.0, adapt_step_size=True,
                 target_accept_prob=0.8, adapt_mass_matrix=True, full_mass=False, mass_update_fn=None):
        self.full_mass = full_mass
        self._mass_update_fn = mass_update_fn
        self.C, self.D = num_chains, dim
        self.dtype, self.device = dtype, device
        self.adapt_step_size = adapt_step_size
        self.adapt_mass_matrix = adapt_mass_matrix
        self.target_accept_prob = target_accept_prob
        self._init_step_size = 1.0 if step_size is None else float(step_size)
        self.step_size = torch.full((num_chains,), self._init_step_size, dtype=dtype, device=device)
        self.inverse_mass = torch.ones(num_chains, dim, dtype=dtype, device=device)
        self._adaptation_disabled = not (adapt_step_size or adapt_mass_matrix)
        self._da = DualAveraging(num_chains, device) if adapt_step_size else None
        self._welford = WelfordFull() if full_mass else WelfordDiag()
        self._warmup_steps = None
        self._schedule = []
        self._current_window = 0
        self._find_reasonable_step_size = None

    @property
    def adaptation_schedule(self):
        return self._schedule

    def configure(self, warmup_steps, find_reasonable_step_size_fn=None):
        self._warmup_steps = warmup_steps
        self._find_reasonable_step_size = find_reasonable_step_size_fn
        if not self._adaptation_disabled:
            self._schedule = build_adaptation_schedule(warmup_steps)
        self._current_window = 0
        if self.adapt_step_size:
            self._da.reset()

    def reset_step_size_adaptation(self, z):
        if self._find_reasonable_step_size is not None:
            self.step_size = self._find_reasonable_step_size(z, self.step_size, self.inverse_mass)
        self._da.prox_center = torch.log(10 * self.step_size.double())
        self._da.reset()

    def _update_step_size(self, accept_prob):
        H = self.target_accept_prob - accept_prob.double()
        self._da.step(H)
        log_step_size, _ = self._da.get_state()
        self.step_size = torch.exp(log_step_size).to(self.dtype)

    def _end_adaptation(self):
        if self.adapt_step_size:
            _, log_step_size_avg = self._da.get_state()
            self.step_size = torch.exp(log_step_size_avg).to(self.dtype)

    def step(self, t, z, accept_prob, z_model=None):
        """``t``: transition index (0-based); ``z`` ``[C, D]`` in kernel coordinates; ``accept_prob`` ``[C]``;
        ``z_model``: the same state in the model's coordinates when they differ (``full_mass``)."""
        if t >= self._warmup_steps or self._adaptation_disabled:
            return
        window = self._schedule[self._current_window]
        num_windows = len(self._schedule)
        mass_phase = self.adapt_mass_matrix and (0 < self._current_window < num_windows - 1)
        if self.adapt_step_size:
            self._update_step_size(accept_prob)
        if mass_phase:
            self._welford.update((z_model if z_model is not None else z).detach())
        if t == window.end:
            if self._current_window == num_windows - 1:
                self._current_window += 1
                self._end_adaptation()
                return
            if self._current_window == 0:
                self._current_window += 1
                return
            if mass_phase and self.full_mass:
                cov = self._welford.get_covariance(regularize=True).to(self.dtype)
                z = self._mass_update_fn(torch.linalg.cholesky(cov))   # kernel state in the new coordinates
                self._welford.reset()
                if self.adapt_step_size:
                    self.reset_step_size_adaptation(z)
            elif mass_phase:
                self.inverse_mass = self._welford.get_covariance(regularize=True).to(self.dtype)
                self._welford.reset()
                if self.adapt_step_size:
                    self.reset_step_size_adaptation(z)
            self._current_window += 1
