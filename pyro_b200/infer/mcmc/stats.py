"""Chain diagnostics: split R-hat and effective sample size.

Restates pyro/ops/stats.py:14-219 (``_compute_chain_variance_stats``, ``gelman_rubin``,
``split_gelman_rubin``, FFT ``autocorrelation``/``autocovariance``, Geyer initial positive +
monotone sequence in ``effective_sample_size``) on tensors laid out ``[chains, samples, ...]``.
Cross-rank merging of per-chain moments follows pyro/ops/streaming.py:231-251 (Chan et al.).
Post-hoc and tiny next to sampling, so plain torch ops (torch.fft) on the device.
"""
import torch


def _chain_variance_stats(x):
    # x: [N, C, ...]
    n = x.size(0)
    var_within = x.var(dim=0).mean(dim=0)
    var_estimator = (n - 1) / n * var_within
    if x.size(1) > 1:
        var_estimator = var_estimator + x.mean(dim=0).var(dim=0)
    else:
        var_within = var_estimator
    return var_within, var_estimator


def gelman_rubin(x):
    """x: [chains, samples, ...] -> r_hat [...]"""
    assert x.dim() >= 2 and x.size(0) >= 2 and x.size(1) >= 2
    xt = x.transpose(0, 1)
    var_within, var_estimator = _chain_variance_stats(xt)
    return (var_estimator / var_within).sqrt()


def split_gelman_rubin(x):
    """x: [chains, samples, ...]; each chain is split in halves first."""
    assert x.dim() >= 2 and x.size(1) >= 4
    half = x.size(1) // 2
    halves = torch.cat([x[:, :half], x[:, -half:]], dim=0)
    return gelman_rubin(halves)


def _next_fast_len(n):
    # smallest 2^a 3^b 5^c >= n
    best = None
    p2 = 1
    while p2 < 2 * n:
        p3 = p2
        while p3 < 2 * n:
            p5 = p3
            while p5 < 2 * n:
                if p5 >= n and (best is None or p5 < best):
                    best = p5
                p5 *= 5
            p3 *= 3
        p2 *= 2
    return best


def autocorrelation(x, dim=0):
    n = x.size(dim)
    m2 = 2 * _next_fast_len(n)
    x = x.transpose(dim, -1)
    centered = x - x.mean(dim=-1, keepdim=True)
    f = torch.fft.rfft(centered, n=m2)
    gram = f.real.pow(2) + f.imag.pow(2)
    ac = torch.fft.irfft(gram, n=m2)[..., :n]
    ac = ac / torch.arange(n, 0, -1, dtype=x.dtype, device=x.device)
    variance = ac[..., :1]
    constant = (variance == 0).expand_as(ac)
    ac = ac / variance.clamp(min=torch.finfo(variance.dtype).tiny)
    ac = torch.where(constant, torch.ones_like(ac), ac)
    return ac.transpose(dim, -1)


def autocovariance(x, dim=0):
    return autocorrelation(x, dim) * x.var(dim, unbiased=False, keepdim=True)


def _cummin(x):
    return torch.cummin(x, dim=0)[0]


def effective_sample_size(x):
    """x: [chains, samples, ...] -> n_eff [...]"""
    assert x.dim() >= 2 and x.size(1) >= 2
    xt = x.transpose(0, 1)  # [N, C, ...]
    n, c = xt.size(0), xt.size(1)
    gamma = autocovariance(xt, dim=0)
    var_within, var_estimator = _chain_variance_stats(xt)
    rho = (var_estimator - var_within + gamma.mean(dim=1)) / var_estimator
    rho[0] = 1
    rho_k = rho if n % 2 == 0 else rho[:-1]
    rho_k = rho_k.reshape((n // 2, 2) + rho_k.shape[1:]).sum(dim=1)
    rho_init = rho_k[0]
    if rho_k.size(0) > 1:
        tau = -1 + 2 * rho_init + 2 * _cummin(rho_k[1:].clamp(min=0)).sum(dim=0)
    else:
        tau = -1 + 2 * rho_init
    return c * n / tau


def merge_moments(n_a, mean_a, m2_a, n_b, mean_b, m2_b):
    """Chan et al. pairwise merge of (count, mean, M2) (pyro/ops/streaming.py:231-251)."""
    n = n_a + n_b
    delta = mean_b - mean_a
    mean = mean_a + delta * (n_b / n)
    m2 = m2_a + m2_b + delta * delta * (n_a * n_b / n)
    return n, mean, m2
