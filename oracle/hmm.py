"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): first-principles marginal likelihood of the
linear-Gaussian state-space model that ``GaussianHMM`` defines.

Model, as documented at pyro/distributions/hmm.py:434-497 (row-vector convention of the reference):
    z_{-1} ~ initial_dist
    z_t = z_{t-1} @ transition_matrix_t + w_t,      w_t ~ transition_dist_t
    x_t = z_t @ observation_matrix_t + v_t,         v_t ~ observation_dist_t,      t = 0..T-1

The reference eliminates the hidden states with a parallel scan of Gaussian tensordots
(pyro/distributions/hmm.py:565-582, pyro/ops/gaussian.py:510-597); the product uses a Kalman filter.
This oracle uses NEITHER: it writes down the joint Gaussian of the stacked observations
``[x_0 .. x_{T-1}]`` -- mean ``mu_t = E[z_t] H_t + E[v_t]`` and covariance blocks
``Cov(x_s, x_t) = H_s^T Cov(z_s, z_t) H_t + [s == t] R_t`` with ``Cov(z_s, z_t) = P_s F_{s+1} .. F_t`` --
and evaluates one dense multivariate-normal log density of dimension T*O.  O((T*O)^3): small cases
only.  Pinned against the reference's own outputs in tests/golden/hmm.npz (tests/test_oracle_golden.py).
"""
import math

import torch


def _at(x, t, nd):
    """Time slice of a possibly time-dependent parameter (time axis first)."""
    return x[t] if x.dim() > nd else x


def gaussian_hmm_log_prob(init_loc, init_cov, F, trans_loc, trans_cov, H, obs_loc, obs_cov, value):
    """log p(value) for ONE sequence ``value [T, O]``; parameters are float64 tensors, time-dependent
    ones carry the time axis first (``F [T, Hd, Hd]`` etc.)."""
    T, O = value.shape
    m, P = init_loc, init_cov
    means, covs, Fs, Hs, mus, Rs = [], [], [], [], [], []
    for t in range(T):
        Ft, Ht = _at(F, t, 2), _at(H, t, 2)
        m = m @ Ft + _at(trans_loc, t, 1)
        P = Ft.transpose(0, 1) @ P @ Ft + _at(trans_cov, t, 2)
        means.append(m)
        covs.append(P)
        Fs.append(Ft)
        Hs.append(Ht)
        mus.append(m @ Ht + _at(obs_loc, t, 1))
        Rs.append(_at(obs_cov, t, 2))
    mu = torch.cat(mus)
    S = torch.zeros(T * O, T * O, dtype=value.dtype)
    for s in range(T):
        C = covs[s]                          # Cov(z_s, z_t), starting at t = s
        for t in range(s, T):
            if t > s:
                C = C @ Fs[t]
            blk = Hs[s].transpose(0, 1) @ C @ Hs[t]
            if t == s:
                blk = blk + Rs[t]
            S[s * O:(s + 1) * O, t * O:(t + 1) * O] = blk
            S[t * O:(t + 1) * O, s * O:(s + 1) * O] = blk.transpose(0, 1)
    L = torch.linalg.cholesky(S)
    d = (value.reshape(-1) - mu).unsqueeze(-1)
    zs = torch.linalg.solve_triangular(L, d, upper=False)
    return -0.5 * (zs * zs).sum() - L.diagonal().log().sum() - 0.5 * T * O * math.log(2 * math.pi)


def gaussian_hmm_filter(init_loc, init_cov, F, trans_loc, trans_cov, H, obs_loc, obs_cov, value):
    """Posterior N(mean, cov) of the FINAL hidden state given the whole sequence (what
    pyro/distributions/hmm.py:604-633 ``GaussianHMM.filter`` returns), from first principles: (z_T, x_1..x_T)
    is jointly Gaussian; condition z_T on x.  Same argument conventions as ``gaussian_hmm_log_prob``."""
    T, O = value.shape
    m, P = init_loc, init_cov
    covs, Fs, Hs, mus, Rs = [], [], [], [], []
    for t in range(T):
        Ft, Ht = _at(F, t, 2), _at(H, t, 2)
        m = m @ Ft + _at(trans_loc, t, 1)
        P = Ft.transpose(0, 1) @ P @ Ft + _at(trans_cov, t, 2)
        covs.append(P)
        Fs.append(Ft)
        Hs.append(Ht)
        mus.append(m @ Ht + _at(obs_loc, t, 1))
        Rs.append(_at(obs_cov, t, 2))
    mu = torch.cat(mus)
    Hd = m.shape[0]
    S = torch.zeros(T * O, T * O, dtype=value.dtype)
    Czx = torch.zeros(Hd, T * O, dtype=value.dtype)          # Cov(z_T, x_s)
    for s in range(T):
        C = covs[s]                                          # Cov(z_s, z_t) for t = s, s+1, ...
        for t in range(s, T):
            if t > s:
                C = C @ Fs[t]
            blk = Hs[s].transpose(0, 1) @ C @ Hs[t]
            if t == s:
                blk = blk + Rs[t]
            S[s * O:(s + 1) * O, t * O:(t + 1) * O] = blk
            S[t * O:(t + 1) * O, s * O:(s + 1) * O] = blk.transpose(0, 1)
        # after the inner loop C = Cov(z_s, z_T)
        Czx[:, s * O:(s + 1) * O] = (Hs[s].transpose(0, 1) @ C).transpose(0, 1)
    d = (value.reshape(-1) - mu).unsqueeze(-1)
    sol = torch.linalg.solve(S, torch.cat([d, Czx.transpose(0, 1)], dim=1))
    mean = m + (Czx @ sol[:, :1]).squeeze(-1)
    cov = P - Czx @ sol[:, 1:]
    return mean, 0.5 * (cov + cov.transpose(0, 1))
