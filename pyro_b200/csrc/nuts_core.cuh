// nuts_core.cuh -- the No-U-Turn sampler transition as an ITERATIVE tree builder, plus the
// native potentials ("compiled model classes") and a counter-based RNG.
//
// Reference semantics: pyro/infer/mcmc/nuts.py:197-522 (recursive _build_tree, multinomial
// sampling, generalised U-turn criterion of Betancourt 2017 A.4.2 at :184-195, divergence
// threshold 1000 at :182).  The reference recursion cannot run on a GPU thread; this restates the
// same tree as a loop over leaves with O(depth) checkpoints:
//   * leaf n of a depth-d subtree is produced by one leapfrog from leaf n-1;
//   * every aligned block of 2^k leaves that ENDS at an odd leaf n is U-turn checked (smallest
//     block first, exactly the order in which the recursion returns); the block's first-leaf
//     momentum and the running momentum sum before it are found in checkpoint slot
//     popcount(n>>1) - j, j = 0..trailing_ones(n)-1;
//   * proposals are drawn by progressive multinomial sampling, which has the same law as the
//     recursion's pairwise Bernoulli merges (nuts.py:303-320).
// Everything is __host__ __device__: tests run the same code on the CPU (hostcheck.cu) and
// compare with the oracle's restatement of the reference recursion.
#pragma once
#include "b2_math.cuh"

namespace b2 {

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: stream = (seed, chain), counter++ -------
struct Philox {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t out[4];
  int have;  // unread words in out

  B2_HD void init(uint64_t seed, uint64_t stream, uint64_t counter) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    ctr[0] = (uint32_t)counter;
    ctr[1] = (uint32_t)(counter >> 32);
    ctr[2] = (uint32_t)stream;
    ctr[3] = (uint32_t)(stream >> 32);
    have = 0;
  }
  B2_HD uint64_t counter() const { return ((uint64_t)ctr[1] << 32) | ctr[0]; }
  static B2_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  B2_HD void refill() {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      const uint32_t n0 = hi1 ^ c1 ^ k0;
      const uint32_t n1 = lo1;
      const uint32_t n2 = hi0 ^ c3 ^ k1;
      const uint32_t n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    have = 4;
    if (++ctr[0] == 0) ++ctr[1];
  }
  B2_HD uint32_t next_u32() {
    if (have == 0) refill();
    return out[4 - (have--)];
  }
  // uniform in [0, 1) with 24 (float) / 53 (double) random bits
  template <typename T>
  B2_HD T uniform() {
    if (sizeof(T) == 4) {
      return (T)((next_u32() >> 8) * (1.0f / 16777216.0f));
    } else {
      const uint64_t a = next_u32() >> 5, b = next_u32() >> 6;
      return (T)((a * 67108864.0 + b) * (1.0 / 9007199254740992.0));
    }
  }
  // standard normal (Box-Muller, one value per call; simple and branch-free)
  template <typename T>
  B2_HD T normal() {
    if (sizeof(T) == 4) {
      // fp32 draws: single-precision Box-Muller (the fp64 transcendental path costs ~10x on this part)
      const float u1 = ((float)(next_u32() >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float u2 = ((float)(next_u32() >> 8) + 0.5f) * (1.0f / 16777216.0f);
      return (T)(sqrtf(-2.0f * logf(u1)) * cosf(6.2831853071795865f * u2));
    }
    const double u1 = ((double)(next_u32() >> 8) + 0.5) * (1.0 / 16777216.0);
    const double u2 = ((double)(next_u32() >> 8) + 0.5) * (1.0 / 16777216.0);
    return (T)(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
  }
};

// ---- native potentials ---------------------------------------------------------------------------
// U(z) = -[ sum_sites log_prob(T^-1(z)) + log|det dT^-1/dz| ]  as pyro/infer/mcmc/util.py:275-286.

// eight_schools family (examples/eight_schools/mcmc.py:27-34):
//   z = [mu, t = log(tau), eta_0..eta_{J-1}]
//   mu ~ Normal(0, s_mu); tau ~ HalfCauchy(s_tau); eta ~ Normal(0,1); y ~ Normal(mu + tau*eta, sigma)
template <typename T>
struct HierNormalModel {
  const T* y;
  const T* sigma;
  int64_t J;
  T s_mu, s_tau;
  B2_HD int64_t dim() const { return J + 2; }
  B2_HD T value_and_grad(const T* z, T* g) const {
    const T mu = z[0], t = z[1];
    const T tau = b2_exp(t);
    const T u = tau / s_tau;
    const T u2 = u * u;
    T U = (T)0.5 * mu * mu / (s_mu * s_mu) + b2_log(s_mu) + Consts<T>::kLogSqrt2Pi;
    U += Consts<T>::kLogPi + b2_log(s_tau) - Consts<T>::kLog2 + b2_log1p(u2) - t;
    T s_res = 0, s_res_eta = 0;
    for (int64_t j = 0; j < J; ++j) {
      const T eta = z[2 + j];
      const T sg = sigma[j];
      const T d = y[j] - mu - tau * eta;
      const T res = d / (sg * sg);
      U += (T)0.5 * eta * eta + Consts<T>::kLogSqrt2Pi;
      U += (T)0.5 * d * res + b2_log(sg) + Consts<T>::kLogSqrt2Pi;
      g[2 + j] = eta - tau * res;
      s_res += res;
      s_res_eta += res * eta;
    }
    g[0] = mu / (s_mu * s_mu) - s_res;
    g[1] = (T)2 * u2 / ((T)1 + u2) - (T)1 - tau * s_res_eta;
    return U;
  }
};

// Bayesian logistic regression (tests/infer/mcmc/test_hmc.py:189-198 family):
//   z = beta[D];  beta ~ Normal(0, s) i.i.d.;  y_n ~ Bernoulli(logits = <X[n,:], beta>)
template <typename T>
struct LogisticModel {
  const T* X;  // [N, D]
  const T* y;  // [N]
  int64_t N;
  int D;
  T s;
  B2_HD int64_t dim() const { return D; }
  B2_HD T value_and_grad(const T* z, T* g) const {
    T U = 0;
    for (int d = 0; d < D; ++d) {
      U += (T)0.5 * z[d] * z[d] / (s * s) + b2_log(s) + Consts<T>::kLogSqrt2Pi;
      g[d] = z[d] / (s * s);
    }
    for (int64_t n = 0; n < N; ++n) {
      T l = 0;
      for (int d = 0; d < D; ++d) l += X[n * D + d] * z[d];
      T sp, sg;
      softplus_sigmoid(l, sp, sg);
      U -= y[n] * l - sp;
      const T r = sg - y[n];
      for (int d = 0; d < D; ++d) g[d] += r * X[n * D + d];
    }
    return U;
  }
};

// ---- tree bookkeeping helpers --------------------------------------------------------------------
B2_HD int popcount32(uint32_t v) {
  int c = 0;
  while (v) { v &= v - 1; ++c; }
  return c;
}
B2_HD int trailing_ones32(uint32_t v) {
  int c = 0;
  while (v & 1u) { v >>= 1; ++c; }
  return c;
}
template <typename T>
B2_HD T logaddexp(T a, T b) {
  // pyro/infer/mcmc/nuts.py:15-17  max + log(exp(a-max) + exp(b-max)); -inf safe
  const T m = b2_max(a, b);
  if (m == -b2_inf<T>()) return m;
  return m + b2_log(b2_exp(a - m) + b2_exp(b - m));
}

constexpr int kNutsMaxDepth = 12;

struct NutsStats {
  float accept_prob;
  int depth;
  int diverging;
  int num_steps;
  int accepted;
};

// One NUTS transition for one chain, everything in the calling thread.
//   z, g: position and grad U(z) (in/out), U: potential (in/out)
//   eps: step size, sminv[d] = sqrt(inverse mass diag)   (whitened momentum r_u; r = r_u / sminv)
template <typename T, typename Model, int MAXD>
B2_HD void nuts_transition(const Model& model, int D, T* z, T* g, T& U, T eps, const T* sminv,
                           int max_depth, T max_delta_energy, Philox& rng, NutsStats& stats) {
  T zc[MAXD], rc[MAXD], gc[MAXD];          // growing edge (current leaf)
  T zl[MAXD], rl[MAXD], gl[MAXD];          // left edge of the whole tree
  T zr[MAXD], rr[MAXD], gr[MAXD];          // right edge
  T zp[MAXD], gp[MAXD];                    // proposal of the whole tree
  T zs[MAXD], gs[MAXD];                    // proposal of the subtree under construction
  T rsum[MAXD], rsub[MAXD];                // momentum sums: whole tree, subtree
  T rck[kNutsMaxDepth][MAXD], sck[kNutsMaxDepth][MAXD];  // checkpoints: leaf momentum, running sum
  T Up = U, Us = U;

  // momentum refresh: r_u ~ N(0, I)   (hmc.py:231-248)
  T ke = 0;
  for (int d = 0; d < D; ++d) {
    const T r = rng.template normal<T>();
    rl[d] = rr[d] = rsum[d] = r;
    zl[d] = zr[d] = zp[d] = z[d];
    gl[d] = gr[d] = gp[d] = g[d];
    ke += r * r;
  }
  const T energy0 = U + (T)0.5 * ke;
  T logw_tree = 0;  // multinomial: tree_weight = -sliced_energy of the initial state = 0
  T sum_accept = 0;
  int num_prop = 0;
  int depth = 0;
  bool accepted = false, diverged = false;

  while (depth < max_depth) {
    // direction ~ Bernoulli(0.5)   (nuts.py:429-433)
    const int dir = (rng.template uniform<T>() < (T)0.5) ? 1 : -1;
    for (int d = 0; d < D; ++d) {
      zc[d] = dir > 0 ? zr[d] : zl[d];
      rc[d] = dir > 0 ? rr[d] : rl[d];
      gc[d] = dir > 0 ? gr[d] : gl[d];
      rsub[d] = 0;
    }
    const T he = (T)0.5 * eps * (T)dir;
    const uint32_t nleaves = 1u << depth;
    T logw_sub = -b2_inf<T>();
    bool turning = false, diverging = false;
    for (uint32_t leaf = 0; leaf < nleaves; ++leaf) {
      // ---- one leapfrog (pyro/ops/integrator.py:45-65) in whitened momentum ----------------
      for (int d = 0; d < D; ++d) {
        rc[d] -= he * sminv[d] * gc[d];
        zc[d] += (T)2 * he * sminv[d] * rc[d];
      }
      T Uc = model.value_and_grad(zc, gc);
      T kec = 0;
      for (int d = 0; d < D; ++d) {
        rc[d] -= he * sminv[d] * gc[d];
        kec += rc[d] * rc[d];
        rsub[d] += rc[d];
      }
      // ---- base tree (nuts.py:197-248) -------------------------------------------------------
      T energy = Uc + (T)0.5 * kec;
      if (energy != energy) energy = b2_inf<T>();
      const T delta = energy - energy0;
      diverging = delta > max_delta_energy;
      sum_accept += b2_min((T)1, b2_exp(-delta));
      ++num_prop;
      const T w_leaf = -delta;
      // progressive multinomial sampling inside the subtree
      bool take;
      if (leaf == 0) {
        logw_sub = w_leaf;
        take = true;
      } else {
        const T nw = logaddexp(logw_sub, w_leaf);
        const T p = b2_exp(w_leaf - nw);
        take = rng.template uniform<T>() < p;
        logw_sub = nw;
      }
      if (take) {
        for (int d = 0; d < D; ++d) { zs[d] = zc[d]; gs[d] = gc[d]; }
        Us = Uc;
      }
      if (diverging) break;
      // ---- U-turn checks of every block that ends at this leaf -------------------------------
      const int idx_max = popcount32(leaf >> 1);
      if ((leaf & 1u) == 0) {
        for (int d = 0; d < D; ++d) { rck[idx_max][d] = rc[d]; sck[idx_max][d] = rsub[d]; }
      } else {
        const int nblk = trailing_ones32(leaf);
        for (int k = idx_max; k > idx_max - nblk && !turning; --k) {
          T a_first = 0, a_last = 0;
          for (int d = 0; d < D; ++d) {
            // momentum sum over the block = running sum - sum before its first leaf
            const T blk = rsub[d] - sck[k][d] + rck[k][d];
            const T rho = blk - (T)0.5 * (rck[k][d] + rc[d]);
            a_first += rck[k][d] * rho;
            a_last += rc[d] * rho;
          }
          turning = (a_first <= (T)0) || (a_last <= (T)0);
        }
        if (turning) break;
      }
    }
    // ---- merge the subtree into the tree (nuts.py:434-503) -----------------------------------
    for (int d = 0; d < D; ++d) {
      if (dir > 0) { zr[d] = zc[d]; rr[d] = rc[d]; gr[d] = gc[d]; }
      else { zl[d] = zc[d]; rl[d] = rc[d]; gl[d] = gc[d]; }
    }
    if (diverging) { diverged = true; break; }
    if (turning) break;
    ++depth;
    const T new_prob = b2_exp(logw_sub - logw_tree);
    if (rng.template uniform<T>() < new_prob) {
      accepted = true;
      for (int d = 0; d < D; ++d) { zp[d] = zs[d]; gp[d] = gs[d]; }
      Up = Us;
    }
    T a_l = 0, a_r = 0;
    for (int d = 0; d < D; ++d) {
      rsum[d] += rsub[d];
      const T rho = rsum[d] - (T)0.5 * (rl[d] + rr[d]);
      a_l += rl[d] * rho;
      a_r += rr[d] * rho;
    }
    if (a_l <= (T)0 || a_r <= (T)0) break;
    logw_tree = logaddexp(logw_tree, logw_sub);
  }
  for (int d = 0; d < D; ++d) { z[d] = zp[d]; g[d] = gp[d]; }
  U = Up;
  stats.accept_prob = num_prop > 0 ? (float)(sum_accept / (T)num_prop) : 0.f;
  stats.depth = depth;
  stats.diverging = diverged ? 1 : 0;
  stats.num_steps = num_prop;
  stats.accepted = accepted ? 1 : 0;
}

}  // namespace b2
