// gamma_sample.cuh -- reparameterised Gamma draws: the sampler and its implicit-reparameterisation gradient.
//
// Replaces torch/distributions/gamma.py:79-87 (``rsample``: ``_standard_gamma(concentration) / rate``, clamped
// away from 0) and the backward of ``_standard_gamma`` (ATen ``_standard_gamma_grad``) -- SURVEY.md 8(f) row 1; at
// BASELINE config 5 that is 29.6 M draws per step.
//
// Sampler: Marsaglia & Tsang (2000), "A simple method for generating gamma variables": for a >= 1,
// d = a - 1/3, c = 1/sqrt(9d), v = (1 + c n)^3 with n ~ N(0,1) accepted when log u < n^2/2 + d - d v + d log v
// (squeeze u < 1 - 0.0331 n^4 first); for a < 1 the draw for a + 1 is multiplied by u^(1/a).
//
// Gradient: x(a) at a fixed quantile satisfies dx/da = -(dP(a,x)/da) / f(x; a), P the regularised lower
// incomplete gamma function, f the density.  Both classical evaluations of P are differentiated term by term
// (the common factor x^a e^-x / Gamma(.) cancels against f, so nothing overflows):
//   x < a + 1   series   P = x^a e^-x / Gamma(a+1) * S,  S = sum_n T_n,  T_n = prod_{k<=n} x/(a+k)
//               dx/da = (x/a) [ sum_n T_n H_n - S (ln x - psi(a+1)) ],  H_n = sum_{k<=n} 1/(a+k)
//   x >= a + 1  Q = 1 - P = x^a e^-x / Gamma(a) * h,  h the continued fraction evaluated with the modified Lentz
//               recurrence, carried together with dh/da:   dx/da = x [ (ln x - psi(a)) h + dh/da ]
// Against a central finite difference of scipy's inverse incomplete gamma function the fp64 evaluation agrees to
// 1e-7 relative over a in [0.05, 1000] (ATen's piecewise approximation: 1.4e-3); the fp32 evaluation agrees with
// the fp64 one to 2e-4 (tests/test_hostcheck_math.py).
#pragma once
#include "b2_math.cuh"
#include "nuts_core.cuh"

namespace b2 {

template <typename T>
B2_HD T standard_gamma_grad(T a, T x) {
  const T one = (T)1;
  const T tol = sizeof(T) == 4 ? (T)1e-7 : (T)1e-16;
  if (!(x > (T)0) || !(a > (T)0)) return (T)0;
  const T lx = b2_log(x);
  if (x < a + one) {
    T Tn = one, S = one, H = (T)0, SH = (T)0;
    for (int n = 1; n < 4000; ++n) {
      const T inv = fast_rcp(a + (T)n);      // fp32 on the device: MUFU.RCP (1 ulp); an IEEE division is ~10 instructions
      Tn *= x * inv;
      H += inv;
      S += Tn;
      SH += Tn * H;
      if (Tn < tol * S) break;
    }
    return (x / a) * (SH - S * (lx - digamma<T>(a + one)));
  }
  T b = x + one - a;
  T c = (T)1e18, cp = (T)0;
  T d = one / b, dp = d * d;            // b' = -1
  T h = d, hp = dp;
  for (int i = 1; i < 4000; ++i) {
    const T an = -(T)i * ((T)i - a), anp = (T)i;
    b += (T)2;
    const T draw = an * d + b, drawp = anp * d + an * dp - one;
    const T ic = fast_rcp(c);
    const T cn = b + an * ic;
    cp = -one + (anp * c - an * cp) * (ic * ic);
    c = cn;
    d = fast_rcp(draw);
    dp = -drawp * d * d;
    const T de = d * c, dep = dp * c + d * cp;
    hp = hp * de + h * dep;
    h *= de;
    if (b2_abs(de - one) < tol && b2_abs(dep) < tol * (one + b2_abs(hp / h))) break;
  }
  return x * ((lx - digamma<T>(a)) * h + hp);
}

// One standard Gamma(a, 1) draw.
template <typename T>
B2_HD T standard_gamma_sample(T a, Philox& rng) {
  T boost = (T)1;
  if (a < (T)1) {
    // u^(1/a) with u in (0, 1]
    const T u = (T)1 - rng.template uniform<T>();
    boost = b2_exp(b2_log(u) / a);
    a += (T)1;
  }
  const T d = a - (T)(1.0 / 3.0);
  const T c = (T)1 / b2_sqrt((T)9 * d);
  for (int it = 0; it < 64; ++it) {
    const T n = rng.template normal<T>();
    T v = (T)1 + c * n;
    if (!(v > (T)0)) continue;
    v = v * v * v;
    const T u = (T)1 - rng.template uniform<T>();     // (0, 1]
    const T n2 = n * n;
    if (u < (T)1 - (T)0.0331 * n2 * n2) return boost * d * v;
    if (b2_log(u) < (T)0.5 * n2 + d * ((T)1 - v + b2_log(v))) return boost * d * v;
  }
  return boost * d;   // not reached in practice (acceptance > 95 % per trial)
}

}  // namespace b2
