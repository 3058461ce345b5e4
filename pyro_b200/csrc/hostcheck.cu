// hostcheck.cu -- TEST INFRASTRUCTURE ONLY.  Builds the __host__ __device__ element functors
// of b2_math.cuh (and the NUTS tree logic of nuts_core.cuh) for the CPU so tests can pin the
// arithmetic against the oracle without a GPU.  It is a separate shared object
// (libpyro_b200_hostcheck.so); nothing in the product path loads it.
#include <stdint.h>

#include "b2_math.cuh"

using namespace b2;

namespace {
template <typename T>
int eval_t(int family, int grad, int64_t n, const T* x, const T* const* p, T* lp, T* dx,
           T* const* dp) {
  for (int64_t i = 0; i < n; ++i) {
    ElemOut<T> o;
    T pl[4] = {p[0] ? p[0][i] : (T)0, p[1] ? p[1][i] : (T)0, p[2] ? p[2][i] : (T)0,
               p[3] ? p[3][i] : (T)0};
    const T xv = x ? x[i] : (T)0;
#define B2H_CASE(F)                                   \
  case F:                                             \
    if (grad) Eval<F, T, true>::run(xv, pl, o);       \
    else Eval<F, T, false>::run(xv, pl, o);           \
    break;
    switch (family) {
      B2H_CASE(kNormal)
      B2H_CASE(kBernoulliLogits)
      B2H_CASE(kGamma)
      B2H_CASE(kBeta)
      B2H_CASE(kPoisson)
      B2H_CASE(kCauchy)
      B2H_CASE(kHalfCauchy)
      B2H_CASE(kExponential)
      B2H_CASE(kLogNormal)
      B2H_CASE(kHalfNormal)
      B2H_CASE(kBernoulliProbs)
      B2H_CASE(kUniform)
      B2H_CASE(kKLNormalNormal)
      B2H_CASE(kKLGammaGamma)
      default:
        return -3;
    }
#undef B2H_CASE
    lp[i] = o.lp;
    if (grad) {
      if (dx) dx[i] = o.dx;
      for (int k = 0; k < 4; ++k)
        if (dp[k]) dp[k][i] = o.dp[k];
    }
  }
  return 0;
}
}  // namespace

extern "C" int b2h_eval_f32(int family, int grad, int64_t n, const float* x, const float* const* p,
                            float* lp, float* dx, float* const* dp) {
  return eval_t<float>(family, grad, n, x, p, lp, dx, dp);
}
extern "C" int b2h_eval_f64(int family, int grad, int64_t n, const double* x,
                            const double* const* p, double* lp, double* dx, double* const* dp) {
  return eval_t<double>(family, grad, n, x, p, lp, dx, dp);
}
extern "C" double b2h_digamma(double x) { return digamma<double>(x); }
extern "C" float b2h_digammaf(float x) { return digamma<float>(x); }
extern "C" double b2h_trigamma(double x) { return trigamma<double>(x); }
