// site_score_fam_a.cu -- instantiations: Normal, Bernoulli(logits), Gamma, Beta, Poisson
#include "site_score.cuh"
namespace b2 {
int dispatch_site_a(int family, int dtype, bool grad, const SiteArgs& a, int kind, cudaStream_t s) {
  switch (family) {
    B2_DISPATCH_CASE(kNormal)
    B2_DISPATCH_CASE(kBernoulliLogits)
    B2_DISPATCH_CASE(kGamma)
    B2_DISPATCH_CASE(kBeta)
    B2_DISPATCH_CASE(kPoisson)
  }
  return B2_ERR_BAD_FAMILY;
}
}  // namespace b2
