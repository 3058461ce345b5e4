// site_score_fam_c.cu -- instantiations: Bernoulli(probs), Uniform, KL(Normal||Normal), KL(Gamma||Gamma), fused Normal rsample fwd/bwd
#include "site_score.cuh"
namespace b2 {
int dispatch_site_c(int family, int dtype, bool grad, const SiteArgs& a, int kind, cudaStream_t s) {
  switch (family) {
    B2_DISPATCH_CASE(kBernoulliProbs)
    B2_DISPATCH_CASE(kUniform)
    B2_DISPATCH_CASE(kKLNormalNormal)
    B2_DISPATCH_CASE(kKLGammaGamma)
    B2_DISPATCH_CASE(kNormalRsample)
    B2_DISPATCH_CASE(kNormalRsampleBwd)
  }
  return B2_ERR_BAD_FAMILY;
}
}  // namespace b2
