// site_score_fam_b.cu -- instantiations: Cauchy, HalfCauchy, Exponential, LogNormal, HalfNormal
#include "site_score.cuh"
namespace b2 {
int dispatch_site_b(int family, int dtype, bool grad, const SiteArgs& a, int kind, cudaStream_t s) {
  switch (family) {
    B2_DISPATCH_CASE(kCauchy)
    B2_DISPATCH_CASE(kHalfCauchy)
    B2_DISPATCH_CASE(kExponential)
    B2_DISPATCH_CASE(kLogNormal)
    B2_DISPATCH_CASE(kHalfNormal)
  }
  return B2_ERR_BAD_FAMILY;
}
}  // namespace b2
