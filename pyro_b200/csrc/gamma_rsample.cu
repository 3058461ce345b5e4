// gamma_rsample.cu -- b2_gamma_rsample: reparameterised Gamma(concentration, rate) draws with the derivative of
// each draw w.r.t. its concentration from the same launch (gamma_sample.cuh).
//   z = max(x / rate, tiny),  x ~ Gamma(concentration, 1)          torch/distributions/gamma.py:79-87
//   dz_dconc = (dx/dconcentration) / rate                             (backward of _standard_gamma, ATen)
// d z / d rate = -z / rate needs no kernel.  Noise: Philox4x32-10, key = seed, stream = element index, counter =
// launch counter << 20 (a draw uses a handful of 128-bit blocks); the launch counter lives in the same device
// array as the Normal draws' and is advanced by a stream-ordered one-thread launch, so replayed CUDA graphs draw
// fresh noise.
#include "b2_common.cuh"
#include "gamma_sample.cuh"

namespace b2 {

struct GammaArgs {
  int ndim;
  int64_t shape[kMaxD];
  int64_t st_conc[kMaxD], st_rate[kMaxD];
  int64_t n;
  const void* conc;
  const void* rate;
  void* z;
  void* dz;
  unsigned long long* state;
};

template <typename T>
__global__ void __launch_bounds__(256) gamma_rsample_kernel(const GammaArgs a) {
  const unsigned long long seed = a.state[0], ctr = a.state[1];
  const T tiny = sizeof(T) == 4 ? (T)1.17549435e-38 : (T)2.2250738585072014e-308;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = i, oc = 0, orr = 0;
    for (int d = a.ndim - 1; d >= 0; --d) {
      const int64_t q = rem / a.shape[d];
      const int64_t idx = rem - q * a.shape[d];
      rem = q;
      oc += idx * a.st_conc[d];
      orr += idx * a.st_rate[d];
    }
    const T al = reinterpret_cast<const T*>(a.conc)[oc];
    const T rt = reinterpret_cast<const T*>(a.rate)[orr];
    Philox rng;
    rng.init(seed, (uint64_t)i, ctr << 20);
    const T x = standard_gamma_sample<T>(al, rng);
    const T zr = x / rt;
    reinterpret_cast<T*>(a.z)[i] = zr > tiny ? zr : tiny;
    if (a.dz) reinterpret_cast<T*>(a.dz)[i] = standard_gamma_grad<T>(al, x > tiny ? x : tiny) / rt;
  }
}

__global__ void gamma_advance_kernel(unsigned long long* state) { state[1] += 1ull; }

}  // namespace b2

using namespace b2;

extern "C" int b2_gamma_rsample(const b2_tensor* conc, const b2_tensor* rate, int ndim, const int64_t* shape,
                                void* z, void* dz_dconc, void* rng_state, void* stream) {
  if (!conc || !rate || !z || !rng_state || (ndim > 0 && !shape)) return B2_ERR_NULL;
  if (ndim < 0 || ndim > kMaxD) return B2_ERR_BAD_SHAPE;
  if (conc->dtype != rate->dtype || (conc->dtype != B2_F32 && conc->dtype != B2_F64)) return B2_ERR_BAD_DTYPE;
  GammaArgs a;
  a.ndim = ndim;
  a.n = 1;
  for (int d = 0; d < kMaxD; ++d) {
    a.shape[d] = 1;
    a.st_conc[d] = a.st_rate[d] = 0;
  }
  for (int d = 0; d < ndim; ++d) {
    if (shape[d] <= 0) return B2_ERR_BAD_SHAPE;
    a.shape[d] = shape[d];
    a.n *= shape[d];
    a.st_conc[d] = conc->stride[d];
    a.st_rate[d] = rate->stride[d];
  }
  a.conc = conc->ptr;
  a.rate = rate->ptr;
  a.z = z;
  a.dz = dz_dconc;
  a.state = reinterpret_cast<unsigned long long*>(rng_state);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int64_t blocks = (a.n + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  if (conc->dtype == B2_F32) gamma_rsample_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(a);
  else gamma_rsample_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(a);
  gamma_advance_kernel<<<1, 1, 0, s>>>(a.state);
  count_launch(2);
  return check_launch();
}
