// rsample.cu -- reparameterised Normal draw with the noise generated IN the kernel (Philox4x32-10),
// fused with the site's own log density.
//
// Replaces, for a guide site `pyro.sample(name, Normal(loc, scale))`:
//   torch.randn (+ the two RNG-offset fill kernels a CUDA-graph replay adds)   torch/distributions/normal.py:82-85
//   loc + eps * scale                                                           (same lines)
//   fn.log_prob(z), scale_and_mask, .sum()                                      pyro/poutine/trace_struct.py:264-278
// One launch writes z, eps (kept for the backward pass) and the 0-d sum of log q(z).
//
// RNG: counter-based; (seed, launch counter) live in a 2 x int64 device array that the kernel itself
// advances, so a captured CUDA graph draws fresh noise on every replay without any host-side state.
// Element i of launch c uses Philox(key = seed, stream = i, counter = c): streams never overlap.  The
// stream is NOT torch's generator stream (documented in DESIGN.md); tests that need given noise inject it.
#include "b2_common.cuh"
#include "b2_math.cuh"
#include "nuts_core.cuh"

namespace b2 {

struct RsampleArgs {
  int ndim;
  int64_t shape[kMaxD];
  int64_t st_loc[kMaxD], st_scale[kMaxD];
  int64_t n;
  const void* loc;
  const void* scale;
  void* z;
  void* eps;
  void* out_sum;
  unsigned long long* state;   // [seed, counter]
};

template <typename T>
__global__ void __launch_bounds__(1024) normal_rsample_philox_kernel(const RsampleArgs a) {
  const unsigned long long seed = a.state[0], ctr = a.state[1];
  double acc = 0.0;
  const unsigned n = (unsigned)a.n;
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned rem = i;
    int64_t ol = 0, os = 0;
    for (int d = a.ndim - 1; d >= 0; --d) {
      const unsigned sd = (unsigned)a.shape[d];
      const unsigned q = rem / sd;
      const int64_t idx = (int64_t)(rem - q * sd);
      rem = q;
      ol += idx * a.st_loc[d];
      os += idx * a.st_scale[d];
    }
    Philox rng;
    rng.init(seed, (uint64_t)i, ctr);
    const T e = rng.template normal<T>();
    T p[2] = {reinterpret_cast<const T*>(a.loc)[ol], reinterpret_cast<const T*>(a.scale)[os]};
    const T zv = p[0] + e * p[1];
    ElemOut<T> o;
    Eval<kNormal, T, false>::run(zv, p, o);
    reinterpret_cast<T*>(a.z)[i] = zv;
    reinterpret_cast<T*>(a.eps)[i] = e;
    acc += (double)o.lp;
  }
  __shared__ double smem[32];
  double v[1] = {acc};
  block_sum<1>(v, smem);
  if (threadIdx.x == 0) {
    *reinterpret_cast<T*>(a.out_sum) = (T)v[0];
    a.state[1] = ctr + 1ull;   // every thread read the counter before the block_sum barriers
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_normal_rsample(const b2_tensor* loc, const b2_tensor* scale, int ndim, const int64_t* shape,
                                 void* z, void* eps, void* out_sum, void* rng_state, void* stream) {
  if (!loc || !scale || !z || !eps || !out_sum || !rng_state || !shape) return B2_ERR_NULL;
  if (ndim < 0 || ndim > kMaxD) return B2_ERR_BAD_SHAPE;
  if (loc->dtype != scale->dtype || (loc->dtype != B2_F32 && loc->dtype != B2_F64)) return B2_ERR_BAD_DTYPE;
  RsampleArgs a;
  a.ndim = ndim;
  a.n = 1;
  for (int d = 0; d < kMaxD; ++d) {
    a.shape[d] = 1;
    a.st_loc[d] = a.st_scale[d] = 0;
  }
  for (int d = 0; d < ndim; ++d) {
    a.shape[d] = shape[d];
    a.n *= shape[d];
    a.st_loc[d] = loc->stride[d];
    a.st_scale[d] = scale->stride[d];
  }
  if (a.n <= 0 || a.n > B2_RSAMPLE_MAX_N) return B2_ERR_TOO_LARGE;
  a.loc = loc->ptr;
  a.scale = scale->ptr;
  a.z = z;
  a.eps = eps;
  a.out_sum = out_sum;
  a.state = reinterpret_cast<unsigned long long*>(rng_state);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int threads = a.n >= 1024 ? 1024 : (int)((a.n + 31) / 32 * 32);
  if (loc->dtype == B2_F32)
    normal_rsample_philox_kernel<float><<<1, threads, 0, s>>>(a);
  else
    normal_rsample_philox_kernel<double><<<1, threads, 0, s>>>(a);
  count_launch();
  return check_launch();
}
