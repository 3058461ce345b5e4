// optim.cu -- multi-tensor fused optimiser updates.
//
// Reference: PyroOptim keeps one torch.optim object per parameter tensor and steps them in a
// Python loop (pyro/optim/optim.py:117-155), each step ~8 ATen launches.  Here ONE advance
// kernel updates every tensor's scalar state (step count, decayed lr, bias-corrected step size)
// on the device, and ONE update kernel (grid.y = tensor index) applies the element updates, so a
// captured CUDA graph can replay the step without host-side scalars.
#include "b2_common.cuh"
#include "b2_math.cuh"

namespace b2 {

constexpr int kAdamHyperStride = 8;  // beta1, beta2, eps, weight_decay, clip_norm, lrd, [step_size], [unused]
constexpr int kAgrHyperStride = 4;   // eta, delta, t, [lr]
constexpr int64_t kAdamFusedAdvanceNumel = 4096;  // largest tensor for the one-launch (advance + update) path

// beta^t for an integer step count by repeated squaring: ~2*log2(t) dependent multiplies instead of
// libm's double pow (hundreds of dependent instructions on ONE thread while the whole update
// waits); agrees with Python's beta ** step to a few ulp.
__device__ __forceinline__ double ipow(double b, int32_t t) {
  double r = 1.0;
  while (t > 0) {
    if (t & 1) r *= b;
    b *= b;
    t >>= 1;
  }
  return r;
}

// pyro/optim/clipped_adam.py:62,80,91-93: lr *= lrd; step += 1;
// step_size = lr * sqrt(1 - beta2^step) / (1 - beta1^step)      (all in double, like Python)
__global__ void adam_advance_kernel(int n, double* hyper, double* lrs, int32_t* steps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* h = hyper + (size_t)i * kAdamHyperStride;
  const double lr = lrs[i] * h[5];
  lrs[i] = lr;
  const int32_t t = steps[i] + 1;
  steps[i] = t;
  const double bc1 = 1.0 - ipow(h[0], t);
  const double bc2 = 1.0 - ipow(h[1], t);
  h[6] = lr * sqrt(bc2) / bc1;
}

// ADV: the launch has ONE CTA per tensor (small parameter sets, e.g. BASELINE config 2's four
// tensors), so that CTA also advances its tensor's scalar state -- no separate advance launch.
template <typename T, bool ADV>
__global__ void __launch_bounds__(256) clipped_adam_kernel(void* const* __restrict__ ps,
                                                           void* const* __restrict__ gs,
                                                           void* const* __restrict__ ms,
                                                           void* const* __restrict__ vs,
                                                           const int64_t* __restrict__ numel,
                                                           double* __restrict__ hyper,
                                                           double* __restrict__ lrs,
                                                           int32_t* __restrict__ steps,
                                                           int zero_grad) {
  pdl_enter();
  const int ti = blockIdx.y;
  const int64_t n = numel[ti];
  double* h = hyper + (size_t)ti * kAdamHyperStride;
  double step_size;
  if (ADV) {
    __shared__ double sh_step;
    if (threadIdx.x == 0) {
      const double lr = lrs[ti] * h[5];
      const int32_t t = steps[ti] + 1;
      const double bc1 = 1.0 - ipow(h[0], t);
      const double bc2 = 1.0 - ipow(h[1], t);
      sh_step = lr * sqrt(bc2) / bc1;
      lrs[ti] = lr;
      steps[ti] = t;
      h[6] = sh_step;
    }
    __syncthreads();
    step_size = sh_step;
  } else {
    step_size = h[6];
  }
  const T b1 = (T)h[0], b2v = (T)h[1], eps = (T)h[2], wd = (T)h[3], clip = (T)h[4];
  const T omb1 = (T)(1.0 - h[0]), omb2 = (T)(1.0 - h[1]);
  const T neg_step = (T)(-step_size);
  T* __restrict__ p = reinterpret_cast<T*>(ps[ti]);
  T* __restrict__ g = reinterpret_cast<T*>(gs[ti]);
  T* __restrict__ m = reinterpret_cast<T*>(ms[ti]);
  T* __restrict__ v = reinterpret_cast<T*>(vs[ti]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    T gi = g[i];
    // grad.clamp_(-clip, clip): NaN propagates like torch.clamp
    gi = (gi != gi) ? gi : b2_min(b2_max(gi, -clip), clip);
    const T pi = p[i];
    if (wd != (T)0) gi = gi + wd * pi;
    const T mi = m[i] * b1 + omb1 * gi;            // exp_avg.mul_(b1).add_(grad, alpha=1-b1)
    const T vi = v[i] * b2v + omb2 * gi * gi;      // exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2)
    const T denom = b2_sqrt(vi) + eps;             // exp_avg_sq.sqrt().add_(eps)
    p[i] = pi + neg_step * (mi / denom);           // p.addcdiv_(exp_avg, denom, value=-step_size)
    m[i] = mi;
    v[i] = vi;
    if (zero_grad) g[i] = (T)0;
  }
}

// pyro/optim/adagrad_rmsprop.py:54-87
__global__ void agr_advance_kernel(int n, double* hyper, int32_t* steps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* h = hyper + (size_t)i * kAgrHyperStride;
  const int32_t t = steps[i] + 1;
  steps[i] = t;
  // lr = eta * step^(-0.5 + delta)
  h[3] = h[0] * pow((double)t, -0.5 + h[1]);
}

template <typename T>
__global__ void __launch_bounds__(256) adagrad_rmsprop_kernel(void* const* __restrict__ ps,
                                                              void* const* __restrict__ gs,
                                                              void* const* __restrict__ ss,
                                                              const int64_t* __restrict__ numel,
                                                              const double* __restrict__ hyper,
                                                              const int32_t* __restrict__ steps,
                                                              int zero_grad) {
  const int ti = blockIdx.y;
  const int64_t n = numel[ti];
  const double* h = hyper + (size_t)ti * kAgrHyperStride;
  const T t = (T)h[2];
  const T omt = (T)(1.0 - h[2]);
  const T lr = (T)h[3];
  const bool first = steps[ti] == 1;  // advance kernel already incremented
  T* __restrict__ p = reinterpret_cast<T*>(ps[ti]);
  T* __restrict__ g = reinterpret_cast<T*>(gs[ti]);
  T* __restrict__ s = reinterpret_cast<T*>(ss[ti]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const T gi = g[i];
    // state["sum"] = grad*grad on the first step, else sum*(1-t) + (t*grad)*grad
    const T si = first ? gi * gi : s[i] * omt + (t * gi) * gi;
    s[i] = si;
    // p.addcdiv_(grad, 1 + sqrt(sum), value=-lr)
    p[i] = p[i] + (-lr) * (gi / ((T)1 + b2_sqrt(si)));
    if (zero_grad) g[i] = (T)0;
  }
}

inline unsigned blocks_for(int64_t max_numel) {
  int64_t b = (max_numel + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_clipped_adam(int n, void* const* p, void* const* g, void* const* m,
                               void* const* v, const int64_t* numel, double* hyper, double* lrs,
                               int32_t* steps, int dtype, int zero_grad, int64_t max_numel,
                               void* stream) {
  if (n <= 0) return B2_OK;
  if (!p || !g || !m || !v || !numel || !hyper || !lrs || !steps) return B2_ERR_NULL;
  if (dtype != B2_F32 && dtype != B2_F64) return B2_ERR_BAD_DTYPE;
  if (n > 65535) return B2_ERR_TOO_LARGE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (max_numel <= kAdamFusedAdvanceNumel) {
    // every tensor fits one CTA's grid-stride loop: one launch does advance + update
    dim3 grid(1, (unsigned)n, 1);
    if (dtype == B2_F32)
      launch_pdl(clipped_adam_kernel<float, true>, grid, dim3(256), 0, s, p, g, m, v, numel, hyper, lrs, steps, zero_grad);
    else
      launch_pdl(clipped_adam_kernel<double, true>, grid, dim3(256), 0, s, p, g, m, v, numel, hyper, lrs, steps, zero_grad);
    count_launch(1);
    return check_launch();
  }
  adam_advance_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, hyper, lrs, steps);
  dim3 grid(blocks_for(max_numel), (unsigned)n, 1);
  if (dtype == B2_F32)
    clipped_adam_kernel<float, false><<<grid, 256, 0, s>>>(p, g, m, v, numel, hyper, lrs, steps, zero_grad);
  else
    clipped_adam_kernel<double, false><<<grid, 256, 0, s>>>(p, g, m, v, numel, hyper, lrs, steps, zero_grad);
  count_launch(2);
  return check_launch();
}

extern "C" int b2_adagrad_rmsprop(int n, void* const* p, void* const* g, void* const* ssum,
                                  const int64_t* numel, double* hyper, int32_t* steps, int dtype,
                                  int zero_grad, int64_t max_numel, void* stream) {
  if (n <= 0) return B2_OK;
  if (!p || !g || !ssum || !numel || !hyper || !steps) return B2_ERR_NULL;
  if (dtype != B2_F32 && dtype != B2_F64) return B2_ERR_BAD_DTYPE;
  if (n > 65535) return B2_ERR_TOO_LARGE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  agr_advance_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, hyper, steps);
  dim3 grid(blocks_for(max_numel), (unsigned)n, 1);
  if (dtype == B2_F32)
    adagrad_rmsprop_kernel<float><<<grid, 256, 0, s>>>(p, g, ssum, numel, hyper, steps, zero_grad);
  else
    adagrad_rmsprop_kernel<double><<<grid, 256, 0, s>>>(p, g, ssum, numel, hyper, steps, zero_grad);
  count_launch(2);
  return check_launch();
}
