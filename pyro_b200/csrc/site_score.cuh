// site_score.cuh -- fused log_prob + score kernels for elementwise families.
//
// One pass over the operands produces: the (scaled, masked) log_prob tensor if wanted, its sum,
// and the gradient w.r.t. value and every parameter, either full shape or summed to a scalar.
// Replaces the ATen chains behind pyro/poutine/trace_struct.py:248-328 (see pyro_b200.h).
//
// Two kernels:
//   site_vec_kernel  -- operands collapse to [R, C] with unit or zero inner stride; 16-byte
//                       vector loads/stores, 2 vectors in flight per operand per thread,
//                       no integer division in the loop.  This is the HBM-roofline kernel.
//   site_gen_kernel  -- any strides up to kMaxD dims; one element per thread iteration.
#pragma once
#include "b2_common.cuh"
#include "b2_math.cuh"

namespace b2 {

struct Opnd {
  const void* ptr;
  int64_t st[kMaxD];
};
struct OutOpnd {
  void* ptr;
  int64_t st[kMaxD];
  int mode;  // 0 = not wanted, 1 = full shape, 2 = scalar (sum over everything)
};

struct SiteArgs {
  int ndim;
  int64_t shape[kMaxD];
  int64_t n;
  Opnd x;
  Opnd p[B2_MAX_PARAMS];
  Opnd mask;  // uint8
  Opnd up;    // upstream gradient (same dtype as value)
  OutOpnd lp;
  OutOpnd gx;
  OutOpnd gp[B2_MAX_PARAMS];
  double scale, weight, sum_coeff;
  int flags;
  void* out_sum;
  double* partials;
  unsigned int* ticket;
  // vector path
  int64_t R, C;
  int tx_log2;  // threads along the column-vector axis = 1 << tx_log2 (block is 256 threads)
};

template <typename T>
__device__ __forceinline__ void finish_outputs(const SiteArgs& a, int k, double tot) {
  constexpr int NPmax = B2_MAX_PARAMS;
  if (k == 0) {
    if (a.out_sum) {
      T* o = reinterpret_cast<T*>(a.out_sum);
      const double s = a.sum_coeff * tot;
      *o = (a.flags & B2_FLAG_ACCUMULATE_SUM) ? (T)((double)*o + s) : (T)s;
    }
  } else if (k == 1) {
    if (a.gx.mode == 2) *reinterpret_cast<T*>(a.gx.ptr) = (T)tot;
  } else if (k - 2 < NPmax) {
    if (a.gp[k - 2].mode == 2) *reinterpret_cast<T*>(a.gp[k - 2].ptr) = (T)tot;
  }
}

// ------------------------------------------------------------------------------------------------
// Vector kernel.  Block = 256 threads arranged TX x TY (TX = 1 << tx_log2 column-vector lanes,
// TY rows).  grid.x tiles column vectors, grid.y tiles rows; both grid-strided.
// ------------------------------------------------------------------------------------------------
template <int FAM, typename T, bool GRAD>
__global__ void __launch_bounds__(256) site_vec_kernel(const SiteArgs a) {
  constexpr int NP = FamilyTraits<FAM>::kNumParams;
  constexpr bool HASV = FamilyTraits<FAM>::kHasValue;
  constexpr int V = VecOf<T>::N;
  constexpr int U = 2;  // vectors in flight per operand
  constexpr int NRED = GRAD ? 2 + NP : 1;

  const int TX = 1 << a.tx_log2;
  const int tx = threadIdx.x & (TX - 1);
  const int ty = threadIdx.x >> a.tx_log2;
  const int TY = 256 >> a.tx_log2;
  const int64_t CV = a.C / V;
  const int64_t cstep = (int64_t)gridDim.x * TX;

  const T* __restrict__ xp = reinterpret_cast<const T*>(a.x.ptr);
  const T* __restrict__ up = reinterpret_cast<const T*>(a.up.ptr);
  const uint8_t* __restrict__ mp = reinterpret_cast<const uint8_t*>(a.mask.ptr);
  T* __restrict__ lpo = reinterpret_cast<T*>(a.lp.ptr);
  const bool want_lp = a.lp.mode == 1;
  const T f0 = (T)(a.weight * a.scale);
  const T scale = (T)a.scale;

  T acc[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) acc[k] = (T)0;

  for (int64_t r = (int64_t)blockIdx.y * TY + ty; r < a.R; r += (int64_t)gridDim.y * TY) {
    // per-row bases
    const T* xr = HASV ? xp + r * a.x.st[0] : nullptr;
    const T* pr[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) pr[k] = reinterpret_cast<const T*>(a.p[k].ptr) + r * a.p[k].st[0];
    const T* ur = up ? up + r * a.up.st[0] : nullptr;
    const uint8_t* mr = mp ? mp + r * a.mask.st[0] : nullptr;

    for (int64_t cv0 = (int64_t)blockIdx.x * TX + tx; cv0 < CV; cv0 += cstep * U) {
      Pack<T> xv[U], pv[U][NP], uv[U];
      uint8_t mv[U][V];
      bool ok[U];
      // ---- issue every load first -----------------------------------------------------------
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t c = (cv0 + u * cstep) * V;
        ok[u] = (cv0 + u * cstep) < CV;
        if (ok[u]) {
          if (HASV) {
            if (a.x.st[1] == 1) {
              xv[u] = (a.x.st[0] != 0 || a.R == 1) ? ld_stream(xr + c) : ld_keep(xr + c);
            } else {
              const T s = __ldg(xr);
#pragma unroll
              for (int j = 0; j < V; ++j) xv[u].v[j] = s;
            }
          }
#pragma unroll
          for (int k = 0; k < NP; ++k) {
            if (a.p[k].st[1] == 1) {
              pv[u][k] = (a.p[k].st[0] != 0 || a.R == 1) ? ld_stream(pr[k] + c) : ld_keep(pr[k] + c);
            } else {
              const T s = __ldg(pr[k]);
#pragma unroll
              for (int j = 0; j < V; ++j) pv[u][k].v[j] = s;
            }
          }
          if (ur) {
            if (a.up.st[1] == 1) {
              uv[u] = ld_stream(ur + c);
            } else {
              const T s = __ldg(ur);
#pragma unroll
              for (int j = 0; j < V; ++j) uv[u].v[j] = s;
            }
          }
          if (mr) {
#pragma unroll
            for (int j = 0; j < V; ++j) mv[u][j] = mr[a.mask.st[1] == 1 ? c + j : 0];
          }
        }
      }
      // ---- compute + store ------------------------------------------------------------------
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const int64_t c = (cv0 + u * cstep) * V;
        Pack<T> lpv, gxv, gpv[NP];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          T pl[NP > 0 ? NP : 1];
#pragma unroll
          for (int k = 0; k < NP; ++k) pl[k] = pv[u][k].v[j];
          ElemOut<T> o;
          Eval<FAM, T, GRAD>::run(HASV ? xv[u].v[j] : (T)0, pl, o);
          const bool m = mr ? (mv[u][j] != 0) : true;
          const T slp = m ? o.lp * scale : (T)0;
          lpv.v[j] = slp;
          acc[0] += slp;
          if (GRAD) {
            T f = m ? f0 : (T)0;
            if (ur) f *= uv[u].v[j];
            const T gxe = m ? f * o.dx : (T)0;
            gxv.v[j] = gxe;
            acc[1] += gxe;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              const T g = m ? f * o.dp[k] : (T)0;
              gpv[k].v[j] = g;
              acc[2 + k] += g;
            }
          }
        }
        if (want_lp) st_stream(lpo + r * a.lp.st[0] + c, lpv);
        if (GRAD) {
          if (a.gx.mode == 1) st_stream(reinterpret_cast<T*>(a.gx.ptr) + r * a.gx.st[0] + c, gxv);
#pragma unroll
          for (int k = 0; k < NP; ++k)
            if (a.gp[k].mode == 1)
              st_stream(reinterpret_cast<T*>(a.gp[k].ptr) + r * a.gp[k].st[0] + c, gpv[k]);
        }
      }
    }
  }

  __shared__ double smem[NRED * 32];
  double red[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) red[k] = (double)acc[k];
  grid_finish<NRED>(red, a.partials, a.ticket, smem,
                    [&](int k, double tot) { finish_outputs<T>(a, k, tot); });
}

// ------------------------------------------------------------------------------------------------
// Generic kernel: arbitrary strides, one element per loop trip.
// ------------------------------------------------------------------------------------------------
template <int FAM, typename T, bool GRAD>
__global__ void __launch_bounds__(256) site_gen_kernel(const SiteArgs a) {
  constexpr int NP = FamilyTraits<FAM>::kNumParams;
  constexpr bool HASV = FamilyTraits<FAM>::kHasValue;
  constexpr int NRED = GRAD ? 2 + NP : 1;
  const T f0 = (T)(a.weight * a.scale);
  const T scale = (T)a.scale;
  T acc[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) acc[k] = (T)0;

  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = i;
    int64_t ox = 0, om = 0, ou = 0, olp = 0, ogx = 0;
    int64_t op[NP > 0 ? NP : 1], ogp[NP > 0 ? NP : 1];
#pragma unroll
    for (int k = 0; k < NP; ++k) op[k] = ogp[k] = 0;
    for (int d = a.ndim - 1; d >= 0; --d) {
      const int64_t q = rem / a.shape[d];
      const int64_t idx = rem - q * a.shape[d];
      rem = q;
      ox += idx * a.x.st[d];
      om += idx * a.mask.st[d];
      ou += idx * a.up.st[d];
      olp += idx * a.lp.st[d];
      ogx += idx * a.gx.st[d];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        op[k] += idx * a.p[k].st[d];
        ogp[k] += idx * a.gp[k].st[d];
      }
    }
    T pl[NP > 0 ? NP : 1];
#pragma unroll
    for (int k = 0; k < NP; ++k) pl[k] = reinterpret_cast<const T*>(a.p[k].ptr)[op[k]];
    const T xv = HASV ? reinterpret_cast<const T*>(a.x.ptr)[ox] : (T)0;
    const bool m = a.mask.ptr ? reinterpret_cast<const uint8_t*>(a.mask.ptr)[om] != 0 : true;
    ElemOut<T> o;
    Eval<FAM, T, GRAD>::run(xv, pl, o);
    const T slp = m ? o.lp * scale : (T)0;
    acc[0] += slp;
    if (a.lp.mode == 1) reinterpret_cast<T*>(a.lp.ptr)[olp] = slp;
    if (GRAD) {
      T f = m ? f0 : (T)0;
      if (a.up.ptr) f *= reinterpret_cast<const T*>(a.up.ptr)[ou];
      const T gxe = m ? f * o.dx : (T)0;
      acc[1] += gxe;
      if (a.gx.mode == 1) reinterpret_cast<T*>(a.gx.ptr)[ogx] = gxe;
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const T g = m ? f * o.dp[k] : (T)0;
        acc[2 + k] += g;
        if (a.gp[k].mode == 1) reinterpret_cast<T*>(a.gp[k].ptr)[ogp[k]] = g;
      }
    }
  }
  __shared__ double smem[NRED * 32];
  double red[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) red[k] = (double)acc[k];
  grid_finish<NRED>(red, a.partials, a.ticket, smem,
                    [&](int k, double tot) { finish_outputs<T>(a, k, tot); });
}

// host-side launcher for one (family, dtype, grad) combination
template <int FAM, typename T, bool GRAD>
int launch_site(const SiteArgs& a, bool vec, cudaStream_t stream) {
  if (vec) {
    constexpr int V = VecOf<T>::N;
    const int64_t CV = a.C / V;
    const int TX = 1 << a.tx_log2, TY = 256 / TX;
    int64_t gx = (CV + (int64_t)TX * 2 - 1) / ((int64_t)TX * 2);
    int64_t gy = (a.R + TY - 1) / TY;
    const int64_t target = (int64_t)kNumSMs * 8;
    if (gx > target) gx = target;
    int64_t gy_cap = target / gx;
    if (gy_cap < 1) gy_cap = 1;
    if (gy > gy_cap) gy = gy_cap;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)gx, (unsigned)gy, 1);
    site_vec_kernel<FAM, T, GRAD><<<grid, 256, 0, stream>>>(a);
  } else {
    int64_t blocks = (a.n + 255) / 256;
    const int64_t target = (int64_t)kNumSMs * 8;
    if (blocks > target) blocks = target;
    if (blocks < 1) blocks = 1;
    site_gen_kernel<FAM, T, GRAD><<<(unsigned)blocks, 256, 0, stream>>>(a);
  }
  count_launch();
  return check_launch();
}

// implemented in site_score_fam*.cu (split so the families compile in parallel)
int dispatch_site_a(int family, int dtype, bool grad, const SiteArgs& a, bool vec, cudaStream_t s);
int dispatch_site_b(int family, int dtype, bool grad, const SiteArgs& a, bool vec, cudaStream_t s);
int dispatch_site_c(int family, int dtype, bool grad, const SiteArgs& a, bool vec, cudaStream_t s);

#define B2_DISPATCH_CASE(FAM)                                                         \
  case FAM:                                                                           \
    if (dtype == B2_F32)                                                              \
      return grad ? launch_site<FAM, float, true>(a, vec, s)                          \
                  : launch_site<FAM, float, false>(a, vec, s);                        \
    else                                                                              \
      return grad ? launch_site<FAM, double, true>(a, vec, s)                         \
                  : launch_site<FAM, double, false>(a, vec, s);

}  // namespace b2
