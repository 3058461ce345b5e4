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
// Inner loop: U vectors per operand are loaded back to back (no predicates, no per-element
// branches), then evaluated, then stored; a one-vector tail loop handles the remainder.
// An operand is either a vector along the columns (column stride 1) or one scalar per row
// (column stride 0, loaded once per row).  MASKUP adds the optional mask / upstream operands.
// ------------------------------------------------------------------------------------------------
template <int FAM, typename T, bool GRAD, bool MASKUP, int NVEC>
struct VecBody {
  static constexpr int NP = FamilyTraits<FAM>::kNumParams;
  static constexpr bool HASV = FamilyTraits<FAM>::kHasValue;
  static constexpr int V = VecOf<T>::N;
  static constexpr int NRED = GRAD ? 2 + NP : 1;
  // two rows at once only where the register budget allows it without spilling: fp32 families
  // with a single parameter (Bernoulli, Poisson, Exponential, Half*): the [particles, N] sites
  static constexpr bool kTwoRows = sizeof(T) == 4 && NP == 1 && !MASKUP && NVEC <= 2;

  // operand classes, fixed for the launch: vector along columns (column stride 1) or one scalar
  // per row; row-invariant (row stride 0) vectors are loaded once per column chunk and reused
  // for every row (e.g. the observations y[N] scored against logits[P, N])
  bool x_vec, p_vec[NP], x_inv, p_inv[NP];
  bool u_vec, m_vec;
  T scale, f0;
  // operand registers: slot [q] belongs to the q-th row in flight; row-invariant operands are
  // loaded once per column chunk into slot 0 and read from there by every row
  static constexpr int NRMAX = kTwoRows ? 2 : 1;
  Pack<T> xv[NRMAX][NVEC], pv[NRMAX][NVEC][NP];

  struct Row {
    const T* xr;
    const T* pr[NP];
    const T* ur;
    const uint8_t* mr;
    T* lpr;
    T* gxr;
    T* gpr[NP];
    T xs, ps[NP], us;
  };

  __device__ __forceinline__ void setup_row(const SiteArgs& a, int64_t r, Row& w) const {
    w.xr = HASV ? reinterpret_cast<const T*>(a.x.ptr) + r * a.x.st[0] : nullptr;
    w.xs = (HASV && !x_vec) ? __ldg(w.xr) : (T)0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      w.pr[k] = reinterpret_cast<const T*>(a.p[k].ptr) + r * a.p[k].st[0];
      w.ps[k] = p_vec[k] ? (T)0 : __ldg(w.pr[k]);
    }
    w.ur = nullptr;
    w.mr = nullptr;
    w.us = (T)1;
    if (MASKUP) {
      w.ur = a.up.ptr ? reinterpret_cast<const T*>(a.up.ptr) + r * a.up.st[0] : nullptr;
      w.us = (w.ur && !u_vec) ? __ldg(w.ur) : (T)1;
      w.mr = a.mask.ptr ? reinterpret_cast<const uint8_t*>(a.mask.ptr) + r * a.mask.st[0] : nullptr;
    }
    w.lpr = a.lp.mode == 1 ? reinterpret_cast<T*>(a.lp.ptr) + r * a.lp.st[0] : nullptr;
    w.gxr = (GRAD && a.gx.mode == 1) ? reinterpret_cast<T*>(a.gx.ptr) + r * a.gx.st[0] : nullptr;
#pragma unroll
    for (int k = 0; k < NP; ++k)
      w.gpr[k] = (GRAD && a.gp[k].mode == 1) ? reinterpret_cast<T*>(a.gp[k].ptr) + r * a.gp[k].st[0] : nullptr;
  }

  __device__ __forceinline__ void load_invariant(const SiteArgs& a, int64_t cv, int64_t cstep) {
#pragma unroll
    for (int u = 0; u < NVEC; ++u) {
      const int64_t c = (cv + u * cstep) * V;
      if (HASV && x_vec && x_inv) xv[0][u] = ld_keep(reinterpret_cast<const T*>(a.x.ptr) + c);
#pragma unroll
      for (int k = 0; k < NP; ++k)
        if (p_vec[k] && p_inv[k]) pv[0][u][k] = ld_keep(reinterpret_cast<const T*>(a.p[k].ptr) + c);
    }
  }

  // NR rows at once: all their loads are issued before any arithmetic
  template <int NR>
  __device__ __forceinline__ void run(const Row (&w)[NR], int64_t cv, int64_t cstep, T (&acc)[NRED]) {
    Pack<T> uv[NR][NVEC];
    uint32_t mbits[NR][NVEC];
#pragma unroll
    for (int q = 0; q < NR; ++q) {
#pragma unroll
      for (int u = 0; u < NVEC; ++u) {
        const int64_t c = (cv + u * cstep) * V;
        if (HASV && x_vec && !x_inv) xv[q][u] = ld_stream(w[q].xr + c);
#pragma unroll
        for (int k = 0; k < NP; ++k)
          if (p_vec[k] && !p_inv[k]) pv[q][u][k] = ld_stream(w[q].pr[k] + c);
        if (MASKUP) {
          if (w[q].ur && u_vec) uv[q][u] = ld_stream(w[q].ur + c);
          mbits[q][u] = 0xffffffffu;
          if (w[q].mr) {
            mbits[q][u] = 0;
#pragma unroll
            for (int j = 0; j < V; ++j)
              mbits[q][u] |= (w[q].mr[m_vec ? c + j : 0] != 0 ? 1u : 0u) << j;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NR; ++q) {
#pragma unroll
      for (int u = 0; u < NVEC; ++u) {
        const int64_t c = (cv + u * cstep) * V;
        Pack<T> lpv, gxv, gpv[NP];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          T pl[NP];
#pragma unroll
          for (int k = 0; k < NP; ++k)
            pl[k] = p_vec[k] ? ((q > 0 && p_inv[k]) ? pv[0][u][k].v[j] : pv[q][u][k].v[j]) : w[q].ps[k];
          const T xe = HASV ? (x_vec ? ((q > 0 && x_inv) ? xv[0][u].v[j] : xv[q][u].v[j]) : w[q].xs) : (T)0;
          ElemOut<T> o;
          Eval<FAM, T, GRAD>::run(xe, pl, o);
          T slp = o.lp * scale;
          T f = f0;
          if (MASKUP) {
            const bool m = (mbits[q][u] >> j) & 1u;
            slp = m ? slp : (T)0;
            f = m ? f0 : (T)0;
            if (w[q].ur) f *= u_vec ? uv[q][u].v[j] : w[q].us;
          }
          lpv.v[j] = slp;
          acc[0] += slp;
          if (GRAD) {
            T gxe = f * o.dx;
            if (MASKUP) gxe = (f == (T)0) ? (T)0 : gxe;  // masked-out NaNs must not leak
            gxv.v[j] = gxe;
            acc[1] += gxe;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              T g = f * o.dp[k];
              if (MASKUP) g = (f == (T)0) ? (T)0 : g;
              gpv[k].v[j] = g;
              acc[2 + k] += g;
            }
          }
        }
        if (w[q].lpr) st_stream(w[q].lpr + c, lpv);
        if (GRAD) {
          if (w[q].gxr) st_stream(w[q].gxr + c, gxv);
#pragma unroll
          for (int k = 0; k < NP; ++k)
            if (w[q].gpr[k]) st_stream(w[q].gpr[k] + c, gpv[k]);
        }
      }
    }
  }

  // all rows of this thread for one column chunk; rows are taken two at a time so that twice the
  // bytes are in flight per thread (the row loop is where a [P, N] site spends its time)
  __device__ __forceinline__ void chunk(const SiteArgs& a, int64_t cv, int64_t cstep, int ty, int TY,
                                        T (&acc)[NRED]) {
    load_invariant(a, cv, cstep);
    const int64_t rstep = (int64_t)gridDim.y * TY;
    int64_t r = (int64_t)blockIdx.y * TY + ty;
    if (kTwoRows) {
      for (; r + rstep < a.R; r += 2 * rstep) {
        Row w[2];
        setup_row(a, r, w[0]);
        setup_row(a, r + rstep, w[1]);
        run<2>(w, cv, cstep, acc);
      }
    }
    for (; r < a.R; r += rstep) {
      Row w[1];
      setup_row(a, r, w[0]);
      run<1>(w, cv, cstep, acc);
    }
  }

  __device__ __forceinline__ void init(const SiteArgs& a) {
    x_vec = a.x.st[1] == 1;
    x_inv = a.x.st[0] == 0 && a.R > 1;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      p_vec[k] = a.p[k].st[1] == 1;
      p_inv[k] = a.p[k].st[0] == 0 && a.R > 1;
    }
    scale = (T)a.scale;
    f0 = (T)(a.weight * a.scale);
    u_vec = MASKUP && a.up.st[1] == 1;
    m_vec = MASKUP && a.mask.st[1] == 1;
  }
};

// Vectors in flight per operand per thread.  Forward-only fp32 kernels are pure streams: 4 vectors
// (measured 87% of the HBM copy peak for Normal).  Kernels that also write gradients carry more
// live registers; 2 vectors keep them at 3 resident CTAs per SM.
template <int FAM, typename T, bool GRAD>
struct VecUnroll {
  static constexpr int U = (sizeof(T) == 4 && !GRAD) ? 4 : 2;
  // families without lgamma/digamma are streams: keep 3 CTAs per SM resident (<= 80 registers,
  // no spills measured with ptxas -v); the special-function families are issue-bound and get the
  // full 128-register budget instead of spilling
  static constexpr bool kLight = FAM == kNormal || FAM == kCauchy || FAM == kHalfCauchy ||
                                 FAM == kExponential || FAM == kHalfNormal || FAM == kUniform;
  static constexpr int kMinBlocks = (sizeof(T) == 4 && !GRAD && kLight) ? 3 : 2;
};

// Loop nest: column chunks outermost (U vectors per thread, then a one-vector tail), rows inside.
template <int FAM, typename T, bool GRAD, bool MASKUP>
__global__ void __launch_bounds__(256, VecUnroll<FAM, T, GRAD>::kMinBlocks) site_vec_kernel(const SiteArgs a) {
  constexpr int NP = FamilyTraits<FAM>::kNumParams;
  constexpr int V = VecOf<T>::N;
  constexpr int U = VecUnroll<FAM, T, GRAD>::U;
  constexpr int NRED = GRAD ? 2 + NP : 1;

  const int TX = 1 << a.tx_log2;
  const int tx = threadIdx.x & (TX - 1);
  const int ty = threadIdx.x >> a.tx_log2;
  const int TY = 256 >> a.tx_log2;
  const int64_t CV = a.C / V;
  const int64_t cstep = (int64_t)gridDim.x * TX;

  T acc[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) acc[k] = (T)0;

  int64_t cv = (int64_t)blockIdx.x * TX + tx;
  {
    VecBody<FAM, T, GRAD, MASKUP, U> body;
    body.init(a);
    for (; cv + (U - 1) * cstep < CV; cv += U * cstep) body.chunk(a, cv, cstep, ty, TY, acc);
  }
  {
    VecBody<FAM, T, GRAD, MASKUP, 1> tail;
    tail.init(a);
    for (; cv < CV; cv += cstep) tail.chunk(a, cv, cstep, ty, TY, acc);
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
    constexpr int U = VecUnroll<FAM, T, GRAD>::U;
    const int64_t CV = a.C / V;
    const int TX = 1 << a.tx_log2, TY = 256 / TX;
    // enough CTAs for ~4 waves of resident blocks; each thread then owns >= U vectors per row
    int64_t gx = (CV + (int64_t)TX * U - 1) / ((int64_t)TX * U);
    int64_t gy = (a.R + TY - 1) / TY;
    const int64_t target = (int64_t)kNumSMs * 12;
    if (gx > target) gx = target;
    int64_t gy_cap = target / gx;
    if (gy_cap < 1) gy_cap = 1;
    if (gy > gy_cap) gy = gy_cap;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)gx, (unsigned)gy, 1);
    if (a.mask.ptr || a.up.ptr)
      site_vec_kernel<FAM, T, GRAD, true><<<grid, 256, 0, stream>>>(a);
    else
      site_vec_kernel<FAM, T, GRAD, false><<<grid, 256, 0, stream>>>(a);
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
