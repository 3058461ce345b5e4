// umma_bench.cu -- how long does ONE tcgen05.mma of a small TF32 / BF16 shape occupy the tensor pipe on B200?
// One CTA per SM; an elected thread issues R MMAs (operands: K-major SWIZZLE_128B tiles of arbitrary
// finite data in shared memory), commits, and the time from first issue to the mbarrier completion is
// measured with clock64.  Variants: the same accumulator for every MMA (dependent chain) or NACC rotating
// accumulators.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc(int fmt, int M, int N) {   // fmt 2 = tf32, 1 = bf16
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

template <int KIND, int M, int N, int NACC, int R>
__global__ void __launch_bounds__(128, 1) bench(long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  __shared__ uint64_t barv;
  __shared__ uint32_t tslot;
  // 64 KB of finite data: A tile at 0, B tile at 32 KB
  for (int i = threadIdx.x; i < 16384; i += 128) reinterpret_cast<float*>(sm)[i] = 0.001f * (float)(i % 977);
  const uint32_t bar = smem_u32(&barv);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (threadIdx.x < 32) {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)::"memory");
    if (pred) {
      const uint64_t da = make_desc(base, 0, 1024, 2), db = make_desc(base + 32768, 0, 1024, 2);
      constexpr uint32_t id = idesc(KIND == 0 ? 2 : 1, M, N);
      t0 = clock64();
#pragma unroll
      for (int r = 0; r < R; ++r)
        mma<KIND>(tmem + (uint32_t)(r % NACC) * (uint32_t)N, da + (uint64_t)((r & 3) * 2), db + (uint64_t)((r & 3) * 2), id,
                  r >= NACC ? 1u : 0u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      t1 = clock64();
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar) : "memory");
      t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int KIND, int M, int N, int NACC, int R>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(bench<KIND, M, N, NACC, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  long long h[2];
  for (int rep = 0; rep < 3; ++rep) {
    bench<KIND, M, N, NACC, R><<<148, 128, 70000>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s M=%3d N=%3d acc=%d R=%3d : issue %6lld cyc (%5.1f / MMA)  done %6lld cyc (%5.1f / MMA)\n", name, M, N, NACC, R,
         h[0], (double)h[0] / R, h[1], (double)h[1] / R);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<0, 128, 64, 1, 64>("tf32 dependent chain", d);
  run<0, 128, 64, 2, 64>("tf32 2 accumulators", d);
  run<0, 128, 64, 4, 64>("tf32 4 accumulators", d);
  run<0, 128, 128, 1, 64>("tf32 dependent chain", d);
  run<0, 128, 128, 2, 64>("tf32 2 accumulators", d);
  run<0, 128, 256, 1, 64>("tf32 dependent chain", d);
  run<0, 64, 40, 1, 64>("tf32 dependent chain", d);
  run<0, 64, 40, 4, 64>("tf32 4 accumulators", d);
  run<0, 64, 64, 4, 64>("tf32 4 accumulators", d);
  run<0, 64, 128, 2, 64>("tf32 2 accumulators", d);
  run<1, 128, 64, 1, 64>("bf16 dependent chain", d);
  run<1, 128, 64, 4, 64>("bf16 4 accumulators", d);
  run<1, 128, 256, 1, 64>("bf16 dependent chain", d);
  run<1, 64, 48, 4, 64>("bf16 4 accumulators", d);
  run<0, 128, 64, 1, 8>("tf32 dependent chain (short)", d);
  run<0, 64, 40, 4, 16>("tf32 4 accumulators (short)", d);
  return 0;
}
