// tcgen05.mma issue-rate probe: clk per instruction for kind::tf32, M = 128, K = 8, A from TMEM (TS) or smem (SS), N = 32..256.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mma_rate tools/mma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int N, bool TS, int NISS, int NCOMMIT = 0, int NMMA = 1>
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int iters) {
  extern __shared__ __align__(1024) float smem[];   // B: [2][N][4] floats (K = 8), A (SS): [2][128][4]
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];   // targets of the per-iteration commits (never waited on)
  __shared__ uint32_t tbase_s;
  for (int i = threadIdx.x; i < 2 * N * 4 + 2 * 128 * 4; i += blockDim.x) smem[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NISS));
    for (int k = 0; k < 4; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2[k])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s;
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < NISS) {
    const int wi = threadIdx.x >> 5;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bd = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF) | ((uint64_t)(((uint32_t)(N * 16) >> 4) & 0x3FFF) << 16) |
                        ((uint64_t)((128u >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
    const uint64_t ad = (uint64_t)((smem_u32(smem + 2 * N * 4) >> 4) & 0x3FFF) | ((uint64_t)(((uint32_t)(128 * 16) >> 4) & 0x3FFF) << 16) |
                        ((uint64_t)((128u >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
     for (int m = 0; m < NMMA; ++m) {
      const uint32_t d = tb + (uint32_t)(wi * 64 + (i & 1) * 32);   // accumulators private to the issuing warp (N <= 32 when NISS > 1)
      if (TS)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d), "r"(tb + 256 + (uint32_t)((i & 3) * 16)), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
     }
#pragma unroll
      for (int k = 0; k < NCOMMIT; ++k)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[(wi * 2 + k) & 3])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    if (wi == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}
template <int N, bool TS, int NISS = 1, int NCOMMIT = 0, int NMMA = 1>
int run() {
  long long* d; CK(cudaMalloc(&d, 148 * 16));
  const int iters = 2000;
  CK(cudaFuncSetAttribute(rate_kernel<N, TS, NISS, NCOMMIT, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int rep = 0; rep < 2; ++rep) rate_kernel<N, TS, NISS, NCOMMIT, NMMA><<<148, 128, (2 * N * 4 + 2 * 128 * 4) * 4>>>(d, iters);
  CK(cudaDeviceSynchronize());
  long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  printf("%d mma + %d commit per iteration, %d issuer(s) %s N=%3d: per issuer per iteration: issue %.1f clk/mma, complete %.1f clk/mma (floor 128*N/256 = %d)\n", NMMA, NCOMMIT, NISS, TS ? "TS" : "SS", N, (double)h[0] / iters, (double)h[1] / iters, N / 2);
  cudaFree(d);
  return 0;
}
int main() {
  run<32, true>(); run<48, true>(); run<64, true>(); run<96, true>(); run<128, true>(); run<256, true>();
  run<32, false>(); run<96, false>(); run<256, false>();
  run<32, true, 2>(); run<32, true, 4>(); run<32, false, 2>();
  run<96, true, 1, 1, 1>(); run<96, true, 1, 2, 1>(); run<96, true, 1, 1, 3>(); run<96, true, 1, 2, 6>(); run<96, true, 1, 2, 9>();
  run<32, true, 2, 1, 1>(); run<32, true, 2, 2, 3>(); run<32, true, 2, 2, 6>();
  return 0;
}
