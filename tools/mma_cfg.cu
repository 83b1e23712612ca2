// tcgen05.mma (kind::tf32, M = 128, K = 8, A from TMEM) execution time against accumulator / operand placement in TMEM:
// number of accumulators in rotation, their column stride, A-operand column offsets, B slices.  One issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mma_cfg tools/mma_cfg.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Cfg { int N, dstride, same_d, same_a, same_b; };
// one iteration = one k-step of the chain kernel: 3 M-tiles x (a_hi w_hi, a_hi w_lo, a_lo w_hi), straight-line
__global__ void __launch_bounds__(128) k(long long* out, int iters, Cfg c) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  for (int i = threadIdx.x; i < 8 * 2 * 96 * 4; i += blockDim.x) smem[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
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
  if (threadIdx.x == 0) {
    const int N = c.N;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bfix = ((uint64_t)(((uint32_t)(N * 16) >> 4) & 0x3FFF) << 16) | ((uint64_t)((128u >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t ds = c.same_d ? 0u : (uint32_t)c.dstride;
    const uint32_t as = c.same_a ? 0u : 16u, al = c.same_a ? 0u : 8u;
    const uint32_t lo = c.same_b ? 0u : (uint32_t)N * 32u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t set = c.same_a ? 0u : (uint32_t)(i & 3), slot = c.same_b ? 0u : (uint32_t)(i & 7);
      const uint64_t dhi = bfix | (uint64_t)(((sbase + slot * 6144u) >> 4) & 0x3FFF);
      const uint64_t dlo = bfix | (uint64_t)(((sbase + slot * 6144u + lo) >> 4) & 0x3FFF);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const uint32_t d = tb + t * ds;
        const uint32_t a = tb + 288 + set * 48u + t * as;
#define MMA(D, A, B) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(D), "r"(A), "l"(B), "r"(idesc), "r"(1u) : "memory")
        MMA(d, a, dhi);
        MMA(d, a, dlo);
        MMA(d, a + al, dhi);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}
int run(Cfg c) {
  long long* d; CK(cudaMalloc(&d, 148 * 8));
  const int iters = 1000;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 8 * 6144>>>(d, iters, c);
  CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
  printf("N=%3d  D stride %3d%s | %s | %s : %.1f clk/mma (floor %d)\n", c.N, c.dstride, c.same_d ? " (one accumulator)" : " (3 accumulators)",
         c.same_a ? "one A tile" : "A: 4 stages x 3 tiles, hi / lo", c.same_b ? "one B slice" : "B: 8 ring slots, hi / lo", (double)h / (iters * 9), c.N / 2);
  cudaFree(d);
  return 0;
}
int main() {
  run({96, 96, 1, 1, 1});
  run({96, 96, 0, 1, 1});
  run({96, 128, 0, 1, 1});
  run({96, 96, 1, 0, 1});
  run({96, 96, 1, 1, 0});
  run({96, 96, 0, 0, 0});
  run({96, 128, 0, 0, 0});
  run({64, 96, 0, 0, 0});
  run({64, 64, 0, 0, 0});
  run({32, 96, 0, 0, 0});
  return 0;
}
