// Round-trip latency of mbarrier hand-offs inside one CTA (B200): producer warps arrive on `full`, a consumer thread waits for it and
// answers on `empty` (plain arrive or tcgen05.commit), the producers wait for `empty`.  Variants: arrivals per thread / per warp,
// number of producer warps, try_wait with / without a suspend hint, reply by tcgen05.commit, an STTM + wait::st before every arrive.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mbar_rtt tools/mbar_rtt.cu ; run: tools/mbar_rtt
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
template <int HINT>
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  if (HINT < 0) {   // test_wait: pure polling
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(addr), "r"(parity) : "memory");
  } else if (HINT == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(addr), "r"(parity) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@!p bra W_%=;\n\t}" ::"r"(addr), "r"(parity), "n"(HINT) : "memory");
  }
}

// MODE bit 0: one arrival per warp (else per thread); bit 1: consumer replies with tcgen05.commit; bit 2: producers do a tcgen05.st + wait::st
// before arriving; bit 3: the consumer is a whole converged warp that waits (lane 0 replies)
template <int HINT, int MODE>
__global__ void rtt_kernel(int iters, int nprod_warps, long long* out) {
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool per_warp = MODE & 1;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], per_warp ? nprod_warps : nprod_warps * 32);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == nprod_warps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64();
  if (warp < nprod_warps) {
    for (int i = 0; i < iters; ++i) {
      if (i > 0) {
        mbar_wait<HINT>(&bars[1], (i - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if ((MODE & 4) && warp < 4) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16)), "r"(i), "r"(i), "r"(i), "r"(i) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      if (per_warp) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[0]);
      } else {
        mbar_arrive(&bars[0]);
      }
    }
    mbar_wait<HINT>(&bars[1], (iters - 1) & 1);
  } else if (warp == nprod_warps && ((MODE & 8) || lane == 0)) {
    for (int i = 0; i < iters; ++i) {
      mbar_wait<HINT>(&bars[0], i & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        if (MODE & 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
        else mbar_arrive(&bars[1]);
      }
      if (MODE & 8) __syncwarp();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[0] = clock64() - t0;
  if (warp == nprod_warps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32));
}

template <int HINT, int MODE>
void run(const char* name, int nprod_warps, long long* d) {
  const int iters = 2000;
  rtt_kernel<HINT, MODE><<<1, (nprod_warps + 1) * 32>>>(iters, nprod_warps, d);
  rtt_kernel<HINT, MODE><<<1, (nprod_warps + 1) * 32>>>(iters, nprod_warps, d);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-58s producers %2d warps: %7.1f clk per round trip%s\n", name, nprod_warps, (double)h / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int w : {1, 4, 16}) {
    run<0, 0>("try_wait, per-thread arrive, mbarrier reply", w, d);
    run<0, 1>("try_wait, per-warp arrive, mbarrier reply", w, d);
    run<-1, 1>("test_wait, per-warp arrive, mbarrier reply", w, d);
    run<256, 1>("try_wait hint 256, per-warp arrive, mbarrier reply", w, d);
    run<0, 3>("try_wait, per-warp arrive, tcgen05.commit reply", w, d);
    run<0, 2>("try_wait, per-thread arrive, tcgen05.commit reply", w, d);
    run<0, 7>("try_wait, per-warp arrive, commit reply, STTM + wait::st", w, d);
    run<0, 6>("try_wait, per-thread arrive, commit reply, STTM + wait::st", w, d);
    run<0, 15>("  same per-warp, consumer = converged warp", w, d);
    run<-1, 7>("test_wait, per-warp arrive, commit reply, STTM + wait::st", w, d);
  }
  return 0;
}
