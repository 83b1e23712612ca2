// tcgen05.st / tcgen05.ld throughput probe: NW warps (warp w -> TMEM lane quarter w % 4) issue back-to-back 32x32b.x16 stores
// (or loads) and report bytes per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_rate tools/tmem_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>   // 0 = st.x16, 1 = ld.x16, 2 = st.x4, 3 = st.x32
__global__ void rate_kernel(long long* out, int iters) {
  __shared__ uint32_t tbase_s;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s + (((threadIdx.x >> 5) & 3) * 32 << 16) + ((threadIdx.x >> 7) * 64);
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t a = tb + (i & 1) * 32;
    if (MODE == 0)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                   ::"r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                   "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    if (MODE == 1) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(a));
    }
    if (MODE == 2)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
    if (MODE == 3)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                   ::"r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                   "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]),
                   "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
  }
  if (MODE == 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += v[i];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) out[1000] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s));
}
template <int MODE>
int run(const char* name, int cols, int nwarps) {
  long long* d; CK(cudaMalloc(&d, 2048 * 8));
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) rate_kernel<MODE><<<148, nwarps * 32>>>(d, iters);
  CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
  printf("%s, %2d warps: %.1f clk per instruction per warp, %.1f B/clk/SM\n", name, nwarps, (double)h / iters, (double)iters * nwarps * 32 * cols * 4 / h);
  cudaFree(d);
  return 0;
}
int main() {
  for (int nw : {1, 4, 8, 12}) run<0>("st.x16", 16, nw);
  for (int nw : {4, 8}) run<2>("st.x4 ", 4, nw);
  for (int nw : {4, 8}) run<3>("st.x32", 32, nw);
  for (int nw : {1, 4, 8}) run<1>("ld.x16", 16, nw);
  return 0;
}
