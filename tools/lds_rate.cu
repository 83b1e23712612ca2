// Shared-memory load throughput per SM (B200): LDS.128 with per-lane distinct addresses (conflict-free), warp-uniform LDS.128 (broadcast),
// 2-way conflicted LDS.128, warp-uniform LDS.32 / LDS.64.  16 warps per CTA, one CTA; clk per warp-instruction per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lds_rate tools/lds_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void lds_kernel(int iters, float* out, long long* clk) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int idx;
  if (MODE == 0) idx = warp * 37 + lane;               // distinct, conflict-free
  else if (MODE == 1) idx = warp * 37;                 // uniform
  else if (MODE == 2) idx = warp * 37 + lane * 2;      // 2-way conflict (even 16-byte chunks only)
  else idx = warp * 37;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = (idx + k * 64 + (i & 7) * 8) & 4095;
      if (MODE <= 2) {
        const float4 v = sm[j];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      } else if (MODE == 3) {
        acc.x += reinterpret_cast<const float*>(sm)[j * 4];
      } else {
        const float2 v = reinterpret_cast<const float2*>(sm)[j * 2];
        acc.x += v.x; acc.y += v.y;
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* clk) {
  const int iters = 2000, warps = 16;
  lds_kernel<MODE><<<1, warps * 32, 65536>>>(iters, out, clk);
  lds_kernel<MODE><<<1, warps * 32, 65536>>>(iters, out, clk);
  long long h = 0;
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-40s %6.2f clk per warp-instruction (16 warps on one SM)\n", name, (double)h / ((double)iters * 16 * warps));
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&clk, 8);
  cudaFuncSetAttribute(lds_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(lds_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(lds_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(lds_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(lds_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  run<0>("LDS.128 distinct, conflict-free", out, clk);
  run<1>("LDS.128 warp-uniform", out, clk);
  run<2>("LDS.128 2-way conflicted", out, clk);
  run<3>("LDS.32 warp-uniform", out, clk);
  run<4>("LDS.64 warp-uniform", out, clk);
  return 0;
}
