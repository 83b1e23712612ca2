// TMA load throughput probe: dense box {C, IW, IH, 1} vs channel-padded box {PS, IW, IH, 1} on an NHWC fp32 tensor.
// Every CTA (1 per SM) keeps NBUF loads in flight and cycles through distinct tiles; reports bytes/clk/SM and GB/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_bench tools/tma_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
template <int NBUF>
__global__ void tma_loop(const __grid_constant__ CUtensorMap tm, uint32_t bytes, int buf_floats, int bands, int BH, int n_tiles, long long* clk, int store,
                         const __grid_constant__ CUtensorMap tm_out, int src_off) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t bar[NBUF];
  if (threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  int issued = 0, done = 0;
  int tile = blockIdx.x;
  auto issue = [&](int t, int b) {
    const int img = t / bands, y0 = (t % bands) * BH;
    mbar_expect_tx(&bar[b], bytes);
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem + b * buf_floats)), "l"(&tm), "r"(smem_u32(&bar[b])), "r"(0), "r"(-1), "r"(y0 - 1), "r"(img) : "memory");
  };
  for (; issued < NBUF && tile < n_tiles; ++issued, tile += gridDim.x) issue(tile, issued % NBUF);
  int t_done = blockIdx.x;
  while (done < issued) {
    const int b = done % NBUF;
    mbar_wait(&bar[b], (done / NBUF) & 1);
    if (store) {
      const int img = t_done / bands, y0 = (t_done % bands) * BH;
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                   ::"l"(&tm_out), "r"(smem_u32(smem + b * buf_floats + src_off)), "r"(0), "r"(0), "r"(y0), "r"(img) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    t_done += gridDim.x;
    ++done;
    if (tile < n_tiles) { issue(tile, b); ++issued; tile += gridDim.x; }
  }
  if (store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  clk[blockIdx.x] = clock64() - t0;
}
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  PFN_encodeTiled enc = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int N = 2048, H = 48, W = 48, C = 24;
  float *x, *y; CK(cudaMalloc(&x, (size_t)N * H * W * C * 4)); CK(cudaMalloc(&y, (size_t)N * H * W * C * 4));
  CK(cudaMemset(x, 0, (size_t)N * H * W * C * 4));
  long long* dclk; CK(cudaMalloc(&dclk, 148 * 8));
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  struct Cfg { int bc, bw, BH; const char* name; };
  Cfg cfgs[] = {{24, 50, 8, "dense  {24,50,10}"}, {28, 55, 8, "padded {28,55,10}"}, {28, 50, 8, "padded {28,50,10}"}, {24, 55, 8, "dense  {24,55,10}"},
                {28, 55, 4, "padded {28,55,6}"}, {24, 50, 4, "dense  {24,50,6}"}, {32, 55, 8, "padded {32,55,10}"}};
  for (auto& c : cfgs) {
    for (int store = 0; store < 2; ++store) {
      if (store && (c.bw + 1) % 8) continue;
      CUtensorMap tm, tmo;
      cuuint32_t box[4] = {(cuuint32_t)c.bc, (cuuint32_t)c.bw, (cuuint32_t)(c.BH + 2), 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      cuuint32_t boxo[4] = {(cuuint32_t)c.bc, (cuuint32_t)c.bw, (cuuint32_t)c.BH, 1};
      CUresult r2 = enc(&tmo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y, dims, strides, boxo, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("%s: encode failed %d %d\n", c.name, (int)r, (int)r2); continue; }
      const uint32_t bytes = (uint32_t)c.bc * c.bw * (c.BH + 2) * 4;
      const int buf_floats = (bytes / 4 + 255) / 256 * 256;
      const int bands = H / c.BH, n_tiles = N * bands;
      for (int nbuf = 1; nbuf <= 3; ++nbuf) {
        if ((size_t)nbuf * buf_floats * 4 > 224 * 1024) continue;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto launch = [&]() {
          const size_t sm = (size_t)nbuf * buf_floats * 4;
          const int src_off = (c.bw + 1) * c.bc;
          if (nbuf == 1) { CK(cudaFuncSetAttribute(tma_loop<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); tma_loop<1><<<148, 32, sm>>>(tm, bytes, buf_floats, bands, c.BH, n_tiles, dclk, store, tmo, src_off); }
          if (nbuf == 2) { CK(cudaFuncSetAttribute(tma_loop<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); tma_loop<2><<<148, 32, sm>>>(tm, bytes, buf_floats, bands, c.BH, n_tiles, dclk, store, tmo, src_off); }
          if (nbuf == 3) { CK(cudaFuncSetAttribute(tma_loop<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); tma_loop<3><<<148, 32, sm>>>(tm, bytes, buf_floats, bands, c.BH, n_tiles, dclk, store, tmo, src_off); }
        };
        launch(); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long hclk[148]; CK(cudaMemcpy(hclk, dclk, sizeof(hclk), cudaMemcpyDeviceToHost));
        const double tiles_per_sm = (double)n_tiles / 148;
        const double hbm_bytes = (double)n_tiles * (c.BH + 2) * W * C * 4 * (store ? 1.0 + (double)c.BH / (c.BH + 2) : 1.0);
        printf("%s store=%d nbuf=%d: %.3f ms, %.0f clk/tile/SM, smem %.1f B/clk/SM, global %.0f GB/s\n", c.name, store, nbuf, ms, hclk[0] / tiles_per_sm,
               bytes * tiles_per_sm / hclk[0], hbm_bytes / ms * 1e-6);
      }
    }
  }
  return 0;
}
