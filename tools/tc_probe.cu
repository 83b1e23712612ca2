// Probe for the tcgen05 building blocks used by the tensor-core BlazeBlock kernel (run on a B200):
//   1. TS-mode tf32 MMA (A in TMEM written with tcgen05.st, B in smem, K-major no-swizzle descriptor),
//      3xTF32 split, commit -> mbarrier, tcgen05.ld of the accumulator.
//   2. TMA load with a box wider than the tensor's channel extent (padded pixel stride in smem) and negative
//      start coordinates; TMA store from the same padded layout (out-of-bounds elements clipped).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 2000000000ll) { printf("mbar timeout\n"); __trap(); }
  }
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// ---------------------------------------------------------------------------- test 1: TS-mode 3xTF32 GEMM
// D[128][N] = A[128][K] * W[K][N];  B operand in smem as [K/4][N][4] (K-major core matrices, LBO = N*16, SBO = 128)
template <int K, int N, int NPROD>
__global__ void __launch_bounds__(160) gemm_probe(const float* __restrict__ A, const float* __restrict__ Bhi,
                                                  const float* __restrict__ Blo, float* __restrict__ D) {
  extern __shared__ __align__(128) float smem[];
  float* s_bhi = smem;
  float* s_blo = smem + K * N;
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < K * N; i += blockDim.x) { s_bhi[i] = Bhi[i]; s_blo[i] = Blo[i]; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(&bar_full, 128);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t col_d = 0, col_ahi = 128, col_alo = 128 + K;   // D: N cols; A hi / lo: K cols each

  if (warp < 4) {
    const int row = tid;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int k = 0; k < K; k += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = A[row * K + k + j];
        hi[j] = to_tf32(a);
        lo[j] = to_tf32(a - __uint_as_float(hi[j]));
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + col_ahi + k),
                   "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + col_alo + k),
                   "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(&bar_full);
    // epilogue
    mbar_wait(&bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int n = 0; n < N; n += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(lane_addr + col_d + n));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; ++j) D[row * N + n + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (lane == 0) {
    mbar_wait(&bar_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto bdesc = [&](const float* base, int kstep) -> uint64_t {
      const uint32_t addr = smem_u32(base) + (uint32_t)kstep * 2u * N * 16u;
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)(((uint32_t)(N * 16) >> 4) & 0x3FFF) << 16;   // LBO: next K chunk (4 floats) of the same rows
      d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;                 // SBO: next group of 8 rows
      d |= (uint64_t)1 << 46;                                      // descriptor version (Blackwell)
      return d;
    };
    auto mma = [&](uint32_t d_addr, uint32_t a_addr, uint64_t bd, uint32_t acc) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(d_addr), "r"(a_addr), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    };
    for (int ks = 0; ks < K / 8; ++ks) {
      mma(tbase + col_d, tbase + col_ahi + ks * 8, bdesc(s_bhi, ks), ks > 0 ? 1u : 0u);
      if (NPROD >= 2) mma(tbase + col_d, tbase + col_ahi + ks * 8, bdesc(s_blo, ks), 1u);
      if (NPROD >= 3) mma(tbase + col_d, tbase + col_alo + ks * 8, bdesc(s_bhi, ks), 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512));
  }
}

template <int K, int N>
static int run_gemm() {
  std::vector<float> A(128 * K), W(K * N), Bhi(K * N), Blo(K * N), D(128 * N);
  srand(1234 + K * 7 + N);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : W) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  auto tf32_rna = [](float x) {
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x1000u; u &= 0xFFFFE000u;
    float r; memcpy(&r, &u, 4); return r;
  };
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) {
      const float w = W[k * N + n];
      const float hi = tf32_rna(w), lo = tf32_rna(w - hi);
      Bhi[((k / 4) * N + n) * 4 + (k % 4)] = hi;
      Blo[((k / 4) * N + n) * 4 + (k % 4)] = lo;
    }
  float *dA, *dBhi, *dBlo, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dBhi, Bhi.size() * 4)); CK(cudaMalloc(&dBlo, Blo.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBhi, Bhi.data(), Bhi.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBlo, Blo.data(), Blo.size() * 4, cudaMemcpyHostToDevice));
  int bad = 0;
  CK(cudaFuncSetAttribute(gemm_probe<K, N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * K * N * 4));
  CK(cudaFuncSetAttribute(gemm_probe<K, N, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * K * N * 4));
  for (int nprod = 1; nprod <= 3; nprod += 2) {
    CK(cudaMemset(dD, 0, D.size() * 4));
    if (nprod == 1) gemm_probe<K, N, 1><<<1, 160, 2 * K * N * 4>>>(dA, dBhi, dBlo, dD);
    else gemm_probe<K, N, 3><<<1, 160, 2 * K * N * 4>>>(dA, dBhi, dBlo, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)W[k * N + n];
        maxerr = fmax(maxerr, fabs(ref - (double)D[m * N + n]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("gemm K=%d N=%d products=%d: max abs err %.3e (max |ref| %.3f) rel %.3e\n", K, N, nprod, maxerr, maxref, maxerr / maxref);
    if (nprod == 3 && maxerr / maxref > 2e-6) bad = 1;
    if (nprod == 1 && maxerr / maxref > 5e-3) bad = 1;
  }
  cudaFree(dA); cudaFree(dBhi); cudaFree(dBlo); cudaFree(dD);
  return bad;
}

// ---------------------------------------------------------------------------- test 2: TMA padded box
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void tma_probe(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, float* dump,
                          int box_floats, int img, int y0, int store_row_off_floats, int store_x0) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, box_floats * 4);
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(&tm_in), "r"(smem_u32(&bar)), "r"(0), "r"(-1), "r"(y0 - 1), "r"(img) : "memory");
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < box_floats; i += blockDim.x) dump[i] = smem[i];
  // modify in place: v -> v + 1000 on every float of the tile (also halo / pad positions), then store rows 1..4
  for (int i = threadIdx.x; i < box_floats; i += blockDim.x) smem[i] += 1000.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && store_row_off_floats >= 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(&tm_out), "r"(smem_u32(smem + store_row_off_floats)), "r"(0), "r"(store_x0), "r"(y0), "r"(img) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// variant 0: load only; 1: + store from a 128B-aligned row offset (box width 16); 2: + store from an unaligned row offset (box width 10)
static int run_tma(int variant) {
  const int N = 2, H = 8, W = 8, C = 24, PS = 28, TR = 4, IW = ((variant == 1 || variant >= 5) ? 16 : W + 2), IH = TR + 2;
  const int PSO = (variant == 6 ? C : PS);   // variant 5: padded box, x0 = 0 (source shifted by one pixel is unaligned -> use x0=0 from col 0: data shifted); 6: dense box C, x0=-1
  PFN_encodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  std::vector<float> x(N * H * W * C);
  for (size_t i = 0; i < x.size(); ++i) x[i] = (float)(i % 977) + 1.f;
  float *dx, *dy, *ddump;
  CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dy, x.size() * 4)); CK(cudaMalloc(&ddump, IH * IW * PS * 4));
  CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy, 0, x.size() * 4));
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tin, tout;
  cuuint32_t box_in[4] = {(cuuint32_t)PS, (cuuint32_t)IW, (cuuint32_t)IH, 1};
  CUresult r = enc(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, dims, strides, box_in, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode padded load box {%d,%d,%d,1} on dims {%d,%d,%d,%d}: CUresult %d\n", PS, IW, IH, C, W, H, N, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  cuuint32_t box_out[4] = {(cuuint32_t)PSO, (cuuint32_t)IW, (cuuint32_t)TR, 1};
  r = enc(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dy, dims, strides, box_out, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode padded store box: CUresult %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  int bad = 0;
  for (int y0 = 0; y0 <= 4; y0 += 4) {
    const int img = 1;
    CK(cudaMemset(dy, 0, x.size() * 4));
    tma_probe<<<1, 128, IH * IW * PS * 4>>>(tin, tout, ddump, IH * IW * PS, img, y0, variant == 0 ? -1 : IW * PS, variant == 5 ? 0 : -1);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> dump(IH * IW * PS), y(x.size());
    CK(cudaMemcpy(dump.data(), ddump, dump.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(y.data(), dy, y.size() * 4, cudaMemcpyDeviceToHost));
    int e1 = 0, e2 = 0;
    for (int hy = 0; hy < IH; ++hy)
      for (int hx = 0; hx < IW; ++hx)
        for (int c = 0; c < PS; ++c) {
          const int gy = y0 - 1 + hy, gx = hx - 1;
          float want = 0.f;
          if (c < C && gy >= 0 && gy < H && gx >= 0 && gx < W) want = x[((img * H + gy) * W + gx) * C + c];
          if (dump[(hy * IW + hx) * PS + c] != want) ++e1;
        }
    for (int n = 0; n < N; ++n)
      for (int gy = 0; gy < H; ++gy)
        for (int gx = 0; gx < W; ++gx)
          for (int c = 0; c < C; ++c) {
            float want = 0.f;
            if (variant > 0 && n == img && gy >= y0 && gy < y0 + TR) want = x[((n * H + gy) * W + gx) * C + c] + 1000.f;
            if (y[((n * H + gy) * W + gx) * C + c] != want) ++e2;
          }
    printf("tma padded box y0=%d: load mismatches %d, store mismatches %d\n", y0, e1, e2);
    bad |= (e1 || e2);
  }
  return bad;
}

int main(int argc, char** argv) {
  int bad = 0;
  const int test = argc > 1 ? atoi(argv[1]) : 0;
  if (test <= 2 || test == 5 || test == 6) bad |= run_tma(test);
  if (test == 3) bad |= run_gemm<24, 32>();
  if (test == 4) { bad |= run_gemm<32, 48>(); bad |= run_gemm<96, 96>(); bad |= run_gemm<40, 16>(); }
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
