// Regressor heads: forward, backward, loss and optimizers for the small fully-convolutional
// yaw/pitch/roll regressors, executed as a register program (hp_head_op) compiled by the Python
// layer from the Keras graphs of the reference:
//   Model-88/attention_model.py:16-80   se_transformer_regr_head (SE gate, MHA, LayerNorm, FF, 1x1 convs)
//   Model-88/attention_model.py:82-169  create_modelC, create_model_complex
//   Model-88/train_88.py:66-253         create_model, create_model_skip_fc, bestmodelV1
//   Model-96/train_96.py:65-110         create_model
// Training semantics (loss, L2, SpatialDropout2D, SGD/Adam/Adamax) follow Keras 2.13 as listed in
// SURVEY.md Appendix B.4-B.5; the step replaces the body of model.fit (train_96.py:175, train_88.py:355).
//
// All tensors are [rows][channels] float32 with rows = images*tokens (or images for per-image
// registers).  Shapes are tiny (<= 128 channels), so kernels are CUDA-core code; the data-parallel
// gradient exchange is one flat-buffer allreduce (comm.cu).
#include <math.h>

#include <algorithm>

#include "common.cuh"

// ============================================================================ activations
__device__ __forceinline__ float act_fwd(int act, float v) { return hp_act_rt(act, v); }
// derivative expressed with the OUTPUT y of the activation (swish has none: rejected when a training program is planned)
__device__ __forceinline__ float act_bwd(int act, float y) {
  switch (act) {
    case HP_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case HP_ACT_TANH: return 1.f - y * y;
    case HP_ACT_SIGMOID: return y * (1.f - y);
    case HP_ACT_SOFTSIGN: {
      float t = 1.f - fabsf(y);
      return t * t;
    }
    case HP_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case HP_ACT_SELU: return y > 0.f ? HP_SELU_SCALE : y + HP_SELU_SCALE * HP_SELU_ALPHA;
    case HP_ACT_SOFTPLUS: return 1.f - expf(-y);                  // sigmoid(x) with y = log(1 + e^x)
    case HP_ACT_LEAKY_RELU: return y > 0.f ? 1.f : 0.2f;
    default: return 1.f;
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ============================================================================ generic dense
// y = act(x W + b) for row-major x [M][K]: persistent CTAs, 128-row tiles.
//   x tile   : cp.async 16-byte chunks (coalesced) into smem rows of XS floats (XS / 4 odd: conflict-free row access)
//   compute  : register tile 8 rows x 4 columns per thread (weights as conflict-free LDS.128, activations as broadcast
//              LDS.128); layers with fewer than 16 output columns use one thread per row instead
//   epilogue : bias + activation into a staging tile that reuses the x tile, then coalesced stores per output segment
//              (the detector heads scatter cls | loc columns of one GEMM to two tensors in anchor order)
#define DENSE_TM 128
struct DenseKParams {
  const float* x;
  const float* W;
  const float* b;
  int M, K, ldx, ldw, N, act, transpose_w, accumulate;
  int KP, NP, XS, OS;        // OS = staging row stride
  int vec_x;                 // x rows can be fetched with 16-byte cp.async
  int n_outs;
  DenseOut outs[2];
  unsigned magic[2];         // ceil(2^32 / width) of each output segment
};

// c += a * w with packed fp32 FMAs (fma.rn.f32x2): same rounding as scalar FMAs, half the FMA issue slots
__device__ __forceinline__ float4 dense_fma4s(float a, float4 w, float4 c) {
  const float2 aa = make_float2(a, a);
  const float2 lo = __ffma2_rn(aa, make_float2(w.x, w.y), make_float2(c.x, c.y));
  const float2 hi = __ffma2_rn(aa, make_float2(w.z, w.w), make_float2(c.z, c.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ void dense_cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}

template <int NPASS>   // passes over the (row group, column group) items: 1 for NP <= 64, 2 up to NP <= 128
__global__ void __launch_bounds__(256, NPASS == 1 ? 3 : 1) dense_kernel(DenseKParams p) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                 // [KP][NP]
  float* bs = Ws + p.KP * p.NP;     // [NP]
  long long* rowoff = reinterpret_cast<long long*>(bs + p.NP);   // [2][TM] destination offset of every tile row per segment
  float* xs = bs + p.NP + 4 * DENSE_TM;   // [TM][XS]; reused as the staging tile [TM][OS]
  const int tid = threadIdx.x;
  for (int i = tid; i < p.KP * p.NP; i += 256) {
    const int k = i / p.NP, n = i - k * p.NP;
    float v = 0.f;
    if (k < p.K && n < p.N) v = p.transpose_w ? p.W[(long long)n * p.ldw + k] : p.W[(long long)k * p.ldw + n];
    Ws[i] = v;
  }
  for (int i = tid; i < p.NP; i += 256) bs[i] = (p.b && i < p.N) ? p.b[i] : 0.f;
  const int ngd = p.NP / 4;
  constexpr int RG = DENSE_TM / 8;  // 16 row groups; a thread's 8 rows are rg + 16 r
  const int n_tiles = (p.M + DENSE_TM - 1) / DENSE_TM;
  const int kc = p.KP / 4;          // 16-byte chunks per row
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long m0 = (long long)tile * DENSE_TM;
    const int rows = (int)((p.M - m0 < DENSE_TM) ? (p.M - m0) : DENSE_TM);
    __syncthreads();   // previous tile fully written out (xs doubles as the staging tile)
    for (int i = tid; i < p.n_outs * DENSE_TM; i += 256) {   // one 64-bit division per row and segment, not per element
      const int o = i / DENSE_TM, r = i - o * DENSE_TM;
      const DenseOut& d = p.outs[o];
      const long long m = m0 + r;
      const long long img = m / d.rows_per_img;
      rowoff[i] = img * d.img_stride + (m - img * d.rows_per_img) * (long long)d.row_stride;
    }
    if (p.vec_x) {
      for (int i = tid; i < DENSE_TM * kc; i += 256) {
        const int r = i / kc, c = i - r * kc;
        const bool ok = r < rows;
        dense_cp_async16(xs + r * p.XS + c * 4, ok ? p.x + (m0 + r) * p.ldx + c * 4 : p.x, ok ? 16 : 0);
      }
      asm volatile("cp.async.commit_group;\n" ::);
      asm volatile("cp.async.wait_group 0;\n" ::);
    } else {
      for (int i = tid; i < DENSE_TM * p.KP; i += 256) {
        const int r = i / p.KP, k = i - r * p.KP;
        xs[r * p.XS + k] = (r < rows && k < p.K) ? p.x[(m0 + r) * p.ldx + k] : 0.f;
      }
    }
    __syncthreads();
    if (p.NP >= 16) {
      // ---- 8 rows x 4 columns per thread; up to two passes over the (row group, column group) items
      float4 acc[NPASS][8];
#pragma unroll
      for (int ps = 0; ps < NPASS; ++ps) {
        const int item = tid + ps * 256;
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[ps][r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < RG * ngd) {
          const int cg = item % ngd, rg = item / ngd;
          const float* arow = xs + rg * p.XS;
          const float* wcol = Ws + cg * 4;
          const int rstep = RG * p.XS, np = p.NP;
#pragma unroll 2
          for (int k = 0; k < p.KP; k += 4, arow += 4, wcol += 4 * np) {
            const float4 w0 = ld4(wcol);
            const float4 w1 = ld4(wcol + np);
            const float4 w2 = ld4(wcol + 2 * np);
            const float4 w3 = ld4(wcol + 3 * np);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const float4 a = ld4(arow + r * rstep);
              float4& c = acc[ps][r];
              c = dense_fma4s(a.x, w0, c);
              c = dense_fma4s(a.y, w1, c);
              c = dense_fma4s(a.z, w2, c);
              c = dense_fma4s(a.w, w3, c);
            }
          }
        }
      }
      __syncthreads();   // every thread is done reading the x tile: it becomes the staging tile
#pragma unroll
      for (int ps = 0; ps < NPASS; ++ps) {
        const int item = tid + ps * 256;
        if (item < RG * ngd) {
          const int cg = item % ngd, rg = item / ngd;
          const float4 bias = ld4(bs + cg * 4);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float4 c = acc[ps][r];
            // the activation is applied at write-out (one copy of the switch instead of 32 per pass: the unrolled
            // epilogue overflowed the instruction cache, 15 % no_instruction stalls in ncu)
            *reinterpret_cast<float4*>(xs + (rg + r * RG) * p.OS + cg * 4) = make_float4(c.x + bias.x, c.y + bias.y, c.z + bias.z, c.w + bias.w);
          }
        }
      }
    } else {
      // ---- narrow layers (NP = 4, 8, 12): one thread per row
      float acc[12];
      if (tid < DENSE_TM) {
#pragma unroll
        for (int c = 0; c < 12; ++c) acc[c] = 0.f;
        const float* arow = xs + tid * p.XS;
        for (int k = 0; k < p.KP; k += 4) {
          const float4 a = ld4(arow + k);
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4)
              if (c4 * 4 < p.NP) {
                const float4 w = ld4(Ws + (k + kk) * p.NP + c4 * 4);
                acc[c4 * 4 + 0] = fmaf(av[kk], w.x, acc[c4 * 4 + 0]); acc[c4 * 4 + 1] = fmaf(av[kk], w.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = fmaf(av[kk], w.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(av[kk], w.w, acc[c4 * 4 + 3]);
              }
        }
      }
      __syncthreads();
      if (tid < DENSE_TM) {
#pragma unroll
        for (int c = 0; c < 12; ++c)
          if (c < p.NP) xs[tid * p.OS + c] = acc[c] + bs[c];
      }
    }
    __syncthreads();
    // ---- coalesced write-out of every output segment
    for (int o = 0; o < p.n_outs; ++o) {
      const DenseOut& d = p.outs[o];
      const int wd = d.col_end - d.col_begin;
      const int total = rows * wd;
      for (int j = tid; j < total; j += 256) {
        const int r = (int)(((unsigned long long)j * p.magic[o]) >> 32);
        const int c = j - r * wd;
        float* dst = d.ptr + rowoff[o * DENSE_TM + r] + c;
        const float val = act_fwd(p.act, xs[r * p.OS + d.col_begin + c]);
        if (p.accumulate) *dst += val; else *dst = val;
      }
    }
  }
}

int hp_launch_dense(hp_ctx* h, const float* x, int M, int K, int ldx, const float* W, int ldw, const float* b,
                    int N, int act, bool transpose_w, const DenseOut* outs, int n_outs, bool accumulate,
                    cudaStream_t st) {
  if (M <= 0) return HP_OK;
  HP_REQUIRE(n_outs >= 1 && n_outs <= 2, HP_ERR_INVALID, "dense: 1 or 2 output segments");
  if (h->impl == HP_IMPL_FAST && h->dense_tc && hp_dense_tc_supported(x, M, K, ldx, N, transpose_w, accumulate))
    return hp_launch_dense_tc(h, x, M, K, ldx, W, ldw, b, N, act, outs, n_outs, st);
  // wide layers whose split weights + staging tile exceed shared memory (the 96 -> 102 detector head of the 8 x 8 map):
  // two tensor-core launches over column halves (x is read twice, the second time mostly from L2)
  if (h->impl == HP_IMPL_FAST && h->dense_tc && !transpose_w && !accumulate && N > 64) {
    const int c_mid = round_up((N + 1) / 2, 4);
    if (hp_dense_tc_supported(x, M, K, ldx, c_mid, false, false) && hp_dense_tc_supported(x, M, K, ldx, N - c_mid, false, false)) {
      for (int half = 0; half < 2; ++half) {
        const int c0 = half ? c_mid : 0, c1 = half ? N : c_mid;
        DenseOut seg[2];
        int ns = 0;
        for (int i = 0; i < n_outs; ++i) {
          const int cb = outs[i].col_begin > c0 ? outs[i].col_begin : c0, ce = outs[i].col_end < c1 ? outs[i].col_end : c1;
          if (cb >= ce) continue;
          seg[ns] = outs[i];
          seg[ns].ptr = outs[i].ptr + (cb - outs[i].col_begin);
          seg[ns].col_begin = cb - c0;
          seg[ns].col_end = ce - c0;
          ++ns;
        }
        if (ns > 0) HP_TRY(hp_launch_dense_tc(h, x, M, K, ldx, W + c0, ldw, b ? b + c0 : nullptr, c1 - c0, act, seg, ns, st));
      }
      return HP_OK;
    }
  }
  DenseKParams p;
  p.x = x; p.W = W; p.b = b; p.M = M; p.K = K; p.ldx = ldx; p.ldw = ldw; p.N = N; p.act = act;
  p.transpose_w = transpose_w ? 1 : 0; p.accumulate = accumulate ? 1 : 0;
  p.KP = round_up(K, 4); p.NP = round_up(N, 4);
  p.XS = ((p.KP / 4) & 1) ? p.KP : p.KP + 4;
  p.OS = ((p.NP / 4) & 1) ? p.NP : p.NP + 4;
  p.vec_x = (K % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)x & 15) == 0) ? 1 : 0;
  p.n_outs = n_outs;
  for (int i = 0; i < n_outs; ++i) {
    p.outs[i] = outs[i];
    const int wd = outs[i].col_end - outs[i].col_begin;
    HP_REQUIRE(wd >= 1 && outs[i].col_end <= N, HP_ERR_INVALID, "dense: bad output segment [%d, %d)", outs[i].col_begin, outs[i].col_end);
    p.magic[i] = (unsigned)((0x100000000ull + wd - 1) / wd);
  }
  const int items = (DENSE_TM / 8) * (p.NP / 4);
  HP_REQUIRE(p.NP >= 16 ? items <= 512 : p.NP <= 12, HP_ERR_UNSUPPORTED, "dense layer with %d outputs too wide", N);
  const bool two_pass = p.NP >= 16 && items > 256;
  const int tile_fl = DENSE_TM * (p.XS > p.OS ? p.XS : p.OS);
  size_t smem = ((size_t)p.KP * p.NP + p.NP + 4 * DENSE_TM + (size_t)tile_fl) * sizeof(float);   // NP % 4 == 0 keeps rowoff 16-byte aligned
  HP_REQUIRE(smem <= 200 * 1024, HP_ERR_UNSUPPORTED, "dense layer %dx%d too large for the head engine", K, N);
  HP_CUDA(cudaFuncSetAttribute(dense_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  HP_CUDA(cudaFuncSetAttribute(dense_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int n_tiles = ceil_div(M, DENSE_TM);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  long long grid = (long long)h->num_sms * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  if (two_pass) dense_kernel<2><<<(unsigned)grid, 256, smem, st>>>(p);
  else dense_kernel<1><<<(unsigned)grid, 256, smem, st>>>(p);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

// gW[K][N] += x^T gz ; gb[N] += sum_rows gz       (gW/gb must be zeroed by the caller)
#define WG_TM 32
__global__ void __launch_bounds__(256) dense_wgrad_kernel(const float* __restrict__ x, int ldx,
                                                          const float* __restrict__ gz, int ldg, float* gW, float* gb,
                                                          int M, int K, int N) {
  extern __shared__ float smem[];
  float* xs = smem;            // [WG_TM][K]
  float* gs = xs + WG_TM * K;  // [WG_TM][N]
  const long long m0 = (long long)blockIdx.x * WG_TM;
  const int rows = (int)((M - m0 < WG_TM) ? (M - m0) : WG_TM);
  for (int i = threadIdx.x; i < WG_TM * K; i += 256) {
    const int r = i / K, k = i - r * K;
    xs[i] = (r < rows) ? x[(m0 + r) * ldx + k] : 0.f;
  }
  for (int i = threadIdx.x; i < WG_TM * N; i += 256) {
    const int r = i / N, n = i - r * N;
    gs[i] = (r < rows) ? gz[(m0 + r) * ldg + n] : 0.f;
  }
  __syncthreads();
  const int total = K * N + N;
  for (int e = threadIdx.x; e < total; e += 256) {
    float acc = 0.f;
    if (e < K * N) {
      const int k = e / N, n = e - k * N;
      for (int r = 0; r < WG_TM; ++r) acc = fmaf(xs[r * K + k], gs[r * N + n], acc);
      atomicAdd(gW + e, acc);
    } else if (gb) {
      const int n = e - K * N;
      for (int r = 0; r < WG_TM; ++r) acc += gs[r * N + n];
      atomicAdd(gb + n, acc);
    }
  }
}

// ============================================================================ elementwise kernels
#define EW_GRID(total) (unsigned)(((total) + 255) / 256 > 65535 * 16 ? 65535 * 16 : ((total) + 255) / 256)

__global__ void act_fwd_kernel(const float* in, float* out, long long total, int act) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) out[i] = act_fwd(act, in[i]);
}
// g_in += g_out * act'(y)
__global__ void act_bwd_acc_kernel(const float* y, const float* gy, float* gx, long long total, int act) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll)
    gx[i] += gy[i] * act_bwd(act, y[i]);
}
// gz = gy * act'(y)  (in place allowed)
__global__ void act_bwd_kernel(const float* y, const float* gy, float* gz, long long total, int act) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll)
    gz[i] = gy[i] * act_bwd(act, y[i]);
}
// out = a + b with row broadcasting (rows_a/rows_b are either `rows` or rows/T)
__global__ void add_kernel(const float* a, const float* b, float* out, long long rows, int C, int T, int a_img, int b_img) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float va = a[(a_img ? r / T : r) * C + c];
    const float vb = b[(b_img ? r / T : r) * C + c];
    out[i] = va + vb;
  }
}
// same-shape fast path of add_kernel: 128-bit accesses, no index arithmetic
__global__ void add_vec4_kernel(const float4* a, const float4* b, float4* out, long long total4) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total4; i += gridDim.x * 256ll) {
    const float4 x = a[i], y = b[i];
    out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}
// out[n,t,c] = x[n,t,c] * g[n,c] for C % 4 == 0 and fewer than 2^31 chunks: 128-bit accesses, 32-bit index arithmetic
__global__ void mulch_vec4_kernel(const float4* x, const float4* g, float4* out, unsigned total4, unsigned C4, unsigned T) {
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total4; i += gridDim.x * 256u) {
    const unsigned r = i / C4, c = i - r * C4;
    const float4 v = x[i], w = g[(r / T) * C4 + c];
    out[i] = make_float4(v.x * w.x, v.y * w.y, v.z * w.z, v.w * w.w);
  }
}
__global__ void acc_kernel(float* dst, const float* src, long long total) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) dst[i] += src[i];
}
// out[n,t,c] = x[n,t,c] * g[n,c]
__global__ void mulch_kernel(const float* x, const float* g, float* out, long long rows, int C, int T) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    out[i] = x[i] * g[(r / T) * C + c];
  }
}
// gx += gy * g ; gg[n,c] += sum_t gy*x          one thread per (n,c)
__global__ void mulch_bwd_kernel(const float* x, const float* g, const float* gy, float* gx, float* gg, int n_img, int C,
                                 int T) {
  const long long total = (long long)n_img * C;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const long long n = i / C;
    const int c = (int)(i - n * C);
    const float gv = g[i];
    float s = 0.f;
    for (int t = 0; t < T; ++t) {
      const long long j = (n * T + t) * C + c;
      const float go = gy[j];
      if (gx) gx[j] += go * gv;
      s = fmaf(go, x[j], s);
    }
    gg[i] += s;
  }
}
__global__ void gap_kernel(const float* x, float* out, int n_img, int C, int T) {
  const long long total = (long long)n_img * C;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const long long n = i / C;
    const int c = (int)(i - n * C);
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += x[(n * T + t) * C + c];
    out[i] = s / (float)T;
  }
}
__global__ void gap_bwd_kernel(const float* gy, float* gx, long long rows, int C, int T) {
  const long long total = rows * C;
  const float inv = 1.f / (float)T;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    gx[i] += gy[(r / T) * C + c] * inv;
  }
}

__host__ __device__ inline uint32_t dropout_hash(uint64_t seed, uint32_t step, uint32_t op_id, uint32_t image,
                                                 uint32_t channel) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(step + 1);
  z ^= 0xBF58476D1CE4E5B9ull * (uint64_t)(op_id + 1);
  z += (((uint64_t)image) << 32 | (uint64_t)channel) * 0x94D049BB133111EBull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
__host__ __device__ inline bool dropout_keep(uint32_t u, float rate) {
  return (float)(u >> 8) * (1.0f / 16777216.0f) >= rate;
}
// mode 0: out = x*m ; mode 1: gx += gy*m  (m = keep/(1-rate), mask over (image, channel))
// step_dev != nullptr: the step index is *step_dev - 1 (device-resident counter of hp_head_train_run, already advanced for
// this step), so that a captured step can be replayed
__global__ void dropout_kernel(const float* in, float* out, long long rows, int C, int T, int per_image, float rate,
                               uint64_t seed, uint32_t step, uint32_t op_id, int mode, const uint32_t* step_dev = nullptr) {
  if (step_dev) step = *step_dev - 1u;
  const long long total = rows * C;
  const float scale = 1.f / (1.f - rate);
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const uint32_t img = (uint32_t)(per_image ? r : r / T);
    const float m = dropout_keep(dropout_hash(seed, step, op_id, img, (uint32_t)c), rate) ? scale : 0.f;
    if (mode == 0) out[i] = in[i] * m; else out[i] += in[i] * m;
  }
}

// ============================================================================ LayerNorm (warp per row)
__global__ void layernorm_kernel(const float* x, const float* gamma, const float* beta, float* out, float* stats,
                                 long long rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = xr[c] - mu;
      v = fmaf(d, d, v);
    }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.f / sqrtf(v / (float)C + eps);
    for (int c = lane; c < C; c += 32) out[r * C + c] = (xr[c] - mu) * rstd * gamma[c] + beta[c];
    if (stats && lane == 0) {
      stats[2 * r] = mu;
      stats[2 * r + 1] = rstd;
    }
  }
}
// Inference LayerNorm with the residual add in front of it fused in: out = LN(x + x2) (x2 may be null).  8 lanes per row, V
// float4 per lane (C = 32 V), the row lives in registers between the two reductions: one read of each input, one write
// (the warp-per-row kernel above re-reads the row three times with scalar loads; add + LayerNorm were 22 + 46 us per
// residual block at batch 4096 against 26 us of HBM time).
template <int V>
__global__ void __launch_bounds__(256) layernorm_add_rows_kernel(const float4* __restrict__ x, const float4* __restrict__ x2,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 float4* __restrict__ out, long long rows, float eps) {
  const int sub = threadIdx.x & 7;
  const long long r0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3;
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 3;
  constexpr int C4 = 8 * V;
  constexpr float invC = 1.f / (float)(4 * C4);
  const long long rows_pad = (rows + stride - 1) / stride * stride;            // whole warps stay in the loop for the shuffles
  for (long long r = r0; r < rows_pad; r += stride) {
    const bool ok = r < rows;
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = ok ? x[r * C4 + sub + 8 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok && x2) {
        const float4 w = x2[r * C4 + sub + 8 * i];
        v[i].x += w.x; v[i].y += w.y; v[i].z += w.z; v[i].w += w.w;
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    for (int o = 4; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
      q = fmaf(v[i].x, v[i].x, q); q = fmaf(v[i].y, v[i].y, q); q = fmaf(v[i].z, v[i].z, q); q = fmaf(v[i].w, v[i].w, q);
    }
    for (int o = 4; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.f / sqrtf(q * invC + eps);
    if (ok) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float* g = gamma + 4 * (sub + 8 * i);          // the parameter vector is not 16-byte aligned in general
        const float* b = beta + 4 * (sub + 8 * i);
        out[r * C4 + sub + 8 * i] = make_float4(fmaf(v[i].x * rstd, g[0], b[0]), fmaf(v[i].y * rstd, g[1], b[1]), fmaf(v[i].z * rstd, g[2], b[2]),
                                                fmaf(v[i].w * rstd, g[3], b[3]));
      }
    }
  }
}
// true when the fast kernel applies: C a multiple of 32 up to 128, everything 16-byte aligned
static bool layernorm_rows_ok(int C, const void* a, const void* b, const void* c) {
  return C % 32 == 0 && C >= 32 && C <= 128 && ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c)) & 15) == 0;
}
static void launch_layernorm_rows(hp_ctx* h, const float* x, const float* x2, const float* gamma, const float* beta, float* out, long long rows,
                                  int C, float eps, cudaStream_t st) {
  const long long blocks = std::min<long long>((rows * 8 + 255) / 256, 148 * 16);
  const float4 *x4 = (const float4*)x, *y4 = (const float4*)x2;
  const float *g4 = gamma, *b4 = beta;
  switch (C / 32) {
    case 1: layernorm_add_rows_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(x4, y4, g4, b4, (float4*)out, rows, eps); break;
    case 2: layernorm_add_rows_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(x4, y4, g4, b4, (float4*)out, rows, eps); break;
    case 3: layernorm_add_rows_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(x4, y4, g4, b4, (float4*)out, rows, eps); break;
    default: layernorm_add_rows_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(x4, y4, g4, b4, (float4*)out, rows, eps); break;
  }
  h->launches++;
}

// gx += rstd*(gh - mean(gh) - xh*mean(gh*xh)), gh = gy*gamma ; ggamma += gy*xh ; gbeta += gy (atomics)
__global__ void layernorm_bwd_kernel(const float* x, const float* gamma, const float* stats, const float* gy, float* gx,
                                     float* ggamma, float* gbeta, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float mu = stats[2 * r], rstd = stats[2 * r + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (x[r * C + c] - mu) * rstd;
      const float gh = gy[r * C + c] * gamma[c];
      s1 += gh;
      s2 = fmaf(gh, xh, s2);
    }
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / (float)C, m2 = s2 / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (x[r * C + c] - mu) * rstd;
      const float g = gy[r * C + c];
      const float gh = g * gamma[c];
      if (gx) gx[r * C + c] += rstd * (gh - m1 - xh * m2);
      atomicAdd(ggamma + c, g * xh);
      atomicAdd(gbeta + c, g);
    }
  }
}

// ============================================================================ attention core
// Q,K,V,O: [n_img*T][HD] with head h at columns [h*d, h*d+d).  One CTA per (image, head), 4 warps,
// warp per query row.  scale = 1/sqrt(d) is applied to Q (Keras MultiHeadAttention).
__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* Q, const float* K, const float* V, float* O, int T,
                                                       int heads, int d, float scale, float* Pout, float* unused) {
  extern __shared__ float smem[];
  const int DS = d + 1;
  float* Ks = smem;                  // [T][DS]
  float* Vs = Ks + T * DS;           // [T][DS]
  float* sc = Vs + T * DS;           // [4][T] probabilities of the row a warp is working on
  float* qs = sc + 4 * T;            // [4][d] scaled query row of each warp
  const int img = blockIdx.x / heads, hh = blockIdx.x - img * heads;
  const int HD = heads * d;
  const long long base = (long long)img * T * HD + hh * d;
  for (int i = threadIdx.x; i < T * d; i += 128) {
    const int s = i / d, j = i - s * d;
    Ks[s * DS + j] = K[base + (long long)s * HD + j];
    Vs[s * DS + j] = V[base + (long long)s * HD + j];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my = sc + warp * T;
  float* myq = qs + warp * d;
  // P V: lane = (half, j): output column j (j < d <= 16 uses two halves of the key range, else one pass per 32 columns)
  const bool split = (d <= 16);
  for (int t = warp; t < T; t += 4) {
    for (int j = lane; j < d; j += 32) myq[j] = Q[base + (long long)t * HD + j] * scale;   // the query row is read once
    __syncwarp();
    float mx = -INFINITY;
    for (int s = lane; s < T; s += 32) {
      float a = 0.f;
      const float* kr = Ks + s * DS;
      for (int j = 0; j < d; ++j) a = fmaf(myq[j], kr[j], a);
      my[s] = a;
      mx = fmaxf(mx, a);
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int s = lane; s < T; s += 32) {
      const float e = expf(my[s] - mx);
      my[s] = e;
      sum += e;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    __syncwarp();
    if (Pout) {
      float* pr = Pout + ((long long)blockIdx.x * T + t) * T;
      for (int s = lane; s < T; s += 32) pr[s] = my[s] * inv;
    }
    if (split) {
      const int j = lane & 15, half = lane >> 4;
      const int s0 = half ? (T + 1) / 2 : 0, s1 = half ? T : (T + 1) / 2;
      float a = 0.f;
      if (j < d)
        for (int s = s0; s < s1; ++s) a = fmaf(my[s], Vs[s * DS + j], a);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      if (half == 0 && j < d) O[base + (long long)t * HD + j] = a * inv;
    } else {
      for (int j = lane; j < d; j += 32) {
        float a = 0.f;
        for (int s = 0; s < T; ++s) a = fmaf(my[s], Vs[s * DS + j], a);
        O[base + (long long)t * HD + j] = a * inv;
      }
    }
    __syncwarp();
  }
}

// Inference variant (no probabilities kept): one CTA per `ipc` images, K and V of all heads staged once in shared memory with
// coalesced float4 loads, one THREAD per (image, head, query row).  The thread keeps its scaled query row and the output
// row in registers and walks the keys: K / V rows are warp-broadcast LDS.128, the scores of the row live in a
// conflict-free shared-memory column ([s][thread]), so there are no shuffles and no per-row warp reductions.  With 36
// tokens (6 x 6 map) the warp-per-row kernel above used 36 of 64 lane slots in the score loops and 16 of 32 in P V:
// 0.33 ms at batch 4096 against 0.15 GB of traffic.
template <int D>
__global__ void __launch_bounds__(1024) attn_rows_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                         const float* __restrict__ V, float* __restrict__ O, int n_img, int T,
                                                         int heads, float scale, int ipc) {
  extern __shared__ float4 smem4[];
  const int HD = heads * D, nthr = blockDim.x, tid = threadIdx.x;
  float* Ks = reinterpret_cast<float*>(smem4);   // [ipc][T][HD]
  float* Vs = Ks + ipc * T * HD;                 // [ipc][T][HD]
  float* sc = Vs + ipc * T * HD;                 // [T][nthr]
  const int img0 = blockIdx.x * ipc;
  const int nim = (n_img - img0 < ipc) ? n_img - img0 : ipc;
  const long long gbase = (long long)img0 * T * HD;
  {
    const float4* K4 = reinterpret_cast<const float4*>(K + gbase);
    const float4* V4 = reinterpret_cast<const float4*>(V + gbase);
    float4* Ks4 = reinterpret_cast<float4*>(Ks);
    float4* Vs4 = reinterpret_cast<float4*>(Vs);
    const int n4 = nim * T * HD / 4;
    for (int i = tid; i < n4; i += nthr) {
      Ks4[i] = K4[i];
      Vs4[i] = V4[i];
    }
  }
  __syncthreads();
  const int per_img = heads * T, pairs = nim * per_img;
  for (int p = tid; p < pairs; p += nthr) {
    const int il = p / per_img, r = p - il * per_img, hh = r / T, t = r - hh * T;
    const long long row = gbase + (long long)(il * T + t) * HD + hh * D;
    float q[D], o[D];
#pragma unroll
    for (int j = 0; j < D; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(Q + row + j);
      q[j] = v.x * scale; q[j + 1] = v.y * scale; q[j + 2] = v.z * scale; q[j + 3] = v.w * scale;
      o[j] = 0.f; o[j + 1] = 0.f; o[j + 2] = 0.f; o[j + 3] = 0.f;
    }
    const float* kb = Ks + il * T * HD + hh * D;
    const float* vb = Vs + il * T * HD + hh * D;
    float* my = sc + tid;
    float mx = -INFINITY;
    for (int s = 0; s < T; ++s) {
      const float* kr = kb + s * HD;
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(kr + j);
        a = fmaf(q[j], kv.x, a); a = fmaf(q[j + 1], kv.y, a); a = fmaf(q[j + 2], kv.z, a); a = fmaf(q[j + 3], kv.w, a);
      }
      my[s * nthr] = a;
      mx = fmaxf(mx, a);
    }
    float sum = 0.f;
    for (int s = 0; s < T; ++s) {
      const float e = expf(my[s * nthr] - mx);
      sum += e;
      const float* vr = vb + s * HD;
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(vr + j);
        o[j] = fmaf(e, vv.x, o[j]); o[j + 1] = fmaf(e, vv.y, o[j + 1]); o[j + 2] = fmaf(e, vv.z, o[j + 2]); o[j + 3] = fmaf(e, vv.w, o[j + 3]);
      }
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < D; j += 4)
      *reinterpret_cast<float4*>(O + row + j) = make_float4(o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv);
  }
}

// images per CTA, threads and shared memory of attn_rows_kernel; false when K, V and one score column per thread do not fit
static bool attn_rows_plan(int T, int heads, int D, int* ipc, int* threads, size_t* smem) {
  const int per_img = heads * T;
  for (int ic = (per_img <= 160 ? 2 : 1); ic >= 1; --ic)
    for (int th = round_up(ic * per_img < 1024 ? ic * per_img : 1024, 32); th >= 64; th = round_up(th / 2, 32)) {
      const size_t sm = ((size_t)2 * ic * T * heads * D + (size_t)T * th) * sizeof(float);
      if (sm <= 200 * 1024) {
        *ipc = ic; *threads = th; *smem = sm;
        return true;
      }
      if (th == 64) break;
    }
  return false;
}

template <int D>
static int launch_attn_rows(hp_ctx* h, const float* Q, const float* K, const float* V, float* O, int n_img, int T, int heads,
                            cudaStream_t st) {
  int ipc = 1, threads = 64;
  size_t smem = 0;
  HP_REQUIRE(attn_rows_plan(T, heads, D, &ipc, &threads, &smem), HP_ERR_UNSUPPORTED, "attention over %d tokens does not fit shared memory", T);
  HP_CUDA(cudaFuncSetAttribute(attn_rows_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  attn_rows_kernel<D><<<ceil_div(n_img, ipc), threads, smem, st>>>(Q, K, V, O, n_img, T, heads, 1.f / sqrtf((float)D), ipc);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

// pass A: per query row: dS = P*(dP - sum(P*dP)), dP = dO V^T ; dQ = scale * dS K ; writes dS to global
__global__ void __launch_bounds__(128) attn_bwd_a_kernel(const float* K, const float* V, const float* dO,
                                                         const float* P, float* dS, float* dQ, int T, int heads, int d,
                                                         float scale) {
  extern __shared__ float smem[];
  const int DS = d + 1;
  float* Ks = smem;
  float* Vs = Ks + T * DS;
  float* sc = Vs + T * DS;  // [4][T]
  const int img = blockIdx.x / heads, hh = blockIdx.x - img * heads;
  const int HD = heads * d;
  const long long base = (long long)img * T * HD + hh * d;
  for (int i = threadIdx.x; i < T * d; i += 128) {
    const int s = i / d, j = i - s * d;
    Ks[s * DS + j] = K[base + (long long)s * HD + j];
    Vs[s * DS + j] = V[base + (long long)s * HD + j];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my = sc + warp * T;
  for (int t = warp; t < T; t += 4) {
    const float* go = dO + base + (long long)t * HD;
    const float* pr = P + ((long long)blockIdx.x * T + t) * T;
    float dot = 0.f;
    for (int s = lane; s < T; s += 32) {
      float a = 0.f;
      for (int j = 0; j < d; ++j) a = fmaf(go[j], Vs[s * DS + j], a);
      my[s] = a;
      dot = fmaf(pr[s], a, dot);
    }
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    float* dsr = dS + ((long long)blockIdx.x * T + t) * T;
    for (int s = lane; s < T; s += 32) {
      const float v = pr[s] * (my[s] - dot);
      my[s] = v;
      dsr[s] = v;
    }
    __syncwarp();
    for (int j = lane; j < d; j += 32) {
      float a = 0.f;
      for (int s = 0; s < T; ++s) a = fmaf(my[s], Ks[s * DS + j], a);
      dQ[base + (long long)t * HD + j] = a * scale;
    }
    __syncwarp();
  }
}
// pass B: thread per key row s: dV[s] = sum_t P[t][s] dO[t] ; dK[s] = scale * sum_t dS[t][s] Q[t]
#define ATTN_DMAX 64
__global__ void __launch_bounds__(128) attn_bwd_b_kernel(const float* Q, const float* dO, const float* P,
                                                         const float* dS, float* dK, float* dV, int T, int heads,
                                                         int d, float scale) {
  extern __shared__ float smem[];
  float* Qs = smem;          // [T][d]
  float* Gs = Qs + T * d;    // [T][d]
  const int nh = blockIdx.x;
  const int img = nh / heads, hh = nh - img * heads;
  const int HD = heads * d;
  const long long base = (long long)img * T * HD + hh * d;
  for (int i = threadIdx.x; i < T * d; i += 128) {
    const int s = i / d, j = i - s * d;
    Qs[i] = Q[base + (long long)s * HD + j];
    Gs[i] = dO[base + (long long)s * HD + j];
  }
  __syncthreads();
  for (int s = blockIdx.y * 128 + threadIdx.x; s < T; s += gridDim.y * 128) {
    float av[ATTN_DMAX], ak[ATTN_DMAX];
#pragma unroll
    for (int j = 0; j < ATTN_DMAX; ++j) { av[j] = 0.f; ak[j] = 0.f; }
    for (int t = 0; t < T; ++t) {
      const float pv = P[((long long)nh * T + t) * T + s];
      const float dv = dS[((long long)nh * T + t) * T + s];
#pragma unroll
      for (int j = 0; j < ATTN_DMAX; ++j)
        if (j < d) {
          av[j] = fmaf(pv, Gs[t * d + j], av[j]);
          ak[j] = fmaf(dv, Qs[t * d + j], ak[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < ATTN_DMAX; ++j)
      if (j < d) {
        dV[base + (long long)s * HD + j] = av[j];
        dK[base + (long long)s * HD + j] = ak[j] * scale;
      }
  }
}

// ============================================================================ loss / optimizer
// g = 2*(pred-y)*inv_count ; sums[0] += (pred-y)^2 ; sums[1] += |pred-y| ; sums[2] += 1 per element
__global__ void mse_loss_kernel(const float* pred, const float* y, float* g, long long total, float inv_count,
                                float* sums) {
  float sq = 0.f, ab = 0.f, cnt = 0.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const float d = pred[i] - y[i];
    if (g) g[i] = 2.f * d * inv_count;
    sq = fmaf(d, d, sq);
    ab += fabsf(d);
    cnt += 1.f;
  }
  for (int o = 16; o; o >>= 1) {
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    ab += __shfl_xor_sync(0xffffffffu, ab, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  __shared__ float red[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = sq; red[1][warp] = ab; red[2][warp] = cnt; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, s);
  }
}
// sums[3] = sum l2coef*w^2 (single block)
__global__ void l2_penalty_kernel(const float* w, const float* l2, int n, float* out) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s = fmaf(l2[i] * w[i], w[i], s);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    *out = t;
  }
}
// Keras 2.13 update rules.  g_total = g + 2*l2*w
struct TrainState {            // device-resident state of hp_head_train_run
  uint32_t t;                  // optimizer steps taken, this one included (must stay the first member: dropout reads it)
  uint32_t k;                  // steps done in this call
  float alpha_t, lr_t;         // Adam / Adamax step sizes of step t
  float acc[2];                // sum over the call's steps of loss * n_global, mae * n_global
};
__global__ void optimizer_kernel(float* w, const float* g, const float* l2, float* m, float* v, int n, int kind,
                                 float lr, float b1, float b2, float eps, float alpha_t, float lr_t, const TrainState* ts = nullptr) {
  if (ts) { alpha_t = ts->alpha_t; lr_t = ts->lr_t; }
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float wi = w[i];
    const float gi = g[i] + 2.f * l2[i] * wi;
    if (kind == HP_OPT_SGD) {
      w[i] = wi - lr * gi;
    } else if (kind == HP_OPT_ADAM) {
      const float mi = m[i] + (gi - m[i]) * (1.f - b1);
      const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
      m[i] = mi;
      v[i] = vi;
      w[i] = wi - (mi * alpha_t) / (sqrtf(vi) + eps);
    } else {
      const float mi = m[i] + (gi - m[i]) * (1.f - b1);
      const float ui = fmaxf(b2 * v[i], fabsf(gi));
      m[i] = mi;
      v[i] = ui;
      w[i] = wi - (lr_t * mi) / (ui + eps);
    }
  }
}

// ============================================================================ head object
struct MhaWs {
  size_t q, k, v, o, gq, gk, gv, go, p, ds;  // float offsets into hp_head::ws
};
struct hp_head {
  std::vector<hp_head_op> ops;
  std::vector<hp_head_reg> regs;
  int out_reg = 0, n_params = 0, in_channels = 0;
  DevBuf params, grads, m, v, l2coef;
  DevBuf acts, gacts, ws;
  std::vector<size_t> reg_off;        // float offsets into acts/gacts (reg 0 unused: external input)
  std::vector<size_t> stat_off;       // per op: layer-norm stats offset in ws
  std::vector<MhaWs> mha;             // per op
  std::vector<int> nreads;            // per register: ops that read it (+ 1 for the output register)
  std::vector<int> alias;             // per register: the register that holds its values after the latest forward
                                      // (inference skips the SpatialDropout2D copies; identity after a training forward)
  uint32_t step = 0;
  bool has_state = false;
  float last_sums[4] = {0, 0, 0, 0};
  // hp_head_train_run: device-resident step state, staging batch, captured step graphs
  DevBuf tstate, xs, ys;
  const uint32_t* step_dev = nullptr;   // set while a train_run step is being issued: dropout reads the step from tstate
  struct RunGraph {
    const void *x = nullptr, *y = nullptr, *idx = nullptr;
    long long first = 0, epoch = 0;
    int batch_global = 0, n_local = 0, rank = 0, world = 0, H = 0, W = 0, impl = 0;
    hp_opt_config opt{};
    uint64_t seed = 0;
    void* comm = nullptr;
    int p2p = 0;
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
  };
  std::vector<RunGraph> run_graphs;
};

static long long reg_rows(const hp_head* hd, int r, int n_img, int T) {
  return hd->regs[r].per_image ? n_img : (long long)n_img * T;
}

int hp_head_create_impl(hp_ctx* h, const hp_head_op* ops, int n_ops, const hp_head_reg* regs, int n_regs, int out_reg,
                        int n_params, hp_head** out) {
  HP_REQUIRE(ops && regs && out && n_ops > 0 && n_regs > 1, HP_ERR_INVALID, "hp_head_create: bad arguments");
  HP_REQUIRE(out_reg > 0 && out_reg < n_regs, HP_ERR_INVALID, "hp_head_create: out_reg %d out of range", out_reg);
  HP_REQUIRE(n_params > 0, HP_ERR_INVALID, "hp_head_create: n_params must be positive");
  hp_head* hd = new hp_head();
  hd->ops.assign(ops, ops + n_ops);
  hd->regs.assign(regs, regs + n_regs);
  hd->out_reg = out_reg;
  hd->nreads.assign(n_regs, 0);
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].in0 >= 0 && ops[i].in0 < n_regs) hd->nreads[ops[i].in0]++;
    if ((ops[i].op == HP_OP_ADD || ops[i].op == HP_OP_MULCH) && ops[i].in1 >= 0 && ops[i].in1 < n_regs) hd->nreads[ops[i].in1]++;
  }
  hd->nreads[out_reg]++;
  hd->alias.resize(n_regs);
  for (int r = 0; r < n_regs; ++r) hd->alias[r] = r;
  hd->n_params = n_params;
  hd->in_channels = regs[0].channels;
  std::vector<float> l2(n_params, 0.f);
  std::vector<char> written(n_regs, 0);
  written[0] = 1;
  for (int i = 0; i < n_ops; ++i) {
    const hp_head_op& o = ops[i];
    auto bad = [&](const char* why) {
      hp_set_error("hp_head_create: op %d (%d): %s", i, o.op, why);
      delete hd;
      return HP_ERR_INVALID;
    };
    if (o.in0 < 0 || o.in0 >= n_regs || o.out <= 0 || o.out >= n_regs) return bad("register index out of range");
    if (!written[o.in0]) return bad("input register read before it is written");
    if (written[o.out]) return bad("registers are single-assignment");
    const int cin = regs[o.in0].channels, cout = regs[o.out].channels;
    switch (o.op) {
      case HP_OP_DENSE:
        if (o.cin != cin || o.cout != cout) return bad("dense shape mismatch");
        if (o.w_off < 0 || o.w_off + (long long)cin * cout > n_params || o.b_off < 0 || o.b_off + cout > n_params)
          return bad("parameter offsets out of range");
        if (regs[o.in0].per_image != regs[o.out].per_image) return bad("dense keeps the row kind");
        for (int j = 0; j < cin * cout; ++j) l2[o.w_off + j] = o.l2_w;
        for (int j = 0; j < cout; ++j) l2[o.b_off + j] = o.l2_b;
        break;
      case HP_OP_ACT:
      case HP_OP_DROPOUT:
        if (cin != cout || regs[o.in0].per_image != regs[o.out].per_image) return bad("shape mismatch");
        if (o.op == HP_OP_DROPOUT && !(o.fparam >= 0.f && o.fparam < 1.f)) return bad("dropout rate must be in [0,1)");
        break;
      case HP_OP_ADD:
        if (o.in1 < 0 || o.in1 >= n_regs || !written[o.in1]) return bad("second input invalid");
        if (regs[o.in1].channels != cin || cout != cin) return bad("add needs equal channel counts");
        if (regs[o.out].per_image != (regs[o.in0].per_image && regs[o.in1].per_image)) return bad("add row kind");
        break;
      case HP_OP_MULCH:
        if (o.in1 < 0 || o.in1 >= n_regs || !written[o.in1]) return bad("second input invalid");
        if (!regs[o.in1].per_image || regs[o.in0].per_image || regs[o.out].per_image) return bad("mulch: x per token, gate per image");
        if (regs[o.in1].channels != cin || cout != cin) return bad("mulch channel mismatch");
        break;
      case HP_OP_GAP:
        if (regs[o.in0].per_image || !regs[o.out].per_image || cin != cout) return bad("gap shapes");
        break;
      case HP_OP_LAYERNORM:
        if (cin != cout || o.w_off < 0 || o.w_off + cin > n_params || o.b_off < 0 || o.b_off + cin > n_params)
          return bad("layernorm parameters");
        break;
      case HP_OP_MHA: {
        const int hdm = o.heads * o.key_dim;
        if (o.heads <= 0 || o.key_dim <= 0 || o.key_dim > ATTN_DMAX) return bad("heads/key_dim unsupported");
        if (cin != cout || regs[o.in0].per_image || regs[o.out].per_image) return bad("mha shapes");
        const long long need = 3ll * (cin * hdm + hdm) + (long long)hdm * cin + cin;
        if (o.w_off < 0 || o.w_off + need > n_params) return bad("mha parameters out of range");
        break;
      }
      default:
        return bad("unknown opcode");
    }
    written[o.out] = 1;
  }
  if (!written[out_reg]) {
    hp_set_error("hp_head_create: output register never written");
    delete hd;
    return HP_ERR_INVALID;
  }
  int rc = hd->params.ensure((size_t)n_params * sizeof(float));
  if (rc == HP_OK) rc = hd->grads.ensure((size_t)(n_params + 4) * sizeof(float));
  if (rc == HP_OK) rc = hd->l2coef.ensure((size_t)n_params * sizeof(float));
  if (rc != HP_OK) { delete hd; return rc; }
  cudaMemset(hd->params.p, 0, (size_t)n_params * sizeof(float));
  cudaMemcpy(hd->l2coef.p, l2.data(), (size_t)n_params * sizeof(float), cudaMemcpyHostToDevice);
  *out = hd;
  return HP_OK;
}

void hp_head_run_graphs_free(hp_head* hd);
void hp_head_free_impl(hp_head* hd) {
  if (hd) hp_head_run_graphs_free(hd);
  if (!hd) return;
  hd->params.release(); hd->grads.release(); hd->m.release(); hd->v.release(); hd->l2coef.release();
  hd->acts.release(); hd->gacts.release(); hd->ws.release();
  hd->tstate.release(); hd->xs.release(); hd->ys.release();
  delete hd;
}

// lay out activation / workspace arenas for (n_img, T)
static int head_plan(hp_head* hd, int n_img, int T, bool training) {
  const int n_regs = (int)hd->regs.size();
  hd->reg_off.assign(n_regs, 0);
  size_t off = 0;
  for (int r = 1; r < n_regs; ++r) {
    hd->reg_off[r] = off;
    off += (size_t)round_up((int)std::min<long long>(reg_rows(hd, r, n_img, T) * hd->regs[r].channels, 1ll << 30), 4);
    HP_REQUIRE(reg_rows(hd, r, n_img, T) * hd->regs[r].channels < (1ll << 30), HP_ERR_INVALID, "head batch too large");
  }
  HP_TRY(hd->acts.ensure(off * sizeof(float)));
  if (training) HP_TRY(hd->gacts.ensure(off * sizeof(float)));
  size_t woff = 0;
  hd->stat_off.assign(hd->ops.size(), 0);
  hd->mha.assign(hd->ops.size(), MhaWs());
  const long long tok = (long long)n_img * T;
  for (size_t i = 0; i < hd->ops.size(); ++i) {
    const hp_head_op& o = hd->ops[i];
    if (o.op == HP_OP_LAYERNORM) {
      hd->stat_off[i] = woff;
      woff += (size_t)round_up((int)(2 * reg_rows(hd, o.in0, n_img, T)), 4);
    } else if (o.op == HP_OP_MHA) {
      const size_t sz = (size_t)round_up((int)(tok * o.heads * o.key_dim), 4);
      MhaWs& w = hd->mha[i];
      w.q = woff; woff += sz; w.k = woff; woff += sz; w.v = woff; woff += sz; w.o = woff; woff += sz;
      if (training) {
        w.gq = woff; woff += sz; w.gk = woff; woff += sz; w.gv = woff; woff += sz; w.go = woff; woff += sz;
        const size_t pp = (size_t)round_up((int)std::min<long long>((long long)n_img * o.heads * T * T, 1ll << 30), 4);
        HP_REQUIRE((long long)n_img * o.heads * T * T < (1ll << 30), HP_ERR_INVALID, "attention training scratch too large");
        w.p = woff; woff += pp; w.ds = woff; woff += pp;
      }
    }
  }
  HP_TRY(hd->ws.ensure((woff + 4) * sizeof(float)));
  return HP_OK;
}

static size_t attn_smem(int T, int d) { return ((size_t)2 * T * (d + 1) + 4 * (size_t)T + 4 * (size_t)d) * sizeof(float); }

static int head_forward(hp_ctx* h, hp_head* hd, const float* feat, int n_img, int T, bool training, uint64_t seed,
                        cudaStream_t st) {
  float* A = hd->acts.f();
  for (size_t r = 0; r < hd->alias.size(); ++r) hd->alias[r] = (int)r;
  auto R = [&](int r) -> const float* { return hd->alias[r] == 0 ? feat : A + hd->reg_off[hd->alias[r]]; };
  auto RW = [&](int r) -> float* { return A + hd->reg_off[r]; };
  // a training step keeps plain fp32 FMA accumulation in forward and backward; inference may use the 3xTF32 kernel
  struct DenseTcScope {
    hp_ctx* h; bool saved;
    DenseTcScope(hp_ctx* c, bool on) : h(c), saved(c->dense_tc) { c->dense_tc = saved && on; }
    ~DenseTcScope() { h->dense_tc = saved; }
  } dense_tc_scope(h, !training);
  for (size_t i = 0; i < hd->ops.size(); ++i) {
    const hp_head_op& o = hd->ops[i];
    const long long rows_out = reg_rows(hd, o.out, n_img, T);
    const int C = hd->regs[o.out].channels;
    const long long total = rows_out * C;
    switch (o.op) {
      case HP_OP_DENSE: {
        if (!training && h->impl == HP_IMPL_FAST && h->dense_tc) {
          // inference: Dense -> [SpatialDropout2D] -> Dense with at most 4 outputs (the yaw / pitch / roll layer) in one
          // kernel when nothing else reads the hidden activations; they are then never written (dense_tc.cu, TAIL)
          size_t k = i + 1;
          int cur = o.out;
          while (k < hd->ops.size() && hd->ops[k].op == HP_OP_DROPOUT && hd->ops[k].in0 == cur && hd->nreads[cur] == 1) cur = hd->ops[k++].out;
          if (k < hd->ops.size() && hd->ops[k].op == HP_OP_DENSE && hd->ops[k].in0 == cur && hd->nreads[cur] == 1 && hd->ops[k].cout <= 4 &&
              hp_dense_tc_tail_supported(R(o.in0), (int)rows_out, o.cin, o.cin, o.cout, hd->ops[k].cout)) {
            const hp_head_op& o2 = hd->ops[k];
            DenseTail tail{hd->params.f() + o2.w_off, hd->params.f() + o2.b_off, o2.cout, o2.cout, o2.act,
                           DenseOut{RW(o2.out), 0, o2.cout, (int)std::max<long long>(rows_out, 1), 0, o2.cout}};
            HP_TRY(hp_launch_dense_tc_tail(h, R(o.in0), (int)rows_out, o.cin, o.cin, hd->params.f() + o.w_off, o.cout,
                                           hd->params.f() + o.b_off, o.cout, o.act, tail, st));
            i = k;
            break;
          }
        }
        DenseOut d{RW(o.out), 0, o.cout, (int)std::max<long long>(rows_out, 1), 0, o.cout};
        HP_TRY(hp_launch_dense(h, R(o.in0), (int)rows_out, o.cin, o.cin, hd->params.f() + o.w_off, o.cout,
                               hd->params.f() + o.b_off, o.cout, o.act, false, &d, 1, false, st));
        break;
      }
      case HP_OP_ACT:
        act_fwd_kernel<<<EW_GRID(total), 256, 0, st>>>(R(o.in0), RW(o.out), total, o.act);
        h->launches++;
        break;
      case HP_OP_ADD:
        // inference: residual add followed by a LayerNorm that is its only reader -> one kernel, the sum is not stored
        if (!training && h->impl == HP_IMPL_FAST && i + 1 < hd->ops.size() && hd->ops[i + 1].op == HP_OP_LAYERNORM &&
            hd->ops[i + 1].in0 == o.out && hd->nreads[o.out] == 1 && !hd->regs[o.in0].per_image == !hd->regs[o.out].per_image &&
            !hd->regs[o.in1].per_image == !hd->regs[o.out].per_image) {
          const hp_head_op& ln = hd->ops[i + 1];
          const float *gm = hd->params.f() + ln.w_off, *bt = hd->params.f() + ln.b_off;
          if (layernorm_rows_ok(C, R(o.in0), R(o.in1), RW(ln.out))) {
            launch_layernorm_rows(h, R(o.in0), R(o.in1), gm, bt, RW(ln.out), rows_out, C, ln.fparam, st);
            ++i;
            break;
          }
        }
        if (!hd->regs[o.in0].per_image == !hd->regs[o.out].per_image && !hd->regs[o.in1].per_image == !hd->regs[o.out].per_image &&
            total % 4 == 0 && (((uintptr_t)R(o.in0) | (uintptr_t)R(o.in1) | (uintptr_t)RW(o.out)) & 15) == 0) {
          add_vec4_kernel<<<EW_GRID(total / 4), 256, 0, st>>>((const float4*)R(o.in0), (const float4*)R(o.in1), (float4*)RW(o.out), total / 4);
          h->launches++;
          break;
        }
        add_kernel<<<EW_GRID(total), 256, 0, st>>>(R(o.in0), R(o.in1), RW(o.out), rows_out, C, T,
                                                   hd->regs[o.in0].per_image && !hd->regs[o.out].per_image,
                                                   hd->regs[o.in1].per_image && !hd->regs[o.out].per_image);
        h->launches++;
        break;
      case HP_OP_MULCH:
        if (C % 4 == 0 && total / 4 < (1ll << 31) && (((uintptr_t)R(o.in0) | (uintptr_t)R(o.in1) | (uintptr_t)RW(o.out)) & 15) == 0) {
          mulch_vec4_kernel<<<EW_GRID(total / 4), 256, 0, st>>>((const float4*)R(o.in0), (const float4*)R(o.in1), (float4*)RW(o.out),
                                                               (unsigned)(total / 4), (unsigned)(C / 4), (unsigned)T);
        } else {
          mulch_kernel<<<EW_GRID(total), 256, 0, st>>>(R(o.in0), R(o.in1), RW(o.out), rows_out, C, T);
        }
        h->launches++;
        break;
      case HP_OP_GAP:
        gap_kernel<<<EW_GRID(total), 256, 0, st>>>(R(o.in0), RW(o.out), n_img, C, T);
        h->launches++;
        break;
      case HP_OP_DROPOUT:
        if (training && o.fparam > 0.f) {
          dropout_kernel<<<EW_GRID(total), 256, 0, st>>>(R(o.in0), RW(o.out), rows_out, C, T, hd->regs[o.out].per_image,
                                                        o.fparam, seed, hd->step, (uint32_t)o.op_id, 0, hd->step_dev);
        } else if (!training) {
          hd->alias[o.out] = hd->alias[o.in0];   // identity at inference: readers of o.out are pointed at the source, no copy
          break;
        } else {
          HP_CUDA(cudaMemcpyAsync(RW(o.out), R(o.in0), total * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
        h->launches++;
        break;
      case HP_OP_LAYERNORM: {
        if (!training && h->impl == HP_IMPL_FAST &&
            layernorm_rows_ok(C, R(o.in0), RW(o.out), nullptr)) {
          launch_layernorm_rows(h, R(o.in0), nullptr, hd->params.f() + o.w_off, hd->params.f() + o.b_off, RW(o.out), rows_out, C, o.fparam, st);
          break;
        }
        long long warps_needed = rows_out;
        unsigned grid = (unsigned)std::min<long long>((warps_needed + 7) / 8, 65535ll * 8);
        layernorm_kernel<<<grid, 256, 0, st>>>(R(o.in0), hd->params.f() + o.w_off, hd->params.f() + o.b_off, RW(o.out),
                                               hd->ws.f() + hd->stat_off[i], rows_out, C, o.fparam);
        h->launches++;
        break;
      }
      case HP_OP_MHA: {
        const int hdm = o.heads * o.key_dim;
        const float* P = hd->params.f() + o.w_off;
        const float *Wq = P, *bq = Wq + C * hdm, *Wk = bq + hdm, *bk = Wk + C * hdm, *Wv = bk + hdm, *bv = Wv + C * hdm,
                    *Wo = bv + hdm, *bo = Wo + hdm * C;
        const MhaWs& w = hd->mha[i];
        float* ws = hd->ws.f();
        const int rows = (int)rows_out;
        DenseOut dq{ws + w.q, 0, hdm, rows, 0, hdm}, dk{ws + w.k, 0, hdm, rows, 0, hdm}, dv{ws + w.v, 0, hdm, rows, 0, hdm};
        HP_TRY(hp_launch_dense(h, R(o.in0), rows, C, C, Wq, hdm, bq, hdm, HP_ACT_LINEAR, false, &dq, 1, false, st));
        HP_TRY(hp_launch_dense(h, R(o.in0), rows, C, C, Wk, hdm, bk, hdm, HP_ACT_LINEAR, false, &dk, 1, false, st));
        HP_TRY(hp_launch_dense(h, R(o.in0), rows, C, C, Wv, hdm, bv, hdm, HP_ACT_LINEAR, false, &dv, 1, false, st));
        int rows_ipc, rows_threads;
        size_t rows_smem;
        const bool rows_ok = !training && h->impl == HP_IMPL_FAST && (hdm % 4) == 0 &&
                             ((((uintptr_t)(ws + w.q)) | ((uintptr_t)(ws + w.k)) | ((uintptr_t)(ws + w.v)) | ((uintptr_t)(ws + w.o))) & 15) == 0 &&
                             attn_rows_plan(T, o.heads, o.key_dim, &rows_ipc, &rows_threads, &rows_smem);
        if (rows_ok && (o.key_dim == 8 || o.key_dim == 16 || o.key_dim == 32)) {
          if (o.key_dim == 8) HP_TRY(launch_attn_rows<8>(h, ws + w.q, ws + w.k, ws + w.v, ws + w.o, n_img, T, o.heads, st));
          else if (o.key_dim == 16) HP_TRY(launch_attn_rows<16>(h, ws + w.q, ws + w.k, ws + w.v, ws + w.o, n_img, T, o.heads, st));
          else HP_TRY(launch_attn_rows<32>(h, ws + w.q, ws + w.k, ws + w.v, ws + w.o, n_img, T, o.heads, st));
          DenseOut dout{RW(o.out), 0, C, rows, 0, C};
          HP_TRY(hp_launch_dense(h, ws + w.o, rows, hdm, hdm, Wo, C, bo, C, HP_ACT_LINEAR, false, &dout, 1, false, st));
          break;
        }
        const size_t smem = attn_smem(T, o.key_dim);
        HP_REQUIRE(smem <= 200 * 1024, HP_ERR_UNSUPPORTED, "attention over %d tokens does not fit shared memory", T);
        HP_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attn_fwd_kernel<<<n_img * o.heads, 128, smem, st>>>(ws + w.q, ws + w.k, ws + w.v, ws + w.o, T, o.heads, o.key_dim,
                                                           1.f / sqrtf((float)o.key_dim), training ? ws + w.p : nullptr,
                                                           nullptr);
        h->launches++;
        DenseOut dout{RW(o.out), 0, C, rows, 0, C};
        HP_TRY(hp_launch_dense(h, ws + w.o, rows, hdm, hdm, Wo, C, bo, C, HP_ACT_LINEAR, false, &dout, 1, false, st));
        break;
      }
    }
    HP_CUDA(cudaGetLastError());
  }
  return HP_OK;
}

static int dense_backward(hp_ctx* h, const float* x, int rows, int K, const float* W, int N, const float* gz,
                          float* gW, float* gb, float* gx, cudaStream_t st) {
  if (rows <= 0) return HP_OK;
  const size_t smem = (size_t)WG_TM * (K + N) * sizeof(float);
  HP_REQUIRE(smem <= 96 * 1024, HP_ERR_UNSUPPORTED, "dense backward %dx%d too large", K, N);
  HP_CUDA(cudaFuncSetAttribute(dense_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  dense_wgrad_kernel<<<ceil_div(rows, WG_TM), 256, smem, st>>>(x, K, gz, N, gW, gb, rows, K, N);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  if (gx) {
    DenseOut d{gx, 0, K, rows, 0, K};
    HP_TRY(hp_launch_dense(h, gz, rows, N, N, W, N, nullptr, K, HP_ACT_LINEAR, true, &d, 1, true, st));
  }
  return HP_OK;
}

static int head_backward(hp_ctx* h, hp_head* hd, const float* feat, int n_img, int T, uint64_t seed, cudaStream_t st) {
  float* A = hd->acts.f();
  float* G = hd->gacts.f();
  float* gp = hd->grads.f();
  auto R = [&](int r) -> const float* { return r == 0 ? feat : A + hd->reg_off[r]; };
  auto GR = [&](int r) -> float* { return r == 0 ? nullptr : G + hd->reg_off[r]; };
  for (int i = (int)hd->ops.size() - 1; i >= 0; --i) {
    const hp_head_op& o = hd->ops[i];
    const long long rows_out = reg_rows(hd, o.out, n_img, T);
    const int C = hd->regs[o.out].channels;
    const long long total = rows_out * C;
    float* gy = GR(o.out);
    if ((o.op == HP_OP_DENSE || o.op == HP_OP_ACT) && o.act == HP_ACT_SWISH) {
      hp_set_error("training through a swish activation is not supported (its derivative is not a function of its output)");
      return HP_ERR_UNSUPPORTED;
    }
    switch (o.op) {
      case HP_OP_DENSE: {
        if (o.act != HP_ACT_LINEAR) {
          act_bwd_kernel<<<EW_GRID(total), 256, 0, st>>>(R(o.out), gy, gy, total, o.act);
          h->launches++;
        }
        HP_TRY(dense_backward(h, R(o.in0), (int)rows_out, o.cin, hd->params.f() + o.w_off, o.cout, gy, gp + o.w_off,
                              gp + o.b_off, GR(o.in0), st));
        break;
      }
      case HP_OP_ACT:
        if (GR(o.in0)) {
          act_bwd_acc_kernel<<<EW_GRID(total), 256, 0, st>>>(R(o.out), gy, GR(o.in0), total, o.act);
          h->launches++;
        }
        break;
      case HP_OP_ADD: {
        const int ins[2] = {o.in0, o.in1};
        for (int k = 0; k < 2; ++k) {
          float* gi = GR(ins[k]);
          if (!gi) continue;
          if (hd->regs[ins[k]].per_image && !hd->regs[o.out].per_image) {
            hp_set_error("backward of broadcasting Add is not supported");
            return HP_ERR_UNSUPPORTED;
          }
          acc_kernel<<<EW_GRID(total), 256, 0, st>>>(gi, gy, total);
          h->launches++;
        }
        break;
      }
      case HP_OP_MULCH: {
        const long long tg = (long long)n_img * C;
        mulch_bwd_kernel<<<EW_GRID(tg), 256, 0, st>>>(R(o.in0), R(o.in1), gy, GR(o.in0), GR(o.in1), n_img, C, T);
        h->launches++;
        break;
      }
      case HP_OP_GAP:
        if (GR(o.in0)) {
          const long long rows_in = (long long)n_img * T;
          gap_bwd_kernel<<<EW_GRID(rows_in * C), 256, 0, st>>>(gy, GR(o.in0), rows_in, C, T);
          h->launches++;
        }
        break;
      case HP_OP_DROPOUT:
        if (GR(o.in0)) {
          if (o.fparam > 0.f)
            dropout_kernel<<<EW_GRID(total), 256, 0, st>>>(gy, GR(o.in0), rows_out, C, T, hd->regs[o.out].per_image,
                                                          o.fparam, seed, hd->step, (uint32_t)o.op_id, 1, hd->step_dev);
          else
            acc_kernel<<<EW_GRID(total), 256, 0, st>>>(GR(o.in0), gy, total);
          h->launches++;
        }
        break;
      case HP_OP_LAYERNORM: {
        unsigned grid = (unsigned)std::min<long long>((rows_out + 7) / 8, 65535ll * 8);
        layernorm_bwd_kernel<<<grid, 256, 0, st>>>(R(o.in0), hd->params.f() + o.w_off, hd->ws.f() + hd->stat_off[i], gy,
                                                   GR(o.in0), gp + o.w_off, gp + o.b_off, rows_out, C);
        h->launches++;
        break;
      }
      case HP_OP_MHA: {
        const int hdm = o.heads * o.key_dim, d = o.key_dim;
        const float* P = hd->params.f() + o.w_off;
        const float *Wq = P, *Wk = Wq + C * hdm + hdm, *Wv = Wk + C * hdm + hdm, *Wo = Wv + C * hdm + hdm;
        float* gP = gp + o.w_off;
        float *gWq = gP, *gbq = gWq + C * hdm, *gWk = gbq + hdm, *gbk = gWk + C * hdm, *gWv = gbk + hdm,
              *gbv = gWv + C * hdm, *gWo = gbv + hdm, *gbo = gWo + hdm * C;
        const MhaWs& w = hd->mha[i];
        float* ws = hd->ws.f();
        const int rows = (int)rows_out;
        const float scale = 1.f / sqrtf((float)d);
        // output projection: out = O Wo + bo
        HP_CUDA(cudaMemsetAsync(ws + w.go, 0, (size_t)rows * hdm * sizeof(float), st));
        HP_TRY(dense_backward(h, ws + w.o, rows, hdm, Wo, C, gy, gWo, gbo, ws + w.go, st));
        const size_t smem_a = attn_smem(T, d);
        HP_CUDA(cudaFuncSetAttribute(attn_bwd_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attn_bwd_a_kernel<<<n_img * o.heads, 128, smem_a, st>>>(ws + w.k, ws + w.v, ws + w.go, ws + w.p, ws + w.ds,
                                                               ws + w.gq, T, o.heads, d, scale);
        h->launches++;
        const size_t smem_b = (size_t)2 * T * d * sizeof(float);
        HP_REQUIRE(smem_b <= 200 * 1024, HP_ERR_UNSUPPORTED, "attention backward over %d tokens too large", T);
        HP_CUDA(cudaFuncSetAttribute(attn_bwd_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        dim3 gb(n_img * o.heads, ceil_div(T, 128));
        attn_bwd_b_kernel<<<gb, 128, smem_b, st>>>(ws + w.q, ws + w.go, ws + w.p, ws + w.ds, ws + w.gk, ws + w.gv, T,
                                                   o.heads, d, scale);
        h->launches++;
        HP_CUDA(cudaGetLastError());
        float* gx = GR(o.in0);
        HP_TRY(dense_backward(h, R(o.in0), rows, C, Wq, hdm, ws + w.gq, gWq, gbq, gx, st));
        HP_TRY(dense_backward(h, R(o.in0), rows, C, Wk, hdm, ws + w.gk, gWk, gbk, gx, st));
        HP_TRY(dense_backward(h, R(o.in0), rows, C, Wv, hdm, ws + w.gv, gWv, gbv, gx, st));
        break;
      }
    }
    HP_CUDA(cudaGetLastError());
  }
  return HP_OK;
}

// ============================================================================ entry points used by api.cu
int hp_head_set_weights_impl(hp_head* hd, const float* src, int n) {
  HP_REQUIRE(hd && src && n == hd->n_params, HP_ERR_INVALID, "hp_head_set_weights: expected %d params", hd ? hd->n_params : -1);
  HP_CUDA(cudaMemcpy(hd->params.p, src, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  return HP_OK;
}
int hp_head_get_weights_impl(hp_head* hd, float* dst, int n) {
  HP_REQUIRE(hd && dst && n == hd->n_params, HP_ERR_INVALID, "hp_head_get_weights: expected %d params", hd ? hd->n_params : -1);
  HP_CUDA(cudaDeviceSynchronize());
  HP_CUDA(cudaMemcpy(dst, hd->params.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return HP_OK;
}
int hp_head_get_grads_impl(hp_head* hd, float* dst, int n) {
  HP_REQUIRE(hd && dst && n == hd->n_params, HP_ERR_INVALID, "hp_head_get_grads: expected %d params", hd ? hd->n_params : -1);
  HP_CUDA(cudaDeviceSynchronize());
  HP_CUDA(cudaMemcpy(dst, hd->grads.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return HP_OK;
}
int hp_head_in_channels(hp_head* hd) { return hd->in_channels; }
int hp_head_out_channels(hp_head* hd) { return hd->regs[hd->out_reg].channels; }

int hp_head_forward_impl(hp_ctx* h, hp_head* hd, const float* feat, int B, int H, int W, float* out, cudaStream_t st) {
  HP_REQUIRE(hd && feat && out && B > 0 && H > 0 && W > 0, HP_ERR_INVALID, "hp_head_forward: bad arguments");
  HP_REQUIRE(!hd->regs[hd->out_reg].per_image, HP_ERR_INVALID, "head output must be per token");
  const int T = H * W;
  HP_TRY(head_plan(hd, B, T, false));
  HP_TRY(head_forward(h, hd, feat, B, T, false, 0, st));
  const long long total = (long long)B * T * hd->regs[hd->out_reg].channels;
  const int src = hd->alias[hd->out_reg];
  HP_CUDA(cudaMemcpyAsync(out, src == 0 ? feat : hd->acts.f() + hd->reg_off[src], total * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return HP_OK;
}

int hp_head_train_step_impl(hp_ctx* h, hp_head* hd, const float* x, const float* y, int n, int H, int W, int n_global,
                            const hp_opt_config* opt, uint64_t seed, float* loss_mae_host, bool update,
                            cudaStream_t st) {
  HP_REQUIRE(hd && x && y && n > 0 && H > 0 && W > 0, HP_ERR_INVALID, "hp_head_train_step: bad arguments");
  HP_REQUIRE(!update || opt, HP_ERR_INVALID, "hp_head_train_step: optimizer config required");
  HP_REQUIRE(n_global >= n, HP_ERR_INVALID, "n_global (%d) must be >= n (%d)", n_global, n);
  const int T = H * W;
  const int Cout = hd->regs[hd->out_reg].channels;
  const int np = hd->n_params;
  HP_TRY(head_plan(hd, n, T, update));
  HP_CUDA(cudaMemsetAsync(hd->grads.p, 0, (size_t)(np + 4) * sizeof(float), st));
  HP_TRY(head_forward(h, hd, x, n, T, update, seed, st));
  const long long total = (long long)n * T * Cout;
  float* sums = hd->grads.f() + np;
  const float* pred = hd->alias[hd->out_reg] == 0 ? x : hd->acts.f() + hd->reg_off[hd->alias[hd->out_reg]];
  if (update) {
    size_t arena = hd->gacts.bytes;
    HP_CUDA(cudaMemsetAsync(hd->gacts.p, 0, arena, st));
  }
  const float inv_count = 1.f / ((float)n_global * (float)T * (float)Cout);
  mse_loss_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 1024), 256, 0, st>>>(
      pred, y, update ? hd->gacts.f() + hd->reg_off[hd->out_reg] : nullptr, total, inv_count, sums);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  if (update) {
    HP_TRY(head_backward(h, hd, x, n, T, seed, st));
    if (h->comm.comm) HP_TRY(hp_comm_allreduce_sum(h, hd->grads.f(), (size_t)np + 3, st));
    if (opt->kind != HP_OPT_SGD && !hd->has_state) {
      HP_TRY(hd->m.ensure((size_t)np * sizeof(float)));
      HP_TRY(hd->v.ensure((size_t)np * sizeof(float)));
      HP_CUDA(cudaMemsetAsync(hd->m.p, 0, (size_t)np * sizeof(float), st));
      HP_CUDA(cudaMemsetAsync(hd->v.p, 0, (size_t)np * sizeof(float), st));
      hd->has_state = true;
    }
  }
  if (loss_mae_host || !update) {
    l2_penalty_kernel<<<1, 256, 0, st>>>(hd->params.f(), hd->l2coef.f(), np, sums + 3);
    h->launches++;
    HP_CUDA(cudaMemcpyAsync(hd->last_sums, sums, 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  if (update) {
    const double t = (double)hd->step + 1.0;
    const double b1 = opt->beta1, b2 = opt->beta2;
    const float alpha_t = (float)(opt->lr * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t)));
    const float lr_t = (float)(opt->lr / (1.0 - pow(b1, t)));
    optimizer_kernel<<<ceil_div(np, 256), 256, 0, st>>>(hd->params.f(), hd->grads.f(), hd->l2coef.f(), hd->m.f(), hd->v.f(),
                                                       np, opt->kind, opt->lr, opt->beta1, opt->beta2, opt->eps, alpha_t,
                                                       lr_t);
    h->launches++;
    HP_CUDA(cudaGetLastError());
    hd->step++;
  }
  if (loss_mae_host) {
    HP_CUDA(cudaStreamSynchronize(st));
    const float cnt = update ? (float)n_global * T * Cout : hd->last_sums[2];
    loss_mae_host[0] = hd->last_sums[0] / cnt + (update ? hd->last_sums[3] : 0.f);
    loss_mae_host[1] = hd->last_sums[1] / cnt;
    if (!update) loss_mae_host[2] = hd->last_sums[3];
  }
  return HP_OK;
}

// ============================================================================ hp_head_train_run: many steps, no host round trips
// Step prologue: advance the device-resident counters and derive the step sizes of optimizer step t (Keras 2.13: alpha_t =
// lr sqrt(1 - b2^t) / (1 - b1^t) for Adam, lr / (1 - b1^t) for Adamax, in double like the host path).
__global__ void train_begin_kernel(TrainState* ts, float lr, float b1, float b2, unsigned int* exchange_seq) {
  if (threadIdx.x == 0) {
    if (exchange_seq) *exchange_seq += 1u;          // tag of this step's peer-memory exchange (context-wide, the same on all ranks)
    const uint32_t t = ts->t + 1u;
    ts->t = t;
    const double td = (double)t, d1 = 1.0 - pow((double)b1, td);
    ts->alpha_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, td)) / d1);
    ts->lr_t = (float)((double)lr / d1);
  }
}
// This rank's rows of global batch k: items idx[first + k * batch_global + rank + j * world], j < n_local (idx == nullptr:
// identity).  One CTA per row.
__global__ void __launch_bounds__(128) train_gather_kernel(const TrainState* ts, const float* __restrict__ x_all, const float* __restrict__ y_all,
                                                          const int32_t* __restrict__ idx, long long first, int batch_global, int rank, int world,
                                                          int xrow, int yrow, float* __restrict__ xs, float* __restrict__ ys) {
  const int j = blockIdx.x;
  const long long pos = first + (long long)ts->k * batch_global + rank + (long long)j * world;
  const long long item = idx ? (long long)idx[pos] : pos;
  const float* xr = x_all + item * xrow;
  const float* yr = y_all + item * yrow;
  for (int i = threadIdx.x; i < xrow; i += blockDim.x) xs[(long long)j * xrow + i] = xr[i];
  for (int i = threadIdx.x; i < yrow; i += blockDim.x) ys[(long long)j * yrow + i] = yr[i];
}
// Step epilogue: loss (mse of the GLOBAL batch + L2 of the pre-update weights) and mae into the call's accumulators.
__global__ void train_end_kernel(TrainState* ts, const float* sums, float count, float n_global) {
  if (threadIdx.x == 0) {
    ts->acc[0] += (sums[0] / count + sums[3]) * n_global;
    ts->acc[1] += (sums[1] / count) * n_global;
    ts->k += 1u;
  }
}

// Fused gradient all-reduce + optimizer over NVLink peer memory (one CTA per slice of the flat buffer [grads..., sum err^2,
// sum |err|, count]; see comm.cu for the inbox layout).  CTA c of rank r at step t:
//   1. pushes slice c of its buffer into slot [t & 1][r] of EVERY rank's inbox (its own included) with 128-bit stores,
//      __threadfence_system(), then stores t into flag c of that slot (the release of the slice),
//   2. spins until flag c of slots [t & 1][0 .. world) of its OWN inbox holds t (acquire),
//   3. sums the world's copies of the slice in rank order -- the same order on every rank, so the sums, hence the weights, are
//      bit-identical everywhere -- writes the sums back to the buffer (loss epilogue, hp_head_get_grads) and applies the
//      optimizer to the parameters of the slice.
// Two parities suffice: a rank cannot start step t + 2 before every peer has pushed step t + 1, i.e. finished reading step t.
// A wait that lasts ~2 s (a peer that died) raises HP_STATUS_P2P_TIMEOUT and goes on, so a lost rank cannot hang the GPU.
__global__ void __launch_bounds__(256) p2p_allreduce_optimizer_kernel(float* buf, int n_total, int np, void* const* peers, int rank, int world,
                                                                       int cap, const TrainState* ts, const unsigned int* seq, float* w,
                                                                       const float* l2, float* m,
                                                                       float* v, int kind, float lr, float b1, float b2, float eps,
                                                                       unsigned int* status) {
  const int c = blockIdx.x;
  const uint32_t t = *seq;                                    // exchange tag, advanced for this step by train_begin_kernel
  const size_t slot_floats = (size_t)cap + HP_P2P_SLICES;
  const int per = (((n_total + HP_P2P_SLICES - 1) / HP_P2P_SLICES) + 3) & ~3;     // floats per slice, a multiple of 4
  const int lo = c * per, hi = min(n_total, lo + per);
  const size_t slot_off = ((size_t)(t & 1u) * world + rank) * slot_floats;
  // ---- 1. push
  for (int r = 0; r < world; ++r) {
    float* dst = (float*)peers[r] + slot_off;
    for (int i = lo + 4 * threadIdx.x; i < hi; i += 4 * blockDim.x) {
      if (i + 4 <= hi) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(buf + i);
      else for (int j = i; j < hi; ++j) dst[j] = buf[j];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    volatile uint32_t* flag = (volatile uint32_t*)((float*)peers[threadIdx.x] + slot_off + cap) + c;
    *flag = t;
  }
  // ---- 2. wait for the world's slices
  float* mine = (float*)peers[rank];
  if (threadIdx.x < world) {
    volatile uint32_t* flag = (volatile uint32_t*)(mine + ((size_t)(t & 1u) * world + threadIdx.x) * slot_floats + cap) + c;
    const long long t0 = clock64();
    while (*flag != t) {
      if (clock64() - t0 > 4000000000ll) {
        if (status) atomicOr(status, HP_STATUS_P2P_TIMEOUT);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  // ---- 3. rank-ordered sum + optimizer
  const float alpha_t = ts->alpha_t, lr_t = ts->lr_t;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float g = 0.f;
    for (int r = 0; r < world; ++r) g += __ldcv(mine + ((size_t)(t & 1u) * world + r) * slot_floats + i);
    buf[i] = g;
    if (i < np) {
      const float wi = w[i];
      const float gi = g + 2.f * l2[i] * wi;
      if (kind == HP_OPT_SGD) {
        w[i] = wi - lr * gi;
      } else if (kind == HP_OPT_ADAM) {
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
        m[i] = mi;
        v[i] = vi;
        w[i] = wi - (mi * alpha_t) / (sqrtf(vi) + eps);
      } else {
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float ui = fmaxf(b2 * v[i], fabsf(gi));
        m[i] = mi;
        v[i] = ui;
        w[i] = wi - (lr_t * mi) / (ui + eps);
      }
    }
  }
}

static int train_run_one_step(hp_ctx* h, hp_head* hd, const float* x_all, const float* y_all, const int32_t* idx, long long first,
                              int batch_global, int n_local, int rank, int world, int H, int W, const hp_opt_config* opt, uint64_t seed,
                              cudaStream_t st) {
  const int T = H * W, Cin = hd->in_channels, Cout = hd->regs[hd->out_reg].channels, np = hd->n_params;
  TrainState* ts = (TrainState*)hd->tstate.p;
  float* sums = hd->grads.f() + np;
  const bool use_p2p = world > 1 && h->comm.p2p_ready && np + 3 <= h->comm.p2p_cap && !h->p2p_off;
  train_begin_kernel<<<1, 32, 0, st>>>(ts, opt->lr, opt->beta1, opt->beta2, use_p2p ? h->comm.p2p_seq : nullptr);
  h->launches++;
  HP_CUDA(cudaMemsetAsync(hd->grads.p, 0, (size_t)(np + 4) * sizeof(float), st));
  if (n_local > 0) {
    train_gather_kernel<<<n_local, 128, 0, st>>>(ts, x_all, y_all, idx, first, batch_global, rank, world, T * Cin, T * Cout, hd->xs.f(), hd->ys.f());
    h->launches++;
    hd->step_dev = &ts->t;
    int rc = head_forward(h, hd, hd->xs.f(), n_local, T, true, seed, st);
    if (rc == HP_OK) {
      const long long total = (long long)n_local * T * Cout;
      const float* pred = hd->alias[hd->out_reg] == 0 ? hd->xs.f() : hd->acts.f() + hd->reg_off[hd->alias[hd->out_reg]];
      rc = cudaMemsetAsync(hd->gacts.p, 0, hd->gacts.bytes, st) == cudaSuccess ? HP_OK : HP_ERR_CUDA;
      const float inv_count = 1.f / ((float)batch_global * (float)T * (float)Cout);
      mse_loss_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 1024), 256, 0, st>>>(pred, hd->ys.f(), hd->gacts.f() + hd->reg_off[hd->out_reg],
                                                                                             total, inv_count, sums);
      h->launches++;
      if (rc == HP_OK) rc = head_backward(h, hd, hd->xs.f(), n_local, T, seed, st);
    }
    hd->step_dev = nullptr;
    HP_TRY(rc);
  }
  l2_penalty_kernel<<<1, 256, 0, st>>>(hd->params.f(), hd->l2coef.f(), np, sums + 3);      // of the pre-update weights (Keras reports loss before the step)
  if (use_p2p) {
    // gradient exchange and optimizer in ONE kernel over NVLink peer memory
    p2p_allreduce_optimizer_kernel<<<HP_P2P_SLICES, 256, 0, st>>>(hd->grads.f(), np + 3, np, h->comm.peers_dev, rank, world, h->comm.p2p_cap, ts,
                                                                 h->comm.p2p_seq, hd->params.f(), hd->l2coef.f(), hd->m.f(), hd->v.f(), opt->kind, opt->lr, opt->beta1,
                                                                 opt->beta2, opt->eps, (unsigned int*)h->status.p);
  } else {
    if (world > 1) HP_TRY(hp_comm_allreduce_sum(h, hd->grads.f(), (size_t)np + 3, st));   // a single-rank run never touches a communicator the context may hold
    optimizer_kernel<<<ceil_div(np, 256), 256, 0, st>>>(hd->params.f(), hd->grads.f(), hd->l2coef.f(), hd->m.f(), hd->v.f(), np, opt->kind, opt->lr,
                                                       opt->beta1, opt->beta2, opt->eps, 0.f, 0.f, ts);
  }
  train_end_kernel<<<1, 32, 0, st>>>(ts, sums, (float)batch_global * T * Cout, (float)batch_global);
  h->launches += 3;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

void hp_head_run_graphs_free(hp_head* hd) {
  for (auto& g : hd->run_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  hd->run_graphs.clear();
}

int hp_head_train_run_impl(hp_ctx* h, hp_head* hd, const float* x_all, const float* y_all, const int32_t* idx, long long n_items,
                           long long first_item, int batch_global, int n_steps, int rank, int world, int H, int W,
                           const hp_opt_config* opt, uint64_t seed, int flags, double* sums_host, cudaStream_t st) {
  HP_REQUIRE(hd && x_all && y_all && opt && H > 0 && W > 0, HP_ERR_INVALID, "hp_head_train_run: bad arguments");
  HP_REQUIRE(world >= 1 && rank >= 0 && rank < world && batch_global >= 1 && n_steps >= 1, HP_ERR_INVALID, "hp_head_train_run: bad batch / rank arguments");
  HP_REQUIRE(first_item >= 0 && first_item + (long long)n_steps * batch_global <= n_items, HP_ERR_INVALID,
             "hp_head_train_run: %d steps of %d items from item %lld exceed the %lld items of the data set", n_steps, batch_global, first_item, n_items);
  HP_REQUIRE(world == 1 || ((h->comm.comm || h->comm.p2p_ready) && h->comm.nranks == world && h->comm.rank == rank), HP_ERR_STATE,
             "hp_head_train_run: %d ranks but the gradient exchange is not initialised for them (hp_comm_init / hp_p2p_open)", world);
  const int n_local = batch_global > rank ? (batch_global - rank + world - 1) / world : 0;     // rows rank, rank + world, ... of the batch
  const int T = H * W, Cin = hd->in_channels, Cout = hd->regs[hd->out_reg].channels, np = hd->n_params;
  const int n_plan = n_local > 0 ? n_local : 1;
  HP_TRY(head_plan(hd, n_plan, T, true));
  HP_TRY(hd->xs.ensure((size_t)n_plan * T * Cin * sizeof(float)));
  HP_TRY(hd->ys.ensure((size_t)n_plan * T * Cout * sizeof(float)));
  HP_TRY(hd->tstate.ensure(sizeof(TrainState)));
  if (!hd->has_state) {        // SGD keeps no moments, but one code path: the optimizer kernel is handed valid pointers either way
    HP_TRY(hd->m.ensure((size_t)np * sizeof(float)));
    HP_TRY(hd->v.ensure((size_t)np * sizeof(float)));
    HP_CUDA(cudaMemsetAsync(hd->m.p, 0, (size_t)np * sizeof(float), st));
    HP_CUDA(cudaMemsetAsync(hd->v.p, 0, (size_t)np * sizeof(float), st));
    hd->has_state = true;
  }
  TrainState init;
  memset(&init, 0, sizeof(init));
  init.t = hd->step;
  HP_CUDA(cudaMemcpyAsync(hd->tstate.p, &init, sizeof(init), cudaMemcpyHostToDevice, st));   // pageable source: staged before the call returns

  int done = 0;
  if ((flags & HP_TRAIN_GRAPH) && st != nullptr) {
    hp_head::RunGraph key;
    key.x = x_all; key.y = y_all; key.idx = idx; key.first = first_item; key.batch_global = batch_global; key.n_local = n_local;
    key.rank = rank; key.world = world; key.H = H; key.W = W; key.impl = h->impl; key.opt = *opt; key.seed = seed; key.comm = h->comm.comm; key.p2p = (h->comm.p2p_ready && !h->p2p_off) ? 1 : 0;
    hp_head::RunGraph* hit = nullptr;
    for (auto& g : hd->run_graphs)
      if (g.x == key.x && g.y == key.y && g.idx == key.idx && g.first == key.first && g.batch_global == key.batch_global && g.n_local == key.n_local &&
          g.rank == key.rank && g.world == key.world && g.H == key.H && g.W == key.W && g.impl == key.impl && g.seed == key.seed && g.comm == key.comm && g.p2p == key.p2p &&
          memcmp(&g.opt, &key.opt, sizeof(hp_opt_config)) == 0)
        hit = &g;
    if (hit && hit->epoch != g_devbuf_epoch) {
      hp_head_run_graphs_free(hd);
      hit = nullptr;
    }
    if (!hit) {
      // first step of a new key runs as plain launches (kernel attributes, lazily sized scratch), the next one is captured
      HP_TRY(train_run_one_step(h, hd, x_all, y_all, idx, first_item, batch_global, n_local, rank, world, H, W, opt, seed, st));
      done = 1;
      if (hd->run_graphs.size() >= 8) hp_head_run_graphs_free(hd);
      key.epoch = g_devbuf_epoch;
      hd->run_graphs.push_back(key);
      hit = &hd->run_graphs.back();
    }
    if (done < n_steps && !hit->exec) {
      const int64_t before = h->launches;
      HP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const int rc = train_run_one_step(h, hd, x_all, y_all, idx, first_item, batch_global, n_local, rank, world, H, W, opt, seed, st);
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc != HP_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      HP_REQUIRE(ce == cudaSuccess && graph, HP_ERR_CUDA, "hp_head_train_run: stream capture failed: %s", cudaGetErrorString(ce));
      HP_REQUIRE(hit->epoch == g_devbuf_epoch, HP_ERR_STATE, "hp_head_train_run: a buffer was reallocated during capture");
      cudaGraphExec_t exec = nullptr;
      const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      HP_REQUIRE(ie == cudaSuccess, HP_ERR_CUDA, "hp_head_train_run: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
      hit->exec = exec;
      hit->launches = (int)(h->launches - before);
      h->launches = before;
    }
    for (; done < n_steps; ++done) {
      HP_CUDA(cudaGraphLaunch(hit->exec, st));
      h->launches += hit->launches;
    }
  }
  for (; done < n_steps; ++done)
    HP_TRY(train_run_one_step(h, hd, x_all, y_all, idx, first_item, batch_global, n_local, rank, world, H, W, opt, seed, st));
  hd->step += (uint32_t)n_steps;
  if (sums_host) {
    TrainState fin;
    HP_CUDA(cudaMemcpyAsync(&fin, hd->tstate.p, sizeof(fin), cudaMemcpyDeviceToHost, st));
    HP_CUDA(cudaStreamSynchronize(st));
    sums_host[0] = fin.acc[0];
    sums_host[1] = fin.acc[1];
  }
  return HP_OK;
}

uint32_t hp_dropout_hash(uint64_t seed, uint32_t step, uint32_t op_id, uint32_t image, uint32_t channel) {
  return dropout_hash(seed, step, op_id, image, channel);
}
