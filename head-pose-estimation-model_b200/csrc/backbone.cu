// BlazeFace "front" backbone for sm_100a: stem 5x5/s2 conv + 16 fused BlazeBlocks + detector heads.
//
// Reference semantics: the Keras graph embedded in BlazePoser/UnifiedModels/*.h5, called at
// BlazePoser/blazeFaceDetectorH5.py:272 (layer table: SURVEY.md Appendix A; padding rules App. B.1).
//
// Data layout in HBM: activations are NHWC float32 with the channel count rounded up to a multiple
// of 4 (only the 42-channel map of block 4 is padded, to 44, with zeros) so every pixel row is
// 16-byte aligned and all global accesses are 128-bit.
//
// Two kernel families, both CUDA:
//   FAST  : stem_kernel (register-tiled 4 px x 12 cout per thread, input halo tile in smem) and
//           blaze_block_kernel<CINP,COUTP,S> (one persistent kernel per block: cp.async halo tile ->
//           depthwise 3x3 in registers -> smem -> pointwise 1x1 register-tiled 8 px x 4 cout ->
//           + bias + channel-padded / max-pooled skip + ReLU -> 128-bit stores)
//   NAIVE : one thread per output element, used as an on-device cross-check of the fast path.
#include "common.cuh"

const BlockShape kBlazeBlocks[16] = {{24, 24, 1}, {24, 28, 1}, {28, 32, 2}, {32, 36, 1}, {36, 42, 1}, {42, 48, 2},
                                     {48, 56, 1}, {56, 64, 1}, {64, 72, 1}, {72, 80, 1}, {80, 88, 1}, {88, 96, 2},
                                     {96, 96, 1}, {96, 96, 1}, {96, 96, 1}, {96, 96, 1}};

#define DET16_NP 36   // 2 cls + 32 loc, padded to a multiple of 4
#define DET8_NP 104   // 6 cls + 96 loc, padded

// ============================================================================ small device helpers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 fma4s(float a, float4 b, float4 c) {
  return make_float4(fmaf(a, b.x, c.x), fmaf(a, b.y, c.y), fmaf(a, b.z, c.z), fmaf(a, b.w, c.w));
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ============================================================================ NAIVE kernels
__global__ void stem_naive_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ b, float* __restrict__ y, int B, int H, int W, int Ho,
                                  int Wo, int pt, int pl) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)B * Ho * Wo * 24;
  if (idx >= total) return;
  int co = (int)(idx % 24);
  long long p = idx / 24;
  int ox = (int)(p % Wo);
  p /= Wo;
  int oy = (int)(p % Ho);
  int n = (int)(p / Ho);
  float acc = b[co];
  for (int ky = 0; ky < 5; ++ky) {
    int iy = oy * 2 - pt + ky;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < 5; ++kx) {
      int ix = ox * 2 - pl + kx;
      if (ix < 0 || ix >= W) continue;
      const float* px = x + (((long long)n * H + iy) * W + ix) * 3;
      const float* pw = w + ((ky * 5 + kx) * 3) * 24 + co;
      acc = fmaf(px[0], pw[0], acc);
      acc = fmaf(px[1], pw[24], acc);
      acc = fmaf(px[2], pw[48], acc);
    }
  }
  y[idx] = fmaxf(acc, 0.f);
}

__global__ void dw_naive_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                const float* __restrict__ b, float* __restrict__ out, int B, int Hi, int Wi, int Ho,
                                int Wo, int CP, int S, int pt, int pl) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)B * Ho * Wo * CP;
  if (idx >= total) return;
  int c = (int)(idx % CP);
  long long p = idx / CP;
  int ox = (int)(p % Wo);
  p /= Wo;
  int oy = (int)(p % Ho);
  int n = (int)(p / Ho);
  float acc = b[c];
  for (int ky = 0; ky < 3; ++ky) {
    int iy = oy * S - pt + ky;
    if (iy < 0 || iy >= Hi) continue;
    for (int kx = 0; kx < 3; ++kx) {
      int ix = ox * S - pl + kx;
      if (ix < 0 || ix >= Wi) continue;
      acc = fmaf(in[(((long long)n * Hi + iy) * Wi + ix) * CP + c], w[(ky * 3 + kx) * CP + c], acc);
    }
  }
  out[idx] = acc;
}

__global__ void pw_naive_kernel(const float* __restrict__ dw, const float* __restrict__ in,
                                const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ out,
                                int B, int Hi, int Wi, int Ho, int Wo, int CINP, int COUTP, int S) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)B * Ho * Wo * COUTP;
  if (idx >= total) return;
  int co = (int)(idx % COUTP);
  long long p = idx / COUTP;
  int ox = (int)(p % Wo);
  long long p2 = p / Wo;
  int oy = (int)(p2 % Ho);
  int n = (int)(p2 / Ho);
  const float* a = dw + p * CINP;
  float acc = b[co];
  for (int k = 0; k < CINP; ++k) acc = fmaf(a[k], w[k * COUTP + co], acc);
  float skip = 0.f;
  if (co < CINP) {
    if (S == 1) {
      skip = in[(((long long)n * Hi + oy) * Wi + ox) * CINP + co];
    } else {
      skip = -INFINITY;
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          int iy = oy * 2 + dy, ix = ox * 2 + dx;
          if (iy < Hi && ix < Wi) skip = fmaxf(skip, in[(((long long)n * Hi + iy) * Wi + ix) * CINP + co]);
        }
    }
  }
  out[idx] = fmaxf(acc + skip, 0.f);
}

// strip channel padding: src [P][CP] -> dst [P][C]
__global__ void unpad_kernel(const float* __restrict__ src, float* __restrict__ dst, long long P, int CP, int C) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= P * C) return;
  long long p = idx / C;
  int c = (int)(idx % C);
  dst[idx] = src[p * CP + c];
}

// ============================================================================ FAST stem
// Output tile 16x16 per CTA iteration, 128 threads, thread = 4 consecutive x pixels x 12 couts.
#define STEM_TILE 16
#define STEM_IH 35          // 2*16 + 3
#define STEM_ROW 108        // 35*3 = 105 floats, padded to 108 so each row is 16-byte aligned
struct StemParams {
  const float* x;
  const float* w;
  const float* b;
  float* y;
  int B, H, W, Ho, Wo, pt, pl;
  int tiles_x, tiles_y, n_tiles;
};

__global__ void __launch_bounds__(128) stem_kernel(StemParams p) {
  __shared__ __align__(16) float s_w[75 * 24];
  __shared__ __align__(16) float s_b[24];
  __shared__ __align__(16) float s_in[STEM_IH * STEM_ROW];
  const int tid = threadIdx.x;
  for (int i = tid; i < 75 * 24; i += 128) s_w[i] = p.w[i];
  if (tid < 24) s_b[tid] = p.b[tid];

  const int half = tid & 1;       // couts [12*half, 12*half+12)
  const int q = tid >> 1;         // 0..63
  const int ty = q >> 2;          // 0..15
  const int txq = q & 3;          // pixel quad in the row
  const int tpi = p.tiles_x * p.tiles_y;

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int n = tile / tpi;
    const int r = tile - n * tpi;
    const int tyi = r / p.tiles_x;
    const int txi = r - tyi * p.tiles_x;
    const int oy0 = tyi * STEM_TILE, ox0 = txi * STEM_TILE;
    const int gy0 = oy0 * 2 - p.pt, gx0 = ox0 * 2 - p.pl;
    __syncthreads();  // previous tile fully consumed (also orders the weight stores on the first trip)
    for (int i = tid; i < STEM_IH * STEM_ROW; i += 128) {
      const int hy = i / STEM_ROW;
      const int c = i - hy * STEM_ROW;
      const int hx = c / 3;
      const int ci = c - hx * 3;
      const int gy = gy0 + hy, gx = gx0 + hx;
      float v = 0.f;
      if (c < 105 && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
        v = __ldg(p.x + (((long long)n * p.H + gy) * p.W + gx) * 3 + ci);
      s_in[i] = v;
    }
    __syncthreads();

    float4 acc[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int m = 0; m < 3; ++m) acc[j][m] = ld4(s_b + half * 12 + m * 4);

#pragma unroll 1
    for (int ky = 0; ky < 5; ++ky) {
      const float* row = s_in + (2 * ty + ky) * STEM_ROW + 24 * txq;
      float a[36];
#pragma unroll
      for (int v = 0; v < 9; ++v) {
        float4 t = ld4(row + 4 * v);
        a[4 * v + 0] = t.x;
        a[4 * v + 1] = t.y;
        a[4 * v + 2] = t.z;
        a[4 * v + 3] = t.w;
      }
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float* wr = s_w + ((ky * 5 + kx) * 3 + ci) * 24 + half * 12;
          const float4 w0 = ld4(wr), w1 = ld4(wr + 4), w2 = ld4(wr + 8);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float av = a[(2 * j + kx) * 3 + ci];
            acc[j][0] = fma4s(av, w0, acc[j][0]);
            acc[j][1] = fma4s(av, w1, acc[j][1]);
            acc[j][2] = fma4s(av, w2, acc[j][2]);
          }
        }
      }
    }
    const int oy = oy0 + ty;
    if (oy < p.Ho) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ox = ox0 + txq * 4 + j;
        if (ox < p.Wo) {
          float* dst = p.y + (((long long)n * p.Ho + oy) * p.Wo + ox) * 24 + half * 12;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int m = 0; m < 3; ++m) st4(dst + 4 * m, max4(acc[j][m], z));
        }
      }
    }
  }
}

// ============================================================================ FAST fused BlazeBlock
struct BlkParams {
  const float* in;
  float* out;
  const float *dww, *dwb, *pww, *pwb;
  int B, Hin, Win, Hout, Wout, pad_t, pad_l;
  int TH, TW, IMGS, PG;              // tile = IMGS images x TH x TW output pixels; PG = ceil(tile/8)
  int IH, IW;                        // halo tile dims: (TH-1)*S+3, (TW-1)*S+3
  int tiles_y, tiles_x, n_tiles;
  int nbuf;                          // 1 or 2 input-tile buffers
};

template <int CINP>
struct DwStride {  // row stride of the depthwise result in smem; odd multiple of 4 floats => conflict-free PW reads
  static constexpr int value = ((CINP / 4) & 1) ? CINP : CINP + 4;
};

template <int CINP, int COUTP, int S>
__global__ void __launch_bounds__(512) blaze_block_kernel(BlkParams p) {
  constexpr int C4 = CINP / 4;
  constexpr int NG = COUTP / 4;
  constexpr int DWS = DwStride<CINP>::value;
  constexpr int R = 4;                       // depthwise: consecutive output pixels per thread pass
  constexpr int NCOL = (R - 1) * S + 3;

  extern __shared__ __align__(16) float smem[];
  float* s_pww = smem;
  float* s_pwb = s_pww + CINP * COUTP;
  float* s_dww = s_pwb + COUTP;
  float* s_dwb = s_dww + 9 * CINP;
  float* s_dw = s_dwb + CINP;
  float* s_in = s_dw + p.PG * 8 * DWS;

  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid * 4; i < CINP * COUTP; i += nthr * 4) st4(s_pww + i, ld4(p.pww + i));
  for (int i = tid * 4; i < COUTP; i += nthr * 4) st4(s_pwb + i, ld4(p.pwb + i));
  for (int i = tid * 4; i < 9 * CINP; i += nthr * 4) st4(s_dww + i, ld4(p.dww + i));
  for (int i = tid * 4; i < CINP; i += nthr * 4) st4(s_dwb + i, ld4(p.dwb + i));

  const int in_tile_floats = p.IMGS * p.IH * p.IW * CINP;
  const int tpi = p.tiles_y * p.tiles_x;
  const int TPI = p.TH * p.TW;
  const int TP = TPI * p.IMGS;

  auto load_tile = [&](int tile, float* dst) {
    const int ig = tile / tpi;
    const int r = tile - ig * tpi;
    const int tyi = r / p.tiles_x;
    const int txi = r - tyi * p.tiles_x;
    const int n0 = ig * p.IMGS;
    const int gy0 = tyi * p.TH * S - p.pad_t, gx0 = txi * p.TW * S - p.pad_l;
    const int npix = p.IMGS * p.IH * p.IW;
    for (int i = tid; i < npix * C4; i += nthr) {
      const int c4 = i % C4;
      const int pix = i / C4;
      const int hx = pix % p.IW;
      const int t2 = pix / p.IW;
      const int hy = t2 % p.IH;
      const int im = t2 / p.IH;
      const int gy = gy0 + hy, gx = gx0 + hx, n = n0 + im;
      const bool ok = (n < p.B) && (gy >= 0) && (gy < p.Hin) && (gx >= 0) && (gx < p.Win);
      const float* src = ok ? p.in + (((long long)n * p.Hin + gy) * p.Win + gx) * CINP + c4 * 4 : p.in;
      cp_async16(dst + pix * CINP + c4 * 4, src, ok ? 16 : 0);
    }
  };

  int tile = blockIdx.x;
  int buf = 0;
  if (tile < p.n_tiles) load_tile(tile, s_in);
  cp_async_commit();

  for (; tile < p.n_tiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    float* cur = s_in + buf * in_tile_floats;
    if (p.nbuf == 2) {
      if (next < p.n_tiles) load_tile(next, s_in + (buf ^ 1) * in_tile_floats);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    const int ig = tile / tpi;
    const int rr = tile - ig * tpi;
    const int tyi = rr / p.tiles_x;
    const int txi = rr - tyi * p.tiles_x;
    const int n0 = ig * p.IMGS, y0 = tyi * p.TH, x0 = txi * p.TW;

    // ---------------- depthwise 3x3 (+bias): halo tile -> s_dw[pixel][CINP]
    {
      const int c4 = tid % C4;
      const int rslot = tid / C4;
      const int RS = nthr / C4;
      if (rslot < RS) {
        float4 w[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = ld4(s_dww + k * CINP + c4 * 4);
        const float4 bias = ld4(s_dwb + c4 * 4);
        const int rpr = (p.TW + R - 1) / R;
        const int nruns = p.IMGS * p.TH * rpr;
        for (int run = rslot; run < nruns; run += RS) {
          const int txq = run % rpr;
          const int t2 = run / rpr;
          const int ty = t2 % p.TH;
          const int im = t2 / p.TH;
          const int tx0 = txq * R;
          float4 acc[R];
#pragma unroll
          for (int j = 0; j < R; ++j) acc[j] = bias;
          const float* base = cur + ((im * p.IH + ty * S) * p.IW + tx0 * S) * CINP + c4 * 4;
          const int cols_left = p.IW - tx0 * S;   // columns available from tx0*S to the end of the halo row
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const float* row = base + ky * p.IW * CINP;
            float4 v[NCOL];
#pragma unroll
            for (int j = 0; j < NCOL; ++j)
              v[j] = (j < cols_left) ? ld4(row + j * CINP) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) acc[j] = fma4(v[j * S + kx], w[ky * 3 + kx], acc[j]);
          }
          float* dst = s_dw + ((im * p.TH + ty) * p.TW + tx0) * DWS + c4 * 4;
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (tx0 + j < p.TW) st4(dst + j * DWS, acc[j]);
        }
      }
    }
    __syncthreads();

    // ---------------- pointwise 1x1 + bias + skip + ReLU
    if (tid < p.PG * NG) {
      const int ng = tid % NG;
      const int pg = tid / NG;
      float4 acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* arow = s_dw + pg * DWS;
      const int astep = p.PG * DWS;
      const float* wcol = s_pww + ng * 4;
#pragma unroll 2
      for (int k = 0; k < CINP; k += 4) {
        const float4 w0 = ld4(wcol + (k + 0) * COUTP);
        const float4 w1 = ld4(wcol + (k + 1) * COUTP);
        const float4 w2 = ld4(wcol + (k + 2) * COUTP);
        const float4 w3 = ld4(wcol + (k + 3) * COUTP);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a = ld4(arow + i * astep + k);
          acc[i] = fma4s(a.x, w0, acc[i]);
          acc[i] = fma4s(a.y, w1, acc[i]);
          acc[i] = fma4s(a.z, w2, acc[i]);
          acc[i] = fma4s(a.w, w3, acc[i]);
        }
      }
      const float4 bias = ld4(s_pwb + ng * 4);
      const bool has_skip = (ng * 4 < CINP);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int pp = pg + i * p.PG;
        if (pp >= TP) continue;
        const int im = pp / TPI;
        const int r2 = pp - im * TPI;
        const int ty = r2 / p.TW;
        const int tx = r2 - ty * p.TW;
        const int n = n0 + im, oy = y0 + ty, ox = x0 + tx;
        if (n >= p.B || oy >= p.Hout || ox >= p.Wout) continue;
        float4 v = add4(acc[i], bias);
        if (has_skip) {
          if (S == 1) {
            const float* sp = cur + ((im * p.IH + ty + p.pad_t) * p.IW + tx + p.pad_l) * CINP + ng * 4;
            v = add4(v, ld4(sp));
          } else {
            const int hy = 2 * ty + p.pad_t, hx = 2 * tx + p.pad_l;
            const float* sp = cur + ((im * p.IH + hy) * p.IW + hx) * CINP + ng * 4;
            float4 m = ld4(sp);
            const bool okx = (2 * ox + 1 < p.Win), oky = (2 * oy + 1 < p.Hin);
            if (okx) m = max4(m, ld4(sp + CINP));
            if (oky) {
              m = max4(m, ld4(sp + p.IW * CINP));
              if (okx) m = max4(m, ld4(sp + p.IW * CINP + CINP));
            }
            v = add4(v, m);
          }
        }
        v = max4(v, make_float4(0.f, 0.f, 0.f, 0.f));
        st4(p.out + (((long long)n * p.Hout + oy) * p.Wout + ox) * COUTP + ng * 4, v);
      }
    }
    __syncthreads();
    if (p.nbuf == 1) {
      if (next < p.n_tiles) load_tile(next, s_in);
      cp_async_commit();
    } else {
      buf ^= 1;
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------- host: tile selection
struct TileCfg {
  int TH, TW, IMGS, PG, threads, nbuf, IH, IW, tiles_y, tiles_x, n_tiles;
  size_t smem;
};

static size_t block_smem_bytes(int CINP, int COUTP, int PG, int in_tile_floats, int nbuf) {
  int C4 = CINP / 4;
  int DWS = (C4 & 1) ? CINP : CINP + 4;
  size_t fl = (size_t)CINP * COUTP + COUTP + 10 * CINP + (size_t)PG * 8 * DWS + (size_t)nbuf * in_tile_floats;
  return fl * sizeof(float);
}

static bool fill_tile(int B, int Hout, int Wout, int S, int CINP, int COUTP, int TH, int TW, int IMGS, int nbuf,
                      TileCfg* tc) {
  const int NG = COUTP / 4;
  const int TP = TH * TW * IMGS;
  if (TH < 1 || TW < 1 || IMGS < 1 || TP < 8 || TP > 256 || (nbuf != 1 && nbuf != 2)) return false;
  const int PG = ceil_div(TP, 8);
  const int threads = round_up(PG * NG, 32);
  if (threads > 512 || threads < CINP / 4) return false;
  const int IH = (TH - 1) * S + 3, IW = (TW - 1) * S + 3;
  const size_t smem = block_smem_bytes(CINP, COUTP, PG, IMGS * IH * IW * CINP, nbuf);
  if (smem > 227 * 1024) return false;
  tc->TH = TH; tc->TW = TW; tc->IMGS = IMGS; tc->PG = PG; tc->threads = threads; tc->nbuf = nbuf;
  tc->IH = IH; tc->IW = IW; tc->tiles_y = ceil_div(Hout, TH); tc->tiles_x = ceil_div(Wout, TW);
  tc->n_tiles = tc->tiles_y * tc->tiles_x * ceil_div(B, IMGS);
  tc->smem = smem;
  return true;
}

// Heuristic cost: wasted pixel slots x halo re-read x an occupancy penalty (resident warps per SM).
static bool choose_tile(int B, int Hout, int Wout, int S, int CINP, int COUTP, TileCfg* best) {
  const size_t kMaxSmem = 227 * 1024;
  double best_cost = 1e30;
  bool found = false;
  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: whole (small) images, several per tile; pass 1: one image split into TH x TW tiles
    const int tw_lo = pass == 0 ? Wout : 4, tw_hi = pass == 0 ? Wout : (Wout < 64 ? Wout : 64);
    for (int TW = tw_lo; TW <= tw_hi; ++TW) {
      for (int TH = (pass == 0 ? Hout : 1); TH <= Hout; ++TH) {
        for (int IMGS = 1; IMGS <= (pass == 0 ? 8 : 1); ++IMGS) {
          if (TH * TW * IMGS < 32 && !(pass == 0 && IMGS == 8)) continue;
          for (int nbuf = 2; nbuf >= 1; --nbuf) {
            TileCfg tc;
            if (!fill_tile(B, Hout, Wout, S, CINP, COUTP, TH, TW, IMGS, nbuf, &tc)) continue;
            const double slots = (double)tc.n_tiles * tc.PG * 8;
            const double waste = slots / ((double)B * Hout * Wout);
            const double halo = (double)(tc.IH * tc.IW) / (double)(TH * TW * S * S);
            int ctas = (int)(kMaxSmem / (tc.smem + 1024));
            if (ctas > 2048 / tc.threads) ctas = 2048 / tc.threads;
            if (ctas > 16) ctas = 16;
            const int warps = ctas * (tc.threads / 32);
            double occ = 1.0;
            if (warps < 4) occ = 2.5;
            else if (warps < 6) occ = 1.6;
            else if (warps < 8) occ = 1.3;
            else if (warps < 12) occ = 1.15;
            else if (warps < 16) occ = 1.05;
            double cost = waste * (1.0 + 0.10 * (halo - 1.0)) * occ;
            if (nbuf == 1) cost *= (ctas >= 2 ? 1.04 : 1.3);
            if (cost < best_cost - 1e-9) {
              best_cost = cost;
              found = true;
              *best = tc;
            }
          }
        }
      }
    }
  }
  return found;
}

template <int CINP, int COUTP, int S>
static int launch_block_t(hp_ctx* h, const BlkParams& bp, const TileCfg& tc, cudaStream_t st) {
  auto kern = blaze_block_kernel<CINP, COUTP, S>;
  static bool attr_set = false;
  if (!attr_set) {
    HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int occ = 0;
  HP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, tc.threads, tc.smem));
  HP_REQUIRE(occ >= 1, HP_ERR_CUDA, "blaze block <%d,%d,%d>: zero occupancy (threads %d smem %zu)", CINP, COUTP, S,
             tc.threads, tc.smem);
  long long grid = (long long)h->num_sms * occ;
  if (grid > tc.n_tiles) grid = tc.n_tiles;
  kern<<<(unsigned)grid, tc.threads, tc.smem, st>>>(bp);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

static int launch_block(hp_ctx* h, int blk, const BlkParams& bp, const TileCfg& tc, cudaStream_t st) {
  switch (blk) {
    case 0: return launch_block_t<24, 24, 1>(h, bp, tc, st);
    case 1: return launch_block_t<24, 28, 1>(h, bp, tc, st);
    case 2: return launch_block_t<28, 32, 2>(h, bp, tc, st);
    case 3: return launch_block_t<32, 36, 1>(h, bp, tc, st);
    case 4: return launch_block_t<36, 44, 1>(h, bp, tc, st);
    case 5: return launch_block_t<44, 48, 2>(h, bp, tc, st);
    case 6: return launch_block_t<48, 56, 1>(h, bp, tc, st);
    case 7: return launch_block_t<56, 64, 1>(h, bp, tc, st);
    case 8: return launch_block_t<64, 72, 1>(h, bp, tc, st);
    case 9: return launch_block_t<72, 80, 1>(h, bp, tc, st);
    case 10: return launch_block_t<80, 88, 1>(h, bp, tc, st);
    case 11: return launch_block_t<88, 96, 2>(h, bp, tc, st);
    default: return launch_block_t<96, 96, 1>(h, bp, tc, st);
  }
}

// ============================================================================ weights
int hp_backbone_load_weights_impl(hp_ctx* h, const float* src, size_t n_floats, int layout_id) {
  HP_REQUIRE(layout_id == 0, HP_ERR_INVALID, "hp_backbone_load_weights: unknown layout_id %d", layout_id);
  HP_REQUIRE(src != nullptr, HP_ERR_INVALID, "hp_backbone_load_weights: null pointer");
  HP_REQUIRE(n_floats == HP_BACKBONE_PARAMS, HP_ERR_INVALID,
             "hp_backbone_load_weights: expected %d floats, got %zu", HP_BACKBONE_PARAMS, n_floats);
  std::vector<float> host;
  auto reserve = [&](size_t n) {
    size_t off = host.size();
    host.resize(off + round_up((int)n, 4), 0.f);
    return off;
  };
  size_t cur = 0;
  size_t o_stem_w = reserve(75 * 24);
  memcpy(&host[o_stem_w], src + cur, 75 * 24 * sizeof(float));
  cur += 75 * 24;
  size_t o_stem_b = reserve(24);
  memcpy(&host[o_stem_b], src + cur, 24 * sizeof(float));
  cur += 24;
  size_t o_stem_bhi = reserve(hp_stem_tc_weight_floats()), o_stem_blo = reserve(hp_stem_tc_weight_floats());
  {
    std::vector<float> sw(host.begin() + o_stem_w, host.begin() + o_stem_w + 75 * 24);
    hp_stem_tc_split_weights(sw.data(), &host[o_stem_bhi], &host[o_stem_blo]);
  }
  size_t o_dww[16], o_dwb[16], o_pww[16], o_pwb[16], o_bhi[16], o_blo[16], o_hhi[16], o_hlo[16];
  float h_unscale[16];
  for (int i = 0; i < 16; ++i) {
    const int cin = kBlazeBlocks[i].cin, cout = kBlazeBlocks[i].cout;
    const int cinp = chan_pad(cin), coutp = chan_pad(cout);
    o_dww[i] = reserve(9 * cinp);
    for (int k = 0; k < 9; ++k)
      for (int c = 0; c < cin; ++c) host[o_dww[i] + k * cinp + c] = src[cur + k * cin + c];
    cur += 9 * cin;
    o_dwb[i] = reserve(cinp);
    for (int c = 0; c < cin; ++c) host[o_dwb[i] + c] = src[cur + c];
    cur += cin;
    o_pww[i] = reserve(cinp * coutp);
    for (int k = 0; k < cin; ++k)
      for (int c = 0; c < cout; ++c) host[o_pww[i] + k * coutp + c] = src[cur + k * cout + c];
    cur += (size_t)cin * cout;
    o_pwb[i] = reserve(coutp);
    for (int c = 0; c < cout; ++c) host[o_pwb[i] + c] = src[cur + c];
    cur += cout;
    const int ntc = hp_tc_weight_floats(cinp, coutp);
    o_bhi[i] = reserve(ntc);
    o_blo[i] = reserve(ntc);
    {
      std::vector<float> pw(host.begin() + o_pww[i], host.begin() + o_pww[i] + (size_t)cinp * coutp);
      hp_tc_split_weights(pw.data(), cinp, coutp, &host[o_bhi[i]], &host[o_blo[i]]);
      const int nh = hp_tc_weight_floats_f16(cinp, coutp);
      o_hhi[i] = reserve(nh);
      o_hlo[i] = reserve(nh);
      h_unscale[i] = hp_tc_split_weights_f16(pw.data(), cinp, coutp, &host[o_hhi[i]], &host[o_hlo[i]]);
    }
  }
  // detector heads: concatenate cls|loc per tap
  const float* cls16_k = src + cur; cur += 88 * 2;
  const float* cls16_b = src + cur; cur += 2;
  const float* cls8_k = src + cur;  cur += 96 * 6;
  const float* cls8_b = src + cur;  cur += 6;
  const float* loc16_k = src + cur; cur += 88 * 32;
  const float* loc16_b = src + cur; cur += 32;
  const float* loc8_k = src + cur;  cur += 96 * 96;
  const float* loc8_b = src + cur;  cur += 96;
  HP_REQUIRE(cur == HP_BACKBONE_PARAMS, HP_ERR_INVALID, "internal: packed size mismatch %zu", cur);
  size_t o_d16w = reserve(88 * DET16_NP), o_d16b = reserve(DET16_NP);
  for (int k = 0; k < 88; ++k) {
    for (int c = 0; c < 2; ++c) host[o_d16w + k * DET16_NP + c] = cls16_k[k * 2 + c];
    for (int c = 0; c < 32; ++c) host[o_d16w + k * DET16_NP + 2 + c] = loc16_k[k * 32 + c];
  }
  for (int c = 0; c < 2; ++c) host[o_d16b + c] = cls16_b[c];
  for (int c = 0; c < 32; ++c) host[o_d16b + 2 + c] = loc16_b[c];
  size_t o_d8w = reserve(96 * DET8_NP), o_d8b = reserve(DET8_NP);
  for (int k = 0; k < 96; ++k) {
    for (int c = 0; c < 6; ++c) host[o_d8w + k * DET8_NP + c] = cls8_k[k * 6 + c];
    for (int c = 0; c < 96; ++c) host[o_d8w + k * DET8_NP + 6 + c] = loc8_k[k * 96 + c];
  }
  for (int c = 0; c < 6; ++c) host[o_d8b + c] = cls8_b[c];
  for (int c = 0; c < 96; ++c) host[o_d8b + 6 + c] = loc8_b[c];

  Backbone& bb = h->bb;
  HP_TRY(bb.arena.ensure(host.size() * sizeof(float)));
  HP_CUDA(cudaMemcpy(bb.arena.p, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  const float* base = bb.arena.f();
  bb.stem_w = base + o_stem_w;
  bb.stem_b = base + o_stem_b;
  bb.stem_bhi = base + o_stem_bhi;
  bb.stem_blo = base + o_stem_blo;
  for (int i = 0; i < 16; ++i) {
    bb.blk[i].dww = base + o_dww[i];
    bb.blk[i].dwb = base + o_dwb[i];
    bb.blk[i].pww = base + o_pww[i];
    bb.blk[i].pwb = base + o_pwb[i];
    bb.blk[i].bhi = base + o_bhi[i];
    bb.blk[i].blo = base + o_blo[i];
    bb.blk[i].hhi = base + o_hhi[i];
    bb.blk[i].hlo = base + o_hlo[i];
    bb.blk[i].h_unscale = h_unscale[i];
  }
  {
    size_t n = 0, off[16];
    for (int i = 0; i < 16; ++i) { off[i] = n; n += 10 * (size_t)chan_pad(kBlazeBlocks[i].cin); }
    bb.host_dw.assign(n, 0.f);
    for (int i = 0; i < 16; ++i) {
      const int cinp = chan_pad(kBlazeBlocks[i].cin);
      memcpy(&bb.host_dw[off[i]], &host[o_dww[i]], 9 * (size_t)cinp * sizeof(float));
      memcpy(&bb.host_dw[off[i] + 9 * (size_t)cinp], &host[o_dwb[i]], (size_t)cinp * sizeof(float));
      bb.blk[i].h_dw = &bb.host_dw[off[i]];
    }
  }
  bb.det16_w = base + o_d16w;
  bb.det16_b = base + o_d16b;
  bb.det8_w = base + o_d8w;
  bb.det8_b = base + o_d8b;
  bb.loaded = true;
  return HP_OK;
}

// ============================================================================ forward
struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
};

int hp_backbone_run(hp_ctx* h, const float* x, int B, int H, int W, float* feat16, float* feat8, float* cls,
                    float* loc, int stop_after_blk, float* dbg_dst, size_t dbg_floats, float* per_layer_ms,
                    int prof_iters, cudaStream_t st) {
  Backbone& bb = h->bb;
  HP_REQUIRE(bb.loaded, HP_ERR_STATE, "backbone weights not loaded (call hp_backbone_load_weights first)");
  HP_REQUIRE(x != nullptr && B > 0 && H >= 16 && W >= 16, HP_ERR_INVALID,
             "hp_backbone_forward: need x != NULL, B > 0 and H, W >= 16 (got B=%d H=%d W=%d)", B, H, W);
  HP_REQUIRE((long long)B * H * W * 28 < (1ll << 40), HP_ERR_INVALID, "batch too large");

  // ---- layer geometry
  int hs[18], ws[18];  // hs[0] = stem output, hs[i+1] = block i output
  int pt_stem, pl_stem;
  same_pad(H, 5, 2, &hs[0], &pt_stem);
  same_pad(W, 5, 2, &ws[0], &pl_stem);
  int pts[16], pls[16];
  size_t max_act = (size_t)B * hs[0] * ws[0] * 24;
  for (int i = 0; i < 16; ++i) {
    same_pad(hs[i], 3, kBlazeBlocks[i].stride, &hs[i + 1], &pts[i]);
    same_pad(ws[i], 3, kBlazeBlocks[i].stride, &ws[i + 1], &pls[i]);
    size_t a = (size_t)B * hs[i + 1] * ws[i + 1] * chan_pad(kBlazeBlocks[i].cout);
    if (a > max_act) max_act = a;
  }
  const int H16 = hs[11], W16 = ws[11], H8 = hs[16], W8 = ws[16];
  HP_TRY(bb.act[0].ensure(max_act * sizeof(float)));
  HP_TRY(bb.act[1].ensure(max_act * sizeof(float)));
  if (!feat16) {
    HP_TRY(bb.feat16.ensure((size_t)B * H16 * W16 * 88 * sizeof(float)));
    feat16 = bb.feat16.f();
  }
  if (!feat8) {
    HP_TRY(bb.feat8.ensure((size_t)B * H8 * W8 * 96 * sizeof(float)));
    feat8 = bb.feat8.f();
  }
  const bool naive = (h->impl == HP_IMPL_NAIVE);
  if (naive) HP_TRY(bb.dwtmp.ensure(max_act * sizeof(float)));

  const bool prof = (per_layer_ms != nullptr);
  if (prof && !h->ev[0]) {
    HP_CUDA(cudaEventCreate(&h->ev[0]));
    HP_CUDA(cudaEventCreate(&h->ev[1]));
  }
  const int iters = prof ? (prof_iters > 0 ? prof_iters : 1) : 1;
  auto dbg_copy = [&](const float* src, long long P, int CP, int C) -> int {
    HP_REQUIRE(dbg_dst && dbg_floats >= (size_t)(P * C), HP_ERR_INVALID, "read_activation: dst too small (need %lld)",
               P * C);
    long long total = P * C;
    unpad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, dbg_dst, P, CP, C);
    h->launches++;
    HP_CUDA(cudaGetLastError());
    return HP_OK;
  };

  // ---- stem
  float* cur = bb.act[0].f();
  {
    for (int it = 0; it < iters; ++it) {
      if (prof && it == 0) HP_CUDA(cudaEventRecord(h->ev[0], st));
      if (naive) {
        long long total = (long long)B * hs[0] * ws[0] * 24;
        stem_naive_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, bb.stem_w, bb.stem_b, cur, B, H, W, hs[0],
                                                                         ws[0], pt_stem, pl_stem);
      } else if (h->impl == HP_IMPL_FAST && h->stem_tc_cfg[0] >= 0 && hp_stem_tc_supported(H, W)) {
        HP_TRY(hp_launch_stem_tc(h, x, cur, B, H, W, bb.stem_bhi, bb.stem_blo, bb.stem_b, h->stem_tc_cfg, st));
        h->launches--;   // counted below
      } else {
        StemParams sp;
        sp.x = x; sp.w = bb.stem_w; sp.b = bb.stem_b; sp.y = cur;
        sp.B = B; sp.H = H; sp.W = W; sp.Ho = hs[0]; sp.Wo = ws[0]; sp.pt = pt_stem; sp.pl = pl_stem;
        sp.tiles_x = ceil_div(ws[0], STEM_TILE); sp.tiles_y = ceil_div(hs[0], STEM_TILE);
        long long nt = (long long)sp.tiles_x * sp.tiles_y * B;
        sp.n_tiles = (int)nt;
        long long grid = (long long)h->num_sms * 8;
        if (grid > nt) grid = nt;
        stem_kernel<<<(unsigned)grid, 128, 0, st>>>(sp);
      }
      h->launches++;
    }
    HP_CUDA(cudaGetLastError());
    if (prof) {
      HP_CUDA(cudaEventRecord(h->ev[1], st));
      HP_CUDA(cudaEventSynchronize(h->ev[1]));
      float ms = 0;
      HP_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
      per_layer_ms[0] = ms / iters;
    }
  }
  if (stop_after_blk == -1) return dbg_copy(cur, (long long)B * hs[0] * ws[0], 24, 24);

  // ---- 16 BlazeBlocks
  int pp = 0;  // index of the ping-pong buffer holding `cur`
  for (int i = 0; i < 16; ++i) {
    const int cin = kBlazeBlocks[i].cin, cout = kBlazeBlocks[i].cout, S = kBlazeBlocks[i].stride;
    const int cinp = chan_pad(cin), coutp = chan_pad(cout);
    // ---- cross-block fusion: blocks 6-10 (tap 16) and 12-15 (tap 8) as one persistent kernel each (blocks_chain.cu)
    if (h->impl == HP_IMPL_FAST && (h->chain_mode & 3) > 0 && (i == 6 || i == 12)) {
      int last = (i == 6) ? 10 : 15;
      const int chain_nblk = last - i + 1;
      if (stop_after_blk >= i && stop_after_blk < last) last = stop_after_blk;
      bool plain = true;
      for (int b = i; b <= last; ++b) plain = plain && h->tc_override[b][0] == 0 && h->tile_override[b][0] == 0;
      ChainCfg ccfg;
      // the split-fp16 kernels need half the weight ring of the 3xTF32 ones (the same test as in hp_launch_chain)
      const bool chain_f16 = !(h->chain_mode & 4) && (h->chain_cfg[0] <= 0 || h->chain_cfg[0] % 2 == 0) && h->chain_cfg[1] <= 2;
      // block 11 (stride 2, 12x12x88 -> 6x6x96 at 96x96 input) rides on the chain 6-10: it is computed from the resident tile
      // while the tile's TMA store is in flight (chain_mode 2 = default; 1 = chains without the tail)
      int tail_blk = -1;
      if (i == 6 && last == 10 && (h->chain_mode & 3) >= 2 && (stop_after_blk < 0 || stop_after_blk > 10) && h->tc_override[11][0] == 0 &&
          h->tile_override[11][0] == 0 && plain && hp_chain_geometry(i, chain_nblk, chain_nblk, hs[i + 1], ws[i + 1], &ccfg, 11, chain_f16))
        tail_blk = 11;
      if (plain && (tail_blk >= 0 || hp_chain_geometry(i, last - i + 1, chain_nblk, hs[i + 1], ws[i + 1], &ccfg, -1, chain_f16))) {
        if (h->chain_cfg[0] > 0) ccfg.nsets = h->chain_cfg[0];
        if (h->chain_cfg[1] > 0) ccfg.niss = h->chain_cfg[1] < ccfg.TR ? h->chain_cfg[1] : ccfg.TR;
        float* cout_buf = (last == 10) ? feat16 : (last == 15) ? feat8 : bb.act[pp ^ 1].f();
        float* tail_buf = tail_blk >= 0 ? bb.act[pp ^ 1].f() : nullptr;
        for (int it = 0; it < iters; ++it) {
          if (prof && it == 0) HP_CUDA(cudaEventRecord(h->ev[0], st));
          HP_TRY(hp_launch_chain(h, i, last - i + 1, cur, cout_buf, B, hs[i + 1], ws[i + 1], ccfg, st, tail_blk, tail_buf));
        }
        const int last_done = tail_blk >= 0 ? tail_blk : last;
        if (prof) {
          HP_CUDA(cudaEventRecord(h->ev[1], st));
          HP_CUDA(cudaEventSynchronize(h->ev[1]));
          float ms = 0;
          HP_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
          // the chain is one launch: its time is booked on its first block, the others report 0
          for (int b = i; b <= last_done; ++b) per_layer_ms[1 + b] = (b == i) ? ms / iters : 0.f;
        }
        if (h->tile_report)
          for (int b = i; b <= last_done; ++b) {
            int* r = h->tile_report + 8 * b;
            r[0] = ccfg.TR; r[1] = ccfg.NI; r[2] = ccfg.PS; r[3] = ccfg.lanes; r[4] = ccfg.nsets; r[5] = (int)ccfg.smem; r[6] = ccfg.niss; r[7] = -100 - i;
          }
        cur = cout_buf;
        if (last != 10 && last != 15) pp ^= 1;
        if (stop_after_blk == last)
          return dbg_copy(cur, (long long)B * hs[last + 1] * ws[last + 1], chan_pad(kBlazeBlocks[last].cout), kBlazeBlocks[last].cout);
        if (tail_blk >= 0) {
          cur = tail_buf;
          pp ^= 1;
          if (stop_after_blk == tail_blk)
            return dbg_copy(cur, (long long)B * hs[tail_blk + 1] * ws[tail_blk + 1], chan_pad(kBlazeBlocks[tail_blk].cout), kBlazeBlocks[tail_blk].cout);
          last = tail_blk;
        }
        i = last;
        continue;
      }
    }
    float* out;
    if (i == 10) out = feat16;
    else if (i == 15) out = feat8;
    else out = bb.act[pp ^ 1].f();
    const int Hi = hs[i], Wi = ws[i], Ho = hs[i + 1], Wo = ws[i + 1];
    TileCfg tc;
    Tile2Cfg tc2;
    TcCfg tcc;
    const bool v1 = (h->impl == HP_IMPL_CPASYNC);
    bool use_tc = false, use_tc_s2 = false;
    if (h->impl == HP_IMPL_FAST && S == 2 && h->tc_override[i][0] >= 0) {
      // stride-2 tensor-core kernel; override: TR 1 selects it explicitly with nsets / epilogue sets / buffers / unit
      const int* tv = h->tc_override[i];
      // default: 2 depthwise + 2 epilogue warp sets (96 registers per thread; 0.317 / 0.144 ms on blocks 2 / 5 against 0.327 / 0.146
      // with 3 sets, profiles/r01/tc_sweep_s2_96.log).  Output maps under 100 pixels (block 11: 36 of 128 lanes per 6 x 6 image,
      // 67 KB of split weights next to 62 KB input bands) are no faster than the CUDA-core kernel and stay there unless asked for.
      if (tv[0] > 0 || Ho * Wo >= 100)
        use_tc_s2 = hp_tcs2_geometry(i, Ho, Wo, tv[0] > 0 && tv[4] > 0 ? tv[4] : 2, tv[0] > 0 && tv[3] > 0 ? tv[3] : 2, &tcc, tv[0] > 0 ? tv[2] : 0);
      HP_REQUIRE(use_tc_s2 || tv[0] == 0, HP_ERR_INVALID, "tc override for block %d: the stride-2 kernel does not fit a %dx%d map", i, Ho, Wo);
      if (use_tc_s2 && tv[0] > 0) {
        if (tv[5] >= 2 && tv[5] < tcc.nbuf) tcc.nbuf = tv[5];
        if (tv[6] > 0) tcc.unit = tv[6];
      }
    }
    if (h->impl == HP_IMPL_FAST && S == 1 && h->tc_override[i][0] >= 0) {
      const int* tv = h->tc_override[i];   // TR, NSTG, BH, npipe (epilogue sets when nbuf > 0), nsets, nbuf, unit, issuers
      if (tv[0] > 0) {                      // explicit geometry (tuning / tests): also on maps the default leaves to the CUDA cores
        use_tc = true;
        if (tv[5] > 0 && tv[0] == 1) {      // pixel-per-lane kernel (maps of at most 128 pixels)
          HP_REQUIRE(hp_tcs_geometry(i, Ho, Wo, tv[4] > 0 ? tv[4] : 3, tv[3] > 0 ? tv[3] : 2, &tcc), HP_ERR_INVALID,
                     "tc override for block %d: the pixel-per-lane kernel does not fit a %dx%d map", i, Ho, Wo);
          if (tv[5] < tcc.nbuf) tcc.nbuf = tv[5];
          if (tv[6] > 0) tcc.unit = tv[6];
          HP_REQUIRE(hp_tc_fits(i, Ho, Wo, tcc), HP_ERR_INVALID, "tc override for block %d does not fit (pixel-per-lane, nsets %d nbuf %d)", i,
                     tcc.nsets, tcc.nbuf);
        } else if (tv[5] > 0) {             // warp-specialised kernel with explicit TR / nsets / esets, rings optional
          HP_REQUIRE(hp_tcd_geometry(i, Ho, Wo, tv[0], tv[4] > 0 ? tv[4] : 2, tv[3] > 0 ? tv[3] : 2, &tcc), HP_ERR_INVALID,
                     "tc override for block %d: TR %d does not fit", i, tv[0]);
          if (tv[1] > 0) tcc.NSTG = tv[1];
          if (tv[5] < tcc.nbuf) tcc.nbuf = tv[5];
          if (tv[6] > 0) tcc.unit = tv[6];
          if (tv[7] > 0) tcc.niss = tv[7];
          tcc.place = tv[8];
          HP_REQUIRE(hp_tc_fits(i, Ho, Wo, tcc), HP_ERR_INVALID, "tc override for block %d does not fit (TR %d NSTG %d nsets %d unit %d nbuf %d)", i,
                     tcc.TR, tcc.NSTG, tcc.nsets, tcc.unit, tcc.nbuf);
        } else {                            // pipelined kernel (halo columns, IWB = 7 mod 8)
          tcc.TR = tv[0]; tcc.nbuf = 0; tcc.ni = 1; tcc.unit = 1; tcc.niss = 1; tcc.place = 0;
          tcc.IWB = ((Wo + 2 + 1 + 7) / 8) * 8 - 1;
          tcc.NSTG = tv[1] > 0 ? tv[1] : 2;
          tcc.BH = tv[2] > 0 ? tv[2] : tv[0];
          tcc.npipe = tv[3] > 0 ? tv[3] : 1;
          tcc.nsets = tv[4] > 0 ? tv[4] : 1;
          HP_REQUIRE(hp_tc_fits(i, Ho, Wo, tcc), HP_ERR_INVALID, "tc override for block %d does not fit (TR %d NSTG %d BH %d npipe %d nsets %d)", i,
                     tcc.TR, tcc.NSTG, tcc.BH, tcc.npipe, tcc.nsets);
        }
      } else {
        use_tc = hp_tc_choose(i, Ho, Wo, &tcc);
      }
      if (use_tc && h->tile_report) {
        int* r = h->tile_report + 8 * i;
        r[0] = tcc.TR; r[1] = tcc.BH; r[2] = tcc.IWB; r[3] = tcc.NSTG; r[4] = tcc.nbuf; r[5] = tcc.npipe; r[6] = tcc.nsets; r[7] = -tcc.ni;
      }
    }
    if (!naive && !use_tc && !use_tc_s2) {
      const int* ov = h->tile_override[i];
      if (v1) {
        if (ov[0] > 0) {
          HP_REQUIRE(fill_tile(B, Ho, Wo, S, cinp, coutp, ov[0], ov[1], ov[2], ov[3], &tc), HP_ERR_INVALID,
                     "tile override %dx%dx%d nbuf %d is not valid for block %d", ov[0], ov[1], ov[2], ov[3], i);
        } else {
          HP_REQUIRE(choose_tile(B, Ho, Wo, S, cinp, coutp, &tc), HP_ERR_UNSUPPORTED,
                     "no tile configuration for block %d at %dx%d", i, Ho, Wo);
        }
        if (h->tile_report) {
          int* r = h->tile_report + 8 * i;
          r[0] = tc.TH; r[1] = tc.TW; r[2] = tc.IMGS; r[3] = tc.nbuf; r[4] = tc.threads; r[5] = (int)tc.smem; r[6] = tc.n_tiles; r[7] = 8;
        }
      } else {
        if (ov[0] > 0) {
          HP_REQUIRE(hp_tile2_fill(B, Ho, Wo, S, cinp, coutp, ov[0], ov[1], ov[2], ov[3], ov[4] ? ov[4] : 8, &tc2), HP_ERR_INVALID,
                     "tile override %dx%dx%d nbuf %d MT %d is not valid for block %d", ov[0], ov[1], ov[2], ov[3], ov[4], i);
        } else {
          HP_REQUIRE(hp_tile2_choose(B, Ho, Wo, S, cinp, coutp, &tc2), HP_ERR_UNSUPPORTED,
                     "no tile configuration for block %d at %dx%d", i, Ho, Wo);
        }
        if (h->tile_report) {
          int* r = h->tile_report + 8 * i;
          r[0] = tc2.TH; r[1] = tc2.TW; r[2] = tc2.IMGS; r[3] = tc2.nbuf; r[4] = tc2.threads; r[5] = (int)tc2.smem; r[6] = tc2.n_tiles; r[7] = tc2.MT;
        }
      }
    }
    for (int it = 0; it < iters; ++it) {
      if (prof && it == 0) HP_CUDA(cudaEventRecord(h->ev[0], st));
      if (naive) {
        long long t1 = (long long)B * Ho * Wo * cinp;
        dw_naive_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(cur, bb.blk[i].dww, bb.blk[i].dwb, bb.dwtmp.f(), B,
                                                                    Hi, Wi, Ho, Wo, cinp, S, pts[i], pls[i]);
        long long t2 = (long long)B * Ho * Wo * coutp;
        pw_naive_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(bb.dwtmp.f(), cur, bb.blk[i].pww, bb.blk[i].pwb, out,
                                                                    B, Hi, Wi, Ho, Wo, cinp, coutp, S);
        h->launches += 2;
      } else if (use_tc) {
        HP_TRY(hp_launch_block_tc(h, i, cur, out, B, Ho, Wo, bb.blk[i], tcc, st));
      } else if (use_tc_s2) {
        HP_TRY(hp_launch_block_tc_s2(h, i, cur, out, B, Hi, Wi, Ho, Wo, pts[i], pls[i], bb.blk[i], tcc, st));
      } else if (!v1) {
        HP_TRY(hp_launch_block_tma(h, i, cur, out, B, Hi, Wi, Ho, Wo, pts[i], pls[i], bb.blk[i], tc2, st));
      } else {
        BlkParams bp;
        bp.in = cur; bp.out = out;
        bp.dww = bb.blk[i].dww; bp.dwb = bb.blk[i].dwb; bp.pww = bb.blk[i].pww; bp.pwb = bb.blk[i].pwb;
        bp.B = B; bp.Hin = Hi; bp.Win = Wi; bp.Hout = Ho; bp.Wout = Wo; bp.pad_t = pts[i]; bp.pad_l = pls[i];
        bp.TH = tc.TH; bp.TW = tc.TW; bp.IMGS = tc.IMGS; bp.PG = tc.PG; bp.IH = tc.IH; bp.IW = tc.IW;
        bp.tiles_y = tc.tiles_y; bp.tiles_x = tc.tiles_x; bp.n_tiles = tc.n_tiles; bp.nbuf = tc.nbuf;
        HP_TRY(launch_block(h, i, bp, tc, st));
      }
    }
    HP_CUDA(cudaGetLastError());
    if (prof) {
      HP_CUDA(cudaEventRecord(h->ev[1], st));
      HP_CUDA(cudaEventSynchronize(h->ev[1]));
      float ms = 0;
      HP_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
      per_layer_ms[1 + i] = ms / iters;
    }
    cur = out;
    if (i != 10 && i != 15) pp ^= 1;
    if (i == 10) {
      // block 11 reads feat16; its output goes to the ping-pong buffer not holding anything live
    }
    if (stop_after_blk == i) return dbg_copy(cur, (long long)B * Ho * Wo, coutp, cout);
  }

  // ---- detector heads: cls16|loc16 from feat16, cls8|loc8 from feat8
  if (cls || loc) {
    const int A16 = H16 * W16 * 2, A8 = H8 * W8 * 6, A = A16 + A8;
    if (prof) HP_CUDA(cudaEventRecord(h->ev[0], st));
    for (int it = 0; it < iters; ++it) {
      DenseOut o16[2], o8[2];
      int n16 = 0, n8 = 0;
      if (cls) {
        o16[n16++] = DenseOut{cls, 0, 2, H16 * W16, (long long)A, 2};
        o8[n8++] = DenseOut{cls + A16, 0, 6, H8 * W8, (long long)A, 6};
      }
      if (loc) {
        o16[n16++] = DenseOut{loc, 2, 34, H16 * W16, (long long)A * 16, 32};
        o8[n8++] = DenseOut{loc + (long long)A16 * 16, 6, 102, H8 * W8, (long long)A * 16, 96};
      }
      HP_TRY(hp_launch_dense(h, feat16, B * H16 * W16, 88, 88, bb.det16_w, DET16_NP, bb.det16_b, 34, HP_ACT_LINEAR, false,
                             o16, n16, false, st));
      HP_TRY(hp_launch_dense(h, feat8, B * H8 * W8, 96, 96, bb.det8_w, DET8_NP, bb.det8_b, 102, HP_ACT_LINEAR, false, o8,
                             n8, false, st));
    }
    if (prof) {
      HP_CUDA(cudaEventRecord(h->ev[1], st));
      HP_CUDA(cudaEventSynchronize(h->ev[1]));
      float ms = 0;
      HP_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
      per_layer_ms[17] = ms / iters;
    }
  } else if (prof) {
    per_layer_ms[17] = 0.f;
  }
  return HP_OK;
}
