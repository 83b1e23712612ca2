// Fused BlazeBlock kernel, second generation (the product path): TMA in, TMA out.
//
//   halo tile  : cp.async.bulk.tensor.4d (TMA) from the NHWC activation; out-of-bounds elements are
//                zero-filled by the hardware, which *is* TensorFlow's SAME zero padding for the
//                depthwise conv.  For the stride-2 blocks the same zero fill stands in for the -inf
//                padding of MaxPooling2D(SAME): their input is always a ReLU output (>= 0), so
//                max(valid >= 0, 0) == max over the valid elements.
//   depthwise  : 3x3 (+bias) from the halo tile into smem, 4-pixel runs sharing input columns,
//                run -> offset tables built once per CTA (no integer divisions in the loops)
//   pointwise  : register-tiled GEMM, thread = MT pixels x 4 couts, activations as broadcast LDS.128,
//                weights as conflict-free LDS.128
//   epilogue   : + bias + skip (identity / channel zero-pad / 2x2 max from the halo tile), ReLU, tile
//                written to smem and stored with one TMA bulk tensor store (the hardware clips
//                partial tiles at the image border and batch end).
// Persistent CTAs, optional double buffering of the halo tile through two mbarriers.
//
// Reference semantics: BlazeBlock = DepthwiseConv2D(3x3, SAME) -> Conv2D(1x1) -> Add(skip) -> ReLU of the
// graph in BlazePoser/UnifiedModels/*.h5, called at blazeFaceDetectorH5.py:272.
#include <cuda.h>

#include "common.cuh"

// ---------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
    if (clock64() - t0 > 4000000000ll) __trap();   // ~2 s: never hang the device on a lost TMA
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// packed fp32 FMAs (fma.rn.f32x2): same rounding as scalar FMAs, half the FMA issue slots
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  const float2 lo = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
  const float2 hi = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), make_float2(c.z, c.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 fma4s(float a, float4 b, float4 c) {
  const float2 aa = make_float2(a, a);
  const float2 lo = __ffma2_rn(aa, make_float2(b.x, b.y), make_float2(c.x, c.y));
  const float2 hi = __ffma2_rn(aa, make_float2(b.z, b.w), make_float2(c.z, c.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ---------------------------------------------------------------------------- kernel
struct Blk2Params {
  const float *dww, *dwb, *pww, *pwb;
  int TH, TW, IMGS, PG, TP;            // tile: IMGS x TH x TW output pixels (TP), PG = ceil(TP/MT)
  int IH, IW;                          // halo tile dims
  int tiles_y, tiles_x, n_tiles;
  int nbuf, pad_t, pad_l;
  int n_runs, rpr;                     // depthwise runs of 4 pixels; runs per tile row
  int in_tile_floats;                  // per halo buffer, padded to 32 floats
  uint32_t in_tile_bytes;              // exact TMA transaction size
  // shared-memory layout (float offsets from the 128B-aligned base)
  int off_w, off_tab, off_dw, off_in;
};

template <int CINP>
struct DwStride2 {
  static constexpr int value = ((CINP / 4) & 1) ? CINP : CINP + 4;
};

template <int CINP, int COUTP, int S, int MT>
__global__ void __launch_bounds__(512, (MT == 4 ? 2 : 1))
blaze_block_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, Blk2Params p) {
  constexpr int C4 = CINP / 4;
  constexpr int NG = COUTP / 4;
  constexpr int DWS = DwStride2<CINP>::value;
  constexpr int R = 4;
  constexpr int NCOL = (R - 1) * S + 3;

  extern __shared__ __align__(128) float smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);   // 2 barriers in the first 128 bytes
  float* s_pww = smem + p.off_w;
  float* s_pwb = s_pww + CINP * COUTP;
  float* s_dww = s_pwb + COUTP;
  float* s_dwb = s_dww + 9 * CINP;
  int* run_tab = reinterpret_cast<int*>(smem + p.off_tab);   // [n_runs][2] : halo offset, dw offset | (valid << 24)
  int* pix_tab = run_tab + 2 * p.n_runs;                     // [PG*MT]    : halo offset of the skip pixel
  float* s_dw = smem + p.off_dw;                             // [PG*MT][DWS]  (aliased by the output tile [TP][COUTP])
  float* s_out = s_dw;
  float* s_in = smem + p.off_in;

  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid * 4; i < CINP * COUTP; i += nthr * 4) st4(s_pww + i, ld4(p.pww + i));
  for (int i = tid * 4; i < COUTP; i += nthr * 4) st4(s_pwb + i, ld4(p.pwb + i));
  for (int i = tid * 4; i < 9 * CINP; i += nthr * 4) st4(s_dww + i, ld4(p.dww + i));
  for (int i = tid * 4; i < CINP; i += nthr * 4) st4(s_dwb + i, ld4(p.dwb + i));
  for (int r = tid; r < p.n_runs; r += nthr) {
    const int txq = r % p.rpr;
    const int t2 = r / p.rpr;
    const int ty = t2 % p.TH;
    const int im = t2 / p.TH;
    const int tx0 = txq * R;
    int valid = p.TW - tx0;
    if (valid > R) valid = R;
    run_tab[2 * r + 0] = ((im * p.IH + ty * S) * p.IW + tx0 * S) * CINP;
    run_tab[2 * r + 1] = (((im * p.TH + ty) * p.TW + tx0) * DWS) | (valid << 24);
  }
  for (int pp = tid; pp < p.PG * MT; pp += nthr) {
    int off = 0;
    if (pp < p.TP) {
      const int TPI = p.TH * p.TW;
      const int im = pp / TPI;
      const int r2 = pp - im * TPI;
      const int ty = r2 / p.TW;
      const int tx = r2 - ty * p.TW;
      off = ((im * p.IH + ty * S + p.pad_t) * p.IW + tx * S + p.pad_l) * CINP;
    }
    pix_tab[pp] = off;
  }
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int tpi = p.tiles_y * p.tiles_x;
  auto issue_load = [&](int tile, int buf) {   // thread 0 only
    const int ig = tile / tpi;
    const int r = tile - ig * tpi;
    const int tyi = r / p.tiles_x;
    const int txi = r - tyi * p.tiles_x;
    mbar_expect_tx(&mbar[buf], p.in_tile_bytes);
    tma_load_4d(s_in + buf * p.in_tile_floats, &tm_in, &mbar[buf], 0, txi * p.TW * S - p.pad_l, tyi * p.TH * S - p.pad_t,
                ig * p.IMGS);
  };

  int tile = blockIdx.x;
  if (tid == 0 && tile < p.n_tiles) issue_load(tile, 0);

  int it = 0;
  for (; tile < p.n_tiles; tile += gridDim.x, ++it) {
    const int next = tile + gridDim.x;
    const int buf = (p.nbuf == 2) ? (it & 1) : 0;
    const uint32_t parity = (p.nbuf == 2) ? ((it >> 1) & 1) : (it & 1);
    const float* cur = s_in + buf * p.in_tile_floats;
    if (tid == 0) {
      if (p.nbuf == 2 && next < p.n_tiles) issue_load(next, buf ^ 1);
      tma_store_wait_read();           // previous tile's output has left smem (s_out aliases s_dw)
    }
    mbar_wait(&mbar[buf], parity);
    __syncthreads();

    // ---------------- depthwise 3x3 (+bias): halo tile -> s_dw[pixel][CINP]
    {
      const int c4 = tid % C4;
      const int rslot = tid / C4;
      const int RS = nthr / C4;
      if (rslot < RS) {
        const float4 bias = ld4(s_dwb + c4 * 4);
        for (int run = rslot; run < p.n_runs; run += RS) {
          const int2 e = *reinterpret_cast<const int2*>(run_tab + 2 * run);
          const float* base = cur + e.x + c4 * 4;
          float4 acc[R];
#pragma unroll
          for (int j = 0; j < R; ++j) acc[j] = bias;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const float* row = base + ky * p.IW * CINP;
            const float4 w0 = ld4(s_dww + (ky * 3 + 0) * CINP + c4 * 4);
            const float4 w1 = ld4(s_dww + (ky * 3 + 1) * CINP + c4 * 4);
            const float4 w2 = ld4(s_dww + (ky * 3 + 2) * CINP + c4 * 4);
            float4 v[NCOL];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) v[j] = ld4(row + j * CINP);
#pragma unroll
            for (int j = 0; j < R; ++j) {
              acc[j] = fma4(v[j * S + 0], w0, acc[j]);
              acc[j] = fma4(v[j * S + 1], w1, acc[j]);
              acc[j] = fma4(v[j * S + 2], w2, acc[j]);
            }
          }
          float* dst = s_dw + (e.y & 0xFFFFFF) + c4 * 4;
          const int valid = e.y >> 24;
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (j < valid) st4(dst + j * DWS, acc[j]);
        }
      }
    }
    __syncthreads();

    // ---------------- pointwise 1x1 (register tiled)
    const bool pw_thread = tid < p.PG * NG;
    const int ng = tid % NG;
    const int pg = tid / NG;
    float4 acc[MT];
    if (pw_thread) {
      const float4 bias = ld4(s_pwb + ng * 4);
#pragma unroll
      for (int i = 0; i < MT; ++i) acc[i] = bias;
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
        for (int i = 0; i < MT; ++i) {
          const float4 a = ld4(arow + i * astep + k);
          acc[i] = fma4s(a.x, w0, acc[i]);
          acc[i] = fma4s(a.y, w1, acc[i]);
          acc[i] = fma4s(a.z, w2, acc[i]);
          acc[i] = fma4s(a.w, w3, acc[i]);
        }
      }
    }
    __syncthreads();   // every thread is done reading s_dw: it becomes the output tile

    // ---------------- epilogue: skip + ReLU -> output tile in smem
    if (pw_thread) {
      const bool has_skip = (ng * 4 < CINP);
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int pp = pg + i * p.PG;
        if (pp < p.TP) {
          float4 v = acc[i];
          if (has_skip) {
            const float* sp = cur + pix_tab[pp] + ng * 4;
            float4 m = ld4(sp);
            if (S == 2) {
              m = max4(m, ld4(sp + CINP));
              m = max4(m, ld4(sp + p.IW * CINP));
              m = max4(m, ld4(sp + p.IW * CINP + CINP));
            }
            v = add4(v, m);
          }
          st4(s_out + pp * COUTP + ng * 4, max4(v, zero));
        }
      }
      fence_async_smem();
    }
    __syncthreads();
    if (tid == 0) {
      const int ig = tile / tpi;
      const int r = tile - ig * tpi;
      const int tyi = r / p.tiles_x;
      const int txi = r - tyi * p.tiles_x;
      tma_store_4d(&tm_out, s_out, 0, txi * p.TW, tyi * p.TH, ig * p.IMGS);
      tma_store_commit();
      if (p.nbuf == 1 && next < p.n_tiles) issue_load(next, 0);
    }
  }
  if (tid == 0) tma_store_wait_all();
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int get_encode() {
  if (g_encode) return HP_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  HP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  HP_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, HP_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
  g_encode = (PFN_encodeTiled)fn;
  return HP_OK;
}

// NHWC float tensor [N][H][W][C] with box [bn][bh][bw][C]
static int make_map(CUtensorMap* tm, const float* base, int N, int H, int W, int C, int bn, int bh, int bw) {
  HP_TRY(get_encode());
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HP_REQUIRE(r == CUDA_SUCCESS, HP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for tensor %dx%dx%dx%d box %dx%dx%d", (int)r, N,
             H, W, C, bn, bh, bw);
  return HP_OK;
}

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

bool hp_tile2_fill(int B, int Hout, int Wout, int S, int CINP, int COUTP, int TH, int TW, int IMGS, int nbuf, int MT,
                   Tile2Cfg* tc) {
  const int NG = COUTP / 4, C4 = CINP / 4;
  const int DWS = (C4 & 1) ? CINP : CINP + 4;
  if (TH < 1 || TW < 1 || IMGS < 1 || TH > 256 || TW > 256 || IMGS > 256 || (nbuf != 1 && nbuf != 2) || (MT != 4 && MT != 8))
    return false;
  const int TP = TH * TW * IMGS;
  if (TP < 8 || TP > 512) return false;
  const int PG = ceil_div(TP, MT);
  const int threads = round_up(PG * NG, 32);
  if (threads > 512 || threads < C4) return false;
  const int IH = (TH - 1) * S + 3, IW = (TW - 1) * S + 3;
  if (IH > 256 || IW > 256) return false;
  tc->TH = TH; tc->TW = TW; tc->IMGS = IMGS; tc->MT = MT; tc->nbuf = nbuf; tc->PG = PG; tc->TP = TP; tc->threads = threads;
  tc->IH = IH; tc->IW = IW;
  tc->tiles_y = ceil_div(Hout, TH); tc->tiles_x = ceil_div(Wout, TW);
  tc->n_tiles = tc->tiles_y * tc->tiles_x * ceil_div(B, IMGS);
  tc->rpr = ceil_div(TW, 4);
  tc->n_runs = IMGS * TH * tc->rpr;
  tc->in_tile_floats = align_up(IMGS * IH * IW * CINP, 32);
  int off = 32;                                         // 128 bytes for the two mbarriers
  tc->off_w = off;
  off = align_up(off + CINP * COUTP + COUTP + 10 * CINP, 32);
  tc->off_tab = off;
  off = align_up(off + 2 * tc->n_runs + PG * MT, 32);
  tc->off_dw = off;
  const int dw_fl = PG * MT * DWS, out_fl = TP * COUTP;
  off = align_up(off + (dw_fl > out_fl ? dw_fl : out_fl), 32);
  tc->off_in = off;
  off += nbuf * tc->in_tile_floats + align_up((4 * S + 3) * CINP, 32);   // tail pad: depthwise tail runs read past the row
  tc->smem = (size_t)off * sizeof(float);
  return tc->smem <= 227 * 1024;
}

// Tuned from the tile sweeps under profiles/ (tools/tile_sweep.py); falls back to a cost heuristic.
bool hp_tile2_choose(int B, int Hout, int Wout, int S, int CINP, int COUTP, Tile2Cfg* best) {
  double best_cost = 1e30;
  bool found = false;
  for (int MT = 4; MT <= 8; MT += 4) {
    for (int pass = 0; pass < 2; ++pass) {
      const int tw_lo = pass == 0 ? Wout : 4, tw_hi = pass == 0 ? Wout : (Wout < 64 ? Wout : 64);
      for (int TW = tw_lo; TW <= tw_hi; ++TW) {
        for (int TH = (pass == 0 ? Hout : 1); TH <= Hout; ++TH) {
          for (int IMGS = 1; IMGS <= (pass == 0 ? 8 : 1); ++IMGS) {
            if (TH * TW * IMGS < 32 && !(pass == 0 && IMGS == 8)) continue;
            for (int nbuf = 2; nbuf >= 1; --nbuf) {
              Tile2Cfg tc;
              if (!hp_tile2_fill(B, Hout, Wout, S, CINP, COUTP, TH, TW, IMGS, nbuf, MT, &tc)) continue;
              const double slots = (double)tc.n_tiles * tc.PG * MT;
              const double waste = slots / ((double)B * Hout * Wout);
              const double halo = (double)(tc.IH * tc.IW) / (double)(TH * TW * S * S);
              int ctas = (int)((227 * 1024) / (tc.smem + 1024));
              const int reg_ctas = (MT == 4 ? 65536 / (64 * tc.threads) : 65536 / (128 * tc.threads));
              if (ctas > reg_ctas) ctas = reg_ctas;
              if (ctas > 2048 / tc.threads) ctas = 2048 / tc.threads;
              if (ctas < 1) ctas = 1;
              const int warps = ctas * (tc.threads / 32);
              double occ = 1.0;
              if (warps < 4) occ = 2.5;
              else if (warps < 6) occ = 1.6;
              else if (warps < 8) occ = 1.3;
              else if (warps < 12) occ = 1.15;
              else if (warps < 16) occ = 1.06;
              else if (warps < 24) occ = 1.02;
              double cost = waste * (1.0 + 0.10 * (halo - 1.0)) * occ;
              if (MT == 4) cost *= 1.03;                       // fewer FFMA per LDS
              if (nbuf == 1) cost *= (ctas >= 2 ? 1.02 : 1.25);
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
  }
  return found;
}

template <int CINP, int COUTP, int S, int MT>
static int launch_t(hp_ctx* h, const CUtensorMap& tin, const CUtensorMap& tout, const Blk2Params& bp, const Tile2Cfg& tc,
                    cudaStream_t st) {
  auto kern = blaze_block_tma_kernel<CINP, COUTP, S, MT>;
  HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int occ = 0;
  HP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, tc.threads, tc.smem));
  HP_REQUIRE(occ >= 1, HP_ERR_CUDA, "blaze block <%d,%d,%d,%d>: zero occupancy (threads %d smem %zu)", CINP, COUTP, S, MT,
             tc.threads, tc.smem);
  long long grid = (long long)h->num_sms * occ;
  if (grid > tc.n_tiles) grid = tc.n_tiles;
  kern<<<(unsigned)grid, tc.threads, tc.smem, st>>>(tin, tout, bp);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

template <int CINP, int COUTP, int S>
static int launch_mt(hp_ctx* h, const CUtensorMap& tin, const CUtensorMap& tout, const Blk2Params& bp, const Tile2Cfg& tc,
                     cudaStream_t st) {
  if (tc.MT == 4) return launch_t<CINP, COUTP, S, 4>(h, tin, tout, bp, tc, st);
  return launch_t<CINP, COUTP, S, 8>(h, tin, tout, bp, tc, st);
}

int hp_launch_block_tma(hp_ctx* h, int blk, const float* in, float* out, int B, int Hin, int Win, int Hout, int Wout,
                        int pad_t, int pad_l, const BlockWeights& w, const Tile2Cfg& tc, cudaStream_t st) {
  const int cinp = chan_pad(kBlazeBlocks[blk].cin), coutp = chan_pad(kBlazeBlocks[blk].cout);
  CUtensorMap tin, tout;
  HP_TRY(make_map(&tin, in, B, Hin, Win, cinp, tc.IMGS, tc.IH, tc.IW));
  HP_TRY(make_map(&tout, out, B, Hout, Wout, coutp, tc.IMGS, tc.TH, tc.TW));
  Blk2Params bp;
  bp.dww = w.dww; bp.dwb = w.dwb; bp.pww = w.pww; bp.pwb = w.pwb;
  bp.TH = tc.TH; bp.TW = tc.TW; bp.IMGS = tc.IMGS; bp.PG = tc.PG; bp.TP = tc.TP; bp.IH = tc.IH; bp.IW = tc.IW;
  bp.tiles_y = tc.tiles_y; bp.tiles_x = tc.tiles_x; bp.n_tiles = tc.n_tiles; bp.nbuf = tc.nbuf; bp.pad_t = pad_t; bp.pad_l = pad_l;
  bp.n_runs = tc.n_runs; bp.rpr = tc.rpr; bp.in_tile_floats = tc.in_tile_floats;
  bp.in_tile_bytes = (uint32_t)((size_t)tc.IMGS * tc.IH * tc.IW * cinp * sizeof(float));
  bp.off_w = tc.off_w; bp.off_tab = tc.off_tab; bp.off_dw = tc.off_dw; bp.off_in = tc.off_in;
  switch (blk) {
    case 0: return launch_mt<24, 24, 1>(h, tin, tout, bp, tc, st);
    case 1: return launch_mt<24, 28, 1>(h, tin, tout, bp, tc, st);
    case 2: return launch_mt<28, 32, 2>(h, tin, tout, bp, tc, st);
    case 3: return launch_mt<32, 36, 1>(h, tin, tout, bp, tc, st);
    case 4: return launch_mt<36, 44, 1>(h, tin, tout, bp, tc, st);
    case 5: return launch_mt<44, 48, 2>(h, tin, tout, bp, tc, st);
    case 6: return launch_mt<48, 56, 1>(h, tin, tout, bp, tc, st);
    case 7: return launch_mt<56, 64, 1>(h, tin, tout, bp, tc, st);
    case 8: return launch_mt<64, 72, 1>(h, tin, tout, bp, tc, st);
    case 9: return launch_mt<72, 80, 1>(h, tin, tout, bp, tc, st);
    case 10: return launch_mt<80, 88, 1>(h, tin, tout, bp, tc, st);
    case 11: return launch_mt<88, 96, 2>(h, tin, tout, bp, tc, st);
    default: return launch_mt<96, 96, 1>(h, tin, tout, bp, tc, st);
  }
}
