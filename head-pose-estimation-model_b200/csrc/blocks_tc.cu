// Fused BlazeBlock kernel, third generation: the pointwise 1x1 conv runs on the tcgen05 tensor cores as a
// 3xTF32 GEMM (fp32-level accuracy), everything else of the block stays fused around it.
//
// Why: ncu on the CUDA-core kernels (profiles/r01) shows the backbone bound by fp32 FMA issue, not HBM: the
// pointwise convs are 83 % of the MACs.  Moving them to the tensor pipe leaves the CUDA cores the depthwise 3x3
// (9 FMA per element) and the HBM roofline becomes reachable.  Plain TF32 (10-bit mantissa) cannot meet the 1e-4
// parity bound, so every product is split: a = a_hi + a_lo, w = w_hi + w_lo (each part TF32-representable),
// D += a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with fp32 accumulation in TMEM; the dropped a_lo*w_lo term is ~2^-22.
//
// Per CTA (persistent, 1 per SM): NPIPE independent pipelines, each = 4 worker warps (thread <-> TMEM lane) + 1 issuer warp.
//   tile     : one band of BH output rows x full width W of one image; lane l <-> (strip yq = l / W, column x = l % W),
//              a lane owns the TR output pixels (yq*TR + t, x), t < TR; pixel t of every lane forms M-tile t.
//   load     : one TMA (cp.async.bulk.tensor.4d) per tile: box {PS, IWB, BH+2, 1} at (0, -1, y0-1, img).  PS > C pads the
//              pixel stride to an odd number of 16-byte chunks (bank-conflict-free column access), out-of-range
//              rows / columns / channels are zero-filled by the hardware (= SAME padding).
//   depthwise: per k-step (8 channels) each lane slides a 3x3 window down its column (TR+2 rows, 3 LDS.128 per row and
//              chunk), splits the result into hi/lo and stores 16 columns per M-tile into a TMEM ring stage (tcgen05.st).
//   pointwise: the issuer thread fires 3 x TR tcgen05.mma (A from TMEM, W_hi / W_lo from smem, K-major no-swizzle
//              descriptors) per k-step into TR accumulators D_t[128 x N16]; tcgen05.commit frees the ring stage.
//   epilogue : tcgen05.ld D_t -> + bias + skip (centre pixel of the halo tile) -> ReLU -> written IN PLACE over the centre
//              pixel (PS >= COUTP) -> one TMA store of the band interior (box {PS, IWB, BH, 1}, clipped to C x W x H).
//
// Reference semantics: BlazeBlock = DepthwiseConv2D(3x3, SAME) -> Conv2D(1x1) -> Add(skip / channel-padded skip) -> ReLU
// (graph called at BlazePoser/blazeFaceDetectorH5.py:272).  Stride-1 blocks only.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace {

// ---------------------------------------------------------------------------- compile-time geometry
template <int CINP, int COUTP>
struct TcGeom {
  static constexpr int C4 = CINP / 4;
  static constexpr int NG = COUTP / 4;
  static constexpr int K8 = (CINP + 7) / 8 * 8;
  static constexpr int KS = K8 / 8;
  static constexpr int N16 = (COUTP + 15) / 16 * 16;
  static constexpr int PSC = ((C4 > NG ? C4 : NG) | 1);   // pixel stride in 16-byte chunks: odd, >= both channel counts
  static constexpr int PS = PSC * 4;                      // ... in floats
};

struct TcParams {
  const float *dww, *dwb, *pwb;   // depthwise [9][CINP], [CINP]; pointwise bias [COUTP]
  const float *bhi, *blo;         // pointwise weights split hi / lo, each [K8/4][N16][4] (K-major core matrices)
  int W, H, BH, IWB, row_pitch;   // row_pitch = IWB * PS floats
  int bands_per_img, n_tiles, lanes;
  uint32_t load_bytes;
  int off_b, off_w, off_pipe, buf_floats;   // shared-memory layout in floats (buffers 1024-byte aligned)
  long long* trace;                         // optional per-tile clock64 stamps of CTA 0 (8 per tile), see hp_debug_tc_trace
  int trace_tiles;
};

#define TC_MAX_PIPE 4
#define TC_MAX_STG 4
#define TC_BAR_FLOATS 128   // 512 bytes: barriers of all pipelines + the TMEM base address



// Depthwise 3x3 (+bias) of one 4-channel chunk for the TR pixels of a lane (sliding window down the column).
template <int CINP, int TR, int PS>
__device__ __forceinline__ void tc_dw_compute(const float* win, const float* dww_c, const float* dwb_c, int row_pitch, bool mask_l, bool mask_r,
                                              float4 (&acc)[TR]) {
  float4 w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = ld4(dww_c + k * CINP);
  // SAME padding at the image's left / right edge when the halo buffer has no padding column there: the tap reads a
  // finite value of a neighbouring row and multiplies it by zero
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (mask_l) { w[0] = z4; w[3] = z4; w[6] = z4; }
  if (mask_r) { w[2] = z4; w[5] = z4; w[8] = z4; }
  const float4 bias = ld4(dwb_c);
#pragma unroll
  for (int t = 0; t < TR; ++t) acc[t] = bias;
#pragma unroll
  for (int r = 0; r < TR + 2; ++r) {
    const float* row = win + r * row_pitch;
    const float4 v0 = ld4(row), v1 = ld4(row + PS), v2 = ld4(row + 2 * PS);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int t = r - ky;
      if (t >= 0 && t < TR) {
        acc[t] = fma4(v0, w[ky * 3 + 0], acc[t]);
        acc[t] = fma4(v1, w[ky * 3 + 1], acc[t]);
        acc[t] = fma4(v2, w[ky * 3 + 2], acc[t]);
      }
    }
  }
}
// TF32 hi / lo split of the chunk and store to the A ring stage: hi at columns [acol + 16 t, +4), lo 8 columns further.
template <int TR>
__device__ __forceinline__ void tc_dw_store(const float4 (&acc)[TR], uint32_t acol) {
#pragma unroll
  for (int t = 0; t < TR; ++t) {
    const uint32_t h0 = tf32_hi(acc[t].x), h1 = tf32_hi(acc[t].y), h2 = tf32_hi(acc[t].z), h3 = tf32_hi(acc[t].w);
    tmem_st4(acol + t * 16, h0, h1, h2, h3);
    tmem_st4(acol + t * 16 + 8, __float_as_uint(acc[t].x - __uint_as_float(h0)), __float_as_uint(acc[t].y - __uint_as_float(h1)),
             __float_as_uint(acc[t].z - __uint_as_float(h2)), __float_as_uint(acc[t].w - __uint_as_float(h3)));
  }
}
template <int CINP, int TR, int PS>
__device__ __forceinline__ void tc_dw_chunk(const float* win, const float* dww_c, const float* dwb_c, int row_pitch, uint32_t acol) {
  float4 acc[TR];
  tc_dw_compute<CINP, TR, PS>(win, dww_c, dwb_c, row_pitch, false, false, acc);
  tc_dw_store<TR>(acc, acol);
}
template <int TR>
__device__ __forceinline__ void tc_zero_chunk(uint32_t acol) {   // K padding (odd number of chunks): zero columns
#pragma unroll
  for (int t = 0; t < TR; ++t) {
    tmem_st4(acol + t * 16, 0u, 0u, 0u, 0u);
    tmem_st4(acol + t * 16 + 8, 0u, 0u, 0u, 0u);
  }
}
// Epilogue of one pixel: accumulator row (N16 columns at dcol) + bias + skip -> ReLU -> in place over the centre pixel.
// All column groups are requested before the single tcgen05.wait::ld (the TMEM load latency is paid once per pixel).
// F16: the accumulator carries the power-of-two scale of the fp16 weights (multiplied out here) and, when a depthwise output did
// not fit fp16, inf / NaN in every column: column 0 is tested (guard).
// BREG = 1: the pointwise bias comes from the caller's registers (breg[NG], loaded once per kernel) instead of one warp-uniform
// LDS.128 per 4 channels and pixel -- a broadcast LDS.128 costs the 4 shared-memory wavefronts of a full-width one, a third of
// the epilogue's shared-memory traffic (bias + skip + store).
template <int C4, int NG, int N16, int F16 = 0, int BREG = 0>
__device__ __forceinline__ void tc_epilogue_pixel(float* cpix, uint32_t dcol, const float* s_pwb, bool active, float us = 1.f,
                                                  uint32_t* guard = nullptr, const float4* breg = nullptr) {
  constexpr int NGRP = (NG + 7) / 8;   // 32-column groups that hold real output channels
  uint32_t v[NGRP][32];
#pragma unroll
  for (int g = 0; g < NGRP; ++g) {
    if (g * 32 + 32 <= N16) {
      tmem_ld32(dcol + g * 32, v[g]);
    } else {                             // N16 is a multiple of 16 only: the last group may be half
      uint32_t h[16];
      tmem_ld16(dcol + g * 32, h);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[g][e] = h[e];
    }
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (F16 && active && (v[0][0] & 0x7F800000u) == 0x7F800000u) *guard = 1u;
#pragma unroll
  for (int g = 0; g < NGRP; ++g) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = g * 8 + jj;
      if (j < NG) {
        const float4 bb = BREG ? breg[j] : ld4(s_pwb + j * 4);
        float4 o = F16 ? make_float4(fmaf(__uint_as_float(v[g][jj * 4 + 0]), us, bb.x), fmaf(__uint_as_float(v[g][jj * 4 + 1]), us, bb.y),
                                     fmaf(__uint_as_float(v[g][jj * 4 + 2]), us, bb.z), fmaf(__uint_as_float(v[g][jj * 4 + 3]), us, bb.w))
                       : make_float4(__uint_as_float(v[g][jj * 4 + 0]) + bb.x, __uint_as_float(v[g][jj * 4 + 1]) + bb.y,
                                     __uint_as_float(v[g][jj * 4 + 2]) + bb.z, __uint_as_float(v[g][jj * 4 + 3]) + bb.w);
        if (j < C4) {
          const float4 sk = ld4(cpix + j * 4);
          o.x += sk.x; o.y += sk.y; o.z += sk.z; o.w += sk.w;
        }
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        if (active) st4(cpix + j * 4, o);
      }
    }
  }
}

template <int TR>
__device__ __forceinline__ void tc_dw_store2(const float4 (&a0)[TR], const float4 (&a1)[TR], uint32_t acol) {
#pragma unroll
  for (int t = 0; t < TR; ++t) {
    const float f[8] = {a0[t].x, a0[t].y, a0[t].z, a0[t].w, a1[t].x, a1[t].y, a1[t].z, a1[t].w};
    uint32_t v[16];   // [hi 8 | lo 8] = the 16 columns of the stage row: one tcgen05.st (each costs ~40 clk of issue)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = tf32_hi(f[e]);
      v[8 + e] = __float_as_uint(f[e] - __uint_as_float(v[e]));
    }
    tmem_st16(acol + t * 16, v);
  }
}

#ifdef HP_LEGACY_KERNELS   // first tensor-core generation (pipelines instead of warp roles): test-only builds, see _lib.build
// NSETS warp sets (4 warps each) share one pipeline: the 4-channel chunks of a tile are dealt round-robin to the sets
// (chunk c -> set c % NSETS), the accumulator rows of the epilogue likewise (row t -> set t % NSETS).  More warps per tile
// without more TMEM or shared memory: the kernel is latency bound with one warp per SM sub-partition.
template <int CINP, int COUTP, int TR, int NSTG, int NPIPE, int NSETS>
__global__ void __launch_bounds__(NPIPE * (128 * NSETS + 32), 1)
blaze_block_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, TcParams p) {
  using G = TcGeom<CINP, COUTP>;
  constexpr int C4 = G::C4, NG = G::NG, K8 = G::K8, KS = G::KS, N16 = G::N16, PS = G::PS;
  constexpr int TCOLS = TR * N16 + NSTG * TR * 16;   // TMEM columns per pipeline: TR accumulators + the A ring
  constexpr int WTHREADS = 128 * NSETS;              // worker threads per pipeline
  constexpr int BARS = 2 + 2 * TC_MAX_STG;           // mbarriers per pipeline
  static_assert(NPIPE * TCOLS <= 512 && NPIPE <= TC_MAX_PIPE, "TMEM budget");
  static_assert(NSTG <= TC_MAX_STG && NPIPE * BARS * 8 + 4 <= TC_BAR_FLOATS * 4, "barrier block");

  extern __shared__ __align__(1024) float smem[];
  // barrier block: per pipeline full, d_full, a_full[NSTG], a_empty[NSTG]; TMEM base in the last word
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem) + (TC_BAR_FLOATS - 1);
  float* s_bhi = smem + p.off_b;
  float* s_blo = s_bhi + K8 * N16;
  float* s_dww = smem + p.off_w;
  float* s_dwb = s_dww + 9 * CINP;
  float* s_pwb = s_dwb + CINP;

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane_id = tid & 31;
  constexpr int n_work_warps = 4 * NSETS * NPIPE;

  for (int i = tid * 4; i < K8 * N16; i += nthr * 4) {
    st4(s_bhi + i, ld4(p.bhi + i));
    st4(s_blo + i, ld4(p.blo + i));
  }
  for (int i = tid * 4; i < 9 * CINP; i += nthr * 4) st4(s_dww + i, ld4(p.dww + i));
  for (int i = tid * 4; i < CINP; i += nthr * 4) st4(s_dwb + i, ld4(p.dwb + i));
  for (int i = tid * 4; i < COUTP; i += nthr * 4) st4(s_pwb + i, ld4(p.pwb + i));
  fence_async_smem();   // the tensor core reads s_bhi / s_blo through the async proxy
  if (tid == 0) {
    for (int q = 0; q < NPIPE; ++q) {
      uint64_t* b = bars + q * BARS;
      mbar_init(&b[0], 1);                                     // full: TMA load landed
      mbar_init(&b[1], 1);                                     // d_full: all MMAs of the tile done
      for (int s = 0; s < NSTG; ++s) {
        mbar_init(&b[2 + s], 256);                             // a_full[s]: both 4-channel halves of the k-step stored (128 lanes each)
        mbar_init(&b[2 + TC_MAX_STG + s], 1);                  // a_empty[s]: MMAs that read stage s are done
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == n_work_warps) {   // first issuer warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;

  const int tile_stride = gridDim.x * NPIPE;

  if (warp < n_work_warps) {
    // =============================================================== worker: depthwise -> TMEM, epilogue
    const int pipe = warp / (4 * NSETS);
    const int set = (warp >> 2) % NSETS;
    const int wq = warp & 3;
    const int lane = wq * 32 + lane_id;          // TMEM lane = column of pixels owned by this thread
    const bool leader = (set == 0 && lane == 0);
    uint64_t* b = bars + pipe * BARS;
    uint64_t* bar_full = &b[0];
    uint64_t* bar_dfull = &b[1];
    uint64_t* bar_afull = &b[2];
    uint64_t* bar_aempty = &b[2 + TC_MAX_STG];
    float* buf = smem + p.off_pipe + pipe * p.buf_floats;
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(pipe * TCOLS);
    constexpr uint32_t colA0 = TR * N16;
    const bool active = lane < p.lanes;
    const bool warp_active = wq * 32 < p.lanes;
    const int l = active ? lane : 0;
    const int yq = l / p.W;
    const int x = l - yq * p.W;
    const int row_pitch = p.row_pitch;
    const int my_off = yq * TR * row_pitch + x * PS;    // top-left pixel of the 3x3 window of output row yq*TR
    const int centre0 = my_off + row_pitch + PS;        // centre pixel of output row yq*TR

    auto tile_coords = [&](int tile, int& img, int& y0) {
      img = tile / p.bands_per_img;
      y0 = (tile - img * p.bands_per_img) * p.BH;
    };
    auto issue_load = [&](int tile) {   // leader only
      int img, y0;
      tile_coords(tile, img, y0);
      mbar_expect_tx(bar_full, p.load_bytes);
      tma_load_4d(buf, &tm_in, bar_full, 0, -1, y0 - 1, img);
    };

    int tile = blockIdx.x * NPIPE + pipe;
    if (leader && tile < p.n_tiles) issue_load(tile);
    for (int it = 0; tile < p.n_tiles; tile += tile_stride, ++it) {
      const int next = tile + tile_stride;
      mbar_wait(bar_full, it & 1);
      if (leader && next < p.n_tiles) {   // warm L2 for the next tile of this pipeline (its buffer is still in use)
        int img, y0;
        tile_coords(next, img, y0);
        tma_prefetch_4d(&tm_in, 0, -1, y0 - 1, img);
      }

      // ---------------- depthwise 3x3 (+bias) -> hi/lo split -> TMEM ring stage; this set's chunks only
#pragma unroll 1
      for (int c4 = set; c4 < 2 * KS; c4 += NSETS) {
        const int ks = c4 >> 1, half = c4 & 1;
        const uint32_t use = (uint32_t)it * KS + ks;   // global k-step counter of this pipeline
        const uint32_t s = use % NSTG;
        if (use >= NSTG) {
          mbar_wait(&bar_aempty[s], ((use / NSTG) - 1) & 1);
          tc_fence_after();
        }
        if (warp_active) {
          const uint32_t acol = tlane + colA0 + s * (TR * 16) + half * 4;
          if (c4 < C4) tc_dw_chunk<CINP, TR, PS>(buf + my_off + c4 * 4, s_dww + c4 * 4, s_dwb + c4 * 4, row_pitch, acol);
          else tc_zero_chunk<TR>(acol);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
        }
        mbar_arrive(&bar_afull[s]);
      }

      // ---------------- epilogue: D_t + bias + skip -> ReLU -> in place over the centre pixel; this set's rows only
      mbar_wait(bar_dfull, it & 1);
      tc_fence_after();
      if (warp_active) {
#pragma unroll
        for (int t = 0; t < TR; ++t) {
          if (t % NSETS != set) continue;
          tc_epilogue_pixel<C4, NG, N16>(buf + centre0 + t * row_pitch, tlane + t * N16, s_pwb, active);
        }
        tc_fence_before();
        fence_async_smem();
      }
      named_bar_sync(1 + pipe, WTHREADS);
      if (leader) {
        int img, y0;
        tile_coords(tile, img, y0);
        tma_store_4d(&tm_out, buf + row_pitch + PS, 0, 0, y0, img);
        tma_store_commit();
        tma_store_wait_read();            // the band has left shared memory: the buffer may be refilled
        if (next < p.n_tiles) issue_load(next);
      }
    }
    if (leader) tma_store_wait_all();
  } else if (lane_id == 0) {
    // =============================================================== issuer: tcgen05.mma for one pipeline
    const int pipe = warp - n_work_warps;
    uint64_t* b = bars + pipe * BARS;
    uint64_t* bar_dfull = &b[1];
    uint64_t* bar_afull = &b[2];
    uint64_t* bar_aempty = &b[2 + TC_MAX_STG];
    const uint32_t tcol = tmem_base + (uint32_t)(pipe * TCOLS);
    constexpr uint32_t colA0 = TR * N16;
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = N16, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // smem descriptor of a [N16 rows][8 k] slice: core matrix = 8 rows x 16 B; LBO (next 4 k) = N16*16 B, SBO (next 8 rows) = 128 B
    const uint64_t desc_fixed = ((uint64_t)(((uint32_t)(N16 * 16) >> 4) & 0x3FFF) << 16) | ((uint64_t)((128u >> 4) & 0x3FFF) << 32) |
                                ((uint64_t)1 << 46);
    const uint32_t bhi_addr = smem_u32(s_bhi), blo_addr = smem_u32(s_blo);
    uint32_t use = 0;
    for (int tile = blockIdx.x * NPIPE + pipe; tile < p.n_tiles; tile += tile_stride) {
#pragma unroll 1
      for (int ks = 0; ks < KS; ++ks, ++use) {
        const uint32_t s = use % NSTG;
        mbar_wait(&bar_afull[s], (use / NSTG) & 1);
        tc_fence_after();
        const uint32_t koff = (uint32_t)ks * 2u * N16 * 16u;
        const uint64_t dhi = desc_fixed | (uint64_t)(((bhi_addr + koff) >> 4) & 0x3FFF);
        const uint64_t dlo = desc_fixed | (uint64_t)(((blo_addr + koff) >> 4) & 0x3FFF);
#pragma unroll
        for (int t = 0; t < TR; ++t) {
          const uint32_t d = tcol + t * N16;
          const uint32_t a = tcol + colA0 + (s * TR + t) * 16;
          mma_tf32_ts(d, a, dhi, idesc, ks > 0 ? 1u : 0u);
          mma_tf32_ts(d, a, dlo, idesc, 1u);
          mma_tf32_ts(d, a + 8, dhi, idesc, 1u);
        }
        tc_commit(&bar_aempty[s]);
      }
      tc_commit(bar_dfull);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == n_work_warps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

#endif  // HP_LEGACY_KERNELS

// ---------------------------------------------------------------------------- deep (warp-specialised) variant
// One pipeline per CTA, every stage of a tile in its own warps so that load, depthwise + MMA, epilogue and store of
// consecutive tiles overlap (measured with tools/tma_bench: a TMA load + store of one band costs ~5000 clk of latency
// when nothing else is in flight, the whole HBM budget of a band is ~3300 clk):
//   loader  (1 thread) : TMA loads into a ring of NBUF halo buffers            free[b]  -> full[b]
//   DW sets (NSETS x 4 warps): depthwise chunks -> TMEM A ring                 full[b], a_empty[s] -> a_full[s]
//   issuer  (1 thread) : 3 x TR tcgen05.mma per k-step into D[i & 1]           a_full[s], d_empty[d] -> a_empty[s], d_full[d]
//   epilogue (NESETS x 4 warps): D[d] + bias + skip -> ReLU in place in buffer b   d_full[d] -> d_empty[d], epi_done[b]
//   storer  (1 thread) : TMA store of the band interior, frees the buffer      epi_done[b] -> free[b]
// Tile = NI images x one band of BH rows x full width.  The halo buffer has NO left halo column: the box starts at x = 0
// and is IWB >= W pixels wide with IWB * PS * 4 a multiple of 128 bytes, so every row (and the store source, the second
// row) is 128-byte aligned.  The left neighbour of x = 0 (and the right neighbour of x = W - 1 when IWB == W) is SAME
// padding: the lane multiplies whatever finite value it reads there by a zeroed depthwise weight.
// Depthwise weights and bias as kernel-parameter constants: every lane of a warp needs the same 16 bytes, and a broadcast
// LDS.128 still costs the 4 passes of a full-width one (measured: with the weights in shared memory the depthwise stage of
// the pixel-per-lane kernel took ~9.5K clk per tile against 3.5K clk of input wavefronts).  From the constant bank they do not touch the
// shared-memory pipe at all.  (Tried in the band kernel too: there the weights are reused over TR rows and the constant
// loads cost more than they save -- blocks 0 / 1 went from 0.43 to 0.45 ms -- so it keeps them in shared memory.)
#ifndef HP_TC_BREG_MAX
#define HP_TC_BREG_MAX 8
#endif
template <int CINP>
struct DwConst {
  float w[9 * CINP];
  float b[CINP];
};

#define TCD_MAXB 4
struct TcdParams {
  const float *dww, *dwb, *pwb, *bhi, *blo;
  int W, H, BH, IWB, row_pitch, img_pitch;   // pitches in floats; img_pitch = (BH + 2) * row_pitch
  int NI, bands_per_img, n_tiles, lanes, lanes_per_img, B;
  int nstg, nbuf;
  uint32_t load_bytes;
  int off_b, off_w, off_pipe, buf_floats;
  long long* trace;
  int trace_tiles;
  float unscale;                   // UNIT 4 (split-fp16 MMAs): inverse of the power-of-two scale of the weights
  unsigned int* status;            // sticky status word of the context (HP_STATUS_CHAIN_RANGE)
};

// UNIT 4: a work unit is 16 channels = one kind::f16 MMA k-step (split fp16 instead of 3xTF32: half the MMAs, half the hand-offs
// between depthwise sets and issuers per tile, half the B-operand traffic; same scheme as the chain kernel, blocks_chain.cu).
// NISS issuer warps: one thread can only issue a tcgen05.mma every ~46-55 clk whatever its size (tools/mma_rate.cu), the
// tensor pipe itself needs 128 * N / 256 clk; the accumulator rows are therefore dealt to NISS issuing threads (t % NISS).
// PLACE = 1: 4 * NISS helper warps with issuer k at warp W_ISSUE + 4 k + 3, i.e. on SM sub-partition 3, which is idle when at most
// 96 TMEM lanes are active; the instructions around each tcgen05.mma then do not queue behind the worker warps (stem: 0.49 -> 0.42 ms).
template <int CINP, int COUTP, int TR, int NSETS, int NESETS, int UNIT, int NISS, int PLACE>
__global__ void __launch_bounds__(128 * NSETS + 128 * NESETS + (PLACE ? 128 * NISS : 32 * (NISS + 2)), 1)
blaze_block_deep_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, TcdParams p) {
  using G = TcGeom<CINP, COUTP>;
  constexpr int C4 = G::C4, NG = G::NG, K8 = G::K8, N16 = G::N16, PS = G::PS;
  constexpr bool F16 = UNIT == 4;
  constexpr int KS = F16 ? (CINP + 15) / 16 : G::KS;              // MMA k-steps per tile
  constexpr int BFLOATS = F16 ? KS * 16 * N16 / 2 : K8 * N16;     // floats of one weight part (hi or lo)
  constexpr uint32_t colA0 = 2 * TR * N16;                        // TMEM: D[0], D[1] (TR * N16 columns each), then the A ring
  static_assert(colA0 + 2 * TR * 16 <= 512, "TMEM budget");
  static_assert(BFLOATS <= K8 * N16, "weight area");

  extern __shared__ __align__(1024) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_full = bars;                                       // [nbuf]
  uint64_t* bar_free = bars + TCD_MAXB;                            // [nbuf]
  uint64_t* bar_epi = bars + 2 * TCD_MAXB;                         // [nbuf]
  uint64_t* bar_afull = bars + 3 * TCD_MAXB;                       // [nstg]
  uint64_t* bar_aempty = bar_afull + TC_MAX_STG;                   // [nstg]
  uint64_t* bar_dfull = bar_aempty + TC_MAX_STG;                   // [2]
  uint64_t* bar_dempty = bar_dfull + 2;                            // [2]
  static_assert((3 * TCD_MAXB + 2 * TC_MAX_STG + 4) * 8 + 4 <= TC_BAR_FLOATS * 4, "barrier block");
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem) + (TC_BAR_FLOATS - 1);
  float* s_bhi = smem + p.off_b;
  float* s_blo = s_bhi + K8 * N16;
  float* s_dww = smem + p.off_w;
  float* s_dwb = s_dww + 9 * CINP;
  float* s_pwb = s_dwb + CINP;
  float* bufs = smem + p.off_pipe;

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane_id = tid & 31;
  constexpr int W_EPI = 4 * NSETS, W_ISSUE = W_EPI + 4 * NESETS;
  constexpr int W_LOAD = PLACE ? W_ISSUE : W_ISSUE + NISS, W_STORE = W_LOAD + 1;
  constexpr int W_ALLOC = PLACE ? W_ISSUE + 3 : W_ISSUE;   // the first issuer warp owns the TMEM allocation
  static_assert(NISS >= 1 && NISS <= TR, "issuer warps");
  const int issuer_of_warp = PLACE ? ((warp >= W_ISSUE && ((warp - W_ISSUE) & 3) == 3) ? (warp - W_ISSUE) >> 2 : -1)
                                   : ((warp >= W_ISSUE && warp < W_ISSUE + NISS) ? warp - W_ISSUE : -1);
  const int NSTG = p.nstg, NBUF = p.nbuf;

  for (int i = tid * 4; i < BFLOATS; i += nthr * 4) {
    st4(s_bhi + i, ld4(p.bhi + i));
    st4(s_blo + i, ld4(p.blo + i));
  }
  for (int i = tid * 4; i < 9 * CINP; i += nthr * 4) st4(s_dww + i, ld4(p.dww + i));
  for (int i = tid * 4; i < CINP; i += nthr * 4) st4(s_dwb + i, ld4(p.dwb + i));
  for (int i = tid * 4; i < COUTP; i += nthr * 4) st4(s_pwb + i, ld4(p.pwb + i));
  // masked taps read (and multiply by zero) floats just outside the rows of a buffer: make every such float finite
  for (int i = p.off_w + 10 * CINP + COUTP + tid * 4; i < p.off_pipe + NBUF * p.buf_floats + 256; i += nthr * 4)
    st4(smem + i, make_float4(0.f, 0.f, 0.f, 0.f));
  fence_async_smem();
  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&bar_full[b], 1);
      mbar_init(&bar_free[b], 1);
      mbar_init(&bar_epi[b], 128 * NESETS);
    }
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&bar_afull[s], UNIT >= 2 ? 128 : 256);   // one set per k-step (UNIT 2, 4) or one set per 4-channel half
      mbar_init(&bar_aempty[s], NISS);                   // one tcgen05.commit per issuer
    }
    for (int d = 0; d < 2; ++d) {
      mbar_init(&bar_dfull[d], NISS);
      mbar_init(&bar_dempty[d], 128 * NESETS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;

  // tile -> (first image, first row); NI > 1 only with one band per image
  auto tile_coords = [&](int tile, int& img, int& y0) {
    const int q = tile / p.bands_per_img;
    img = q * p.NI;
    y0 = (tile - q * p.bands_per_img) * p.BH;
  };
  const int row_pitch = p.row_pitch;
  // trace slots: 0 load issued, 1 DW sees full, 2 DW set 0 done, 3 epilogue sees d_full, 4 epilogue done, 5 store issued, 6 store read done,
  // 7 MMAs issued, 8 last DW set done, 9 issuer sees last a_full, 10 issuer sees first a_full, 11 issuer has D free
  auto stamp = [&](int i, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && i < p.trace_tiles) p.trace[i * 12 + slot] = clock64();
  };

  if (warp < W_ISSUE) {
    // lane geometry shared by the depthwise and the epilogue warps: lane -> (image in tile, strip, column)
    const int wq = warp & 3;
    const int lane = wq * 32 + lane_id;
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const bool active = lane < p.lanes;
    const bool warp_active = wq * 32 < p.lanes;
    const int l = active ? lane : 0;
    const int im = l / p.lanes_per_img;
    const int l2 = l - im * p.lanes_per_img;
    const int yq = l2 / p.W;
    const int x = l2 - yq * p.W;
    const int my_off = im * p.img_pitch + yq * TR * row_pitch + (x - 1) * PS;   // top-left pixel of the 3x3 window of output row yq*TR
    const bool mask_l = (x == 0), mask_r = (x == p.W - 1 && p.IWB == p.W);
    if (warp < W_EPI) {
      // =============================================================== depthwise sets
      // Work units (a 4-channel chunk, or with UNIT 2 both chunks of a k-step) are dealt round-robin to the sets over the
      // GLOBAL unit counter of the CTA, not per tile: consecutive units of a set are then always NSETS units apart, which
      // bounds how far one set can run ahead of the others (an mbarrier parity wait is only safe one phase ahead: the
      // host checks that consecutive units of a set are at most NSTG k-steps apart).  The TMEM stores of a unit are waited for, and the unit published, only after
      // the next unit has been computed: tcgen05.st latency and the a_empty round trip hide behind the LDS / FFMA work.
      const int set = warp >> 2;
      constexpr int UPT = (UNIT >= 2) ? KS : 2 * KS;                       // units per tile
      const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const uint32_t n_units = (uint32_t)my_tiles * UPT;
      uint64_t* pending = nullptr;
      int cur_i = -1;
      const float* buf = bufs;
#pragma unroll 1
      for (uint32_t g = set; g < n_units; g += NSETS) {
        const int i = (int)(g / UPT);
        const int u = (int)(g - (uint32_t)i * UPT);
        const int ks = (UNIT >= 2) ? u : (u >> 1), half = (UNIT >= 2) ? 0 : (u & 1);
        const int c4 = F16 ? 4 * u : (UNIT == 2) ? 2 * u : u;
        const uint32_t use = (uint32_t)i * KS + ks;
        const uint32_t s = use % NSTG;
        if (i != cur_i) {                                                    // first unit of this set in tile i
          cur_i = i;
          const int b = i % NBUF;
          buf = bufs + b * p.buf_floats;
          mbar_wait(&bar_full[b], (i / NBUF) & 1);
          if (tid == 0) stamp(i, 1);
        }
        float4 acc[TR], acc1[TR];
        uint32_t hv[F16 ? TR : 1][16];                                       // F16: the 16 columns of every stage row
        if (F16) {
          if (warp_active) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (c4 + q < C4) {
                tc_dw_compute<CINP, TR, PS>(buf + my_off + (c4 + q) * 4, s_dww + (c4 + q) * 4, s_dwb + (c4 + q) * 4, row_pitch, mask_l, mask_r, acc);
#pragma unroll
                for (int t = 0; t < TR; ++t) {
                  const __half2 h0 = __floats2half2_rn(acc[t].x, acc[t].y), h1 = __floats2half2_rn(acc[t].z, acc[t].w);
                  const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
                  const __half2 l0 = __floats2half2_rn(acc[t].x - b0.x, acc[t].y - b0.y), l1 = __floats2half2_rn(acc[t].z - b1.x, acc[t].w - b1.y);
                  hv[F16 ? t : 0][2 * q] = *reinterpret_cast<const uint32_t*>(&h0);
                  hv[F16 ? t : 0][2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                  hv[F16 ? t : 0][8 + 2 * q] = *reinterpret_cast<const uint32_t*>(&l0);
                  hv[F16 ? t : 0][8 + 2 * q + 1] = *reinterpret_cast<const uint32_t*>(&l1);
                }
              } else {                                                       // K padding
#pragma unroll
                for (int t = 0; t < TR; ++t) hv[F16 ? t : 0][2 * q] = hv[F16 ? t : 0][2 * q + 1] = hv[F16 ? t : 0][8 + 2 * q] = hv[F16 ? t : 0][8 + 2 * q + 1] = 0u;
              }
            }
          }
        } else if (warp_active && c4 < C4)
          tc_dw_compute<CINP, TR, PS>(buf + my_off + c4 * 4, s_dww + c4 * 4, s_dwb + c4 * 4, row_pitch, mask_l, mask_r, acc);
        if (UNIT == 2) {
          if (warp_active && c4 + 1 < C4) {
            tc_dw_compute<CINP, TR, PS>(buf + my_off + c4 * 4 + 4, s_dww + c4 * 4 + 4, s_dwb + c4 * 4 + 4, row_pitch, mask_l, mask_r, acc1);
          } else {
#pragma unroll
            for (int t = 0; t < TR; ++t) acc1[t] = make_float4(0.f, 0.f, 0.f, 0.f);   // K padding
          }
        }
        if (pending != nullptr) {
          if (warp_active) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
          }
          mbar_arrive(pending);
        }
        if (use >= (uint32_t)NSTG) {
          mbar_wait(&bar_aempty[s], ((use / NSTG) - 1) & 1);
          tc_fence_after();
        }
        if (warp_active) {
          const uint32_t acol = tlane + colA0 + s * (TR * 16) + half * 4;
          if (F16) {
#pragma unroll
            for (int t = 0; t < TR; ++t) tmem_st16(acol + t * 16, hv[F16 ? t : 0]);
          } else if (UNIT == 2) tc_dw_store2<TR>(acc, acc1, acol);
          else if (c4 < C4) tc_dw_store<TR>(acc, acol);
          else tc_zero_chunk<TR>(acol);
        }
        pending = &bar_afull[s];
        if (g + NSETS >= n_units || (int)((g + NSETS) / UPT) != i) {
          // last unit of this set in tile i: publish now (the next unit may have to wait for a TMA load)
          if (warp_active) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
          }
          mbar_arrive(pending);
          pending = nullptr;
          if (tid == 0) stamp(i, 2);
          if (tid == (NSETS - 1) * 128) stamp(i, 8);
        }
      }
    } else {
      // =============================================================== epilogue warps
      const int centre0 = my_off + row_pitch + PS;
      const int eset = (warp - W_EPI) >> 2;
      uint32_t guard = 0u;
      constexpr int BREG = (NG <= HP_TC_BREG_MAX) ? 1 : 0;          // the bias of up to 32 output channels lives in registers
      float4 breg[BREG ? NG : 1];
      if (BREG) {
#pragma unroll
        for (int j = 0; j < (BREG ? NG : 1); ++j) breg[j] = ld4(s_pwb + j * 4);
      }
      int i = 0, b = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        const int d = i & 1;
        float* buf = bufs + b * p.buf_floats;
        mbar_wait(&bar_dfull[d], (i >> 1) & 1);
        tc_fence_after();
        if (tid == W_EPI * 32) stamp(i, 3);
        mbar_wait(&bar_full[b], (i / NBUF) & 1);   // already complete; orders the TMA-written skip pixels for this thread
        if (warp_active) {
#pragma unroll
          for (int t = 0; t < TR; ++t)
            if (t % NESETS == eset)
              tc_epilogue_pixel<C4, NG, N16, F16 ? 1 : 0, BREG>(buf + centre0 + t * row_pitch, tlane + d * (TR * N16) + t * N16, s_pwb, active, p.unscale,
                                                                &guard, breg);
          tc_fence_before();
          fence_async_smem();
        }
        mbar_arrive(&bar_dempty[d]);
        mbar_arrive(&bar_epi[b]);
        if (tid == W_EPI * 32) stamp(i, 4);
        if (++b == NBUF) b = 0;
      }
      if (F16 && guard && p.status) atomicOr(p.status, 4u);
    }
  } else if (lane_id == 0) {
    if (issuer_of_warp >= 0) {
      // =============================================================== MMA issuers (accumulator rows t % NISS == issuer)
      const int issuer = issuer_of_warp;
      const uint32_t idesc = F16 ? ((1u << 4) | ((uint32_t)(N16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)) : tc_idesc_tf32(N16);
      const uint64_t desc_fixed = tc_bdesc_fixed(N16);
      const uint32_t bhi_addr = smem_u32(s_bhi), blo_addr = smem_u32(s_blo);
      uint32_t use = 0;
      int i = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        const int d = i & 1;
        if (i >= 2) {
          mbar_wait(&bar_dempty[d], ((i >> 1) - 1) & 1);
          tc_fence_after();
        }
        if (issuer == 0) stamp(i, 11);
        uint64_t dhi = desc_fixed | (uint64_t)((bhi_addr >> 4) & 0x3FFF), dlo = desc_fixed | (uint64_t)((blo_addr >> 4) & 0x3FFF);
#pragma unroll 1
        for (int ks = 0; ks < KS; ++ks, ++use, dhi += (2u * N16 * 16u) >> 4, dlo += (2u * N16 * 16u) >> 4) {
          const uint32_t s = use % NSTG;
          mbar_wait_lean(&bar_afull[s], (use / NSTG) & 1);
          tc_fence_after();
          if (issuer == 0 && p.trace != nullptr) {            // off the critical path of an untraced run: one predictable branch
            if (ks == 0) stamp(i, 10);
            if (ks == KS - 1) stamp(i, 9);
          }
#pragma unroll
          for (int t = 0; t < TR; ++t) {
            if (t % NISS != issuer) continue;
            const uint32_t dc = tmem_base + d * (TR * N16) + t * N16;
            const uint32_t a = tmem_base + colA0 + (s * TR + t) * 16;
            if (F16) {
              mma_f16_ts(dc, a, dhi, idesc, ks > 0 ? 1u : 0u);
              mma_f16_ts(dc, a, dlo, idesc, 1u);
              mma_f16_ts(dc, a + 8, dhi, idesc, 1u);
            } else {
              mma_tf32_ts(dc, a, dhi, idesc, ks > 0 ? 1u : 0u);
              mma_tf32_ts(dc, a, dlo, idesc, 1u);
              mma_tf32_ts(dc, a + 8, dhi, idesc, 1u);
            }
          }
          tc_commit(&bar_aempty[s]);
        }
        tc_commit(&bar_dfull[d]);
        if (issuer == 0) stamp(i, 7);
      }
    } else if (warp == W_LOAD) {
      // =============================================================== TMA loader
      int i = 0, b = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        if (i >= NBUF) mbar_wait(&bar_free[b], ((i / NBUF) - 1) & 1);
        int img, y0;
        tile_coords(tile, img, y0);
        mbar_expect_tx(&bar_full[b], p.load_bytes);
        tma_load_4d(bufs + b * p.buf_floats, &tm_in, &bar_full[b], 0, 0, y0 - 1, img);
        stamp(i, 0);
        if (++b == NBUF) b = 0;
      }
    } else if (warp == W_STORE) {
      // =============================================================== TMA storer
      int i = 0, b = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        mbar_wait(&bar_epi[b], (i / NBUF) & 1);
        int img, y0;
        tile_coords(tile, img, y0);
        for (int im = 0; im < p.NI; ++im)
          if (img + im < p.B) tma_store_4d(&tm_out, bufs + b * p.buf_floats + im * p.img_pitch + row_pitch, 0, 0, y0, img + im);
        tma_store_commit();
        stamp(i, 5);
        tma_store_wait_read();
        stamp(i, 6);
        mbar_arrive(&bar_free[b]);
        if (++b == NBUF) b = 0;
      }
      tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}



// ============================================================================ small maps: one pixel per lane
// Maps of at most 128 pixels (blocks 12-15: 6 x 6 at 88 / 96 input, 8 x 8 at 128; blocks 6-10: 11 x 11 at 88 input).  The
// band kernel above needs halo rows / padded columns around every image, so only 2-4 images of 96 channels fit a tile
// and the per-tile latency chain dominates.  Here a tile is NI = 128 / (H W) WHOLE images, stored compactly as
// [NI H W pixels][PS floats] (one TMA box over the tensor viewed as [B H W][C]), lane <-> pixel, and every lane keeps
// the shared-memory offsets of its 9 taps: a tap outside the image points at zeros behind the tile (SAME padding), so
// no weight masking and no halo.  Same warp roles and barriers as blaze_block_deep_kernel with one
// accumulator row per lane (TR = 1); a work unit is a k-step (8 channels: 18 LDS.128 of inputs, 18 of weights).
// The epilogue sets alternate tiles (set e owns accumulator buffer D[e] when NESETS == 2).
struct TcsParams {
  const float *dww, *dwb, *pwb, *bhi, *blo;
  int W, H, P, NI, rows, n_tiles;            // P = H * W pixels per image, rows = NI * P lanes in use
  int nstg, nbuf, ku, upt;                   // ku k-steps per depthwise unit (ring stage = 16 ku TMEM columns), units per tile
  uint32_t load_bytes;
  int off_b, off_w, off_pipe, buf_floats;
  long long* trace;                          // optional clock stamps of CTA 0 (same 12 slots per tile as the band kernel)
  int trace_tiles;
};

template <int CINP, int COUTP, int NSETS, int NESETS>
__global__ void __launch_bounds__(128 * NSETS + 128 * NESETS + 96, 1)
blaze_block_small_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                         const __grid_constant__ DwConst<CINP> dwc, TcsParams p) {
  using G = TcGeom<CINP, COUTP>;
  constexpr int C4 = G::C4, NG = G::NG, K8 = G::K8, KS = G::KS, N16 = G::N16, PS = G::PS;
  constexpr uint32_t colA0 = 2 * N16;                             // TMEM: D[0], D[1], then the A ring (16 columns per stage)
  static_assert(colA0 + TC_MAX_STG * 32 <= 512, "TMEM budget");

  extern __shared__ __align__(1024) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_full = bars;                                       // [nbuf]
  uint64_t* bar_free = bars + TCD_MAXB;                            // [nbuf]
  uint64_t* bar_epi = bars + 2 * TCD_MAXB;                         // [nbuf]
  uint64_t* bar_afull = bars + 3 * TCD_MAXB;                       // [nstg]
  uint64_t* bar_aempty = bar_afull + TC_MAX_STG;                   // [nstg]
  uint64_t* bar_dfull = bar_aempty + TC_MAX_STG;                   // [2]
  uint64_t* bar_dempty = bar_dfull + 2;                            // [2]
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem) + (TC_BAR_FLOATS - 1);
  float* s_bhi = smem + p.off_b;
  float* s_blo = s_bhi + K8 * N16;
  float* s_pwb = smem + p.off_w + 10 * CINP;
  float* bufs = smem + p.off_pipe;

  const int tid = threadIdx.x, nthr = blockDim.x;
  if (tid == 0 && p.trace != nullptr && blockIdx.x == 0 && p.trace_tiles > 0) p.trace[11] = clock64();
  const int warp = tid >> 5, lane_id = tid & 31;
  constexpr int W_EPI = 4 * NSETS, W_ISSUE = W_EPI + 4 * NESETS, W_LOAD = W_ISSUE + 1, W_STORE = W_ISSUE + 2;
  const int NSTG = p.nstg, NBUF = p.nbuf;

  for (int i = tid * 4; i < K8 * N16; i += nthr * 4) {
    st4(s_bhi + i, ld4(p.bhi + i));
    st4(s_blo + i, ld4(p.blo + i));
  }
  for (int i = tid * 4; i < COUTP; i += nthr * 4) st4(s_pwb + i, ld4(p.pwb + i));
  // the zeros behind every tile are written once; TMA and the epilogue never touch them
  {
    const int zf = p.buf_floats - p.rows * PS;
    for (int i = tid * 4; i < NBUF * zf; i += nthr * 4) {
      const int b = i / zf;
      st4(bufs + b * p.buf_floats + p.rows * PS + (i - b * zf), make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  fence_async_smem();
  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&bar_full[b], 1);
      mbar_init(&bar_free[b], 1);
      mbar_init(&bar_epi[b], 128);
    }
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&bar_afull[s], 128);
      mbar_init(&bar_aempty[s], 1);
    }
    for (int d = 0; d < 2; ++d) {
      mbar_init(&bar_dfull[d], 1);
      mbar_init(&bar_dempty[d], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ISSUE) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto stamp = [&](int i, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && i < p.trace_tiles) p.trace[i * 12 + slot] = clock64();
  };

  if (warp < W_ISSUE) {
    const int wq = warp & 3;
    const int lane = wq * 32 + lane_id;                            // pixel of the tile
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const bool active = lane < p.rows;
    const bool warp_active = wq * 32 < p.rows;
    if (warp < W_EPI) {
      // =============================================================== depthwise sets (k-step units, global round-robin)
      const int set = warp >> 2;
      // tap offsets in floats.  A tap outside the image reads zeros behind the tile, at the 16-byte chunk (mod 8) the
      // in-bounds address would have had: the 8 lanes of a quarter warp then still hit 8 different bank groups (with one
      // shared pixel of zeros the edge lanes collided with their neighbours: 2 wavefronts per LDS.128 quarter, DW 1.8x slower)
      int off[9];
      {
        const int r = lane % p.P, y = r / p.W, x = r - y * p.W;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3 - 1, dx = t % 3 - 1;
          const bool ok = active && y + dy >= 0 && y + dy < p.H && x + dx >= 0 && x + dx < p.W;
          const int q = lane + dy * p.W + dx;                      // neighbour pixel of the tile (may be < 0 or another image)
          off[t] = ok ? q * PS : p.rows * PS + ((q * (PS / 4)) & 7) * 4;
        }
      }
      // a unit = ku k-steps: the per-unit hand-off (a_empty wait, tcgen05.wait::st, fence, arrive) costs ~1K clk whatever it carries
      const int KU = p.ku, UPT = p.upt;
      const uint32_t n_units = (uint32_t)my_tiles * UPT;
      int cur_i = -1;
      const float* buf = bufs;
#pragma unroll 1
      for (uint32_t g = set; g < n_units; g += NSETS) {
        const int i = (int)(g / UPT);
        const int ks0 = (int)(g - (uint32_t)i * UPT) * KU;
        const int nk = (KS - ks0 < KU) ? KS - ks0 : KU;
        const uint32_t s = g % NSTG;
        if (i != cur_i) {                                          // first unit of this set in tile i
          cur_i = i;
          const int b = i % NBUF;
          buf = bufs + b * p.buf_floats;
          mbar_wait(&bar_full[b], (i / NBUF) & 1);
          if (tid == 0) stamp(i, 1);
        }
        if (g >= (uint32_t)NSTG) {
          mbar_wait(&bar_aempty[s], ((g / NSTG) - 1) & 1);
          tc_fence_after();
        }
        if (warp_active) {
#pragma unroll 1
          for (int kk = 0; kk < nk; ++kk) {
            const int c = 8 * (ks0 + kk);
            const bool two = (2 * (ks0 + kk) + 1 < C4);            // the second 4-channel chunk of the k-step exists
            float4 a0 = ld4(dwc.b + c), a1 = two ? ld4(dwc.b + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const float* q = buf + off[t] + c;
              a0 = fma4(ld4(q), ld4(dwc.w + t * CINP + c), a0);
              if (two) a1 = fma4(ld4(q + 4), ld4(dwc.w + t * CINP + c + 4), a1);
            }
            const float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            uint32_t v[16];                                        // [hi 8 | lo 8]
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              v[e] = tf32_hi(f[e]);
              v[8 + e] = __float_as_uint(f[e] - __uint_as_float(v[e]));
            }
            tmem_st16(tlane + colA0 + s * (16 * KU) + kk * 16, v);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
        }
        mbar_arrive(&bar_afull[s]);
        if (g + NSETS >= n_units || (int)((g + NSETS) / UPT) != i) {
          if (tid == 0) stamp(i, 2);
          if (tid == (NSETS - 1) * 128) stamp(i, 8);
        }
      }
    } else {
      // =============================================================== epilogue sets: tile i belongs to set i % NESETS
      const int eset = (warp - W_EPI) >> 2;
      for (int i = eset; i < my_tiles; i += NESETS) {
        const int d = i & 1, b = i % NBUF;
        float* buf = bufs + b * p.buf_floats;
        mbar_wait(&bar_dfull[d], (i >> 1) & 1);
        tc_fence_after();
        if ((tid & 127) == 0) stamp(i, 3);
        if (warp_active) {
          tc_epilogue_pixel<C4, NG, N16>(buf + lane * PS, tlane + d * N16, s_pwb, active);
          tc_fence_before();
          fence_async_smem();
        }
        mbar_arrive(&bar_dempty[d]);
        mbar_arrive(&bar_epi[b]);
        if ((tid & 127) == 0) stamp(i, 4);
      }
    }
  } else if (lane_id == 0) {
    if (warp == W_ISSUE) {
      // =============================================================== MMA issuer
      const uint32_t idesc = tc_idesc_tf32(N16);
      const uint64_t desc_fixed = tc_bdesc_fixed(N16);
      const uint32_t bhi_addr = smem_u32(s_bhi), blo_addr = smem_u32(s_blo);
      uint32_t use = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int d = i & 1;
        if (i >= 2) {
          mbar_wait(&bar_dempty[d], ((i >> 1) - 1) & 1);
          tc_fence_after();
        }
        if (i > 0) stamp(i, 11);                                     // slot 11 of tile 0: kernel entry
#pragma unroll 1
        for (int u = 0; u < p.upt; ++u, ++use) {
          const uint32_t s = use % NSTG;
          mbar_wait_lean(&bar_afull[s], (use / NSTG) & 1);
          tc_fence_after();
          if (u == 0) stamp(i, 10);
          if (u == p.upt - 1) stamp(i, 9);
          const uint32_t dc = tmem_base + d * N16;
          for (int kk = 0; kk < p.ku && u * p.ku + kk < KS; ++kk) {
            const int ks = u * p.ku + kk;
            const uint32_t koff = (uint32_t)ks * 2u * N16 * 16u;
            const uint64_t dhi = desc_fixed | (uint64_t)(((bhi_addr + koff) >> 4) & 0x3FFF);
            const uint64_t dlo = desc_fixed | (uint64_t)(((blo_addr + koff) >> 4) & 0x3FFF);
            const uint32_t a = tmem_base + colA0 + s * (16 * p.ku) + kk * 16;
            mma_tf32_ts(dc, a, dhi, idesc, ks > 0 ? 1u : 0u);
            mma_tf32_ts(dc, a, dlo, idesc, 1u);
            mma_tf32_ts(dc, a + 8, dhi, idesc, 1u);
          }
          tc_commit(&bar_aempty[s]);
        }
        tc_commit(&bar_dfull[d]);
        stamp(i, 7);
      }
    } else if (warp == W_LOAD) {
      // =============================================================== TMA loader
      int i = 0, b = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        if (i >= NBUF) mbar_wait(&bar_free[b], ((i / NBUF) - 1) & 1);
        mbar_expect_tx(&bar_full[b], p.load_bytes);
        tma_load_4d(bufs + b * p.buf_floats, &tm_in, &bar_full[b], 0, tile * p.rows, 0, 0);
        stamp(i, 0);
        if (++b == NBUF) b = 0;
      }
    } else if (warp == W_STORE) {
      // =============================================================== TMA storer
      int i = 0, b = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        mbar_wait(&bar_epi[b], (i / NBUF) & 1);
        tma_store_4d(&tm_out, bufs + b * p.buf_floats, 0, tile * p.rows, 0, 0);
        tma_store_commit();
        stamp(i, 5);
        tma_store_wait_read();
        stamp(i, 6);
        mbar_arrive(&bar_free[b]);
        if (++b == NBUF) b = 0;
      }
      tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ISSUE) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}


// ============================================================================ stride-2 blocks (2, 5, 11): one output pixel per lane
// out = ReLU(pointwise(depthwise3x3 stride 2 (x)) + channel_pad(maxpool2x2(x))).  A tile is a band of R output rows of one
// image (R Wo <= 128 lanes, lane <-> output pixel).  The input band -- 2R+1 rows x 2Wo+1 pixels starting at image
// row 2 y0 - pad_t, column -pad_l -- comes in one TMA box: everything outside the image is zero-filled by the TMA, so the
// 9 tap offsets of a lane need no boundary logic at all, and the max-pool of the skip path reads 4 of the same pixels
// (rows / columns + pad_t / + pad_l; the zero fill is harmless there because block inputs are ReLU outputs, >= 0).
// Depthwise -> TF32 hi / lo -> TMEM A ring (work unit = `ku` k-steps), one accumulator row per lane, epilogue into a
// separate output staging buffer (the output has different geometry: not in place) -> TMA store.
// Neighbouring lanes are TWO input pixels apart, and with a pixel stride of an odd number of 16-byte chunks that is an even
// number of chunks: every depthwise / max-pool LDS.128 of the dense band was 2-way bank conflicted (8 wavefronts instead of
// 4; ncu: 77 % of the shared-memory wavefront budget on block 2).  The band therefore arrives as TWO boxes with a traversal
// stride of 2 pixels (cuTensorMapEncodeTiled elementStrides): the EVEN band columns 0, 2, .., 2 Wo form plane E, the ODD ones
// plane O, each stored densely [2R+1][Wo+1][PSI].  Tap kx = 0 of lane x is E[x], kx = 1 is O[x], kx = 2 is E[x+1]:
// neighbouring lanes are ONE pixel apart in every access.
struct Tcs2Params {
  const float *pwb, *bhi, *blo;
  int Wo, Ho, R, rows, bands_per_img, n_tiles, pad_t, pad_l;
  int IWB, row_pitch;                        // planes: IWB = Wo + 1 pixels per row, row_pitch = IWB * PSI floats
  int plane_floats;                          // plane O follows plane E at this offset
  int nstg, nbuf, ku, upt;
  uint32_t load_bytes;
  int off_b, off_w, off_out, out_floats, off_in, in_floats;
  long long* trace;
  int trace_tiles;
};

template <int CINP, int COUTP, int NSETS, int NESETS>
__global__ void __launch_bounds__(128 * NSETS + 128 * NESETS + 96, 1)
blaze_block_s2_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                      const __grid_constant__ DwConst<CINP> dwc, Tcs2Params p) {
  using G = TcGeom<CINP, COUTP>;
  constexpr int C4 = G::C4, NG = G::NG, K8 = G::K8, KS = G::KS, N16 = G::N16;
  constexpr int PSI = (C4 | 1) * 4, PSO = (NG | 1) * 4;           // pixel strides of the input band / the output staging tile
  constexpr uint32_t colA0 = 2 * N16;                             // TMEM: D[0], D[1], then the A ring (16 ku columns per stage)

  extern __shared__ __align__(1024) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_full = bars;                                       // [nbuf]  TMA load landed
  uint64_t* bar_free = bars + TCD_MAXB;                            // [nbuf]  epilogue has read the skip pixels
  uint64_t* bar_afull = bars + 2 * TCD_MAXB;                       // [nstg]
  uint64_t* bar_aempty = bar_afull + TC_MAX_STG;                   // [nstg]
  uint64_t* bar_dfull = bar_aempty + TC_MAX_STG;                   // [2]
  uint64_t* bar_dempty = bar_dfull + 2;                            // [2]
  uint64_t* bar_ofull = bar_dempty + 2;                            // [2]     output staging tile written
  uint64_t* bar_ofree = bar_ofull + 2;                             // [2]     ... and read by the TMA store
  static_assert((2 * TCD_MAXB + 2 * TC_MAX_STG + 8) * 8 + 4 <= TC_BAR_FLOATS * 4, "barrier block");
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem) + (TC_BAR_FLOATS - 1);
  float* s_bhi = smem + p.off_b;
  float* s_blo = s_bhi + K8 * N16;
  float* s_pwb = smem + p.off_w;
  float* obufs = smem + p.off_out;
  float* bufs = smem + p.off_in;

  const int tid = threadIdx.x, nthr = blockDim.x;
  if (tid == 0 && p.trace != nullptr && blockIdx.x == 0 && p.trace_tiles > 0) p.trace[11] = clock64();
  const int warp = tid >> 5, lane_id = tid & 31;
  constexpr int W_EPI = 4 * NSETS, W_ISSUE = W_EPI + 4 * NESETS, W_LOAD = W_ISSUE + 1, W_STORE = W_ISSUE + 2;
  const int NSTG = p.nstg, NBUF = p.nbuf;

  for (int i = tid * 4; i < K8 * N16; i += nthr * 4) {
    st4(s_bhi + i, ld4(p.bhi + i));
    st4(s_blo + i, ld4(p.blo + i));
  }
  for (int i = tid * 4; i < COUTP; i += nthr * 4) st4(s_pwb + i, ld4(p.pwb + i));
  for (int i = tid * 4; i < 2 * p.out_floats; i += nthr * 4) st4(obufs + i, make_float4(0.f, 0.f, 0.f, 0.f));
  fence_async_smem();
  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&bar_full[b], 1);
      mbar_init(&bar_free[b], 128);
    }
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&bar_afull[s], 128);
      mbar_init(&bar_aempty[s], 1);
    }
    for (int d = 0; d < 2; ++d) {
      mbar_init(&bar_dfull[d], 1);
      mbar_init(&bar_dempty[d], 128);
      mbar_init(&bar_ofull[d], 128);
      mbar_init(&bar_ofree[d], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ISSUE) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto stamp = [&](int i, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && i < p.trace_tiles) p.trace[i * 12 + slot] = clock64();
  };

  if (warp < W_ISSUE) {
    const int wq = warp & 3;
    const int lane = wq * 32 + lane_id;                            // output pixel of the band: (lane / Wo, lane % Wo)
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const bool active = lane < p.rows;
    const bool warp_active = wq * 32 < p.rows;
    const int yl = active ? lane / p.Wo : 0, xl = active ? lane - yl * p.Wo : 0;
    const int win0 = (2 * yl * p.IWB + xl) * PSI;                 // top-left tap of the 3x3 window: band row 2 yl, E[xl]
    const int koff1 = p.plane_floats, koff2 = PSI;                // taps kx = 1 (O[xl]) and kx = 2 (E[xl + 1]) relative to kx = 0
    if (warp < W_EPI) {
      // =============================================================== depthwise sets (units of ku k-steps, global round-robin)
      const int set = warp >> 2;
      const int KU = p.ku, UPT = p.upt;
      const uint32_t n_units = (uint32_t)my_tiles * UPT;
      int cur_i = -1;
      const float* buf = bufs;
#pragma unroll 1
      for (uint32_t g = set; g < n_units; g += NSETS) {
        const int i = (int)(g / UPT);
        const int ks0 = (int)(g - (uint32_t)i * UPT) * KU;
        const int nk = (KS - ks0 < KU) ? KS - ks0 : KU;
        const uint32_t s = g % NSTG;
        if (i != cur_i) {                                          // first unit of this set in tile i
          cur_i = i;
          const int b = i % NBUF;
          buf = bufs + b * p.in_floats + win0;
          mbar_wait(&bar_full[b], (i / NBUF) & 1);
          if (tid == 0) stamp(i, 1);
        }
        if (g >= (uint32_t)NSTG) {
          mbar_wait(&bar_aempty[s], ((g / NSTG) - 1) & 1);
          tc_fence_after();
        }
        if (warp_active) {
#pragma unroll 1
          for (int kk = 0; kk < nk; ++kk) {
            const int c = 8 * (ks0 + kk);
            const bool two = (2 * (ks0 + kk) + 1 < C4);            // the second 4-channel chunk of the k-step exists
            float4 a0 = ld4(dwc.b + c), a1 = two ? ld4(dwc.b + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const float* q = buf + (t / 3) * p.row_pitch + ((t % 3) == 0 ? 0 : (t % 3) == 1 ? koff1 : koff2) + c;
              a0 = fma4(ld4(q), ld4(dwc.w + t * CINP + c), a0);
              if (two) a1 = fma4(ld4(q + 4), ld4(dwc.w + t * CINP + c + 4), a1);
            }
            const float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            uint32_t v[16];                                        // [hi 8 | lo 8]
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              v[e] = tf32_hi(f[e]);
              v[8 + e] = __float_as_uint(f[e] - __uint_as_float(v[e]));
            }
            tmem_st16(tlane + colA0 + s * (16 * KU) + kk * 16, v);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
        }
        mbar_arrive(&bar_afull[s]);
        if (g + NSETS >= n_units || (int)((g + NSETS) / UPT) != i) {
          if (tid == 0) stamp(i, 2);
          if (tid == (NSETS - 1) * 128) stamp(i, 8);
        }
      }
    } else {
      // =============================================================== epilogue sets: tile i belongs to set i % NESETS
      const int eset = (warp - W_EPI) >> 2;
      // 2x2 max-pool window of the skip path: image pixels (2 y + {0, 1}, 2 x + {0, 1}) = band columns 2 xl + pad_l + {0, 1}:
      // E[xl], O[xl] (pad_l = 0) or O[xl], E[xl + 1] (pad_l = 1)
      const int skip0 = win0 + p.pad_t * p.row_pitch + (p.pad_l ? koff1 : 0);
      const int skip1 = win0 + p.pad_t * p.row_pitch + (p.pad_l ? koff2 : koff1);
      constexpr int NGRP = (NG + 7) / 8;
      for (int i = eset; i < my_tiles; i += NESETS) {
        const int d = i & 1, b = i % NBUF, ob = i & 1;
        const float* win = bufs + b * p.in_floats + skip0;
        const float* win1 = bufs + b * p.in_floats + skip1;
        float* opix = obufs + ob * p.out_floats + lane * PSO;
        mbar_wait(&bar_dfull[d], (i >> 1) & 1);
        tc_fence_after();
        if ((tid & 127) == 0) stamp(i, 3);
        if (i >= 2) mbar_wait(&bar_ofree[ob], ((i >> 1) - 1) & 1);
        if (warp_active) {
          uint32_t v[NGRP][32];
#pragma unroll
          for (int g = 0; g < NGRP; ++g) {
            if (g * 32 + 32 <= N16) {
              tmem_ld32(tlane + d * N16 + g * 32, v[g]);
            } else {
              uint32_t hlf[16];
              tmem_ld16(tlane + d * N16 + g * 32, hlf);
#pragma unroll
              for (int e = 0; e < 16; ++e) v[g][e] = hlf[e];
            }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tc_fence_before();
#pragma unroll
          for (int g = 0; g < NGRP; ++g) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int j = g * 8 + jj;
              if (j < NG) {
                const float4 bb = ld4(s_pwb + j * 4);
                float4 o = make_float4(__uint_as_float(v[g][jj * 4 + 0]) + bb.x, __uint_as_float(v[g][jj * 4 + 1]) + bb.y,
                                       __uint_as_float(v[g][jj * 4 + 2]) + bb.z, __uint_as_float(v[g][jj * 4 + 3]) + bb.w);
                if (j < C4) {
                  const float4 s0 = ld4(win + j * 4), s1 = ld4(win1 + j * 4);
                  const float4 s2 = ld4(win + p.row_pitch + j * 4), s3 = ld4(win1 + p.row_pitch + j * 4);
                  o.x += fmaxf(fmaxf(s0.x, s1.x), fmaxf(s2.x, s3.x)); o.y += fmaxf(fmaxf(s0.y, s1.y), fmaxf(s2.y, s3.y));
                  o.z += fmaxf(fmaxf(s0.z, s1.z), fmaxf(s2.z, s3.z)); o.w += fmaxf(fmaxf(s0.w, s1.w), fmaxf(s2.w, s3.w));
                }
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                if (active) st4(opix + j * 4, o);
              }
            }
          }
          fence_async_smem();
        }
        mbar_arrive(&bar_dempty[d]);
        mbar_arrive(&bar_free[b]);
        mbar_arrive(&bar_ofull[ob]);
        if ((tid & 127) == 0) stamp(i, 4);
      }
    }
  } else if (lane_id == 0) {
    auto tile_coords = [&](int tile, int& img, int& y0) {
      img = tile / p.bands_per_img;
      y0 = (tile - img * p.bands_per_img) * p.R;
    };
    if (warp == W_ISSUE) {
      // =============================================================== MMA issuer
      const uint32_t idesc = tc_idesc_tf32(N16);
      const uint64_t desc_fixed = tc_bdesc_fixed(N16);
      const uint32_t bhi_addr = smem_u32(s_bhi), blo_addr = smem_u32(s_blo);
      uint32_t use = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int d = i & 1;
        if (i >= 2) {
          mbar_wait(&bar_dempty[d], ((i >> 1) - 1) & 1);
          tc_fence_after();
        }
        if (i > 0) stamp(i, 11);
#pragma unroll 1
        for (int u = 0; u < p.upt; ++u, ++use) {
          const uint32_t s = use % NSTG;
          mbar_wait_lean(&bar_afull[s], (use / NSTG) & 1);
          tc_fence_after();
          if (u == 0) stamp(i, 10);
          if (u == p.upt - 1) stamp(i, 9);
          const uint32_t dc = tmem_base + d * N16;
          for (int kk = 0; kk < p.ku && u * p.ku + kk < KS; ++kk) {
            const int ks = u * p.ku + kk;
            const uint32_t koff = (uint32_t)ks * 2u * N16 * 16u;
            const uint64_t dhi = desc_fixed | (uint64_t)(((bhi_addr + koff) >> 4) & 0x3FFF);
            const uint64_t dlo = desc_fixed | (uint64_t)(((blo_addr + koff) >> 4) & 0x3FFF);
            const uint32_t a = tmem_base + colA0 + s * (16 * p.ku) + kk * 16;
            mma_tf32_ts(dc, a, dhi, idesc, ks > 0 ? 1u : 0u);
            mma_tf32_ts(dc, a, dlo, idesc, 1u);
            mma_tf32_ts(dc, a + 8, dhi, idesc, 1u);
          }
          tc_commit(&bar_aempty[s]);
        }
        tc_commit(&bar_dfull[d]);
        stamp(i, 7);
      }
    } else if (warp == W_LOAD) {
      // =============================================================== TMA loader
      int i = 0, b = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        if (i >= NBUF) mbar_wait(&bar_free[b], ((i / NBUF) - 1) & 1);
        int img, y0;
        tile_coords(tile, img, y0);
        mbar_expect_tx(&bar_full[b], 2u * p.load_bytes);
        tma_load_4d(bufs + b * p.in_floats, &tm_in, &bar_full[b], 0, -p.pad_l, 2 * y0 - p.pad_t, img);                       // plane E
        tma_load_4d(bufs + b * p.in_floats + p.plane_floats, &tm_in, &bar_full[b], 0, 1 - p.pad_l, 2 * y0 - p.pad_t, img);   // plane O
        stamp(i, 0);
        if (++b == NBUF) b = 0;
      }
    } else if (warp == W_STORE) {
      // =============================================================== TMA storer
      int i = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        const int ob = i & 1;
        mbar_wait(&bar_ofull[ob], (i >> 1) & 1);
        int img, y0;
        tile_coords(tile, img, y0);
        tma_store_4d(&tm_out, obufs + ob * p.out_floats, 0, 0, y0, img);
        tma_store_commit();
        stamp(i, 5);
        tma_store_wait_read();
        stamp(i, 6);
        mbar_arrive(&bar_ofree[ob]);
      }
      tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ISSUE) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;

int get_encode() {
  if (g_encode) return HP_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  HP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  HP_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, HP_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
  g_encode = (PFN_encodeTiled)fn;
  return HP_OK;
}

// NHWC float tensor [N][H][W][C] with box [bn][bh][bw][bc]; bc may exceed C (padded pixel stride in smem)
// wstride > 1: every wstride-th pixel of a row is loaded (bw pixels land densely in shared memory)
int make_map(CUtensorMap* tm, const float* base, int N, int H, int W, int C, int bn, int bh, int bw, int bc, int wstride = 1) {
  HP_TRY(get_encode());
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)(bw * wstride), (cuuint32_t)bh, (cuuint32_t)bn};   // traversal extent: bw * wstride
  cuuint32_t estr[4] = {1, (cuuint32_t)wstride, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HP_REQUIRE(r == CUDA_SUCCESS, HP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for tensor %dx%dx%dx%d box %dx%dx%dx%d", (int)r, N,
             H, W, C, bn, bh, bw, bc);
  return HP_OK;
}

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// shared-memory layout in floats: [barriers][B hi | B lo][dw weights, dw bias, pw bias][pipeline buffers ...]
inline void tc_layout(int cinp, int coutp, int K8, int N16, int* off_b, int* off_w, int* off_pipe) {
  int off = TC_BAR_FLOATS;
  *off_b = off;
  off = align_up(off + 2 * K8 * N16, 32);
  *off_w = off;
  off = align_up(off + 10 * cinp + coutp, 256);
  *off_pipe = off;
}

#ifdef HP_LEGACY_KERNELS
template <int CINP, int COUTP, int TR, int NSTG, int NPIPE, int NSETS>
int launch_tc(hp_ctx* h, const float* in, float* out, int B, int H, int W, const BlockWeights& w, const TcCfg& tc, cudaStream_t st) {
  using G = TcGeom<CINP, COUTP>;
  constexpr int TCOLS = TR * G::N16 + NSTG * TR * 16;
  static_assert(NPIPE * TCOLS <= 512, "TMEM budget");
  TcParams p;
  p.dww = w.dww; p.dwb = w.dwb; p.pwb = w.pwb; p.bhi = w.bhi; p.blo = w.blo;
  p.W = W; p.H = H; p.BH = tc.BH; p.IWB = tc.IWB; p.row_pitch = tc.IWB * G::PS;
  p.bands_per_img = ceil_div(H, tc.BH);
  p.n_tiles = B * p.bands_per_img;
  p.lanes = (tc.BH / TR) * W;
  p.load_bytes = (uint32_t)((size_t)G::PS * tc.IWB * (tc.BH + 2) * sizeof(float));
  p.trace = h->tc_trace; p.trace_tiles = h->tc_trace_tiles;
  tc_layout(CINP, COUTP, G::K8, G::N16, &p.off_b, &p.off_w, &p.off_pipe);
  const int off = p.off_pipe;
  p.buf_floats = align_up(G::PS * tc.IWB * (tc.BH + 2), 256);
  const size_t smem = (size_t)(off + tc.npipe * p.buf_floats) * sizeof(float);
  HP_REQUIRE(smem <= 227 * 1024, HP_ERR_INVALID, "tc block <%d,%d>: %zu bytes of shared memory needed", CINP, COUTP, smem);
  HP_REQUIRE(p.lanes >= 1 && p.lanes <= 128 && tc.BH % TR == 0 && (tc.IWB + 1) % 8 == 0 && tc.IWB >= W + 2 && tc.IWB <= 256 &&
                 tc.BH + 2 <= 256,
             HP_ERR_INVALID, "tc block <%d,%d>: bad band geometry BH %d IWB %d W %d", CINP, COUTP, tc.BH, tc.IWB, W);
  CUtensorMap tin, tout;
  HP_TRY(make_map(&tin, in, B, H, W, CINP, 1, tc.BH + 2, tc.IWB, G::PS));
  HP_TRY(make_map(&tout, out, B, H, W, COUTP, 1, tc.BH, tc.IWB, G::PS));
  auto kern = blaze_block_tc_kernel<CINP, COUTP, TR, NSTG, NPIPE, NSETS>;
  HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  long long grid = h->num_sms;
  const long long need = ceil_div(p.n_tiles, tc.npipe);
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, NPIPE * (128 * NSETS + 32), smem, st>>>(tin, tout, p);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

#endif  // HP_LEGACY_KERNELS

template <int CINP, int COUTP, int TR, int NSETS, int NESETS, int UNIT, int NISS, int PLACE>
int launch_deep(hp_ctx* h, const float* in, float* out, int B, int H, int W, const BlockWeights& w, const TcCfg& tc, cudaStream_t st) {
  using G = TcGeom<CINP, COUTP>;
  TcdParams p;
  p.dww = w.dww; p.dwb = w.dwb; p.pwb = w.pwb;
  p.bhi = UNIT == 4 ? w.hhi : w.bhi; p.blo = UNIT == 4 ? w.hlo : w.blo;
  p.unscale = UNIT == 4 ? w.h_unscale : 1.f;
  p.status = (unsigned int*)h->status.p;
  p.W = W; p.H = H; p.BH = tc.BH; p.IWB = tc.IWB; p.row_pitch = tc.IWB * G::PS; p.img_pitch = (tc.BH + 2) * p.row_pitch;
  p.NI = tc.ni; p.B = B;
  p.bands_per_img = ceil_div(H, tc.BH);
  p.n_tiles = ceil_div(B, tc.ni) * p.bands_per_img;
  p.lanes_per_img = (tc.BH / TR) * W;
  p.lanes = tc.ni * p.lanes_per_img;
  p.nstg = tc.NSTG; p.nbuf = tc.nbuf;
  p.load_bytes = (uint32_t)((size_t)G::PS * tc.IWB * (tc.BH + 2) * tc.ni * sizeof(float));
  p.trace = h->tc_trace; p.trace_tiles = h->tc_trace_tiles;
  tc_layout(CINP, COUTP, G::K8, G::N16, &p.off_b, &p.off_w, &p.off_pipe);
  p.buf_floats = align_up(G::PS * tc.IWB * (tc.BH + 2) * tc.ni, 256);
  const size_t smem = (size_t)(p.off_pipe + tc.nbuf * p.buf_floats + 256) * sizeof(float);
  HP_REQUIRE(smem <= 227 * 1024, HP_ERR_INVALID, "tc deep block <%d,%d>: %zu bytes of shared memory needed", CINP, COUTP, smem);
  HP_REQUIRE(p.lanes >= 1 && p.lanes <= 128 && tc.BH % TR == 0 && (tc.IWB * G::PS) % 32 == 0 && tc.IWB >= W && tc.IWB <= 256 &&
                 tc.BH + 2 <= 256 && tc.ni >= 1 && tc.ni <= 256 && (tc.ni == 1 || p.bands_per_img == 1) && tc.nbuf >= 2 &&
                 tc.nbuf <= TCD_MAXB && tc.NSTG >= 2 && tc.NSTG <= TC_MAX_STG && 2 * TR * G::N16 + tc.NSTG * TR * 16 <= 512 &&
                 (tc.IWB > W || W % 8 == 0) && tc.nsets <= tc.NSTG * (tc.unit >= 2 ? 1 : 2),
             HP_ERR_INVALID, "tc deep block <%d,%d>: bad geometry TR %d BH %d IWB %d W %d NI %d nbuf %d nstg %d", CINP, COUTP, TR, tc.BH,
             tc.IWB, W, tc.ni, tc.nbuf, tc.NSTG);
  CUtensorMap tin, tout;
  HP_TRY(make_map(&tin, in, B, H, W, CINP, tc.ni, tc.BH + 2, tc.IWB, G::PS));
  HP_TRY(make_map(&tout, out, B, H, W, COUTP, 1, tc.BH, tc.IWB, G::PS));
  auto kern = blaze_block_deep_kernel<CINP, COUTP, TR, NSETS, NESETS, UNIT, NISS, PLACE>;
  HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  long long grid = h->num_sms;
  if (grid > p.n_tiles) grid = p.n_tiles;
  kern<<<(unsigned)grid, 128 * NSETS + 128 * NESETS + (PLACE ? 128 * NISS : 32 * (NISS + 2)), smem, st>>>(tin, tout, p);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}


// buffer of the pixel-per-lane kernel: rows pixels + zeros for the out-of-image taps (a pixel + 8 chunks), padded to 1 KB
inline int tcs_buf_floats(int rows, int PS) { return align_up((rows + 1) * PS + 32, 256); }

template <int CINP, int COUTP, int NSETS, int NESETS>
int launch_small(hp_ctx* h, const float* in, float* out, int B, int H, int W, const BlockWeights& w, const TcCfg& tc, cudaStream_t st) {
  using G = TcGeom<CINP, COUTP>;
  TcsParams p;
  p.dww = w.dww; p.dwb = w.dwb; p.pwb = w.pwb; p.bhi = w.bhi; p.blo = w.blo;
  p.W = W; p.H = H; p.P = H * W; p.NI = tc.ni; p.rows = tc.ni * p.P;
  p.n_tiles = ceil_div(B, tc.ni);
  p.nstg = tc.NSTG; p.nbuf = tc.nbuf;
  p.ku = (tc.unit >= 1 && tc.unit <= 2) ? tc.unit : 1; p.upt = ceil_div(G::KS, p.ku);
  p.load_bytes = (uint32_t)((size_t)G::PS * p.rows * sizeof(float));
  p.trace = h->tc_trace; p.trace_tiles = h->tc_trace_tiles;
  tc_layout(CINP, COUTP, G::K8, G::N16, &p.off_b, &p.off_w, &p.off_pipe);
  p.buf_floats = tcs_buf_floats(p.rows, G::PS);
  const size_t smem = (size_t)(p.off_pipe + tc.nbuf * p.buf_floats) * sizeof(float);
  HP_REQUIRE(smem <= 227 * 1024, HP_ERR_INVALID, "tc small block <%d,%d>: %zu bytes of shared memory needed", CINP, COUTP, smem);
  HP_REQUIRE(p.rows >= 1 && p.rows <= 128 && tc.nbuf >= 2 && tc.nbuf <= TCD_MAXB && tc.NSTG >= 2 && tc.NSTG <= TC_MAX_STG &&
                 tc.nsets <= tc.NSTG && 2 * G::N16 + tc.NSTG * 16 * p.ku <= 512 && (long long)B * p.P < (1ll << 31),
             HP_ERR_INVALID, "tc small block <%d,%d>: bad geometry %dx%d NI %d nbuf %d nstg %d", CINP, COUTP, H, W, tc.ni, tc.nbuf, tc.NSTG);
  // the feature maps as [B H W pixels][C]: one box = the pixels of NI images, PS >= C floats per pixel (zero fill / clipped)
  CUtensorMap tin, tout;
  HP_TRY(make_map(&tin, in, 1, 1, B * p.P, CINP, 1, 1, p.rows, G::PS));
  HP_TRY(make_map(&tout, out, 1, 1, B * p.P, COUTP, 1, 1, p.rows, G::PS));
  HP_REQUIRE(w.h_dw != nullptr, HP_ERR_STATE, "tc small block: host copy of the depthwise weights missing");
  DwConst<CINP> dwc;
  memcpy(dwc.w, w.h_dw, sizeof(dwc));
  auto kern = blaze_block_small_kernel<CINP, COUTP, NSETS, NESETS>;
  HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  long long grid = h->num_sms;
  if (grid > p.n_tiles) grid = p.n_tiles;
  kern<<<(unsigned)grid, 128 * NSETS + 128 * NESETS + 96, smem, st>>>(tin, tout, dwc, p);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}


// shared-memory layout of the stride-2 kernel (floats); returns the bytes needed
inline size_t tcs2_layout(int cinp, int coutp, int Wo, int R, int nbuf, Tcs2Params* p) {
  const int C4 = cinp / 4, NG = coutp / 4, K8 = (cinp + 7) / 8 * 8, N16 = (coutp + 15) / 16 * 16;
  const int PSI = (C4 | 1) * 4, PSO = (NG | 1) * 4;
  p->IWB = Wo + 1;
  p->row_pitch = p->IWB * PSI;
  p->plane_floats = align_up((2 * R + 1) * p->row_pitch, 32);   // TMA destinations are 128-byte aligned
  int off = TC_BAR_FLOATS;
  p->off_b = off;
  off = align_up(off + 2 * K8 * N16, 32);
  p->off_w = off;
  off = align_up(off + coutp, 256);
  p->off_out = off;
  p->out_floats = align_up(R * Wo * PSO, 256);
  off += 2 * p->out_floats;
  p->off_in = off;
  p->in_floats = align_up(2 * p->plane_floats, 256);
  p->load_bytes = (uint32_t)((size_t)(2 * R + 1) * p->row_pitch * sizeof(float));
  return (size_t)(off + nbuf * p->in_floats) * sizeof(float);
}

template <int CINP, int COUTP, int NSETS, int NESETS>
int launch_s2(hp_ctx* h, const float* in, float* out, int B, int Hi, int Wi, int Ho, int Wo, int pad_t, int pad_l, const BlockWeights& w,
              const TcCfg& tc, cudaStream_t st) {
  using G = TcGeom<CINP, COUTP>;
  constexpr int PSI = (G::C4 | 1) * 4, PSO = (G::NG | 1) * 4;
  Tcs2Params p;
  p.pwb = w.pwb; p.bhi = w.bhi; p.blo = w.blo;
  p.Wo = Wo; p.Ho = Ho; p.R = tc.BH; p.rows = tc.BH * Wo;
  p.bands_per_img = ceil_div(Ho, tc.BH);
  p.n_tiles = B * p.bands_per_img;
  p.pad_t = pad_t; p.pad_l = pad_l;
  p.nstg = tc.NSTG; p.nbuf = tc.nbuf; p.ku = tc.unit; p.upt = ceil_div(G::KS, tc.unit);
  p.trace = h->tc_trace; p.trace_tiles = h->tc_trace_tiles;
  const size_t smem = tcs2_layout(CINP, COUTP, Wo, tc.BH, tc.nbuf, &p);
  HP_REQUIRE(smem <= 227 * 1024, HP_ERR_INVALID, "tc stride-2 block <%d,%d>: %zu bytes of shared memory needed", CINP, COUTP, smem);
  HP_REQUIRE(p.rows >= 1 && p.rows <= 128 && tc.nbuf >= 2 && tc.nbuf <= TCD_MAXB && tc.NSTG >= 2 && tc.NSTG <= TC_MAX_STG &&
                 tc.unit >= 1 && tc.unit <= 2 && tc.nsets <= tc.NSTG && 2 * G::N16 + tc.NSTG * 16 * tc.unit <= 512 && 2 * p.IWB <= 256 &&
                 2 * tc.BH + 1 <= 256 && pad_t >= 0 && pad_t <= 1 && pad_l >= 0 && pad_l <= 1,
             HP_ERR_INVALID, "tc stride-2 block <%d,%d>: bad geometry %dx%d R %d nbuf %d nstg %d", CINP, COUTP, Ho, Wo, tc.BH, tc.nbuf, tc.NSTG);
  HP_REQUIRE(w.h_dw != nullptr, HP_ERR_STATE, "tc stride-2 block: host copy of the depthwise weights missing");
  CUtensorMap tin, tout;
  HP_TRY(make_map(&tin, in, B, Hi, Wi, CINP, 1, 2 * tc.BH + 1, p.IWB, PSI, 2));
  HP_TRY(make_map(&tout, out, B, Ho, Wo, COUTP, 1, tc.BH, Wo, PSO));
  DwConst<CINP> dwc;
  memcpy(dwc.w, w.h_dw, sizeof(dwc));
  auto kern = blaze_block_s2_kernel<CINP, COUTP, NSETS, NESETS>;
  HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  long long grid = h->num_sms;
  if (grid > p.n_tiles) grid = p.n_tiles;
  kern<<<(unsigned)grid, 128 * NSETS + 128 * NESETS + 96, smem, st>>>(tin, tout, dwc, p);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

template <int CINP, int COUTP>
int launch_tc_cfg(hp_ctx* h, const float* in, float* out, int B, int H, int W, const BlockWeights& w, const TcCfg& tc,
                  cudaStream_t st) {
  constexpr int N16 = TcGeom<CINP, COUTP>::N16;
  if (tc.TR == 1) {   // pixel-per-lane kernel for maps of at most 128 pixels (npipe = epilogue warp sets)
    if constexpr (CINP >= 48) {
      if (tc.nsets == 3 && tc.npipe == 2) return launch_small<CINP, COUTP, 3, 2>(h, in, out, B, H, W, w, tc, st);
      if (tc.nsets == 3 && tc.npipe == 1) return launch_small<CINP, COUTP, 3, 1>(h, in, out, B, H, W, w, tc, st);
      if (tc.nsets == 2 && tc.npipe == 1) return launch_small<CINP, COUTP, 2, 1>(h, in, out, B, H, W, w, tc, st);
      if (tc.nsets == 4 && tc.npipe == 2) return launch_small<CINP, COUTP, 4, 2>(h, in, out, B, H, W, w, tc, st);
    }
    hp_set_error("tc small block <%d,%d>: no kernel for nsets %d esets %d", CINP, COUTP, tc.nsets, tc.npipe);
    return HP_ERR_UNSUPPORTED;
  }
#define TCD_CASE(TR_, NSETS_, NESETS_, UNIT_, NISS_, PLACE_)                             \
  if constexpr (2 * TR_ * N16 + 2 * TR_ * 16 <= 512)                                     \
    if (tc.TR == TR_ && tc.nsets == NSETS_ && tc.npipe == NESETS_ && tc.unit == UNIT_ && tc.niss == NISS_ && tc.place == PLACE_) \
      return launch_deep<CINP, COUTP, TR_, NSETS_, NESETS_, UNIT_, NISS_, PLACE_>(h, in, out, B, H, W, w, tc, st);
  if (tc.nbuf > 0) {   // warp-specialised kernel: npipe carries the number of epilogue warp sets
    // (the split-fp16 form of this kernel, UNIT 4, is not instantiated: measured at 96 x 96, batch 4096, blocks 0-5 took 2.30 ms
    // instead of 1.69 ms -- 16-channel units leave one of the two depthwise sets idle a third of the time at K = 24 and these
    // blocks are bound by the depthwise shared-memory traffic, not by the hand-offs; profiles/r02/chain_f16_96b.log)
    TCD_CASE(4, 2, 2, 1, 1, 0) TCD_CASE(4, 3, 2, 1, 1, 0) TCD_CASE(2, 2, 2, 1, 1, 0) TCD_CASE(2, 3, 2, 1, 1, 0) TCD_CASE(4, 2, 1, 1, 1, 0)
    TCD_CASE(2, 3, 1, 1, 1, 0) TCD_CASE(2, 2, 2, 2, 1, 0) TCD_CASE(2, 3, 2, 2, 1, 0) TCD_CASE(2, 4, 1, 2, 1, 0) TCD_CASE(4, 2, 2, 2, 1, 0)
    TCD_CASE(4, 3, 2, 2, 1, 0) TCD_CASE(4, 2, 2, 2, 2, 0) TCD_CASE(4, 2, 2, 2, 4, 0) TCD_CASE(4, 3, 2, 2, 2, 0) TCD_CASE(2, 3, 2, 2, 2, 0)
    TCD_CASE(2, 2, 2, 2, 2, 0) TCD_CASE(2, 3, 2, 1, 2, 0)
    TCD_CASE(4, 2, 1, 2, 2, 1) TCD_CASE(4, 2, 2, 2, 2, 1) TCD_CASE(4, 3, 1, 2, 2, 1) TCD_CASE(2, 3, 1, 2, 2, 1) TCD_CASE(2, 2, 2, 2, 2, 1)
    TCD_CASE(2, 3, 2, 2, 2, 1)
    hp_set_error("tc deep block: no kernel for TR %d nsets %d esets %d unit %d issuers %d placement %d", tc.TR, tc.nsets, tc.npipe, tc.unit,
                 tc.niss, tc.place);
    return HP_ERR_UNSUPPORTED;
  }
#undef TCD_CASE
#ifdef HP_LEGACY_KERNELS
#define TC_CASE(TR_, NSTG_, NPIPE_, NSETS_)                                              \
  if constexpr (NPIPE_ * TR_ * (N16 + 16 * NSTG_) <= 512)                                \
    if (tc.TR == TR_ && tc.NSTG == NSTG_ && tc.npipe == NPIPE_ && tc.nsets == NSETS_)    \
      return launch_tc<CINP, COUTP, TR_, NSTG_, NPIPE_, NSETS_>(h, in, out, B, H, W, w, tc, st);
  if constexpr (CINP <= 36) {   // pipelined variant: kept for the four early blocks only (comparison / fallback)
    TC_CASE(4, 2, 2, 2) TC_CASE(4, 1, 2, 2) TC_CASE(4, 1, 1, 1) TC_CASE(2, 2, 3, 2) TC_CASE(2, 2, 4, 1) TC_CASE(2, 1, 4, 1)
  }
#undef TC_CASE
#else
  if (tc.nbuf == 0) {
    hp_set_error("the pipelined tensor-core block kernel is not part of this build (test-only: rebuild with -DHP_LEGACY_KERNELS)");
    return HP_ERR_UNSUPPORTED;
  }
#endif
  hp_set_error("tc block: no kernel for TR %d NSTG %d npipe %d nsets %d", tc.TR, tc.NSTG, tc.npipe, tc.nsets);
  return HP_ERR_UNSUPPORTED;
}

}  // namespace

// Split pointwise weights [CINP][COUTP] (row-major, zero padded) into TF32 hi / lo parts laid out [K8/4][N16][4].
void hp_tc_split_weights(const float* pww, int cinp, int coutp, float* bhi, float* blo) {
  const int K8 = (cinp + 7) / 8 * 8, N16 = (coutp + 15) / 16 * 16;
  auto rna = [](float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u += 0x1000u;   // round to nearest, ties away (cvt.rna.tf32.f32)
    u &= 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
  };
  for (int k = 0; k < K8; ++k)
    for (int n = 0; n < N16; ++n) {
      const float wv = (k < cinp && n < coutp) ? pww[k * coutp + n] : 0.f;
      const float hi = rna(wv), lo = rna(wv - hi);
      const size_t idx = ((size_t)(k / 4) * N16 + n) * 4 + (k % 4);
      bhi[idx] = hi;
      blo[idx] = lo;
    }
}

int hp_tc_weight_floats(int cinp, int coutp) { return ((cinp + 7) / 8 * 8) * ((coutp + 15) / 16 * 16); }

int hp_tc_weight_floats_f16(int cinp, int coutp) { return ((cinp + 15) / 16 * 16) * ((coutp + 15) / 16 * 16) / 2; }

float hp_tc_split_weights_f16(const float* pww, int cinp, int coutp, float* hhi_f, float* hlo_f) {
  const int K16 = (cinp + 15) / 16 * 16, N16 = (coutp + 15) / 16 * 16;
  uint16_t* hhi = reinterpret_cast<uint16_t*>(hhi_f);
  uint16_t* hlo = reinterpret_cast<uint16_t*>(hlo_f);
  float wmax = 0.f;
  for (int i = 0; i < cinp * coutp; ++i) {
    const float a = fabsf(pww[i]);
    if (a > wmax && a < INFINITY) wmax = a;
  }
  int shift = 0;
  if (wmax > 0.f) {
    int e;
    frexpf(wmax, &e);                      // wmax = m * 2^e, m in [0.5, 1)
    shift = 13 - e;                        // wmax * 2^shift in [2^12, 2^13)
    if (shift > 100) shift = 100;
    if (shift < -100) shift = -100;
  }
  const float scale = ldexpf(1.f, shift);
  for (int k = 0; k < K16; ++k)
    for (int n = 0; n < N16; ++n) {
      const float wv = (k < cinp && n < coutp) ? pww[k * coutp + n] * scale : 0.f;
      const uint16_t hi = hp_f32_to_f16_rn(wv);
      const uint16_t lo = hp_f32_to_f16_rn(wv - hp_f16_to_f32(hi));
      const size_t idx = ((size_t)(k / 8) * N16 + n) * 8 + (k % 8);
      hhi[idx] = hi;
      hlo[idx] = lo;
    }
  return ldexpf(1.f, -shift);
}

// Row width (pixels) of a halo buffer of the warp-specialised kernel: no halo columns when W is a multiple of 8,
// otherwise at least one zero column on the right and rows of a multiple of 128 bytes.
static int tcd_row_pixels(int W) { return (W % 8 == 0) ? W : round_up(W + 1, 8); }

// Shared memory / TMEM feasibility of one geometry (the kernel instantiations are listed in launch_tc_cfg).
bool hp_tc_fits(int blk, int H, int W, const TcCfg& tc) {
  const int cinp = chan_pad(kBlazeBlocks[blk].cin), coutp = chan_pad(kBlazeBlocks[blk].cout);
  const int C4 = cinp / 4, NG = coutp / 4, N16 = (coutp + 15) / 16 * 16, K8 = (cinp + 7) / 8 * 8;
  const int PS = ((C4 > NG ? C4 : NG) | 1) * 4;
  const int ni = tc.nbuf > 0 ? tc.ni : 1;
  if (tc.TR == 1 && tc.nbuf > 0) {   // pixel-per-lane kernel
    if (kBlazeBlocks[blk].stride != 1 || cinp < 48 || ni < 1 || ni * H * W > 128 || tc.nbuf < 2 || tc.nbuf > TCD_MAXB || tc.NSTG < 2 ||
        tc.NSTG > TC_MAX_STG || tc.nsets < 1 || tc.nsets > tc.NSTG || tc.unit < 1 || tc.unit > 2 || 2 * N16 + tc.NSTG * 16 * tc.unit > 512 ||
        tc.npipe < 1 || tc.npipe > 2)
      return false;
    int ob, ow, op;
    tc_layout(cinp, coutp, K8, N16, &ob, &ow, &op);
    return (size_t)(op + tc.nbuf * tcs_buf_floats(ni * H * W, PS)) * 4 <= 227 * 1024;
  }
  if (tc.TR < 1 || tc.BH < tc.TR || tc.BH % tc.TR || ni < 1 || (tc.BH / tc.TR) * W * ni > 128 || tc.BH + 2 > 256) return false;
  if (tc.nbuf > 0) {
    if (2 * tc.TR * N16 + tc.NSTG * tc.TR * 16 > 512 || tc.niss < 1 || tc.niss > tc.TR ||
        128 * tc.nsets + 128 * tc.npipe + (tc.place ? 128 * tc.niss : 32 * (tc.niss + 2)) > 1024 ||
        tc.npipe < 1 || tc.npipe > 2)
      return false;
    if (tc.nbuf < 2 || tc.nbuf > TCD_MAXB || tc.NSTG < 2 || tc.NSTG > TC_MAX_STG || (ni > 1 && tc.BH < H)) return false;
    if (tc.unit < 1 || tc.unit > 2 || tc.nsets > tc.NSTG * (tc.unit == 2 ? 1 : 2)) return false;   // a set may run at most one ring phase ahead
  } else {
    if (tc.npipe * tc.TR * (N16 + 16 * tc.NSTG) > 512) return false;
    if (tc.npipe * (128 * tc.nsets + 32) > 1024) return false;
  }
  int off_b, off_w, off_pipe;
  tc_layout(cinp, coutp, K8, N16, &off_b, &off_w, &off_pipe);
  const size_t buf = (size_t)align_up(PS * tc.IWB * (tc.BH + 2) * ni, 256) * 4;
  return (size_t)off_pipe * 4 + (size_t)(tc.nbuf > 0 ? tc.nbuf : tc.npipe) * buf + 1024 <= 227 * 1024;
}

// Geometry of the warp-specialised kernel for one (TR, nsets, esets): largest ring of halo buffers that fits, whole
// images per tile when a whole image needs at most 128 lanes.
bool hp_tcd_geometry(int blk, int H, int W, int TR, int nsets, int esets, TcCfg* tc) {
  const int cinp = chan_pad(kBlazeBlocks[blk].cin), coutp = chan_pad(kBlazeBlocks[blk].cout);
  const int N16 = (coutp + 15) / 16 * 16, K8 = (cinp + 7) / 8 * 8;
  if (kBlazeBlocks[blk].stride != 1 || W > 128 || H < 1 || 2 * TR * N16 + 2 * TR * 16 > 512) return false;
  TcCfg t;
  t.TR = TR; t.nsets = nsets; t.npipe = esets; t.unit = 1; t.niss = 1; t.place = 0;
  t.NSTG = (512 - 2 * TR * N16) / (TR * 16);
  if (t.NSTG > TC_MAX_STG) t.NSTG = TC_MAX_STG;
  if (t.NSTG > K8 / 8 && K8 / 8 >= 2) t.NSTG = K8 / 8;
  t.IWB = tcd_row_pixels(W);
  const int strips = ceil_div(H, TR);
  if (strips * W <= 128) {          // whole images: as many per tile as lanes and a ring of >= 3 (then >= 2) buffers allow
    t.BH = strips * TR;
    for (int want = 3; want >= 2; --want)
      for (t.ni = 128 / (strips * W); t.ni >= 1; --t.ni)
        for (t.nbuf = TCD_MAXB; t.nbuf >= want; --t.nbuf)
          if (hp_tc_fits(blk, H, W, t)) {
            *tc = t;
            return true;
          }
    return false;
  }
  t.ni = 1;
  for (int want = 3; want >= 2; --want)
    for (int max_strips = 128 / W; max_strips >= 1; --max_strips) {
      const int bands = ceil_div(strips, max_strips);
      t.BH = ceil_div(strips, bands) * TR;
      for (t.nbuf = TCD_MAXB; t.nbuf >= want; --t.nbuf)
        if (hp_tc_fits(blk, H, W, t)) {
          *tc = t;
          return true;
        }
    }
  return false;
}

// Geometry of the pixel-per-lane kernel (maps of at most 128 pixels): whole images per tile, the deepest ring that fits.
bool hp_tcs_geometry(int blk, int H, int W, int nsets, int esets, TcCfg* tc) {
  if (H < 1 || W < 1 || H * W > 128) return false;
  TcCfg t;
  t.TR = 1; t.nsets = nsets; t.npipe = esets; t.unit = 2; t.niss = 1; t.place = 0;
  t.NSTG = TC_MAX_STG; t.BH = H; t.IWB = W;
  t.ni = 128 / (H * W);
  for (t.nbuf = TCD_MAXB; t.nbuf >= 2; --t.nbuf)
    if (hp_tc_fits(blk, H, W, t)) {
      *tc = t;
      return true;
    }
  return false;
}

// Geometry of the stride-2 tensor-core kernel: the tallest band of output rows (R Wo <= 128 lanes) whose input band fits
// shared memory at least twice next to the split weights and two output staging tiles.
bool hp_tcs2_geometry(int blk, int Ho, int Wo, int nsets, int esets, TcCfg* tc, int force_R) {
  const int cinp = chan_pad(kBlazeBlocks[blk].cin), coutp = chan_pad(kBlazeBlocks[blk].cout);
  const int N16 = (coutp + 15) / 16 * 16;
  if (kBlazeBlocks[blk].stride != 2 || Wo < 1 || Wo > 127 || Ho < 1) return false;
  TcCfg t;
  t.TR = 1; t.nsets = nsets; t.npipe = esets; t.niss = 1; t.place = 0; t.ni = 1; t.IWB = Wo + 1;
  t.unit = 2;
  t.NSTG = TC_MAX_STG;
  if (2 * N16 + t.NSTG * 16 * t.unit > 512) return false;
  int Rmax = 128 / Wo;
  if (Rmax > Ho) Rmax = Ho;
  if (force_R > 0) {                                               // tuning: explicit band height, as many buffers as fit (at most 4)
    if (force_R > Rmax) return false;
    for (int want = TCD_MAXB; want >= 2; --want) {
      Tcs2Params p;
      if (tcs2_layout(cinp, coutp, Wo, force_R, want, &p) <= 227 * 1024) {
        t.BH = force_R; t.nbuf = want;
        *tc = t;
        return true;
      }
    }
    return false;
  }
  // three input bands in flight beat taller bands (the band cycle load -> depthwise -> epilogue is latency bound: block 2
  // went from 5.3K clk per 120-pixel tile with 2 buffers to 3 buffers of 96 pixels), as long as >= 40 % of the rows remain
  // (tools/s2_sweep.py: block 5 at 88 x 88 input 0.179 ms with 2 bands of 6 rows in 2 buffers -> 0.130 in 3 buffers; at 128 x 128
  // 0.266 ms with 8 rows in 2 buffers -> 0.242 with 4 rows in 3)
  for (int want = 3; want >= 2; --want)
    for (int R = Rmax; R >= 1; --R) {
      const int bands = ceil_div(Ho, R);
      const int Rb = ceil_div(Ho, bands);                       // equally tall bands
      if (want == 3 && Rb * 10 < Rmax * 4) break;
      Tcs2Params p;
      if (tcs2_layout(cinp, coutp, Wo, Rb, want, &p) <= 227 * 1024) {
        t.BH = Rb; t.nbuf = want;
        *tc = t;
        return true;
      }
    }
  return false;
}

int hp_launch_block_tc_s2(hp_ctx* h, int blk, const float* in, float* out, int B, int Hi, int Wi, int Ho, int Wo, int pad_t, int pad_l,
                          const BlockWeights& w, const TcCfg& tc, cudaStream_t st) {
  HP_REQUIRE(w.bhi && w.blo, HP_ERR_STATE, "tc stride-2 block %d: split weights missing", blk);
#define S2_CASE(CI, CO)                                                                                                            \
  if (tc.nsets == 3 && tc.npipe == 2) return launch_s2<CI, CO, 3, 2>(h, in, out, B, Hi, Wi, Ho, Wo, pad_t, pad_l, w, tc, st);      \
  if (tc.nsets == 2 && tc.npipe == 2) return launch_s2<CI, CO, 2, 2>(h, in, out, B, Hi, Wi, Ho, Wo, pad_t, pad_l, w, tc, st);      \
  if (tc.nsets == 3 && tc.npipe == 1) return launch_s2<CI, CO, 3, 1>(h, in, out, B, Hi, Wi, Ho, Wo, pad_t, pad_l, w, tc, st);      \
  if (tc.nsets == 4 && tc.npipe == 1) return launch_s2<CI, CO, 4, 1>(h, in, out, B, Hi, Wi, Ho, Wo, pad_t, pad_l, w, tc, st);
  switch (blk) {
    case 2: S2_CASE(28, 32) break;
    case 5: S2_CASE(44, 48) break;
    case 11: S2_CASE(88, 96) break;
    default: break;
  }
#undef S2_CASE
  hp_set_error("tc stride-2 block %d: no kernel for nsets %d esets %d", blk, tc.nsets, tc.npipe);
  return HP_ERR_UNSUPPORTED;
}

// Default geometry for a stride-1 block with an H x W map; false when the tensor-core kernel does not apply.
// Choices measured with tools/tc_sweep.py (profiles/): k-step work units everywhere; 4 rows per lane where two accumulator
// sets of 4 x N16 columns fit TMEM next to a ring of >= 2 stages (N16 <= 48), else 2; maps of at most 128 pixels (the
// 6 x 6 / 8 x 8 blocks 12-15, 11 x 11 at 88 input) use the pixel-per-lane kernel: with halo buffers and 96 channels only 2-4
// images fit a tile of the band kernel and the per-tile latency chain dominates (0.21 ms against 0.12 ms per block on the
// CUDA cores at 96 x 96 input, profiles/r01/tc_sweep_96_deep.log).
bool hp_tc_choose(int blk, int H, int W, TcCfg* tc) {
  if (kBlazeBlocks[blk].stride != 1 || W > 128 || H < 1) return false;
  TcCfg a;
  if (H * W <= 128) return hp_tcs_geometry(blk, H, W, 3, 2, tc);   // pixel-per-lane kernel (blocks 6-15 on 6 x 6 ... 11 x 11 maps)
  if (W < 12) return false;
  const int n16 = (chan_pad(kBlazeBlocks[blk].cout) + 15) / 16 * 16;
  if (n16 == 48 && hp_tcd_geometry(blk, H, W, 2, 3, 1, &a)) {   // blocks 3, 4: 2 rows, 3 depthwise sets
    // up to 96 lanes: one epilogue set, issuers on the free sub-partition 3 (0.181 / 0.203 ms at 96 x 96 input against 0.181 / 0.216 with
    // two epilogue sets); with 128 lanes (128 x 128 input) that sub-partition is busy: two epilogue sets, unplaced issuers (0.277 / 0.379
    // -> 0.269 / 0.354 ms)
    const bool wide = (a.BH / a.TR) * W * a.ni > 96;
    if (wide && !hp_tcd_geometry(blk, H, W, 2, 3, 2, &a)) return false;
    a.unit = 2; a.niss = 2; a.place = wide ? 0 : 1;
    if (hp_tc_fits(blk, H, W, a)) { *tc = a; return true; }
  }
  // tiles of at most 96 lanes leave TMEM sub-partition 3 (and its scheduler) to the issuer warps: one epilogue set is then enough and
  // faster (blocks 0 / 1 at 96 x 96 input: 0.414 / 0.415 -> 0.403 / 0.406 ms; with 128 lanes -- 128 x 128 input -- it is slower: 0.674 -> 0.707)
  if (hp_tcd_geometry(blk, H, W, 4, 2, 1, &a) && (a.BH / a.TR) * W * a.ni <= 96) {
    a.unit = 2; a.niss = 2; a.place = 1;
    if (hp_tc_fits(blk, H, W, a)) { *tc = a; return true; }
  }
  if (hp_tcd_geometry(blk, H, W, 4, 2, 2, &a)) {
    a.unit = 2; a.niss = 2;
    if (hp_tc_fits(blk, H, W, a)) { *tc = a; return true; }
  }
  if (hp_tcd_geometry(blk, H, W, 2, 3, 2, &a)) {
    a.niss = 2;
    a.unit = 2;
    if (hp_tc_fits(blk, H, W, a)) { *tc = a; return true; }
    a.unit = 1;
    if (hp_tc_fits(blk, H, W, a)) { *tc = a; return true; }
  }
  return false;
}

int hp_launch_block_tc(hp_ctx* h, int blk, const float* in, float* out, int B, int H, int W, const BlockWeights& w,
                       const TcCfg& tc, cudaStream_t st) {
  HP_REQUIRE(w.bhi && w.blo, HP_ERR_STATE, "tc block %d: split weights missing", blk);
  switch (blk) {
    case 0: return launch_tc_cfg<24, 24>(h, in, out, B, H, W, w, tc, st);
    case 1: return launch_tc_cfg<24, 28>(h, in, out, B, H, W, w, tc, st);
    case 3: return launch_tc_cfg<32, 36>(h, in, out, B, H, W, w, tc, st);
    case 4: return launch_tc_cfg<36, 44>(h, in, out, B, H, W, w, tc, st);
    case 6: return launch_tc_cfg<48, 56>(h, in, out, B, H, W, w, tc, st);
    case 7: return launch_tc_cfg<56, 64>(h, in, out, B, H, W, w, tc, st);
    case 8: return launch_tc_cfg<64, 72>(h, in, out, B, H, W, w, tc, st);
    case 9: return launch_tc_cfg<72, 80>(h, in, out, B, H, W, w, tc, st);
    case 10: return launch_tc_cfg<80, 88>(h, in, out, B, H, W, w, tc, st);
    case 12: case 13: case 14: case 15: return launch_tc_cfg<96, 96>(h, in, out, B, H, W, w, tc, st);
    default: break;
  }
  hp_set_error("tc block: block %d has no tensor-core instantiation", blk);
  return HP_ERR_UNSUPPORTED;
}
