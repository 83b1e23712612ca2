// Stem (Conv2D 5x5, stride 2, SAME, 3 -> 24, + bias, ReLU) as an implicit split-fp16 GEMM on the tcgen05 tensor cores.
//
// Reference semantics: first layer `conv2d` of the Keras graph in BlazePoser/UnifiedModels/*.h5,
// called at BlazePoser/blazeFaceDetectorH5.py:272.  The CUDA-core stem (backbone.cu) runs at 17 % of the HBM roofline
// because 1800 MACs per output pixel sit on the fp32 pipe; here they go to the tensor pipe and the CUDA cores only
// gather and split the operands.
//
// GEMM view: M = output pixels, N = 24 (padded to 32), K = 80.  With NHWC and 3 input channels the 5 x 3 taps of one
// kernel row are 15 CONTIGUOUS floats of an input row, so the A row of an output pixel is 5 runs of 16 floats (the 15
// taps preceded by one zero-weighted float that keeps the run 8-byte aligned): k = ky * 16 + 1 + kx * 3 + ci.  No
// im2col buffer exists: each lane copies its runs from the input band in shared memory (LDS.64, conflict free)
// straight into the TMEM A ring, split into fp16 hi / lo parts.
//
// Precision: x = x_hi + x_lo and w = w_hi + w_lo with all four parts in fp16 (11-bit significands, lo parts may be fp16
// subnormals: absolute error <= 2^-25), D += x_hi w_hi + x_hi w_lo + x_lo w_hi in fp32.  fp16 rather than TF32 because one
// thread can issue a tcgen05.mma only every ~45 clk whatever its size (tools/mma_rate.cu) and kind::f16 covers K = 16 per
// instruction instead of 8: 60 instead of 120 MMAs per band.  fp16 needs |x| <= 65504: inputs are normalised pixels
// ([-1, 1], blazeFaceDetectorH5.py:247-269); larger magnitudes must use the CUDA-core stem (hp_debug_set_stem_tc(-1)).
//
// Warp-specialised pipeline per CTA (persistent, 1 per SM), same scheme as blaze_block_deep_kernel:
//   loader   (1 thread): TMA load of the input band (2 BH + 3 rows x W x 3) into a ring of NBUF buffers; rows above /
//                        below the image are zero-filled by the hardware (SAME padding), left / right padding is
//                        applied in registers (lanes x == 0 and x == Wo - 1)
//   gather sets (NSETS x 4 warps): lane <-> output column x of a strip of TR output rows; unit = one k-step (8 floats of a run)
//   issuers  (NISS threads in NISS warps): 3 tcgen05.mma per k-step and accumulator row into D[i & 1]
//   epilogue (NESETS x 4 warps): D + bias -> ReLU -> output staging ring (pixel stride 28 floats: conflict free)
//   storer   (1 thread): TMA store of the staged band (box wider than the 24 channels: clipped by the hardware)
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace {

constexpr int ST_K8 = 80, ST_KS = 5, ST_N16 = 32, ST_COUT = 24, ST_PSO = 28;   // ST_KS k-steps of 16 (one kernel row each)
constexpr int ST_MAXB = 4, ST_MAXO = 3, ST_MAXSTG = 4;
constexpr int ST_BAR_FLOATS = 128;

struct StemTcParams {
  const float *bhi, *blo, *bias;   // weights as fp16 [K/8][32][8] hi / lo (each 1280 floats of storage), bias [24]
  int W, H, Wo, Ho, BH, IR;        // IR = 2 BH + 3 input rows per band
  int row_floats;                  // W * 3
  int bands_per_img, n_tiles, lanes;
  int nstg, nbuf, nout;
  int IWBO;                        // staged output row width in pixels (multiple of 8, >= Wo)
  uint32_t load_bytes;
  int off_b, off_bias, off_in, in_floats, off_out, out_floats;
  long long* trace;                // optional per-tile clock64 stamps of CTA 0 (12 per tile, same slots as the block kernel)
  int trace_tiles;
  unsigned int* status;            // device word of the context: bit 0 is set when an input does not fit fp16 (|x| > 65504, inf, NaN)
};

// PLACE = 1: 4 * NISS helper warps so that issuer k is warp W_ISSUE + 4 k + 3, i.e. on SM sub-partition 3, which only hosts
// the (mostly idle) warps of TMEM lane quarter 3 when fewer than 97 lanes are active: the ~15 instructions around every
// tcgen05.mma then do not queue behind the gather / epilogue warps (133 -> ~60 clk per MMA).
template <int TR, int NSETS, int NESETS, int NISS, int PLACE>
__global__ void __launch_bounds__(128 * NSETS + 128 * NESETS + (PLACE ? 128 * NISS : 32 * (NISS + 2)), 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, StemTcParams p) {
  constexpr uint32_t colA0 = 2 * TR * ST_N16;
  static_assert(colA0 + 2 * TR * 16 <= 512, "TMEM budget");
  extern __shared__ __align__(1024) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_full = bars;                       // [nbuf]  input band landed
  uint64_t* bar_infree = bars + ST_MAXB;           // [nbuf]  all gather threads are done with the band
  uint64_t* bar_epi = bars + 2 * ST_MAXB;          // [nout]  staged output complete
  uint64_t* bar_outfree = bar_epi + ST_MAXO;       // [nout]  staged output has left shared memory
  uint64_t* bar_afull = bar_outfree + ST_MAXO;     // [nstg]
  uint64_t* bar_aempty = bar_afull + ST_MAXSTG;    // [nstg]
  uint64_t* bar_dfull = bar_aempty + ST_MAXSTG;    // [2]
  uint64_t* bar_dempty = bar_dfull + 2;            // [2]
  static_assert((2 * ST_MAXB + 2 * ST_MAXO + 2 * ST_MAXSTG + 4) * 8 + 4 <= ST_BAR_FLOATS * 4, "barrier block");
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem) + (ST_BAR_FLOATS - 1);
  float* s_bhi = smem + p.off_b;
  float* s_blo = s_bhi + ST_K8 * ST_N16 / 2;
  float* s_bias = smem + p.off_bias;
  float* in_bufs = smem + p.off_in;
  float* out_bufs = smem + p.off_out;

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane_id = tid & 31;
  constexpr int W_EPI = 4 * NSETS, W_ISSUE = W_EPI + 4 * NESETS;
  constexpr int W_LOAD = PLACE ? W_ISSUE : W_ISSUE + NISS, W_STORE = W_LOAD + 1;
  // issuer index of this warp (-1: not an issuer)
  const int issuer_of_warp = PLACE ? ((warp >= W_ISSUE && ((warp - W_ISSUE) & 3) == 3) ? (warp - W_ISSUE) >> 2 : -1)
                                   : ((warp >= W_ISSUE && warp < W_ISSUE + NISS) ? warp - W_ISSUE : -1);
  constexpr int W_ALLOC = PLACE ? W_ISSUE + 3 : W_ISSUE;   // the first issuer warp owns the TMEM allocation
  const int NSTG = p.nstg, NBUF = p.nbuf, NOUT = p.nout;

  for (int i = tid * 4; i < ST_K8 * ST_N16 / 2; i += nthr * 4) {
    st4(s_bhi + i, ld4(p.bhi + i));
    st4(s_blo + i, ld4(p.blo + i));
  }
  if (tid < ST_N16) s_bias[tid] = tid < ST_COUT ? p.bias[tid] : 0.f;
  // the zero-weighted lead float of a run and masked taps may lie just outside a band buffer: make those floats finite
  for (int i = p.off_bias + ST_N16 + tid; i < p.off_in; i += nthr) smem[i] = 0.f;
  for (int i = p.off_in + tid * 4; i < p.off_in + NBUF * p.in_floats; i += nthr * 4) st4(smem + i, make_float4(0.f, 0.f, 0.f, 0.f));
  fence_async_smem();
  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&bar_full[b], 1);
      mbar_init(&bar_infree[b], 128 * NSETS);
    }
    for (int o = 0; o < NOUT; ++o) {
      mbar_init(&bar_epi[o], 128 * NESETS);
      mbar_init(&bar_outfree[o], 1);
    }
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&bar_afull[s], 128);
      mbar_init(&bar_aempty[s], NISS);
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

  auto tile_coords = [&](int tile, int& img, int& y0) {
    img = tile / p.bands_per_img;
    y0 = (tile - img * p.bands_per_img) * p.BH;
  };
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto stamp = [&](int i, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && i < p.trace_tiles) p.trace[i * 12 + slot] = clock64();
  };

  if (warp < W_ISSUE) {
    const int wq = warp & 3;
    const int lane = wq * 32 + lane_id;
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const bool active = lane < p.lanes;
    const bool warp_active = wq * 32 < p.lanes;
    const int l = active ? lane : 0;
    const int yq = l / p.Wo;
    const int x = l - yq * p.Wo;
    if (warp < W_EPI) {
      // =============================================================== gather sets: input band -> TF32 hi / lo -> TMEM A ring
      const int set = warp >> 2;
      // work unit = one k-step = one kernel row (16 floats per output row); units are dealt round-robin to the sets over the
      // global unit counter (see blaze_block_deep_kernel: consecutive units of a set must be at most NSTG k-steps apart)
      constexpr int UPT = ST_KS;
      const uint32_t n_units = (uint32_t)my_tiles * UPT;
      // run of kernel row ky for output row t: band row 2 (yq TR + t) + ky, floats [6 x - 4, 6 x + 12)
      const int base_off = (2 * yq * TR) * p.row_floats + 6 * x - 4;
      const bool first_col = (x == 0), last_col = (x == p.Wo - 1);
      uint64_t* pending = nullptr;
      int cur_i = -1, cur_b = 0;
      const float* buf = in_bufs;
      __half2 amax = __floats2half2_rn(0.f, 0.f);   // running NaN-propagating max of |hi|: inf / NaN <=> an input that does not fit fp16
#pragma unroll 1
      for (uint32_t g = set; g < n_units; g += NSETS) {
        const int i = (int)(g / UPT);
        const int ky = (int)(g - (uint32_t)i * UPT);
        const uint32_t use = g;
        const uint32_t s = use % NSTG;
        if (i != cur_i) {
          cur_i = i;
          cur_b = i % NBUF;
          buf = in_bufs + cur_b * p.in_floats;
          mbar_wait(&bar_full[cur_b], (i / NBUF) & 1);
          if (tid == 0) stamp(i, 1);
        }
        uint32_t v[TR][16];   // per output row: 8 columns of fp16 pairs hi (k = 2c, 2c + 1), then 8 columns lo
        if (warp_active) {
          const float* src = buf + base_off + ky * p.row_floats;
#pragma unroll
          for (int t = 0; t < TR; ++t) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float2 q = *reinterpret_cast<const float2*>(src + t * 2 * p.row_floats + 2 * e);
              if (e == 0) q.x = 0.f;                                        // zero-weighted lead float: keep it finite
              if (first_col && e == 0) q.y = 0.f;                           // kx = 0 taps (floats 1..3) left of the image
              if (first_col && e == 1) q = make_float2(0.f, 0.f);
              if (last_col && e >= 5) q = make_float2(0.f, 0.f);            // kx = 3, 4 taps (floats 10..15) right of the image
              const __half2 hi = __floats2half2_rn(q.x, q.y);
              const float2 back = __half22float2(hi);
              const __half2 lo = __floats2half2_rn(q.x - back.x, q.y - back.y);
              v[t][e] = *reinterpret_cast<const uint32_t*>(&hi);
              amax = __hmax2_nan(amax, __habs2(hi));                        // one HMNMX2 per pair (three integer ops per pair cost 0.06 ms)
              v[t][8 + e] = *reinterpret_cast<const uint32_t*>(&lo);
            }
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
          const uint32_t acol = tlane + colA0 + s * (TR * 16);
#pragma unroll
          for (int t = 0; t < TR; ++t) tmem_st16(acol + t * 16, v[t]);
        }
        pending = &bar_afull[s];
        if (g + NSETS >= n_units || (int)((g + NSETS) / UPT) != i) {
          // last unit of this set in tile i: publish it and release the band (this thread reads nothing more from it)
          if (warp_active) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
          }
          mbar_arrive(pending);
          pending = nullptr;
          mbar_arrive(&bar_infree[cur_b]);
          if (tid == 0) stamp(i, 2);
          if (tid == (NSETS - 1) * 128) stamp(i, 8);
        }
      }
      {
        const uint32_t ab = *reinterpret_cast<const uint32_t*>(&amax);
        if ((((ab & 0x7C007C00u) + 0x04000400u) & 0x80008000u) && p.status) atomicOr(p.status, 1u);   // an all-ones fp16 exponent
      }
    } else {
      // =============================================================== epilogue warps
      const int eset = (warp - W_EPI) >> 2;
      int o = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int d = i & 1;
        float* ob = out_bufs + o * p.out_floats;
        mbar_wait(&bar_dfull[d], (i >> 1) & 1);
        tc_fence_after();
        if (tid == W_EPI * 32) stamp(i, 3);
        if (i >= NOUT) mbar_wait(&bar_outfree[o], ((i / NOUT) - 1) & 1);
        if (tid == W_EPI * 32) stamp(i, 11);
        if (warp_active) {
#pragma unroll
          for (int t = 0; t < TR; ++t) {
            if (t % NESETS != eset) continue;
            uint32_t v[32];
            tmem_ld32(tlane + d * (TR * ST_N16) + t * ST_N16, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float* dst = ob + ((yq * TR + t) * p.IWBO + x) * ST_PSO;
#pragma unroll
            for (int j = 0; j < ST_COUT / 4; ++j) {
              const float4 bb = ld4(s_bias + j * 4);
              const uint32_t* vv = &v[j * 4];
              float4 r4 = make_float4(fmaxf(__uint_as_float(vv[0]) + bb.x, 0.f), fmaxf(__uint_as_float(vv[1]) + bb.y, 0.f),
                                      fmaxf(__uint_as_float(vv[2]) + bb.z, 0.f), fmaxf(__uint_as_float(vv[3]) + bb.w, 0.f));
              if (active) st4(dst + j * 4, r4);
            }
          }
          tc_fence_before();
          fence_async_smem();
        }
        mbar_arrive(&bar_dempty[d]);
        mbar_arrive(&bar_epi[o]);
        if (tid == W_EPI * 32) stamp(i, 4);
        if (++o == NOUT) o = 0;
      }
    }
  } else if (lane_id == 0) {
    if (issuer_of_warp >= 0) {
      // =============================================================== MMA issuers: accumulator rows t % NISS == issuer (one thread
      // issues a tcgen05.mma only every ~46-55 clk, the tensor pipe needs 16 clk at N = 32: tools/mma_rate.cu)
      const int issuer = issuer_of_warp;
      // kind::f16: A / B fp16 (format 0), D fp32, K = 16 per instruction; B slice [32 rows][16 k] = 2 core matrices of 8 rows x 16 B
      const uint32_t idesc = (1u << 4) | ((uint32_t)(ST_N16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t desc_fixed = tc_bdesc_fixed(ST_N16);
      const uint32_t bhi_addr = smem_u32(s_bhi), blo_addr = smem_u32(s_blo);
      uint32_t use = 0;
      // every instruction of this loop is on the critical path (a lone warp retires a dependent instruction every ~5 clk):
      // descriptors advance by additions, the trace test is hoisted, the wait is the bare try_wait loop
      const bool traced = issuer == 0 && p.trace != nullptr && blockIdx.x == 0;
      const uint64_t dhi0 = desc_fixed | (uint64_t)((bhi_addr >> 4) & 0x3FFF), dlo0 = desc_fixed | (uint64_t)((blo_addr >> 4) & 0x3FFF);
      constexpr uint32_t KSTEP_DESC = (2u * ST_N16 * 16u) >> 4;       // one k-step of B in descriptor units (16 B); no carry out of the address field
      for (int i = 0; i < my_tiles; ++i) {
        const int d = i & 1;
        if (i >= 2) {
          mbar_wait_lean(&bar_dempty[d], ((i >> 1) - 1) & 1);
          tc_fence_after();
        }
        uint64_t dhi = dhi0, dlo = dlo0;
#pragma unroll 1
        for (int ks = 0; ks < ST_KS; ++ks, ++use, dhi += KSTEP_DESC, dlo += KSTEP_DESC) {
          const uint32_t s = use % NSTG;
          mbar_wait_lean(&bar_afull[s], (use / NSTG) & 1);
          tc_fence_after();
          if (traced) {
            if (ks == 0) stamp(i, 10);
            if (ks == ST_KS - 1) stamp(i, 9);
          }
#pragma unroll
          for (int t = 0; t < TR; ++t) {
            if (t % NISS != issuer) continue;
            const uint32_t dc = tmem_base + d * (TR * ST_N16) + t * ST_N16;
            const uint32_t a = tmem_base + colA0 + (s * TR + t) * 16;
            mma_f16_ts(dc, a, dhi, idesc, ks > 0 ? 1u : 0u);
            mma_f16_ts(dc, a, dlo, idesc, 1u);
            mma_f16_ts(dc, a + 8, dhi, idesc, 1u);
          }
          tc_commit(&bar_aempty[s]);
        }
        tc_commit(&bar_dfull[d]);
        if (traced) stamp(i, 7);
      }
    } else if (warp == W_LOAD) {
      // =============================================================== TMA loader
      int b = 0, i = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        if (i >= NBUF) mbar_wait(&bar_infree[b], ((i / NBUF) - 1) & 1);
        int img, y0;
        tile_coords(tile, img, y0);
        mbar_expect_tx(&bar_full[b], p.load_bytes);
        tma_load_4d(in_bufs + b * p.in_floats, &tm_in, &bar_full[b], 0, 0, 2 * y0 - 1, img);
        stamp(i, 0);
        if (++b == NBUF) b = 0;
      }
    } else if (warp == W_STORE) {
      // =============================================================== TMA storer
      int o = 0, i = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        mbar_wait(&bar_epi[o], (i / NOUT) & 1);
        int img, y0;
        tile_coords(tile, img, y0);
        tma_store_4d(&tm_out, out_bufs + o * p.out_floats, 0, 0, y0, img);
        tma_store_commit();
        stamp(i, 5);
        tma_store_wait_read();
        stamp(i, 6);
        mbar_arrive(&bar_outfree[o]);
        if (++o == NOUT) o = 0;
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


// ---------------------------------------------------------------------------------------------------------------------------
// Second generation ("flat") of the tensor-core stem.  The kernel above is instruction-issue bound on three of the four SM
// sub-partitions (ncu, profiles/r02/ncu_r02_hot_stem.txt: 9.8 K warp instructions per band of 384 pixels, lane quarter 3 idle
// because a strip of 2 x 48 columns fills 96 of the 128 TMEM lanes).  Here
//   * lane l of M-tile t owns output pixel t * 128 + l of the band in row-major order: every M-tile is full (3 instead of 4
//     M-tiles per band at 96 x 96: 45 instead of 60 MMAs, all four sub-partitions share the gather work);
//   * the band is loaded as TWO overlapping TMA boxes (left / right half of the rows) that start 4 floats left of the image
//     and end beyond its right edge: the hardware zero-fills SAME padding on all four sides, so a run is 8 unconditional
//     LDS.64 -- no per-element masks;
//   * the fp16 range guard is a GEMM column: output column 24 (N is padded to 32 anyway) has the weight 2^-10 for every k, so
//     it is inf / NaN exactly when some fp16 hi part of the pixel's 80 inputs is; the epilogue tests that one value per pixel
//     instead of a max over every gathered pair.
struct StemFlatParams {
  const float *bhi, *blo, *bias;
  int Wo, Ho, BH, IR;
  int WB;                          // floats per row of one half band (TMA box width)
  int xs;                          // output columns < xs read the left half, the others the right half
  int offR;                        // added to the offsets of right-half lanes: half_floats - (first padded column of the right box)
  int xr0;                         // row coordinate (floats) of the right box
  int np, trn;                     // output pixels / M-tiles per band
  int bands_per_img, n_tiles;
  int nstg, nbuf, nout;
  int IWBO;
  uint32_t load_bytes;             // both boxes
  int half_floats;
  int off_b, off_bias, off_in, in_floats, off_out, out_floats;
  long long* trace;
  int trace_tiles;
  unsigned int* status;
  int exp_;                        // timing experiments (env HP_STEM_EXP; results are wrong): 1 = one box per band, 2 = no negative
                                   // column coordinate, 4 = no MMAs, 8 = no gather loads / conversions
};

// Work unit = one M-tile (128 consecutive output pixels of a band) with ALL of K: the first version of this kernel handed a
// k-step of every M-tile of the band from the gather sets to the issuers (5 hand-offs per band, as the kernel above) and ran at
// 6.5 K clk per band whatever the band height, the number of sets / issuers / buffers, with the MMAs or the gather loads
// switched off (tools/stem_exp.py): every tcgen05.st -> wait::st -> mbarrier -> issuer -> tcgen05.commit -> mbarrier round trip
// costs ~1.3 K clk of latency and a ring of 4 stages cannot hide five of them per band.  Now a unit carries 15 MMAs, an
// issuer owns whole units (accumulators are never shared between issuing threads), the epilogue releases an accumulator as soon
// as it is in registers, and the four rings (A stage, accumulator) are indexed by the unit counter.
template <int TR, int NSETS, int NESETS, int NISS>
__global__ void __launch_bounds__(128 * NSETS + 128 * NESETS + 32 * (NISS + 2), 1)
stem_flat_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ StemFlatParams p) {
  constexpr int NSTG = ST_MAXSTG;                         // A stages = accumulators = units in flight
  constexpr uint32_t colA0 = NSTG * ST_N16;               // TMEM: NSTG accumulators of 32 columns, then NSTG A stages of 80
  constexpr uint32_t STAGE = ST_KS * 16;
  static_assert(colA0 + NSTG * STAGE <= 512, "TMEM budget");
  static_assert(TR == 4, "in_off selection");
  extern __shared__ __align__(1024) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_full = bars;
  uint64_t* bar_infree = bars + ST_MAXB;
  uint64_t* bar_epi = bars + 2 * ST_MAXB;
  uint64_t* bar_outfree = bar_epi + ST_MAXO;
  uint64_t* bar_afull = bar_outfree + ST_MAXO;
  uint64_t* bar_aempty = bar_afull + ST_MAXSTG;
  uint64_t* bar_dfull = bar_aempty + ST_MAXSTG;
  uint64_t* bar_dempty = bar_dfull + ST_MAXSTG;
  static_assert((2 * ST_MAXB + 2 * ST_MAXO + 4 * ST_MAXSTG) * 8 + 4 <= ST_BAR_FLOATS * 4, "barrier block");
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem) + (ST_BAR_FLOATS - 1);
  float* s_bhi = smem + p.off_b;
  float* s_blo = s_bhi + ST_K8 * ST_N16 / 2;
  float* s_bias = smem + p.off_bias;
  float* in_bufs = smem + p.off_in;
  float* out_bufs = smem + p.off_out;

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane_id = tid & 31;
  constexpr int W_EPI = 4 * NSETS, W_ISSUE = W_EPI + 4 * NESETS, W_LOAD = W_ISSUE + NISS, W_STORE = W_LOAD + 1;
  const int NBUF = p.nbuf, NOUT = p.nout;
  const int trn = p.trn;

  for (int i = tid * 4; i < ST_K8 * ST_N16 / 2; i += nthr * 4) {
    st4(s_bhi + i, ld4(p.bhi + i));
    st4(s_blo + i, ld4(p.blo + i));
  }
  if (tid < ST_N16) s_bias[tid] = tid < ST_COUT ? p.bias[tid] : 0.f;
  fence_async_smem();
  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&bar_full[b], 1);
      mbar_init(&bar_infree[b], 128 * trn);
    }
    for (int o = 0; o < NOUT; ++o) {
      mbar_init(&bar_epi[o], 128 * trn);
      mbar_init(&bar_outfree[o], 1);
    }
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&bar_afull[s], 128);
      mbar_init(&bar_aempty[s], 1);
      mbar_init(&bar_dfull[s], 1);
      mbar_init(&bar_dempty[s], 128);
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

  auto tile_coords = [&](int tile, int& img, int& y0) {
    img = tile / p.bands_per_img;
    y0 = (tile - img * p.bands_per_img) * p.BH;
  };
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t n_units = (uint32_t)my_tiles * (uint32_t)trn;      // unit u = M-tile u % trn of this CTA's band u / trn
  auto stamp = [&](int i, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && i < p.trace_tiles) p.trace[i * 12 + slot] = clock64();
  };

  if (warp < W_ISSUE) {
    const int wq = warp & 3;
    const int lane = wq * 32 + lane_id;
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
    if (warp < W_EPI) {
      // =============================================================== gather sets: band -> fp16 hi / lo -> TMEM A stage
      const int set = warp >> 2;
      int in_off[TR];                       // run of kernel row 0 of this lane's pixel in M-tile t (floats from the band buffer)
#pragma unroll
      for (int t = 0; t < TR; ++t) {
        const int pix = t * 128 + lane;
        const int pp = pix < p.np ? pix : 0;
        const int row = pp / p.Wo, x = pp - row * p.Wo;
        in_off[t] = 2 * row * p.WB + 6 * x + (x >= p.xs ? p.offR : 0);
      }
      int i = 0, t = set;
      while (t >= trn) { t -= trn; ++i; }
      int cur_i = -1, cur_b = 0;
      const float* buf = in_bufs;
#pragma unroll 1
      for (uint32_t u = set; u < n_units; u += NSETS) {
        const uint32_t s = u % NSTG;
        if (i != cur_i) {
          cur_i = i;
          cur_b = i % NBUF;
          buf = in_bufs + cur_b * p.in_floats;
          mbar_wait(&bar_full[cur_b], (i / NBUF) & 1);
          if (tid == 0) stamp(i, 1);
        }
        const float* sr = buf + (t == 0 ? in_off[0] : t == 1 ? in_off[1] : t == 2 ? in_off[2] : in_off[3]);
        if (u >= (uint32_t)NSTG) {
          mbar_wait(&bar_aempty[s], ((u / NSTG) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acol = tlane + colA0 + s * STAGE;
#pragma unroll
        for (int ky = 0; ky < ST_KS; ++ky) {
          uint32_t v[16];   // 8 columns of fp16 pairs hi (k = 2c, 2c + 1), then 8 columns lo
          if (!(p.exp_ & 8)) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float2 q = *reinterpret_cast<const float2*>(sr + ky * p.WB + 2 * e);
              const __half2 hi = __floats2half2_rn(q.x, q.y);
              const float2 back = __half22float2(hi);
              // q - back as ONE packed FMA (back * -1 is exact, so the rounding is that of the subtraction): the gather warps are
              // bound by issue slots (ncu: 75 % issue-active), not by the FMA pipe
              const float2 rest = __ffma2_rn(back, make_float2(-1.f, -1.f), q);
              const __half2 lo = __floats2half2_rn(rest.x, rest.y);
              v[e] = *reinterpret_cast<const uint32_t*>(&hi);
              v[8 + e] = *reinterpret_cast<const uint32_t*>(&lo);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0u;
          }
          tmem_st16(acol + ky * 16, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(&bar_afull[s]);
        mbar_arrive(&bar_infree[cur_b]);                    // this thread reads nothing more of the band for this unit
        if (tid == 0) stamp(i, 2);
        if (tid == (NSETS - 1) * 128) stamp(i, 8);
        t += NSETS;
        while (t >= trn) { t -= trn; ++i; }
      }
    } else {
      // =============================================================== epilogue sets
      const int eset = (warp - W_EPI) >> 2;
      int out_off[TR];
      bool act[TR];
#pragma unroll
      for (int t = 0; t < TR; ++t) {
        const int pix = t * 128 + lane;
        act[t] = pix < p.np;
        const int pp = act[t] ? pix : 0;
        const int row = pp / p.Wo, x = pp - row * p.Wo;
        out_off[t] = (row * p.IWBO + x) * ST_PSO;
      }
      uint32_t guard = 0u;                  // largest exponent field seen in the guard column
      int i = 0, t = eset;
      while (t >= trn) { t -= trn; ++i; }
      int cur_i = -1, o = 0;
#pragma unroll 1
      for (uint32_t u = eset; u < n_units; u += NESETS) {
        const uint32_t s = u % NSTG;
        if (i != cur_i) {
          cur_i = i;
          o = i % NOUT;
          if (i >= NOUT) mbar_wait(&bar_outfree[o], ((i / NOUT) - 1) & 1);
          if (tid == W_EPI * 32) stamp(i, 11);
        }
        mbar_wait(&bar_dfull[s], (u / NSTG) & 1);
        tc_fence_after();
        if (tid == W_EPI * 32 && t == 0) stamp(i, 3);
        uint32_t v[32];
        tmem_ld32(tlane + s * ST_N16, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(&bar_dempty[s]);                        // the accumulator is in registers
        const bool a = t == 0 ? act[0] : t == 1 ? act[1] : t == 2 ? act[2] : act[3];
        if (a) {
          guard = max(guard, v[ST_COUT] & 0x7F800000u);
          float* dst = out_bufs + o * p.out_floats + (t == 0 ? out_off[0] : t == 1 ? out_off[1] : t == 2 ? out_off[2] : out_off[3]);
#pragma unroll
          for (int j = 0; j < ST_COUT / 4; ++j) {
            const float4 bb = ld4(s_bias + j * 4);
            const uint32_t* vv = &v[j * 4];
            st4(dst + j * 4, make_float4(fmaxf(__uint_as_float(vv[0]) + bb.x, 0.f), fmaxf(__uint_as_float(vv[1]) + bb.y, 0.f),
                                         fmaxf(__uint_as_float(vv[2]) + bb.z, 0.f), fmaxf(__uint_as_float(vv[3]) + bb.w, 0.f)));
          }
        }
        fence_async_smem();
        mbar_arrive(&bar_epi[o]);
        if (tid == W_EPI * 32 && t == trn - 1) stamp(i, 4);
        t += NESETS;
        while (t >= trn) { t -= trn; ++i; }
      }
      if (guard == 0x7F800000u && p.status) atomicOr(p.status, 1u);
    }
  } else if (lane_id == 0) {
    if (warp < W_LOAD) {
      // =============================================================== MMA issuers: units u % NISS == issuer, 15 MMAs each
      const int issuer = warp - W_ISSUE;
      const uint32_t idesc = (1u << 4) | ((uint32_t)(ST_N16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t desc_fixed = tc_bdesc_fixed(ST_N16);
      const uint32_t bhi_addr = smem_u32(s_bhi), blo_addr = smem_u32(s_blo);
      const bool traced = p.trace != nullptr && blockIdx.x == 0;
      const uint64_t dhi0 = desc_fixed | (uint64_t)((bhi_addr >> 4) & 0x3FFF), dlo0 = desc_fixed | (uint64_t)((blo_addr >> 4) & 0x3FFF);
      constexpr uint32_t KSTEP_DESC = (2u * ST_N16 * 16u) >> 4;
      const bool no_mma = (p.exp_ & 4) != 0;
      int i = 0, t = issuer;
      while (t >= trn) { t -= trn; ++i; }
#pragma unroll 1
      for (uint32_t u = issuer; u < n_units; u += NISS) {
        const uint32_t s = u % NSTG;
        if (u >= (uint32_t)NSTG) mbar_wait_lean(&bar_dempty[s], ((u / NSTG) - 1) & 1);
        mbar_wait_lean(&bar_afull[s], (u / NSTG) & 1);
        tc_fence_after();
        if (traced && t == 0) stamp(i, 10);
        if (traced && t == trn - 1) stamp(i, 9);
        const uint32_t dc = tmem_base + s * ST_N16;
        const uint32_t a0 = tmem_base + colA0 + s * STAGE;
        if (!no_mma) {
#pragma unroll
          for (int ks = 0; ks < ST_KS; ++ks) {
            mma_f16_ts(dc, a0 + ks * 16, dhi0 + ks * KSTEP_DESC, idesc, ks > 0 ? 1u : 0u);
            mma_f16_ts(dc, a0 + ks * 16, dlo0 + ks * KSTEP_DESC, idesc, 1u);
            mma_f16_ts(dc, a0 + ks * 16 + 8, dhi0 + ks * KSTEP_DESC, idesc, 1u);
          }
        }
        tc_commit(&bar_aempty[s]);
        tc_commit(&bar_dfull[s]);
        if (traced && t == trn - 1) stamp(i, 7);
        t += NISS;
        while (t >= trn) { t -= trn; ++i; }
      }
    } else if (warp == W_LOAD) {
      // =============================================================== TMA loader: two boxes per band
      int b = 0, i = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        if (i >= NBUF) mbar_wait(&bar_infree[b], ((i / NBUF) - 1) & 1);
        int img, y0;
        tile_coords(tile, img, y0);
        float* dst = in_bufs + b * p.in_floats;
        mbar_expect_tx(&bar_full[b], (p.exp_ & 1) ? p.load_bytes / 2 : p.load_bytes);
        tma_load_4d(dst, &tm_in, &bar_full[b], (p.exp_ & 2) ? 0 : -4, 2 * y0 - 1, img, 0);
        if (!(p.exp_ & 1)) tma_load_4d(dst + p.half_floats, &tm_in, &bar_full[b], p.xr0, 2 * y0 - 1, img, 0);
        stamp(i, 0);
        if (++b == NBUF) b = 0;
      }
    } else if (warp == W_STORE) {
      // =============================================================== TMA storer
      int o = 0, i = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        mbar_wait(&bar_epi[o], (i / NOUT) & 1);
        int img, y0;
        tile_coords(tile, img, y0);
        tma_store_4d(&tm_out, out_bufs + o * p.out_floats, 0, 0, y0, img);
        tma_store_commit();
        stamp(i, 5);
        tma_store_wait_read();
        stamp(i, 6);
        mbar_arrive(&bar_outfree[o]);
        if (++o == NOUT) o = 0;
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

}  // namespace

int hp_stem_tc_weight_floats() { return ST_K8 * ST_N16 / 2; }   // fp16 elements stored two per float slot

static uint16_t f32_to_f16_rn(float f);
unsigned short hp_f32_to_f16_rn(float f) { return f32_to_f16_rn(f); }
static float f16_to_f32(uint16_t h);
float hp_f16_to_f32(unsigned short h) { return f16_to_f32(h); }
static uint16_t f32_to_f16_rn(float f) {   // round to nearest even, subnormals kept, overflow to infinity
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (x > 0x7F800000u ? 0x200u : 0u));
  if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                       // rounds to >= 65520: infinity
  if (x < 0x38800000u) {                                                         // below 2^-14: subnormal half
    if (x < 0x33000000u) return (uint16_t)sign;                                  // below 2^-25: zero
    const int e = (int)(x >> 23);
    const uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
    const int shift = 126 - e;                                                   // 14 .. 24
    uint32_t h = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((x - 0x38000000u) >> 13);
  const uint32_t rem = x & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
  return (uint16_t)(sign | h);
}
static float f16_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
  uint32_t x;
  if (e == 0) {
    if (m == 0) x = sign;
    else {
      int sh = 0;
      uint32_t mm = m;
      while (!(mm & 0x400u)) { mm <<= 1; ++sh; }
      x = sign | ((uint32_t)(113 - sh) << 23) | ((mm & 0x3FFu) << 13);
    }
  } else if (e == 31) x = sign | 0x7F800000u | (m << 13);
  else x = sign | ((e + 112u) << 23) | (m << 13);
  float f;
  memcpy(&f, &x, 4);
  return f;
}

// Stem kernel [5][5][3][24] (HWIO, = packed rows (ky*5+kx)*3+ci) -> fp16 hi / lo parts in the GEMM layout [K/8][32][8 halves]
// (K-major core matrices of 8 rows x 16 bytes), k = ky * 16 + 1 + kx * 3 + ci (k = ky * 16 is the zero-weighted alignment float).
void hp_stem_tc_split_weights(const float* w75x24, float* bhi_f, float* blo_f) {
  uint16_t* bhi = reinterpret_cast<uint16_t*>(bhi_f);
  uint16_t* blo = reinterpret_cast<uint16_t*>(blo_f);
  for (int i = 0; i < ST_K8 * ST_N16; ++i) bhi[i] = blo[i] = 0;
  for (int ky = 0; ky < 5; ++ky)
    for (int j = 0; j < 15; ++j)
      for (int n = 0; n < ST_COUT; ++n) {
        const int k = ky * 16 + 1 + j;
        const float wv = w75x24[(ky * 15 + j) * ST_COUT + n];
        const uint16_t hi = f32_to_f16_rn(wv);
        const uint16_t lo = f32_to_f16_rn(wv - f16_to_f32(hi));
        const size_t idx = ((size_t)(k / 8) * ST_N16 + n) * 8 + (k % 8);
        bhi[idx] = hi;
        blo[idx] = lo;
      }
  // guard column (stem_flat_kernel): weight 2^-10 for every k, so that output column 24 is inf / NaN exactly when an fp16 hi
  // part of the pixel's inputs is (|x| >= 65520, inf, NaN); 80 x 65504 x 2^-10 stays far inside fp32
  for (int k = 0; k < ST_K8; ++k) bhi[((size_t)(k / 8) * ST_N16 + ST_COUT) * 8 + (k % 8)] = 0x1400;
}

bool hp_stem_tc_supported(int H, int W) {
  // even sizes (pad 1 before / 2 after in both directions), rows of whole 4-pixel groups for the TMA box
  return H >= 4 && W >= 8 && H % 2 == 0 && W % 4 == 0 && W / 2 <= 128 && W / 4 <= 256;
}

// Geometry of stem_flat_kernel: the two half-band boxes and the band height with the fewest M-tiles per image.
static bool stem_flat_geometry(int H, int W, const int* cfg, StemFlatParams* p, size_t* smem_bytes) {
  constexpr int TR = 4;
  const int Ho = H / 2, Wo = W / 2;
  p->Wo = Wo; p->Ho = Ho;
  // padded row coordinates: float f of an input row sits at column f + 4; the run of output column x is columns [6 x, 6 x + 16)
  p->xs = (Wo + 1) / 2;
  const int XR = (6 * p->xs) / 4 * 4;
  p->xr0 = XR - 4;
  int wb = 6 * p->xs + 10;
  if (3 * W + 10 - XR > wb) wb = 3 * W + 10 - XR;
  p->WB = (wb + 3) / 4 * 4;
  if (p->WB > 256) return false;
  p->IWBO = round_up(Wo, 8);
  p->nstg = 4;
  int best_bh = 0, best_tiles = 1 << 30;
  size_t best_smem = 0;
  int best_nbuf = 0, best_nout = 0;
  for (int BH = 1; BH <= Ho; ++BH) {
    if (cfg && cfg[0] > 0 && BH != cfg[0]) continue;
    const int trn = ceil_div(BH * Wo, 128);
    if (trn > TR) break;
    const int IR = 2 * BH + 3;
    if (IR > 256) break;
    const int half_floats = tc_align_up(IR * p->WB, 32);
    const int in_floats = tc_align_up(half_floats + IR * p->WB, 256);
    const int out_floats = tc_align_up(BH * p->IWBO * ST_PSO, 256);
    int off = ST_BAR_FLOATS + ST_K8 * ST_N16;
    off = tc_align_up(off + ST_N16 + 16, 256);
    int nbuf = (cfg && cfg[1] > 0) ? cfg[1] : ST_MAXB, nout = (cfg && cfg[2] > 0) ? cfg[2] : ST_MAXO;
    if (nbuf > ST_MAXB) nbuf = ST_MAXB;
    if (nout > ST_MAXO) nout = ST_MAXO;
    auto total = [&]() { return (size_t)(off + nbuf * in_floats + nout * out_floats) * sizeof(float); };
    while (total() > 227 * 1024 && nout > 2) --nout;
    while (total() > 227 * 1024 && nbuf > 2) --nbuf;
    if (total() > 227 * 1024 || nbuf < 2 || nout < 2) continue;
    const int tiles = ceil_div(Ho, BH) * trn;
    if (tiles <= best_tiles) {               // ties: the taller band re-reads fewer halo rows
      best_tiles = tiles; best_bh = BH; best_smem = total(); best_nbuf = nbuf; best_nout = nout;
    }
  }
  if (best_bh == 0) return false;
  p->BH = best_bh;
  p->IR = 2 * best_bh + 3;
  p->np = best_bh * Wo;
  p->trn = ceil_div(p->np, 128);
  p->bands_per_img = ceil_div(Ho, best_bh);
  p->nbuf = best_nbuf; p->nout = best_nout;
  p->half_floats = tc_align_up(p->IR * p->WB, 32);
  p->in_floats = tc_align_up(p->half_floats + p->IR * p->WB, 256);
  p->out_floats = tc_align_up(best_bh * p->IWBO * ST_PSO, 256);
  p->offR = p->half_floats - XR;
  p->load_bytes = (uint32_t)(2u * (uint32_t)p->IR * (uint32_t)p->WB * sizeof(float));
  p->off_b = ST_BAR_FLOATS;
  p->off_bias = ST_BAR_FLOATS + ST_K8 * ST_N16;
  p->off_in = tc_align_up(p->off_bias + ST_N16 + 16, 256);
  p->off_out = p->off_in + p->nbuf * p->in_floats;
  *smem_bytes = best_smem;
  return true;
}

static int launch_stem_flat(hp_ctx* h, const float* x, float* out, int B, int H, int W, const float* bhi, const float* blo, const float* bias,
                            const int* cfg, StemFlatParams& p, size_t smem, cudaStream_t st) {
  p.bhi = bhi; p.blo = blo; p.bias = bias;
  p.trace = h->tc_trace; p.trace_tiles = h->tc_trace_tiles;
  p.status = (unsigned int*)h->status.p;
  p.n_tiles = B * p.bands_per_img;
  {
    static int exp_env = -1;
    if (exp_env < 0) {
      const char* e = getenv("HP_STEM_EXP");
      exp_env = e ? atoi(e) : 0;
    }
    p.exp_ = exp_env;
  }
  CUtensorMap tin, tout;
  {
    // input seen as [B][H][3 W floats]; the boxes reach outside on all four sides (zero fill = SAME padding)
    const cuuint64_t dims[4] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)B, 1};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 12, (cuuint64_t)H * W * 12, (cuuint64_t)B * H * W * 12};
    const cuuint32_t box[4] = {(cuuint32_t)p.WB, (cuuint32_t)p.IR, 1, 1};
    HP_TRY(tc_make_map4(&tin, x, dims, strides, box));
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)ST_COUT, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)ST_COUT * 4, (cuuint64_t)p.Wo * ST_COUT * 4, (cuuint64_t)p.Ho * p.Wo * ST_COUT * 4};
    const cuuint32_t box[4] = {(cuuint32_t)ST_PSO, (cuuint32_t)p.IWBO, (cuuint32_t)p.BH, 1};
    HP_TRY(tc_make_map4(&tout, out, dims, strides, box));
  }
  // cfg[3] = gather sets + 16 x issuers (+ 128: two epilogue sets, + 512: three); 0 = the measured best (tools/stem_sweep.py)
  int nsets = 4, niss = 3, nesets = 2;
  if (cfg && cfg[3] > 0) {
    nsets = cfg[3] % 16;
    niss = (cfg[3] / 16) % 8;
    if (niss == 0) niss = 2;
    nesets = (cfg[3] & 512) ? 3 : (cfg[3] & 128) ? 2 : 1;
  }
  long long grid = h->num_sms;
  if (grid > p.n_tiles) grid = p.n_tiles;
#define STEM_FLAT_CASE(NSETS_, NESETS_, NISS_)                                                                         \
  if (nsets == NSETS_ && nesets == NESETS_ && niss == NISS_) {                                                         \
    auto kern = stem_flat_kernel<4, NSETS_, NESETS_, NISS_>;                                                           \
    HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));                      \
    kern<<<(unsigned)grid, 128 * NSETS_ + 128 * NESETS_ + 32 * (NISS_ + 2), smem, st>>>(tin, tout, p);                 \
    h->launches++;                                                                                                     \
    HP_CUDA(cudaGetLastError());                                                                                       \
    return HP_OK;                                                                                                      \
  }
  STEM_FLAT_CASE(2, 1, 2) STEM_FLAT_CASE(3, 1, 2) STEM_FLAT_CASE(4, 1, 2) STEM_FLAT_CASE(2, 1, 1) STEM_FLAT_CASE(2, 1, 4)
  STEM_FLAT_CASE(2, 2, 2) STEM_FLAT_CASE(3, 2, 2) STEM_FLAT_CASE(4, 2, 2) STEM_FLAT_CASE(3, 2, 3) STEM_FLAT_CASE(4, 2, 3)
  STEM_FLAT_CASE(2, 2, 4) STEM_FLAT_CASE(3, 2, 4) STEM_FLAT_CASE(4, 2, 4) STEM_FLAT_CASE(3, 3, 3) STEM_FLAT_CASE(3, 3, 4) STEM_FLAT_CASE(4, 3, 2)
  STEM_FLAT_CASE(2, 3, 2) STEM_FLAT_CASE(2, 3, 4)
#undef STEM_FLAT_CASE
  hp_set_error("stem tc (flat): no kernel for %d gather sets, %d epilogue sets, %d issuers", nsets, nesets, niss);
  return HP_ERR_UNSUPPORTED;
}

int hp_launch_stem_tc(hp_ctx* h, const float* x, float* out, int B, int H, int W, const float* bhi, const float* blo, const float* bias,
                      const int* cfg, cudaStream_t st) {
  HP_REQUIRE(hp_stem_tc_supported(H, W), HP_ERR_UNSUPPORTED, "stem tc: unsupported input size %dx%d", H, W);
  // cfg[3] + 256 selects the first-generation kernel (strips of rows, lane <-> column); it also serves rows too wide for two boxes
  if (!(cfg && (cfg[3] & 256))) {
    StemFlatParams fp;
    size_t fsmem = 0;
    if (stem_flat_geometry(H, W, cfg, &fp, &fsmem)) return launch_stem_flat(h, x, out, B, H, W, bhi, blo, bias, cfg, fp, fsmem, st);
    HP_REQUIRE(!(cfg && cfg[0] > 0), HP_ERR_INVALID, "stem tc: bad band height %d", cfg[0]);
  }
  int cfg_old[4] = {cfg ? cfg[0] : 0, cfg ? cfg[1] : 0, cfg ? cfg[2] : 0, cfg ? cfg[3] & 255 : 0};
  cfg = cfg_old;
  constexpr int TR = 4;
  const int Ho = H / 2, Wo = W / 2;
  StemTcParams p;
  p.bhi = bhi; p.blo = blo; p.bias = bias;
  p.trace = h->tc_trace; p.trace_tiles = h->tc_trace_tiles;
  p.status = (unsigned int*)h->status.p;
  p.W = W; p.H = H; p.Wo = Wo; p.Ho = Ho;
  const int strips = ceil_div(Ho, TR);
  const int max_strips = 128 / Wo;
  const int bands = ceil_div(strips, max_strips);
  p.BH = ceil_div(strips, bands) * TR;
  if (cfg && cfg[0] > 0) p.BH = cfg[0];
  HP_REQUIRE(p.BH % TR == 0 && (p.BH / TR) * Wo <= 128 && p.BH >= TR, HP_ERR_INVALID, "stem tc: bad band height %d", p.BH);
  p.IR = 2 * p.BH + 3;
  p.row_floats = W * 3;
  p.bands_per_img = ceil_div(Ho, p.BH);
  p.n_tiles = B * p.bands_per_img;
  p.lanes = (p.BH / TR) * Wo;
  p.nstg = 4;
  p.IWBO = round_up(Wo, 8);
  p.load_bytes = (uint32_t)((size_t)p.row_floats * p.IR * sizeof(float));
  int off = ST_BAR_FLOATS;
  p.off_b = off;
  off += ST_K8 * ST_N16;          // hi + lo, fp16
  p.off_bias = off;
  off = tc_align_up(off + ST_N16 + 16, 256);      // >= 16 zeroed floats in front of the first band buffer
  p.off_in = off;
  p.in_floats = tc_align_up(p.row_floats * p.IR + 16, 256);
  p.out_floats = tc_align_up(p.BH * p.IWBO * ST_PSO, 256);
  p.nbuf = (cfg && cfg[1] > 0) ? cfg[1] : ST_MAXB;
  p.nout = (cfg && cfg[2] > 0) ? cfg[2] : ST_MAXO;
  auto total = [&]() { return (size_t)(p.off_in + p.nbuf * p.in_floats + p.nout * p.out_floats) * sizeof(float); };
  while (total() > 227 * 1024 && p.nout > 2) --p.nout;
  while (total() > 227 * 1024 && p.nbuf > 2) --p.nbuf;
  HP_REQUIRE(total() <= 227 * 1024 && p.nbuf >= 2 && p.nbuf <= ST_MAXB && p.nout >= 2 && p.nout <= ST_MAXO, HP_ERR_UNSUPPORTED,
             "stem tc: %zu bytes of shared memory needed for %dx%d", total(), H, W);
  p.off_out = p.off_in + p.nbuf * p.in_floats;
  // input seen as [B][H][g][W*3/g floats]: a band row is g contiguous pieces (as few as the 256-element box limit allows)
  CUtensorMap tin, tout;
  {
    int g = 1;
    while (g <= 16 && !((W * 3) % g == 0 && (W * 3 / g) <= 256 && (W * 3 / g) % 4 == 0)) ++g;
    HP_REQUIRE(g <= 16, HP_ERR_UNSUPPORTED, "stem tc: cannot split a row of %d floats into TMA boxes", W * 3);
    const int inner = W * 3 / g;
    const cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)g, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)inner * 4, (cuuint64_t)W * 12, (cuuint64_t)H * W * 12};
    const cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)g, (cuuint32_t)p.IR, 1};
    HP_TRY(tc_make_map4(&tin, x, dims, strides, box));
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)ST_COUT, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)ST_COUT * 4, (cuuint64_t)Wo * ST_COUT * 4, (cuuint64_t)Ho * Wo * ST_COUT * 4};
    const cuuint32_t box[4] = {(cuuint32_t)ST_PSO, (cuuint32_t)p.IWBO, (cuuint32_t)p.BH, 1};
    HP_TRY(tc_make_map4(&tout, out, dims, strides, box));
  }
  const int nsets = (cfg && cfg[3] > 0) ? cfg[3] % 16 : 2;
  const int niss = (cfg && cfg[3] >= 16) ? (cfg[3] / 16) % 8 : 2;
  const int place = (cfg && cfg[3] > 0) ? (cfg[3] >= 128 ? 1 : 0) : 1;   // + 128: issuers on SM sub-partition 3 (default)
  long long grid = h->num_sms;
  if (grid > p.n_tiles) grid = p.n_tiles;
  const size_t smem = total();
#define STEM_CASE(NSETS_, NESETS_, NISS_, PLACE_)                                                                      \
  if (nsets == NSETS_ && niss == NISS_ && place == PLACE_) {                                                           \
    auto kern = stem_tc_kernel<TR, NSETS_, NESETS_, NISS_, PLACE_>;                                                    \
    HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));                      \
    kern<<<(unsigned)grid, 128 * NSETS_ + 128 * NESETS_ + (PLACE_ ? 128 * NISS_ : 32 * (NISS_ + 2)), smem, st>>>(tin, tout, p); \
    h->launches++;                                                                                                     \
    HP_CUDA(cudaGetLastError());                                                                                       \
    return HP_OK;                                                                                                      \
  }
  // consecutive units of a gather set are nsets k-steps apart: at most the ring depth (see blaze_block_deep_kernel)
  HP_REQUIRE(nsets <= p.nstg, HP_ERR_INVALID, "stem tc: %d gather sets need a ring of %d stages", nsets, nsets);
  STEM_CASE(4, 1, 4, 0) STEM_CASE(4, 1, 2, 0) STEM_CASE(3, 1, 4, 0) STEM_CASE(3, 2, 2, 0) STEM_CASE(2, 2, 4, 0) STEM_CASE(2, 2, 2, 0)
  STEM_CASE(2, 1, 2, 1) STEM_CASE(3, 1, 2, 1) STEM_CASE(2, 2, 2, 1) STEM_CASE(2, 1, 1, 1)
#undef STEM_CASE
  hp_set_error("stem tc: no kernel for %d gather sets, %d issuers, placement %d", nsets, niss, place);
  return HP_ERR_UNSUPPORTED;
}
