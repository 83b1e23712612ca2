// Cross-block fusion: a CHAIN of stride-1 BlazeBlocks on one spatial size (blocks 6-10 on the H/8 map, 12-15 on the
// H/16 map) runs as ONE persistent kernel.  A tile = NI whole images; it is loaded once by TMA, stays resident in
// shared memory while every block of the chain is applied to it IN PLACE, and is stored once.  HBM traffic drops to
// "first input + last output" (12x12: 78 KB per image instead of 391 KB), the per-block TMA round trips, launches and
// weight prologues of the per-block kernels disappear.
//
// Reference semantics per block (graph called at BlazePoser/blazeFaceDetectorH5.py:272):
//   out = ReLU(Conv1x1(DepthwiseConv3x3_SAME(x) + b_dw) + b_pw + channel_pad(x)).
//
// Tile layout in shared memory: [NI][H+1][W][PS floats] -- one zero row BELOW every image (TMA zero-fills it: the box is
// H+1 rows from row 0; a TMA store must not start at a negative coordinate -- measured: illegal instruction -- so the padding
// sits on the high side) and a zero row + pixel in front of the first image.  Every vertical tap of every pixel is then a
// plain address: row -1 is the zero row of the image before (or the lead row), row H the image's own.  There is NO padding
// pixel between rows: the left tap of x = 0 / the right tap of x = W-1 read the neighbouring row's end pixel and are
// multiplied by a zeroed depthwise weight (per-lane factors on the three weights of that column).  Why: with PS = an odd
// number of 16-byte chunks (>= the widest block of the chain) the bank group of a pixel is its linear index mod 8, and
// with W + 1 pixels per row the strips of a 12 x 12 map collided (5.3 wavefronts per LDS.128 instead of 4; 7 at 6 x 6).
// Lanes are dealt to pixel columns by a host-side table (chain_lane_table) so that every quarter warp -- the 8 lanes that
// share a shared-memory wavefront of a 128-bit access -- touches 8 different bank groups whenever the geometry allows it.
// The output of a block overwrites its input pixel (C_out >= C_in).
//
// Work decomposition (lane <-> TMEM lane <-> pixel column, as in blaze_block_deep_kernel): lane l owns column x of strip
// yq of image im ((im, yq, x) = lane_tab[l]) and the TR output pixels (yq*TR + t, x); pixel t of every lane forms
// M-tile t, so one tile of the chain is TR M-tiles of at most 128 rows.  Per block ("step"):
//   DW phase : unit = one k-step (8 channels) for all TR M-tiles: sliding 3x3 window down the column -> TF32 hi / lo ->
//              tcgen05.st into the A stage of the warp set; units are dealt round-robin to NSETS sets of 4 warps.
//   MMA      : issuer thread(s): 3 tcgen05.mma per M-tile and k-step (a_hi w_hi + a_hi w_lo + a_lo w_hi) into D_t; the
//              pointwise weights stream through a ring of k-step slices (cp.async.bulk from L2, one slice = W_hi | W_lo
//              rows of 8 input channels), so nothing but 8 slices of <= 6 KB is resident.
//   EPI phase: (after all MMAs of the step: every depthwise read of the tile is done, in-place writes are safe) unit =
//              (M-tile, 32 accumulator columns): tcgen05.ld + bias + skip -> ReLU -> st.shared over the input pixel.
// The same worker warps run both phases; the next step's DW phase starts when all epilogue units have arrived.
#include <cuda_fp16.h>

#include <vector>

#include "tc_common.cuh"

namespace {

#define CH_MAXBLK 5
#define CH_RING 8                 // weight ring: k-step slices in flight
#define CH_SLOT_FLOATS 1536       // one slice: W_hi | W_lo, each [2][N16 <= 96][4] floats
#define CH_DSTRIDE 96             // TMEM columns per accumulator D_t
#define CH_BAR_FLOATS 128

struct ChainBlk {
  const float *bhi, *blo;         // pointwise weights split hi / lo: TF32 [K8/4][N16][4] floats, or (F16 kernels) fp16 [K16/8][N16][8] halves
  float unscale;                  // F16: the weights carry a power-of-two scale, the epilogue multiplies the accumulator by its inverse
  const float *dww, *pwb;         // depthwise [9][cin] followed by its bias [cin]; pointwise bias [cout]
  int cin, cout, ks, n16;         // padded channel counts (cin % 8 == 0), k-steps, MMA N
  int w_off;                      // float offset of this block's [10 cin | cout] in the shared weight area
};

struct ChainParams {
  ChainBlk blk[CH_MAXBLK];
  int nblk;
  // optional stride-2 TAIL block (block 11 behind the chain 6-10): computed from the resident tile right after the last
  // chain block, while the TMA store of the tile (the 88-channel tap) is in flight; its output goes straight to global memory
  int tail;                       // 0 / 1
  ChainBlk tblk;
  float* tail_out;                // [B][Ho][Wo][tblk.cout]
  int Ho, Wo, tail_lpi, tail_lanes, pad_t, pad_l;   // output map, lanes per image / in use, SAME padding of the depthwise conv
  int H, W, NI, B, n_tiles;
  int lanes, lpi;                 // TMEM lanes in use, lanes per image (= strips * W)
  int row_pitch;                  // W * PS floats
  // lane -> pixel column: im | strip << 8 | x << 16 (chain blocks) and im | oy << 8 | ox << 16 | swapped order << 31 (tail block), dealt so that the
  // quarter warps are free of shared-memory bank conflicts where the geometry allows it (chain_lane_table)
  uint32_t lane_tab[128], tail_tab[128];
  uint32_t load_bytes;
  int off_w, w_floats, off_ring, off_zero, zero_floats, off_tile;   // shared-memory layout in floats
  long long* trace;               // optional clock stamps of CTA 0: 8 per step
  int trace_steps;
  unsigned int* status;           // sticky status word of the context: HP_STATUS_CHAIN_RANGE when a depthwise output does not fit fp16
  int exp_;                       // timing experiments (env HP_CHAIN_EXP): 1 = no depthwise loads / math, 2 = no epilogue shared-memory traffic, 4 = no MMAs
  int dbg;                        // bring-up aid (env HP_CHAIN_DBG): 1 = setup only, 2 = + tile load / store, 3 = + first weight slices
};

// Watchdog record of the chain kernel: a wait that does not complete within ~0.25 s records {1, barrier id, parity, thread, CTA, step},
// raises the CTA-wide abort flag (every later wait of the CTA then falls through) and the kernel runs to its end with garbage
// instead of trapping: the host reads the record (hp_debug_chain_status) and the context stays usable.
__device__ unsigned int g_chain_timeout[8];

__device__ __forceinline__ void ch_wait(uint64_t* bar, uint32_t parity, int id, volatile uint32_t* s_abort, int step) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, n = 0;
  long long t0 = 0;
  while (true) {
    done = mbar_try_wait(addr, parity);
    if (done) break;
    if (*s_abort) break;
    if ((++n & 63u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 500000000ll) {
        if (atomicCAS(&g_chain_timeout[0], 0u, 1u) == 0u) {
          g_chain_timeout[1] = (unsigned)id; g_chain_timeout[2] = parity; g_chain_timeout[3] = threadIdx.x;
          g_chain_timeout[4] = blockIdx.x; g_chain_timeout[5] = (unsigned)step;
        }
        *s_abort = 1u;
        break;
      }
    }
  }
}

// Wait of the issuer loops: nothing but the try_wait loop (every instruction of an issuer's k-step is on the critical path: a lone
// warp retires a dependent instruction every ~5 clk, so the ~25 instructions of the watchdog wait cost more than the MMAs).
__device__ __forceinline__ void ch_wait_lean(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "CH_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra CH_WAIT_%=;\n\t}"
      ::"r"(addr), "r"(parity) : "memory");
}

// One arrival per WARP: 512 per-thread arrivals on one mbarrier are 512 serialised shared-memory atomics (measured: a hand-off
// round trip of the empty pipeline cost ~1.4 K clk).  __syncwarp orders the lanes' shared-memory / TMEM accesses (each lane has
// executed its own tcgen05.wait / fences) before the elected lane's releasing arrive.
// HP_CHAIN_SETSYNC = 1: the warp sets run through the blocks of a tile independently.  Set s owns the channel groups
// u = s, s + NSETS, ... of EVERY block, in the epilogue (accumulator columns 8 u ..) as in the depthwise phase (k-step u), and the
// depthwise conv of a channel group only reads that group: a set needs nothing but its own epilogue output, so a named barrier
// of its 128 threads replaces the CTA-wide "epilogue done" barrier between the blocks.  Sets that finish their epilogue units
// early start the next block's depthwise round while the others are still in the (latency-bound) epilogue; only the issuer waits
// for all of them (the first MMA of the next block overwrites the accumulators they read).
#ifndef HP_CHAIN_SETSYNC
#define HP_CHAIN_SETSYNC 1
#endif
#ifndef HP_CHAIN_DWEXP
#define HP_CHAIN_DWEXP 0          // timing experiments on the depthwise phase (see the uses); 0 in every shipped build
#endif
#ifndef HP_CHAIN_WARP_ARRIVE
#define HP_CHAIN_WARP_ARRIVE 1
#endif
__device__ __forceinline__ void ch_arrive(uint64_t* bar, int lane_id) {
#if HP_CHAIN_WARP_ARRIVE
  __syncwarp();
  if (lane_id == 0) mbar_arrive(bar);
#else
  mbar_arrive(bar);
#endif
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Thread layout: NSETS * 4 worker warps, then 4 utility warps: +0 tile loader / storer, +1 weight loader, +2 second
// issuer (NISS == 2), +3 first issuer (owns the TMEM allocation; sub-partition 3 is the least loaded one when fewer than
// 97 lanes are in use).
// F16 = 1: the pointwise conv runs as split-fp16 MMAs (kind::f16, K = 16 per instruction) instead of 3xTF32 (K = 8): the MMA count
// -- one thread issues ~110 clk per MMA next to the depthwise warps, 360 MMAs per tile of the chain 6-10 were on the critical
// path of every block -- and the B-operand shared-memory traffic halve.  x = hi + lo with both parts fp16 (22 significant bits,
// lo may be subnormal: absolute error <= 2^-25), weights pre-scaled by a power of two so that their lo parts stay normal.  Two
// worker sets (8 channels each) fill one A stage of 16 channels.  A depthwise output beyond the fp16 range (|a| >= 65520) gives an
// inf / NaN accumulator row, which the epilogue reports in the status word (the host re-runs the batch with the TF32 kernels).
template <int TR, int PS, int NSETS, int NISS, int F16>
__global__ void __launch_bounds__(128 * NSETS + 64 + 32 * NISS, 1)
blaze_chain_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ ChainParams p) {
  constexpr uint32_t colA0 = TR * CH_DSTRIDE;          // TMEM: D_0 .. D_{TR-1}, then one A stage of TR * 16 columns per set
  constexpr uint32_t STAGE = TR * 16;
  static_assert(colA0 + NSETS * STAGE <= 512, "TMEM budget");
  // F16: only NSETS / 2 A stages are in use; the columns behind them hold the max-pooled skip values of the tail block (one column
  // per input channel of the lane's output pixel)
  constexpr uint32_t colPool = colA0 + (NSETS / 2) * STAGE;
  static_assert(!F16 || colPool + 96 <= 512, "TMEM budget (pool columns)");
  // Early release of the tile by the tail block (see there).  Only in the TR = 2 instantiations (128 x 128 input: 0.983 -> 0.955 ms).
  // With TR = 3 (96 x 96 input) it is slower, 0.606 -> 0.69 .. 0.74 ms, with or without extra spills (-DHP_CHAIN_EARLY_TR3=1): the sets
  // drift further apart across the tile boundary, the epilogue of one set then runs against the depthwise loads of the others (every
  // epilogue of the tile takes 5.6-8.4 K clk instead of 3.2-4.6 K) and the slowest set paces every round.
#ifndef HP_CHAIN_EARLY_TR3
#define HP_CHAIN_EARLY_TR3 0
#endif
  constexpr bool EARLY = F16 && (TR == 2 || HP_CHAIN_EARLY_TR3);
  static_assert(!EARLY || HP_CHAIN_SETSYNC, "the parked pool values rely on epilogue unit u belonging to the set that ran k-step u");
  static_assert(NISS >= 1 && NISS <= 3 && NISS <= TR, "issuers");
  static_assert(!F16 || NSETS % 2 == 0, "two worker sets share an A stage of 16 channels");
  constexpr int NWORK = 128 * NSETS;
  constexpr int RING_HALF = F16 ? NSETS / 2 : NSETS;   // weight slices per round (= per ring half)

  extern __shared__ __align__(1024) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_tile_full = bars + 0;
  uint64_t* bar_tile_done = bars + 1;
  uint64_t* bar_dfull = bars + 2;
  uint64_t* bar_epi = bars + 3;
  // hand-offs between workers, issuers and the weight loader happen once per ROUND (= one k-step per warp set): the sets
  // work on NSETS k-steps at the same time and finish together anyway, and every mbarrier / tcgen05.commit round trip costs
  // several hundred cycles of latency whatever it carries (measured: tools/chain_check.py with HP_CHAIN_EXP)
  uint64_t* bar_afull = bars + 4;                        // all worker threads: the A stages of round R are written
  uint64_t* bar_aempty = bars + 5;                       // issuers (commit): the MMAs of round R have read them
  uint64_t* bar_wfull = bars + 6;                        // [2] weight slices of round R (ring half R & 1) have landed
  uint64_t* bar_wempty = bars + 8;                       // [2] issuers (commit): ring half R & 1 may be overwritten
  uint64_t* bar_tail_done = bars + 10;                   // all worker threads: the tail block has read the tile for the last time
  // bar_dfull[t] = bars + 2 (t = 0), bars + 11, bars + 12: the accumulator of M-tile t is complete.  The last round of a step is
  // issued tile by tile with one commit per tile, so the epilogue of tile 0 runs while the MMAs of tiles 1, 2 drain.
  static_assert(TR <= 3, "d_full barriers");
  static_assert(2 * NSETS <= CH_RING && 13 * 8 + 8 <= CH_BAR_FLOATS * 4, "weight ring / barrier block");
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem) + (CH_BAR_FLOATS - 1);
  volatile uint32_t* s_abort = reinterpret_cast<uint32_t*>(smem) + (CH_BAR_FLOATS - 2);
  float* s_w = smem + p.off_w;
  float* s_ring = smem + p.off_ring;
  float* tile = smem + p.off_tile;

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane_id = tid & 31;
  constexpr int W_UTIL = 4 * NSETS, W_TLOAD = W_UTIL, W_WLOAD = W_UTIL + 1, W_ISS0 = W_UTIL + 2;   // issuer i = warp W_ISS0 + i; issuer 0 owns the TMEM allocation

  // depthwise weights + biases of the whole chain, zeros around the tile (pads, lead pixel, trailing rows)
  for (int b = 0; b < p.nblk; ++b) {
    const ChainBlk& cb = p.blk[b];
    float* d = s_w + cb.w_off;
    for (int i = tid * 4; i < 10 * cb.cin; i += nthr * 4) st4(d + i, ld4(cb.dww + i));
    for (int i = tid * 4; i < cb.cout; i += nthr * 4) st4(d + 10 * cb.cin + i, ld4(cb.pwb + i));
  }
  if (p.tail) {
    float* d = s_w + p.tblk.w_off;
    for (int i = tid * 4; i < 10 * p.tblk.cin; i += nthr * 4) st4(d + i, ld4(p.tblk.dww + i));
    for (int i = tid * 4; i < p.tblk.cout; i += nthr * 4) st4(d + 10 * p.tblk.cin + i, ld4(p.tblk.pwb + i));
  }
  for (int i = tid * 4; i < p.zero_floats; i += nthr * 4) st4(smem + p.off_zero + i, make_float4(0.f, 0.f, 0.f, 0.f));
  fence_async_smem();
  if (tid == 0) {
    *s_abort = 0u;
    mbar_init(bar_tile_full, 1);
    constexpr uint32_t NARRIVE = HP_CHAIN_WARP_ARRIVE ? NWORK / 32 : NWORK;   // arrivals of the worker warps per phase
    mbar_init(bar_tile_done, NARRIVE);
    mbar_init(bar_dfull, 1);
    mbar_init(bars + 11, 1);
    mbar_init(bars + 12, 1);
    mbar_init(bar_epi, NARRIVE);
    mbar_init(bar_tail_done, NARRIVE);
    mbar_init(bar_afull, NARRIVE);
    mbar_init(bar_aempty, NISS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_wfull[s], 1);
      mbar_init(&bar_wempty[s], NISS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ISS0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nblk = p.nblk;
  auto stamp = [&](int step, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && step < p.trace_steps) p.trace[step * 8 + slot] = clock64();
  };

  if (p.dbg != 0) {
    if (p.dbg >= 2 && warp == W_TLOAD && lane_id == 0 && my_tiles > 0) {
      mbar_expect_tx(bar_tile_full, p.load_bytes);
      tma_load_4d(tile, &tm_in, bar_tile_full, 0, 0, 0, (int)blockIdx.x * p.NI);
      ch_wait(bar_tile_full, 0, 1, s_abort, 0);
      if (p.dbg >= 5 || (p.dbg == 4 && ((int)blockIdx.x + 1) * p.NI <= p.B)) {
        tma_store_4d(&tm_out, tile, 0, 0, 0, (int)blockIdx.x * p.NI);
        tma_store_commit();
        tma_store_wait_all();
      }
    }
    if (p.dbg >= 3 && warp == W_WLOAD && lane_id == 0) {
      const ChainBlk& cb = p.blk[0];
      const uint32_t half_bytes = (uint32_t)cb.n16 * 32u;
      const int nk = cb.ks < NSETS ? cb.ks : NSETS;
      mbar_expect_tx(&bar_wfull[0], (uint32_t)nk * 2u * half_bytes);
      for (int ks = 0; ks < nk; ++ks) {
        float* dst = s_ring + ks * CH_SLOT_FLOATS;
        bulk_g2s(dst, cb.bhi + (size_t)ks * cb.n16 * 8, half_bytes, &bar_wfull[0]);
        bulk_g2s(dst + cb.n16 * 8, cb.blo + (size_t)ks * cb.n16 * 8, half_bytes, &bar_wfull[0]);
      }
      ch_wait(&bar_wfull[0], 0, 5, s_abort, 0);
    }
  } else if (warp < W_UTIL) {
    // =============================================================== workers: depthwise units, then epilogue units
    const int set = warp >> 2, wq = warp & 3;
    const int lane = wq * 32 + lane_id;
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const bool active = lane < p.lanes;
    const bool warp_active = wq * 32 < p.lanes;
    const uint32_t lt = p.lane_tab[active ? lane : 0];
    const int im = (int)(lt & 255u), yq = (int)((lt >> 8) & 255u), x = (int)(lt >> 16);
    const int row_pitch = p.row_pitch;
    // SAME padding left / right: the taps of column x - 1 (x + 1) of the first (last) column get a zero weight
    const float fL = x == 0 ? 0.f : 1.f, fR = x == p.W - 1 ? 0.f : 1.f;
    // top-left tap of the window of output row yq*TR: buffer row im*(H+1) + yq*TR - 1 (= image row yq*TR - 1; row -1 of the
    // buffer is the lead zero row), column x - 1
    const float* win = tile + ((im * (p.H + 1) + yq * TR - 1) * p.W + x - 1) * PS;
    float* centre0 = const_cast<float*>(win) + row_pitch + PS;
    uint32_t g0 = 0, e0 = 0;     // global round counter / epilogue unit counter
    uint32_t guard = 0u;         // F16: set when an accumulator of this lane is inf / NaN
    int step = 0;
    for (int it = 0; it < my_tiles; ++it) {
      for (int b = 0; b < nblk; ++b, ++step) {
        const ChainBlk& cb = p.blk[b];
        const int cin = cb.cin, KS = cb.ks, n16 = cb.n16;
        const float* s_dww = s_w + cb.w_off;
        const float* s_pwb = s_dww + 10 * cin;
        if (b == 0) ch_wait(bar_tile_full, it & 1, 1, s_abort, step);
        else if (!HP_CHAIN_SETSYNC) ch_wait(bar_epi, (step - 1) & 1, 2, s_abort, step);
        tc_fence_after();
        if (tid == 0) stamp(step, 0);
        // ---------------- depthwise rounds: in round r this set computes k-step r * NSETS + set (if the block has it)
        const int rounds = (KS + NSETS - 1) / NSETS;
#pragma unroll 1
        for (int r = 0; r < rounds; ++r, ++g0) {
          const int ks = r * NSETS + set;
          const bool has = ks < KS && warp_active;
          float4 acc[2][TR];
          if (p.exp_ & 1) {
#pragma unroll
            for (int t = 0; t < TR; ++t) acc[0][t] = acc[1][t] = make_float4(1.f, 2.f, 3.f, 4.f);
          } else if (has) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int c = ks * 8 + half * 4;
              const float* wp = s_dww + c;
              float4 w[9];
#pragma unroll
              for (int k = 0; k < 9; ++k) w[k] = ld4(wp + k * cin);
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                w[ky * 3 + 0] = scale4(w[ky * 3 + 0], fL);
                w[ky * 3 + 2] = scale4(w[ky * 3 + 2], fR);
              }
              const float4 bias = ld4(wp + 9 * cin);
#pragma unroll
              for (int t = 0; t < TR; ++t) acc[half][t] = bias;
              const float* wc = win + c;
#pragma unroll
              for (int rr = 0; rr < TR + 2; ++rr) {
                const float* row = wc + rr * row_pitch;
#if HP_CHAIN_DWEXP == 2      // timing experiment: arithmetic without the data loads
                const float4 v0 = make_float4(acc[half][0].y, acc[half][0].x, (float)rr, acc[half][0].w), v1 = v0, v2 = v0;
#else
                const float4 v0 = ld4(row), v1 = ld4(row + PS), v2 = ld4(row + 2 * PS);
#endif
#if HP_CHAIN_DWEXP == 1      // timing experiment: data loads with next to no arithmetic
                {
                  const int t = rr < TR ? rr : TR - 1;
                  acc[half][t].x = __uint_as_float((__float_as_uint(acc[half][t].x) ^ __float_as_uint(v0.x) ^ __float_as_uint(v1.x)) ^ (__float_as_uint(v2.x) ^ __float_as_uint(v0.y) ^ __float_as_uint(v1.y)));
                  acc[half][t].y = __uint_as_float((__float_as_uint(acc[half][t].y) ^ __float_as_uint(v2.y) ^ __float_as_uint(v0.z)) ^ (__float_as_uint(v1.z) ^ __float_as_uint(v2.z) ^ __float_as_uint(v0.w)));
                  acc[half][t].z = __uint_as_float(__float_as_uint(acc[half][t].z) ^ __float_as_uint(v1.w) ^ __float_as_uint(v2.w));
                }
#else
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                  const int t = rr - ky;
                  if (t >= 0 && t < TR) {
                    acc[half][t] = fma4(v0, w[ky * 3 + 0], acc[half][t]);
                    acc[half][t] = fma4(v1, w[ky * 3 + 1], acc[half][t]);
                    acc[half][t] = fma4(v2, w[ky * 3 + 2], acc[half][t]);
                  }
                }
#endif
              }
            }
          }
          if (g0 >= 1) {                                  // the MMAs of the previous round have read the A stages
            ch_wait(bar_aempty, (g0 - 1) & 1, 3, s_abort, step);
            tc_fence_after();
          }
          if (F16) {
            // stage set / 2 holds 16 channels: columns [0, 8) fp16 pairs hi, [8, 16) lo; this set's 8 channels are pairs half * 4 ...
            const uint32_t acol = tlane + colA0 + (set >> 1) * STAGE + (set & 1) * 4;
            if (has) {
#pragma unroll
              for (int t = 0; t < TR; ++t) {
                const float f[8] = {acc[0][t].x, acc[0][t].y, acc[0][t].z, acc[0][t].w, acc[1][t].x, acc[1][t].y, acc[1][t].z, acc[1][t].w};
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
#if HP_CHAIN_DWEXP == 3      // timing experiment: no fp16 split
                  hi[e] = __float_as_uint(f[2 * e]) & 0x3fff3fffu;
                  lo[e] = __float_as_uint(f[2 * e + 1]) & 0x3fff3fffu;
#else
                  const __half2 h2 = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                  const float2 back = __half22float2(h2);
                  const float2 rest = __ffma2_rn(back, make_float2(-1.f, -1.f), make_float2(f[2 * e], f[2 * e + 1]));   // f - back, one packed FMA
                  const __half2 l2 = __floats2half2_rn(rest.x, rest.y);
                  hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
                  lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
#endif
                }
                tmem_st4(acol + t * 16, hi[0], hi[1], hi[2], hi[3]);
                tmem_st4(acol + t * 16 + 8, lo[0], lo[1], lo[2], lo[3]);
              }
            } else if (warp_active && ks >= KS && (ks ^ 1) < KS) {
              // the block ends in the middle of this stage's 16 channels: the other half multiplies zero weights, keep it finite
#pragma unroll
              for (int t = 0; t < TR; ++t) {
                tmem_st4(acol + t * 16, 0u, 0u, 0u, 0u);
                tmem_st4(acol + t * 16 + 8, 0u, 0u, 0u, 0u);
              }
            }
            if (warp_active) {
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              tc_fence_before();
            }
          } else if (has) {
            const uint32_t acol = tlane + colA0 + set * STAGE;
#pragma unroll
            for (int t = 0; t < TR; ++t) {
              const float f[8] = {acc[0][t].x, acc[0][t].y, acc[0][t].z, acc[0][t].w, acc[1][t].x, acc[1][t].y, acc[1][t].z, acc[1][t].w};
              uint32_t v[16];                             // [hi 8 | lo 8]
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                v[e] = tf32_hi(f[e]);
                v[8 + e] = __float_as_uint(f[e] - __uint_as_float(v[e]));
              }
              tmem_st16(acol + t * 16, v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
          }
          ch_arrive(bar_afull, lane_id);
        }
        if (tid == 0) stamp(step, 1);
        // ---------------- epilogue units: 8 accumulator columns of ALL TR M-tiles, dealt to the sets with (e0 + unit) % NSETS == set.
        // The bias of a unit is loaded once for the TR pixels (a warp-uniform LDS.128 costs the 4 wavefronts of a full-width one:
        // with one bias load per pixel a third of the epilogue's shared-memory traffic was bias).  The first unit of a set takes
        // the M-tiles as their accumulators complete (d_full per tile), later units load all TR rows before one tcgen05.wait.
        const int C4 = cin >> 2, NG = cb.cout >> 2;
        const int n_eu = (NG + 1) >> 1;
        (void)n16;
        if (warp_active) {
          const float us = cb.unscale;
          auto finish = [&](int t, int j, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, const float4& bb) {
            if (j < NG && !(p.exp_ & 2)) {
              float* cpix = centre0 + t * row_pitch;
              float4 o = F16 ? make_float4(fmaf(__uint_as_float(x0), us, bb.x), fmaf(__uint_as_float(x1), us, bb.y), fmaf(__uint_as_float(x2), us, bb.z),
                                           fmaf(__uint_as_float(x3), us, bb.w))
                             : make_float4(__uint_as_float(x0) + bb.x, __uint_as_float(x1) + bb.y, __uint_as_float(x2) + bb.z, __uint_as_float(x3) + bb.w);
              if (j < C4) {
                const float4 sk = ld4(cpix + j * 4);
                o.x += sk.x; o.y += sk.y; o.z += sk.z; o.w += sk.w;
              }
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
              if (active && (yq * TR + t < p.H)) st4(cpix + j * 4, o);
            }
          };
          int u = HP_CHAIN_SETSYNC ? set : (int)((NSETS + set - (e0 % NSETS)) % NSETS);
          bool first = true;
#pragma unroll 1
          for (; u < n_eu; u += NSETS) {
            const int j0 = 2 * u;
            const float4 b0 = ld4(s_pwb + j0 * 4);
            const float4 b1 = (j0 + 1 < NG) ? ld4(s_pwb + j0 * 4 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const uint32_t dcol = tlane + (uint32_t)u * 8u;
            if (first) {
              first = false;
#pragma unroll
              for (int t = 0; t < TR; ++t) {
                ch_wait(t == 0 ? bar_dfull : bars + 10 + t, step & 1, 4, s_abort, step);
                tc_fence_after();
                if (t == 0 && tid == 0) stamp(step, 2);
                uint32_t v[8];
                tmem_ld8(dcol + t * CH_DSTRIDE, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // an fp16 infinity among the A operands of this pixel makes every accumulator column inf or NaN: test one
                if (F16 && u == 0 && active && (yq * TR + t < p.H) && (v[0] & 0x7F800000u) == 0x7F800000u) guard = 1u;
                finish(t, j0, v[0], v[1], v[2], v[3], b0);
                finish(t, j0 + 1, v[4], v[5], v[6], v[7], b1);
              }
            } else {
              uint32_t v[TR][8];
#pragma unroll
              for (int t = 0; t < TR; ++t) tmem_ld8(dcol + t * CH_DSTRIDE, v[t]);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int t = 0; t < TR; ++t) {
                if (F16 && u == 0 && active && (yq * TR + t < p.H) && (v[t][0] & 0x7F800000u) == 0x7F800000u) guard = 1u;
                finish(t, j0, v[t][0], v[t][1], v[t][2], v[t][3], b0);
                finish(t, j0 + 1, v[t][4], v[t][5], v[t][6], v[t][7], b1);
              }
            }
          }
          if (first && tid == 0) stamp(step, 2);
          tc_fence_before();
          if (b == nblk - 1) fence_async_smem();           // the tile leaves through the async proxy (TMA store)
        }
        e0 += (uint32_t)n_eu;
        if (HP_CHAIN_SETSYNC) named_bar_sync(1 + set, 128);   // this set's channel groups of the block are in the tile (all 4 warps, active or not)
        ch_arrive(bar_epi, lane_id);
        if (b == nblk - 1) ch_arrive(bar_tile_done, lane_id);
        if (tid == 0) stamp(step, 3);
      }
      if (p.tail) {
        // ------------------------------------------------------------ stride-2 tail block on the finished tile (read only)
        // lane <-> OUTPUT pixel (im2, oy, ox): one M-tile.  out = ReLU(PW(DW3x3_s2(x) + b_dw) + b_pw + pad_c(maxpool2x2(x)));
        // SAME padding: the depthwise window of output y starts at input row 2 y - pad_t (TensorFlow puts the odd padding
        // pixel after), the pool window at row 2 y; everything outside the image reads the zeros around it (inputs are
        // ReLU outputs, so a zero is neutral for the max).
        const ChainBlk& cb = p.tblk;
        const int cin = cb.cin, KS = cb.ks, n16 = cb.n16;
        const float* s_dww = s_w + cb.w_off;
        const float* s_pwb = s_dww + 10 * cin;
        const bool active2 = lane < p.tail_lanes;
        const bool warp_active2 = wq * 32 < p.tail_lanes;
        const uint32_t lt2 = p.tail_tab[active2 ? lane : 0];
        const int im2 = (int)(lt2 & 255u), oy = (int)((lt2 >> 8) & 255u), ox = (int)((lt2 >> 16) & 255u);
        // neighbouring lanes are two input pixels = an even number of 16-byte chunks apart: only 4 of the 8 bank groups are hit.
        // Every second lane of a bank group (swp) takes the 4-channel halves of a k-step, and the two columns of the pool window,
        // in the opposite order -- one chunk / one pixel further is the other parity -- which makes the LDS.128 conflict-free.
        const int swp = (int)(lt2 >> 31);
        const float* win2 = tile + ((im2 * (p.H + 1) + 2 * oy - p.pad_t) * p.W + 2 * ox - p.pad_l) * PS;
        const float* pool = tile + ((im2 * (p.H + 1) + 2 * oy) * p.W + 2 * ox) * PS;
        // window columns 2 ox - pad_l + kx outside the image take a zero weight; the pool window's second column may fall off an odd map
        float fk[3];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int col = 2 * ox - p.pad_l + kx;
          fk[kx] = (col >= 0 && col < p.W) ? 1.f : 0.f;
        }
        const bool pool2 = 2 * ox + 1 < p.W;
        const float* pool_a = swp ? pool + PS : pool;             // first / second column of the pool window in this lane's order
        const float* pool_b = swp ? pool : pool + PS;
        const bool pool_va = swp ? pool2 : true, pool_vb = swp ? true : pool2;
        if (!HP_CHAIN_SETSYNC) ch_wait(bar_epi, (step - 1) & 1, 2, s_abort, step);
        tc_fence_after();
        if (tid == 0) stamp(step, 0);
        const int rounds = (KS + NSETS - 1) / NSETS;
#pragma unroll 1
        for (int r = 0; r < rounds; ++r, ++g0) {
          const int ks = r * NSETS + set;
          const bool has = ks < KS && warp_active2;
          float4 acc[2];
          if (has) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int c = ks * 8 + (half ^ swp) * 4;
              const float* wp = s_dww + c;
              float4 a = ld4(wp + 9 * cin);
#pragma unroll
              for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) a = fma4(ld4(win2 + ky * row_pitch + kx * PS + c), scale4(ld4(wp + (ky * 3 + kx) * cin), fk[kx]), a);
              acc[half] = a;
            }
            if (swp) {                                      // back to channel order
              const float4 t4 = acc[0];
              acc[0] = acc[1];
              acc[1] = t4;
            }
          }
          if (g0 >= 1) {
            ch_wait(bar_aempty, (g0 - 1) & 1, 3, s_abort, step);
            tc_fence_after();
          }
          if (F16) {
            const uint32_t acol = tlane + colA0 + (set >> 1) * STAGE + (set & 1) * 4;
            if (has) {
              const float f[8] = {acc[0].x, acc[0].y, acc[0].z, acc[0].w, acc[1].x, acc[1].y, acc[1].z, acc[1].w};
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 h2 = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                const float2 back = __half22float2(h2);
                const float2 rest = __ffma2_rn(back, make_float2(-1.f, -1.f), make_float2(f[2 * e], f[2 * e + 1]));
                const __half2 l2 = __floats2half2_rn(rest.x, rest.y);
                hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
                lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              tmem_st4(acol, hi[0], hi[1], hi[2], hi[3]);
              tmem_st4(acol + 8, lo[0], lo[1], lo[2], lo[3]);
            } else if (warp_active2 && ks >= KS && (ks ^ 1) < KS) {
              tmem_st4(acol, 0u, 0u, 0u, 0u);
              tmem_st4(acol + 8, 0u, 0u, 0u, 0u);
            }
            if (warp_active2) {
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              tc_fence_before();
            }
          } else if (has) {
            const float f[8] = {acc[0].x, acc[0].y, acc[0].z, acc[0].w, acc[1].x, acc[1].y, acc[1].z, acc[1].w};
            uint32_t v[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              v[e] = tf32_hi(f[e]);
              v[8 + e] = __float_as_uint(f[e] - __uint_as_float(v[e]));
            }
            tmem_st16(tlane + colA0 + set * STAGE, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
          }
          ch_arrive(bar_afull, lane_id);
        }
        (void)n16;
        const int C4 = cin >> 2, NG = cb.cout >> 2;
        const int n_eu = (NG + 1) >> 1;                           // units of 8 accumulator columns, dealt to the sets like the chain blocks' units
        if (EARLY) {
          // The 2x2 max-pool of the skip path is taken now, while the MMAs of the last round drain, and parked in spare TMEM columns of
          // the lane: the epilogue below then reads nothing of the tile, which is released here -- the next tile's TMA load (~3 K clk)
          // hides behind the tail block's epilogue instead of following it.
          if (warp_active2) {
#pragma unroll 1
            for (int j = 2 * set; j < C4; j += (j & 1) ? 2 * NSETS - 1 : 1) {   // the float4 groups 2 u, 2 u + 1 of this set's units u
              const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
              const float4 p00 = pool_va ? ld4(pool_a + j * 4) : z4, p01 = pool_vb ? ld4(pool_b + j * 4) : z4;
              const float4 p10 = pool_va ? ld4(pool_a + row_pitch + j * 4) : z4, p11 = pool_vb ? ld4(pool_b + row_pitch + j * 4) : z4;
              tmem_st4(tlane + colPool + (uint32_t)j * 4u, __float_as_uint(fmaxf(fmaxf(p00.x, p01.x), fmaxf(p10.x, p11.x))),
                       __float_as_uint(fmaxf(fmaxf(p00.y, p01.y), fmaxf(p10.y, p11.y))), __float_as_uint(fmaxf(fmaxf(p00.z, p01.z), fmaxf(p10.z, p11.z))),
                       __float_as_uint(fmaxf(fmaxf(p00.w, p01.w), fmaxf(p10.w, p11.w))));
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
          ch_arrive(bar_tail_done, lane_id);                      // nothing below reads the tile
        }
        if (tid == 0) stamp(step, 1);
        ch_wait(bar_dfull, step & 1, 4, s_abort, step);
        tc_fence_after();
        if (tid == 0) stamp(step, 2);
        const long long img = (long long)(blockIdx.x + (long long)it * gridDim.x) * p.NI + im2;
        const bool valid2 = active2 && img < p.B;
        float* dst = p.tail_out + ((img * p.Ho + oy) * p.Wo + ox) * (long long)cb.cout;
        if (warp_active2) {
          const float us = cb.unscale;
          int u = HP_CHAIN_SETSYNC ? set : (int)((NSETS + set - (e0 % NSETS)) % NSETS);
#pragma unroll 1
          for (; u < n_eu; u += NSETS) {
            uint32_t v[8], pk[8];
            tmem_ld8(tlane + (uint32_t)u * 8u, v);
            if (EARLY && 2 * u < C4) tmem_ld8(tlane + colPool + (uint32_t)u * 8u, pk);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (F16 && u == 0 && valid2 && (v[0] & 0x7F800000u) == 0x7F800000u) guard = 1u;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int j = 2 * u + q;
              if (j < NG) {
                const float4 bb = ld4(s_pwb + j * 4);
                float4 o = F16 ? make_float4(fmaf(__uint_as_float(v[q * 4 + 0]), us, bb.x), fmaf(__uint_as_float(v[q * 4 + 1]), us, bb.y),
                                             fmaf(__uint_as_float(v[q * 4 + 2]), us, bb.z), fmaf(__uint_as_float(v[q * 4 + 3]), us, bb.w))
                               : make_float4(__uint_as_float(v[q * 4 + 0]) + bb.x, __uint_as_float(v[q * 4 + 1]) + bb.y,
                                             __uint_as_float(v[q * 4 + 2]) + bb.z, __uint_as_float(v[q * 4 + 3]) + bb.w);
                if (EARLY) {
                  if (j < C4) {
                    o.x += __uint_as_float(pk[q * 4 + 0]); o.y += __uint_as_float(pk[q * 4 + 1]);
                    o.z += __uint_as_float(pk[q * 4 + 2]); o.w += __uint_as_float(pk[q * 4 + 3]);
                  }
                } else if (j < C4) {
                  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                  const float4 p00 = pool_va ? ld4(pool_a + j * 4) : z4, p01 = pool_vb ? ld4(pool_b + j * 4) : z4;
                  const float4 p10 = pool_va ? ld4(pool_a + row_pitch + j * 4) : z4, p11 = pool_vb ? ld4(pool_b + row_pitch + j * 4) : z4;
                  o.x += fmaxf(fmaxf(p00.x, p01.x), fmaxf(p10.x, p11.x));
                  o.y += fmaxf(fmaxf(p00.y, p01.y), fmaxf(p10.y, p11.y));
                  o.z += fmaxf(fmaxf(p00.z, p01.z), fmaxf(p10.z, p11.z));
                  o.w += fmaxf(fmaxf(p00.w, p01.w), fmaxf(p10.w, p11.w));
                }
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                if (valid2) *reinterpret_cast<float4*>(dst + j * 4) = o;
              }
            }
          }
          tc_fence_before();
        }
        e0 += (uint32_t)n_eu;
        ch_arrive(bar_epi, lane_id);
        if (!EARLY) ch_arrive(bar_tail_done, lane_id);
        if (tid == 0) stamp(step, 3);
        ++step;
      }
    }
    if (F16 && guard && p.status) atomicOr(p.status, 4u);
  } else {
    // utility warps: all 32 lanes walk the role loops together (waits included), lane 0 issues the asynchronous instructions.
    // (Keeping the warp converged measured the same as `if (lane_id == 0) {...}` with lanes 1-31 parked at the final barrier;
    // what costs is the instruction count of the issuing lane: a lone warp retires a dependent instruction every ~5 clk.)
    const bool leader = lane_id == 0;
    if (warp >= W_ISS0) {
      // =============================================================== MMA issuers (M-tiles t % NISS == issuer)
      const int issuer = warp - W_ISS0;
      const uint32_t ring_addr = smem_u32(s_ring);
      const uint32_t wfull_addr = smem_u32(bar_wfull), afull_addr = smem_u32(bar_afull);
      const uint32_t wempty_addr = smem_u32(bar_wempty), aempty_addr = smem_u32(bar_aempty), dfull_addr = smem_u32(bar_dfull), epi_addr = smem_u32(bar_epi);
      uint32_t R = 0;                                                     // global round counter
      int step = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int b = 0; b < nblk + p.tail; ++b, ++step) {
          const ChainBlk& cbi = b < nblk ? p.blk[b] : p.tblk;
          const int KS = cbi.ks, n16 = cbi.n16;
          const int trn = b < nblk ? TR : 1;                               // the tail block has one M-tile
          const uint32_t idesc = F16 ? ((1u << 4) | ((uint32_t)(n16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)) : tc_idesc_tf32(n16);
          const uint64_t desc_hi0 = tc_bdesc_fixed(n16) | (uint64_t)((ring_addr >> 4) & 0x3FFF);
          const uint32_t lo_off = ((uint32_t)n16 * 32u) >> 4;              // W_lo follows W_hi in the slot (descriptor units of 16 B)
          const bool traced = p.trace != nullptr && blockIdx.x == 0 && leader && issuer == 0;
          const bool no_mma = (p.exp_ & 4) != 0;
          const int rounds = (KS + NSETS - 1) / NSETS;
          // the first MMA of the step overwrites accumulators that the epilogue of the previous step reads (the workers
          // themselves no longer wait for each other between the blocks of a tile)
          // (also across tiles: the tail block releases the tile before its epilogue, so the first block of the next tile can be
          // under way while a set still reads the tail's accumulator)
          if (HP_CHAIN_SETSYNC && step > 0) ch_wait_lean(epi_addr, (uint32_t)(step - 1) & 1u);
#pragma unroll 1
          for (int r = 0; r < rounds; ++r, ++R) {
            const uint32_t half = R & 1u;
            ch_wait_lean(wfull_addr + half * 8, (R >> 1) & 1);
            ch_wait_lean(afull_addr, R & 1);
            tc_fence_after();
            if (leader) {
              if (traced && r == 0) stamp(step, 4);
              const int nk8 = KS - r * NSETS < NSETS ? KS - r * NSETS : NSETS;
              const int nk = F16 ? (nk8 + 1) >> 1 : nk8;                     // MMA k-steps (A stages, weight slices) of this round
              const uint64_t dhi0 = desc_hi0 + (uint64_t)(half * RING_HALF * ((CH_SLOT_FLOATS * 4) >> 4));   // the ring stays below 256 KB: no carry
              auto mma3 = [&](int t, int j, uint64_t dhi, uint32_t a0) {
                const uint64_t dlo = dhi + lo_off;
                const uint32_t acc_flag = (r | j) ? 1u : 0u;
                const uint32_t dc = tmem_base + t * CH_DSTRIDE;
                const uint32_t a = a0 + t * 16;
                if (F16) {
                  mma_f16_ts(dc, a, dhi, idesc, acc_flag);
                  mma_f16_ts(dc, a, dlo, idesc, 1u);
                  mma_f16_ts(dc, a + 8, dhi, idesc, 1u);
                } else {
                  mma_tf32_ts(dc, a, dhi, idesc, acc_flag);
                  mma_tf32_ts(dc, a, dlo, idesc, 1u);
                  mma_tf32_ts(dc, a + 8, dhi, idesc, 1u);
                }
              };
              if (r + 1 < rounds) {
                uint64_t dhi = dhi0;
                uint32_t a0 = tmem_base + colA0;
#pragma unroll 1
                for (int j = 0; j < nk; ++j, dhi += (CH_SLOT_FLOATS * 4) >> 4, a0 += STAGE) {
#pragma unroll
                  for (int t = 0; t < TR; ++t) {
                    if (t % NISS != issuer || no_mma || t >= trn) continue;
                    mma3(t, j, dhi, a0);
                  }
                }
              } else {
                // last round of the step: M-tile by M-tile (the k order within a tile is unchanged), one d_full per tile
#pragma unroll
                for (int t = 0; t < TR; ++t) {
                  if (t % NISS != issuer) continue;
                  if (!no_mma && t < trn) {
                    uint64_t dhi = dhi0;
                    uint32_t a0 = tmem_base + colA0;
#pragma unroll 1
                    for (int j = 0; j < nk; ++j, dhi += (CH_SLOT_FLOATS * 4) >> 4, a0 += STAGE) mma3(t, j, dhi, a0);
                  }
                  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t == 0 ? dfull_addr : dfull_addr + (8 + t) * 8) : "memory");
                }
                if (traced) stamp(step, 5);
              }
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(aempty_addr) : "memory");
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(wempty_addr + half * 8) : "memory");
            }
            __syncwarp();
          }
        }
      }
    } else if (warp == W_WLOAD) {
      // =============================================================== weight ring loader (the slices of one round per barrier)
      uint32_t R = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int b = 0; b < nblk + p.tail; ++b) {
          const ChainBlk& cb = b < nblk ? p.blk[b] : p.tblk;
          const uint32_t half_bytes = (uint32_t)cb.n16 * 32u;          // [2][n16][4] floats
          const int rounds = (cb.ks + NSETS - 1) / NSETS;
#pragma unroll 1
          for (int r = 0; r < rounds; ++r, ++R) {
            const uint32_t half = R & 1u;
            if (R >= 2) ch_wait(&bar_wempty[half], ((R >> 1) - 1) & 1, 7, s_abort, (int)R);
            if (leader) {
              const int nk8 = cb.ks - r * NSETS < NSETS ? cb.ks - r * NSETS : NSETS;
              const int nk = F16 ? (nk8 + 1) >> 1 : nk8;
              mbar_expect_tx(&bar_wfull[half], (uint32_t)nk * 2u * half_bytes);
              for (int j = 0; j < nk; ++j) {
                const int ks = (F16 ? (r * NSETS) >> 1 : r * NSETS) + j;
                float* dst = s_ring + (half * RING_HALF + j) * CH_SLOT_FLOATS;
                bulk_g2s(dst, cb.bhi + (size_t)ks * cb.n16 * 8, half_bytes, &bar_wfull[half]);
                bulk_g2s(dst + cb.n16 * 8, cb.blo + (size_t)ks * cb.n16 * 8, half_bytes, &bar_wfull[half]);
              }
            }
            __syncwarp();
          }
        }
      }
    } else if (warp == W_TLOAD) {
      // =============================================================== tile loader / storer
      int tile_idx = blockIdx.x;
      if (my_tiles > 0 && leader) {
        mbar_expect_tx(bar_tile_full, p.load_bytes);
        tma_load_4d(tile, &tm_in, bar_tile_full, 0, 0, 0, tile_idx * p.NI);
      }
      for (int it = 0; it < my_tiles; ++it, tile_idx += gridDim.x) {
        const int next = tile_idx + (int)gridDim.x;
        if (it + 1 < my_tiles && leader) tma_prefetch_4d(&tm_in, 0, 0, 0, next * p.NI);   // warm L2 while this tile is computed
        ch_wait(bar_tile_done, it & 1, 8, s_abort, it);
        if (leader) {
          tma_store_4d(&tm_out, tile, 0, 0, 0, tile_idx * p.NI);
          tma_store_commit();
          stamp(it * (nblk + p.tail) + nblk - 1, 6);
        }
        if (p.tail) ch_wait(bar_tail_done, it & 1, 9, s_abort, it);          // the tail block reads the tile while the store is in flight
        if (leader) {
          tma_store_wait_read();
          stamp(it * (nblk + p.tail) + nblk - 1, 7);
          if (it + 1 < my_tiles) {
            mbar_expect_tx(bar_tile_full, p.load_bytes);
            tma_load_4d(tile, &tm_in, bar_tile_full, 0, 0, 0, next * p.NI);
          }
        }
        __syncwarp();
      }
      if (leader) tma_store_wait_all();
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ISS0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int TR, int PS, int NSETS, int NISS, int F16>
int launch_chain(hp_ctx* h, const CUtensorMap& tin, const CUtensorMap& tout, const ChainParams& p, size_t smem, cudaStream_t st) {
  auto kern = blaze_chain_kernel<TR, PS, NSETS, NISS, F16>;
  HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  long long grid = h->num_sms;
  if (grid > p.n_tiles) grid = p.n_tiles;
  kern<<<(unsigned)grid, 128 * NSETS + 64 + 32 * NISS, smem, st>>>(tin, tout, p);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

}  // namespace

// Watchdog record of the last chain kernels (see ch_wait): out[0] != 0 means a wait timed out; clears the record.
int hp_chain_status(unsigned int out[8]) {
  HP_CUDA(cudaDeviceSynchronize());
  HP_CUDA(cudaMemcpyFromSymbol(out, g_chain_timeout, 8 * sizeof(unsigned int)));
  unsigned int z[8] = {};
  HP_CUDA(cudaMemcpyToSymbol(g_chain_timeout, z, sizeof(z)));
  return HP_OK;
}

// Geometry of the chain kernel for blocks [first, first + nblk) on an H x W map: rows per lane TR, images per tile NI.
// Returns false when the chain kernel does not apply (the per-block kernels are used instead).
bool hp_chain_geometry(int first, int nblk, int chain_nblk, int H, int W, ChainCfg* cfg, int tail_blk, bool f16) {
  if (nblk < 1 || nblk > CH_MAXBLK || first < 0 || first + nblk > 16 || H < 2 || W < 2 || W > 64) return false;
  // chain_nblk >= nblk: length of the full chain (a truncated chain, used to read intermediate activations, keeps the geometry
  // -- pixel stride, rows per lane, images per tile -- of the full one)
  if (chain_nblk < nblk || first + chain_nblk > 16) return false;
  int cmax = 0;
  for (int b = first; b < first + chain_nblk; ++b) {
    const int cin = chan_pad(kBlazeBlocks[b].cin), cout = chan_pad(kBlazeBlocks[b].cout);
    if (kBlazeBlocks[b].stride != 1 || cin % 8 != 0 || cout < cin || cout > 96) return false;
    if (b > first && kBlazeBlocks[b].cin != kBlazeBlocks[b - 1].cout) return false;
    if (cout > cmax) cmax = cout;
  }
  const int PS = ((cmax / 4) | 1) * 4;
  if (PS != 92 && PS != 100) return false;               // instantiated pixel strides (chains ending at 88 / 96 channels)
  int w_floats = 0;
  for (int b = first; b < first + chain_nblk; ++b) w_floats += 10 * chan_pad(kBlazeBlocks[b].cin) + chan_pad(kBlazeBlocks[b].cout);
  if (tail_blk >= 0) {
    // a stride-2 block behind the chain: same input channels as the chain's output, at most 96 outputs, one M-tile of output pixels
    if (tail_blk != first + chain_nblk || tail_blk > 15 || kBlazeBlocks[tail_blk].stride != 2 || kBlazeBlocks[tail_blk].cin != kBlazeBlocks[tail_blk - 1].cout ||
        chan_pad(kBlazeBlocks[tail_blk].cin) % 8 != 0 || chan_pad(kBlazeBlocks[tail_blk].cout) > 96 || nblk != chain_nblk)
      return false;
    w_floats += 10 * chan_pad(kBlazeBlocks[tail_blk].cin) + chan_pad(kBlazeBlocks[tail_blk].cout);
  }
  ChainCfg best;
  double best_cost = 1e30;
  for (int TR = 2; TR <= 3; ++TR) {
    const int strips = ceil_div(H, TR), lpi = strips * W;
    if (lpi > 128) continue;
    for (int NI = 128 / lpi; NI >= 1; --NI) {
      ChainCfg c;
      c.TR = TR; c.NI = NI; c.PS = PS; c.lanes = NI * lpi; c.lpi = lpi;
      c.nsets = 4; c.niss = 1;   // measured at 96 x 96, batch 4096: 1 issuer 0.589 ms, 2 issuers 0.609, 3 issuers 0.659 (every extra issuer adds its commits to each round)
      const int lead = tc_align_up((W + 1) * PS, 32);     // zero row above the first image + the left neighbour of its first pixel
      const int rows = NI * (H + 1) + TR;                 // image rows + their zero rows, slack for partial strips
      int off = CH_BAR_FLOATS;
      c.off_w = off; off = tc_align_up(off + w_floats, 32);
      c.ring_slots = f16 ? CH_RING / 2 : CH_RING;        // split fp16: NSETS / 2 slices of 16 channels per round, two rounds in flight
      c.off_ring = off; off += c.ring_slots * CH_SLOT_FLOATS;
      c.off_zero = off; off += lead;
      c.off_tile = off;
      c.tile_floats = rows * W * PS;
      off += c.tile_floats + PS;                          // one more pixel: the right neighbour of the last pixel
      c.zero_floats = off - c.off_zero;
      c.w_floats = w_floats;
      c.smem = (size_t)off * sizeof(float);
      if (c.smem > 227 * 1024) continue;
      if (tail_blk >= 0 && NI * ceil_div(H, 2) * ceil_div(W, 2) > 128) continue;
      // cost: warp-instructions ~ active warps * M-tiles per image; more images per tile amortise the per-step bubbles
      // (preferring TR = 3 for its fewer shared-memory loads per output was measured: 6 x 6 maps with 10 images per tile 0.151 -> 0.149 ms,
      // 16 x 16 and 8 x 8 maps at 128 x 128 input 0.961 -> 1.169 / 0.258 -> 0.279 ms: the cost below stays)
      const double cost = (double)ceil_div(c.lanes, 32) * TR / NI + 0.05 * TR / NI;
      if (cost < best_cost) { best_cost = cost; best = c; }
      break;                                              // the largest NI that fits is the best one for this TR
    }
  }
  if (best_cost >= 1e30) return false;
  *cfg = best;
  return true;
}

// lane -> pixel column tables of the chain kernel for an H x W map (host only)
static void chain_tables(int H, int W, int TR, int NI, bool tail, uint32_t* lane_tab, uint32_t* tail_tab) {
  const int strips = ceil_div(H, TR);
  std::vector<int> pix;
  std::vector<uint32_t> code;
  for (int im = 0; im < NI; ++im)
    for (int yq = 0; yq < strips; ++yq)
      for (int x = 0; x < W; ++x) {
        pix.push_back((im * (H + 1) + yq * TR) * W + x);
        code.push_back((uint32_t)im | ((uint32_t)yq << 8) | ((uint32_t)x << 16));
      }
  tc_lane_table(pix, code, lane_tab);
  if (!tail) return;
  pix.clear();
  code.clear();
  const int Ho = ceil_div(H, 2), Wo = ceil_div(W, 2);
  for (int im = 0; im < NI; ++im)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox) {
        pix.push_back((im * (H + 1) + 2 * oy) * W + 2 * ox);
        code.push_back((uint32_t)im | ((uint32_t)oy << 8) | ((uint32_t)ox << 16));
      }
  tc_lane_table(pix, code, tail_tab, 0x80000000u);
}

// Host-side geometry of the chain kernel for blocks [first, first + nblk) (+ the stride-2 block behind them when tail != 0) on an
// H x W map, without touching the device: out[0..7] = {TR, NI, PS, lanes, tail lanes, shared-memory bytes, 0, 0}, out[8..135] = lane table,
// out[136..263] = tail table.  Returns HP_ERR_UNSUPPORTED when the chain kernel does not apply.
int hp_chain_describe(int first, int nblk, int H, int W, int tail, unsigned int* out264) {
  ChainCfg cfg;
  const int tail_blk = tail ? first + nblk : -1;
  if (!hp_chain_geometry(first, nblk, nblk, H, W, &cfg, tail_blk)) {
    hp_set_error("chain %d..%d does not apply to a %dx%d map", first, first + nblk - 1, H, W);
    return HP_ERR_UNSUPPORTED;
  }
  memset(out264, 0, 264 * sizeof(unsigned int));
  out264[0] = cfg.TR; out264[1] = cfg.NI; out264[2] = cfg.PS; out264[3] = cfg.lanes;
  out264[4] = tail ? cfg.NI * ceil_div(H, 2) * ceil_div(W, 2) : 0;
  out264[5] = (unsigned int)cfg.smem;
  chain_tables(H, W, cfg.TR, cfg.NI, tail != 0, out264 + 8, out264 + 136);
  return HP_OK;
}

int hp_launch_chain(hp_ctx* h, int first, int nblk, const float* in, float* out, int B, int H, int W, const ChainCfg& cfg,
                    cudaStream_t st, int tail_blk, float* tail_out) {
  const Backbone& bb = h->bb;
  ChainParams p;
  memset(&p, 0, sizeof(p));
  const bool f16 = !(h->chain_mode & 4) && cfg.nsets % 2 == 0 && cfg.niss <= 2;   // chain_mode + 4: 3xTF32 products
  HP_REQUIRE(cfg.ring_slots >= (f16 ? cfg.nsets : 2 * cfg.nsets), HP_ERR_STATE, "chain: weight ring of %d slices is too small for the %s kernel",
             cfg.ring_slots, f16 ? "split-fp16" : "3xTF32");
  p.status = (unsigned int*)h->status.p;
  p.nblk = nblk;
  int w_off = 0;
  for (int b = 0; b < nblk; ++b) {
    const BlockWeights& w = bb.blk[first + b];
    HP_REQUIRE(w.bhi && w.blo, HP_ERR_STATE, "chain: split weights of block %d missing", first + b);
    ChainBlk& cb = p.blk[b];
    cb.bhi = f16 ? w.hhi : w.bhi; cb.blo = f16 ? w.hlo : w.blo; cb.dww = w.dww; cb.pwb = w.pwb;
    cb.unscale = f16 ? w.h_unscale : 1.f;
    cb.cin = chan_pad(kBlazeBlocks[first + b].cin);
    cb.cout = chan_pad(kBlazeBlocks[first + b].cout);
    cb.ks = cb.cin / 8;
    cb.n16 = (cb.cout + 15) / 16 * 16;
    cb.w_off = w_off;
    w_off += 10 * cb.cin + cb.cout;
    HP_REQUIRE(w.dwb == w.dww + 9 * cb.cin, HP_ERR_STATE, "chain: depthwise bias of block %d does not follow its kernel", first + b);
  }
  if (tail_blk >= 0) {
    const BlockWeights& w = bb.blk[tail_blk];
    HP_REQUIRE(w.bhi && w.blo && tail_out, HP_ERR_STATE, "chain: tail block %d needs split weights and an output buffer", tail_blk);
    ChainBlk& cb = p.tblk;
    cb.bhi = f16 ? w.hhi : w.bhi; cb.blo = f16 ? w.hlo : w.blo; cb.dww = w.dww; cb.pwb = w.pwb;
    cb.unscale = f16 ? w.h_unscale : 1.f;
    cb.cin = chan_pad(kBlazeBlocks[tail_blk].cin);
    cb.cout = chan_pad(kBlazeBlocks[tail_blk].cout);
    cb.ks = cb.cin / 8;
    cb.n16 = (cb.cout + 15) / 16 * 16;
    cb.w_off = w_off;
    w_off += 10 * cb.cin + cb.cout;
    HP_REQUIRE(w.dwb == w.dww + 9 * cb.cin, HP_ERR_STATE, "chain: depthwise bias of block %d does not follow its kernel", tail_blk);
    p.tail = 1;
    p.tail_out = tail_out;
    p.Ho = ceil_div(H, 2); p.Wo = ceil_div(W, 2);
    p.tail_lpi = p.Ho * p.Wo;
    p.tail_lanes = cfg.NI * p.tail_lpi;
    int o_, pb;
    same_pad(H, 3, 2, &o_, &pb); p.pad_t = pb;
    same_pad(W, 3, 2, &o_, &pb); p.pad_l = pb;
    HP_REQUIRE(p.tail_lanes <= 128 && p.pad_t <= 1 && p.pad_l <= 1, HP_ERR_INVALID, "chain: tail block on %dx%d does not fit one M-tile", H, W);
  }
  {
    const int strips = ceil_div(H, cfg.TR);
    HP_REQUIRE(cfg.NI * strips * W == cfg.lanes && cfg.NI < 256 && strips < 256 && W < 256, HP_ERR_INVALID, "chain: lane count %d does not match %d x %d x %d", cfg.lanes, cfg.NI, strips, W);
    chain_tables(H, W, cfg.TR, cfg.NI, tail_blk >= 0, p.lane_tab, p.tail_tab);
  }
  HP_REQUIRE(w_off <= cfg.w_floats, HP_ERR_STATE, "chain: weight area too small (%d > %d floats)", w_off, cfg.w_floats);
  p.H = H; p.W = W; p.NI = cfg.NI; p.B = B;
  p.n_tiles = ceil_div(B, cfg.NI);
  p.lanes = cfg.lanes; p.lpi = cfg.lpi;
  p.row_pitch = W * cfg.PS;
  p.load_bytes = (uint32_t)((size_t)cfg.NI * (H + 1) * W * cfg.PS * sizeof(float));
  p.off_w = cfg.off_w; p.w_floats = cfg.w_floats; p.off_ring = cfg.off_ring; p.off_zero = cfg.off_zero; p.zero_floats = cfg.zero_floats;
  p.off_tile = cfg.off_tile;
  p.trace = h->tc_trace; p.trace_steps = h->tc_trace_tiles;
  {
    static int dbg_env = -1, exp_env = -1;             // read once: bring-up / timing-experiment switches
    if (dbg_env < 0) {
      const char* e = getenv("HP_CHAIN_DBG");
      dbg_env = e ? atoi(e) : 0;
      e = getenv("HP_CHAIN_EXP");
      exp_env = e ? atoi(e) : 0;
    }
    p.dbg = dbg_env;
    p.exp_ = exp_env;
  }
  HP_REQUIRE(cfg.smem <= 227 * 1024 && cfg.lanes >= 1 && cfg.lanes <= 128 && (cfg.off_tile * 4) % 128 == 0 && H + 1 <= 256 && W + 1 <= 256 &&
                 cfg.NI <= 256 && p.load_bytes < (1u << 20),
             HP_ERR_INVALID, "chain %d..%d: bad geometry TR %d NI %d PS %d on %dx%d", first, first + nblk - 1, cfg.TR, cfg.NI, cfg.PS, H, W);
  const int cin0 = p.blk[0].cin, coutL = p.blk[nblk - 1].cout;
  CUtensorMap tin, tout;
  {
    const cuuint64_t din[4] = {(cuuint64_t)cin0, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t sin_[3] = {(cuuint64_t)cin0 * 4, (cuuint64_t)W * cin0 * 4, (cuuint64_t)H * W * cin0 * 4};
    const cuuint64_t dout[4] = {(cuuint64_t)coutL, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t sout[3] = {(cuuint64_t)coutL * 4, (cuuint64_t)W * coutL * 4, (cuuint64_t)H * W * coutL * 4};
    const cuuint32_t box[4] = {(cuuint32_t)cfg.PS, (cuuint32_t)W, (cuuint32_t)(H + 1), (cuuint32_t)cfg.NI};
    HP_TRY(tc_make_map4(&tin, in, din, sin_, box));
    HP_TRY(tc_make_map4(&tout, out, dout, sout, box));
  }
#define CHAIN_CASE(TR_, PS_, NSETS_, NISS_, F16_)                                                             \
  if (cfg.TR == TR_ && cfg.PS == PS_ && cfg.nsets == NSETS_ && cfg.niss == NISS_ && (int)f16 == F16_)           \
    return launch_chain<TR_, PS_, NSETS_, NISS_, F16_>(h, tin, tout, p, cfg.smem, st);
  CHAIN_CASE(3, 92, 4, 1, 1) CHAIN_CASE(2, 92, 4, 1, 1) CHAIN_CASE(3, 100, 4, 1, 1) CHAIN_CASE(2, 100, 4, 1, 1)
  CHAIN_CASE(3, 92, 4, 2, 1) CHAIN_CASE(2, 92, 4, 2, 1) CHAIN_CASE(3, 100, 4, 2, 1) CHAIN_CASE(2, 100, 4, 2, 1)
  CHAIN_CASE(3, 92, 4, 2, 0) CHAIN_CASE(2, 92, 4, 2, 0) CHAIN_CASE(3, 100, 4, 2, 0) CHAIN_CASE(2, 100, 4, 2, 0)
  CHAIN_CASE(3, 92, 4, 1, 0) CHAIN_CASE(2, 92, 4, 1, 0) CHAIN_CASE(3, 100, 4, 1, 0) CHAIN_CASE(2, 100, 4, 1, 0)
  CHAIN_CASE(3, 92, 4, 3, 0) CHAIN_CASE(3, 100, 4, 3, 0)
#undef CHAIN_CASE
  hp_set_error("chain: no kernel for TR %d PS %d nsets %d issuers %d", cfg.TR, cfg.PS, cfg.nsets, cfg.niss);
  return HP_ERR_UNSUPPORTED;
}
