// Dense / 1x1-conv layer y = act(x W + b) of the detector and regressor heads as a 3xTF32 GEMM on the tcgen05 tensor cores
// (forward inference only; training, transposed weights and accumulation stay on dense_kernel in heads.cu).
//
// Reference semantics: Conv2D(1x1) / Dense layers of the unified graph's detector heads and of the
// regressor heads (Model-88/attention_model.py:16-169, Model-88/train_88.py:66-253, Model-96/train_96.py:65-110).
// north_star (c): "dense and attention contractions on tcgen05 tensor cores only where the GEMM shapes justify it" -- these
// have M = B * H * W rows (147k-590k at batch 4096) against K, N <= 128; on the CUDA cores they were 1.0 ms of a 5.7 ms step.
//
// Same warp-specialised scheme and precision split as blaze_block_deep_kernel (blocks_tc.cu), with one accumulator row
// per lane (lane <-> token row):
//   loader   (1 thread): TMA box {KPAD, 128 rows} of x into a ring of buffers; KPAD > K pads the row stride to an odd
//                        number of 16-byte chunks (conflict-free row access), columns >= K are zero-filled (K padding)
//   gather sets        : lane copies 16 floats (2 k-steps) of its row, splits them into TF32 hi / lo, tcgen05.st
//   issuer   (1 thread): 3 tcgen05.mma per k-step (A from TMEM, W_hi / W_lo from smem; W is split in the kernel prologue)
//   epilogue sets      : tcgen05.ld D + bias -> staging tile -> activation + coalesced scatter to the output segments
// TAIL variant: a following narrow layer z = act2(y W2 + b2) with at most 4 outputs (the 64 -> 3 / 128 -> 3 regressor
// outputs) is applied to the accumulator row in registers; y is never written (at batch 4096 the 64-channel
// intermediate of the Model-88 head is 268 MB written and read back).  The epilogue sets then alternate tiles.
#include "tc_common.cuh"

namespace {

constexpr int DT_MAXB = 4, DT_MAXSTG = 4, DT_BAR_FLOATS = 96, DT_MAXKU = 4, DT_MAXD = 4;
constexpr int DT_ROWS = 128;

struct DenseTcParams {
  const float *W, *b;
  int M, K, N, ldw, act;
  int K8, KS, N16, KPAD, OS;          // OS: staging row stride (odd number of 16-byte chunks)
  int n_tiles, nstg, nbuf;
  int ku, upt;                        // k-steps per gather unit (ring stage = 16 ku TMEM columns), units per tile
  int nd;                             // accumulator buffers in TMEM
  int niss;                           // issuing threads = A rings (nstg stages each)
  uint32_t load_bytes;
  int off_b, off_bias, off_rowoff, off_stage, off_in, in_floats;
  int n_outs;
  DenseOut outs[2];
  unsigned magic[2], magic4[2];
  int vec4[2];                        // the segment can be written as float4 (width, offsets and strides multiples of 4)
  // TAIL variant
  const float *W2, *b2;               // [N][ldw2], [n2]
  int n2, ldw2, act2, off_tail;       // off_tail: W2 rows padded to float4, [N16][4]
  DenseOut out2;
  long long* trace;                   // optional clock64 stamps of CTA 0, 12 slots per tile (tools/dense_trace.py)
  int trace_tiles;
};

// compile-time activation for the unrolled epilogue loops (a switch per element became a jump table per element)
template <int ACT>
__device__ __forceinline__ float dt_act_c(float v) { return hp_act_c<ACT>(v); }
// hidden channels [c0, c0 + 4 nq) of one row: y = act(D + bias), acc += y W2 (W2 rows as float4 in shared memory)
template <int ACT>
__device__ __forceinline__ void dt_tail_group(const uint32_t (&v)[32], int nq, const float* s_bias_c, const float* s_tail_c, float4& acc) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < nq) {
      const float4 bb = ld4(s_bias_c + j * 4);
      const float y[4] = {dt_act_c<ACT>(__uint_as_float(v[j * 4 + 0]) + bb.x), dt_act_c<ACT>(__uint_as_float(v[j * 4 + 1]) + bb.y),
                          dt_act_c<ACT>(__uint_as_float(v[j * 4 + 2]) + bb.z), dt_act_c<ACT>(__uint_as_float(v[j * 4 + 3]) + bb.w)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 w = ld4(s_tail_c + (j * 4 + e) * 4);
        acc.x = fmaf(y[e], w.x, acc.x); acc.y = fmaf(y[e], w.y, acc.y); acc.z = fmaf(y[e], w.z, acc.z); acc.w = fmaf(y[e], w.w, acc.w);
      }
    }
  }
}
__device__ __forceinline__ float dt_act(int act, float v) { return hp_act_rt(act, v); }

template <int ACT>
__device__ __forceinline__ float4 dt_act4_c(float4 v) { return make_float4(hp_act_c<ACT>(v.x), hp_act_c<ACT>(v.y), hp_act_c<ACT>(v.z), hp_act_c<ACT>(v.w)); }
__device__ __forceinline__ float4 dt_act4(int act, float4 v) {
  switch (act) {
    case HP_ACT_RELU: return dt_act4_c<HP_ACT_RELU>(v);
    case HP_ACT_TANH: return dt_act4_c<HP_ACT_TANH>(v);
    case HP_ACT_SIGMOID: return dt_act4_c<HP_ACT_SIGMOID>(v);
    case HP_ACT_SOFTSIGN: return dt_act4_c<HP_ACT_SOFTSIGN>(v);
    case HP_ACT_ELU: return dt_act4_c<HP_ACT_ELU>(v);
    case HP_ACT_SELU: return dt_act4_c<HP_ACT_SELU>(v);
    case HP_ACT_SOFTPLUS: return dt_act4_c<HP_ACT_SOFTPLUS>(v);
    case HP_ACT_SWISH: return dt_act4_c<HP_ACT_SWISH>(v);
    case HP_ACT_LEAKY_RELU: return dt_act4_c<HP_ACT_LEAKY_RELU>(v);
    default: return v;
  }
}

template <int NSETS, int NESETS, bool TAIL>
__global__ void __launch_bounds__(128 * NSETS + 128 * NESETS + 96, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap tm_in, DenseTcParams p) {
  extern __shared__ __align__(1024) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_full = bars;                      // [nbuf]
  uint64_t* bar_infree = bars + DT_MAXB;          // [nbuf]
  uint64_t* bar_afull = bars + 2 * DT_MAXB;       // [niss][DT_MAXSTG]: one A ring per issuing thread
  uint64_t* bar_aempty = bar_afull + 2 * DT_MAXSTG;
  uint64_t* bar_dfull = bar_aempty + 2 * DT_MAXSTG;   // [nd]
  uint64_t* bar_dempty = bar_dfull + DT_MAXD;     // [nd]
  static_assert((2 * DT_MAXB + 4 * DT_MAXSTG + 2 * DT_MAXD) * 8 + 4 <= DT_BAR_FLOATS * 4, "barrier block");
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem) + (DT_BAR_FLOATS - 1);
  float* s_bhi = smem + p.off_b;
  float* s_blo = s_bhi + p.K8 * p.N16;
  float* s_bias = smem + p.off_bias;
  long long* rowoff = reinterpret_cast<long long*>(smem + p.off_rowoff);   // [2][128]
  float* stage = smem + p.off_stage;                                        // [128][OS]
  float* in_bufs = smem + p.off_in;

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane_id = tid & 31;
  // slots: 0 load issued, 1 gather sees full, 2 gather set 0 done, 3 epilogue sees d_full, 4 epilogue done with D, 5 epilogue
  // written out, 7 MMAs issued, 8 last gather set done, 9 / 10 issuer sees last / first a_full, 11 issuer has D (tile 0: entry)
  if (tid == 0 && p.trace != nullptr && blockIdx.x == 0 && p.trace_tiles > 0) p.trace[11] = clock64();
  auto stamp = [&](int i, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && i < p.trace_tiles) p.trace[i * 12 + slot] = clock64();
  };
  constexpr int W_EPI = 4 * NSETS, W_ISSUE = W_EPI + 4 * NESETS, W_LOAD = W_ISSUE + 1;
  const int NSTG = p.nstg, NBUF = p.nbuf, KS = p.KS, N16 = p.N16;

  // weights: split into TF32 hi / lo in the UMMA layout [K8/4][N16][4] (k and n padding are zeros)
  for (int i = tid; i < p.K8 * N16; i += nthr) {
    const int k = i / N16, n = i - k * N16;
    const float w = (k < p.K && n < p.N) ? p.W[(long long)k * p.ldw + n] : 0.f;
    const uint32_t hi = tf32_hi(w);
    const int idx = ((k >> 2) * N16 + n) * 4 + (k & 3);
    s_bhi[idx] = __uint_as_float(hi);
    s_blo[idx] = w - __uint_as_float(hi);
  }
  for (int i = tid; i < N16; i += nthr) s_bias[i] = (p.b && i < p.N) ? p.b[i] : 0.f;
  float* s_tail = smem + p.off_tail;                                        // [N16][4]: W2 row of every hidden channel
  if (TAIL)
    for (int i = tid; i < N16 * 4; i += nthr) {
      const int c = i >> 2, j = i & 3;
      s_tail[i] = (c < p.N && j < p.n2) ? p.W2[(long long)c * p.ldw2 + j] : 0.f;
    }
  fence_async_smem();
  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&bar_full[b], 1);
      mbar_init(&bar_infree[b], 128 * NSETS);
    }
    for (int r = 0; r < p.niss; ++r)
      for (int s = 0; s < NSTG; ++s) {
        mbar_init(&bar_afull[r * DT_MAXSTG + s], 128);
        mbar_init(&bar_aempty[r * DT_MAXSTG + s], 1);
      }
    for (int d = 0; d < p.nd; ++d) {
      mbar_init(&bar_dfull[d], 1);
      mbar_init(&bar_dempty[d], TAIL ? 128 : 128 * NESETS);
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
  const int ND = p.nd;                            // accumulator buffers: 2, or 4 for the fused narrow layer (its epilogue holds D longest)
  const uint32_t colA0 = ND * N16;                // D[0 .. nd) (N16 columns each), then the A ring (16 ku columns per stage)
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp < W_ISSUE) {
    const int wq = warp & 3;
    const int lane = wq * 32 + lane_id;           // token row of the tile
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
    if (warp < W_EPI) {
      // =============================================================== gather sets: x tile -> TF32 hi / lo -> TMEM A ring
      // work unit = KU k-steps (8 KU floats of the row, one ring stage of 16 KU columns); global round-robin over the sets as
      // in blaze_block_deep_kernel.  The per-unit protocol (a_empty wait, tcgen05.wait::st, fence, arrive) costs ~1K clk
      // whatever the unit holds (measured with one k-step per unit: 1.1-1.3K clk per unit, the gather was the bottleneck of
      // every layer), so a unit carries up to 4 k-steps.
      const int set = warp >> 2;
      const int KU = p.ku, UPT = p.upt;
      const uint32_t n_units = (uint32_t)my_tiles * UPT;
      int cur_i = -1, cur_b = 0;
      const float* row = in_bufs;
#pragma unroll 1
      for (uint32_t g = set; g < n_units; g += NSETS) {
        const int i = (int)(g / UPT);
        const int ks0 = (int)(g - (uint32_t)i * UPT) * KU;
        const int nk = (KS - ks0 < KU) ? KS - ks0 : KU;
        // two issuing threads (p.niss == 2): tiles alternate between two A rings with their own barriers; q = index of the
        // unit within its ring
        const int ring = (p.niss == 2) ? (i & 1) : 0;
        const uint32_t q = (p.niss == 2) ? (uint32_t)(i >> 1) * UPT + (uint32_t)(ks0 / KU) : g;
        const uint32_t s = q % NSTG;
        uint64_t* a_full = &bar_afull[ring * DT_MAXSTG + s];
        if (i != cur_i) {
          cur_i = i;
          cur_b = i % NBUF;
          row = in_bufs + cur_b * p.in_floats + lane * p.KPAD;
          mbar_wait(&bar_full[cur_b], (i / NBUF) & 1);
          if (tid == 0) stamp(i, 1);
        }
        if (q >= (uint32_t)NSTG) {
          mbar_wait(&bar_aempty[ring * DT_MAXSTG + s], ((q / NSTG) - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int kk = 0; kk < DT_MAXKU; ++kk) {
          if (kk < nk) {
            const float* q = row + 8 * (ks0 + kk);
            const float4 q0 = ld4(q), q1 = ld4(q + 4);
            const float f[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint32_t v[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              v[e] = tf32_hi(f[e]);
              v[8 + e] = __float_as_uint(f[e] - __uint_as_float(v[e]));
            }
            tmem_st16(tlane + colA0 + (ring * NSTG + s) * (16 * KU) + kk * 16, v);
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(a_full);
        if (g + NSETS >= n_units || (int)((g + NSETS) / UPT) != i) {
          mbar_arrive(&bar_infree[cur_b]);   // this thread reads nothing more from the tile
          if (tid == 0) stamp(i, 2);
          if (tid == (NSETS - 1) * 128) stamp(i, 8);
        }
      }
    } else if (TAIL) {
      // =============================================================== epilogue sets, fused narrow layer: tile i -> set i % NESETS
      const int eset = (warp - W_EPI) >> 2;
      float b2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b2[j] = (p.b2 && j < p.n2) ? p.b2[j] : 0.f;
      for (int i = eset; i < my_tiles; i += NESETS) {
        const int d = i % ND;
        const long long m = ((long long)blockIdx.x + (long long)i * gridDim.x) * DT_ROWS + lane;
        mbar_wait(&bar_dfull[d], (i / ND) & 1);
        tc_fence_after();
        if ((tid & 127) == 0) stamp(i, 3);
        float4 acc = make_float4(b2[0], b2[1], b2[2], b2[3]);
        for (int g = 0; g * 32 < N16; ++g) {
          uint32_t v[32];
          if (g * 32 + 32 <= N16) {
            tmem_ld32(tlane + d * N16 + g * 32, v);
          } else {
            uint32_t hlf[16];
            tmem_ld16(tlane + d * N16 + g * 32, hlf);
#pragma unroll
            for (int e = 0; e < 16; ++e) { v[e] = hlf[e]; v[16 + e] = 0u; }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int nq = (N16 - g * 32 < 32) ? (N16 - g * 32) / 4 : 8;   // padded hidden channels (>= N) have zero W2 rows
          const float* sb = s_bias + g * 32;
          const float* stl = s_tail + g * 128;
          switch (p.act) {
            case HP_ACT_RELU: dt_tail_group<HP_ACT_RELU>(v, nq, sb, stl, acc); break;
            case HP_ACT_TANH: dt_tail_group<HP_ACT_TANH>(v, nq, sb, stl, acc); break;
            case HP_ACT_SIGMOID: dt_tail_group<HP_ACT_SIGMOID>(v, nq, sb, stl, acc); break;
            case HP_ACT_SOFTSIGN: dt_tail_group<HP_ACT_SOFTSIGN>(v, nq, sb, stl, acc); break;
            case HP_ACT_ELU: dt_tail_group<HP_ACT_ELU>(v, nq, sb, stl, acc); break;
            case HP_ACT_SELU: dt_tail_group<HP_ACT_SELU>(v, nq, sb, stl, acc); break;
            case HP_ACT_SOFTPLUS: dt_tail_group<HP_ACT_SOFTPLUS>(v, nq, sb, stl, acc); break;
            case HP_ACT_SWISH: dt_tail_group<HP_ACT_SWISH>(v, nq, sb, stl, acc); break;
            case HP_ACT_LEAKY_RELU: dt_tail_group<HP_ACT_LEAKY_RELU>(v, nq, sb, stl, acc); break;
            default: dt_tail_group<HP_ACT_LINEAR>(v, nq, sb, stl, acc); break;
          }
        }
        tc_fence_before();
        mbar_arrive(&bar_dempty[d]);
        if ((tid & 127) == 0) stamp(i, 4);
        if (m < p.M) {
          const DenseOut& dd = p.out2;
          const long long img = m / dd.rows_per_img;
          float* dst = dd.ptr + img * dd.img_stride + (m - img * dd.rows_per_img) * (long long)dd.row_stride;
          const float z[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < p.n2) dst[j] = dt_act(p.act2, z[j]);
        }
        if ((tid & 127) == 0) stamp(i, 5);
      }
    } else {
      // =============================================================== epilogue sets
      const int eset = (warp - W_EPI) >> 2;
      const int etid = tid - W_EPI * 32;           // 0 .. 128 * NESETS - 1
      const bool linear_act = p.act == HP_ACT_LINEAR;
      int tile = blockIdx.x;
      for (int i = 0; i < my_tiles; ++i, tile += gridDim.x) {
        const int d = i % ND;
        const long long m0 = (long long)tile * DT_ROWS;
        const int rows = (int)((p.M - m0 < DT_ROWS) ? (p.M - m0) : DT_ROWS);
        // destination offset of every tile row per output segment (one 64-bit division per row, none per element)
        for (int j = etid; j < p.n_outs * DT_ROWS; j += 128 * NESETS) {
          const int o = j / DT_ROWS, r = j - o * DT_ROWS;
          const DenseOut& dd = p.outs[o];
          const unsigned m = (unsigned)(m0 + r);                 // M is an int: 32-bit division
          const unsigned img = m / (unsigned)dd.rows_per_img;
          rowoff[j] = (long long)img * dd.img_stride + (long long)(m - img * (unsigned)dd.rows_per_img) * dd.row_stride;
        }
        mbar_wait(&bar_dfull[d], (i / ND) & 1);
        tc_fence_after();
        if (etid == 0) stamp(i, 3);
        // D row + bias -> staging tile; the 32-column groups are dealt to the epilogue sets
        for (int g = eset; g * 32 < N16; g += NESETS) {
          uint32_t v[32];
          if (g * 32 + 32 <= N16) {
            tmem_ld32(tlane + d * N16 + g * 32, v);
          } else {
            uint32_t hlf[16];
            tmem_ld16(tlane + d * N16 + g * 32, hlf);
#pragma unroll
            for (int e = 0; e < 16; ++e) { v[e] = hlf[e]; v[16 + e] = 0u; }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = g * 32 + j * 4;
            if (c < N16) {
              const float4 bb = ld4(s_bias + c);
              st4(stage + lane * p.OS + c, make_float4(__uint_as_float(v[j * 4 + 0]) + bb.x, __uint_as_float(v[j * 4 + 1]) + bb.y,
                                                       __uint_as_float(v[j * 4 + 2]) + bb.z, __uint_as_float(v[j * 4 + 3]) + bb.w));
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&bar_dempty[d]);
        if (etid == 0) stamp(i, 4);
        named_bar_sync(1, 128 * NESETS);
        // activation + coalesced write-out of every output segment
        for (int o = 0; o < p.n_outs; ++o) {
          const DenseOut& dd = p.outs[o];
          const int wd = dd.col_end - dd.col_begin;
          const int total = rows * wd;
          if (p.vec4[o]) {                      // 16-byte aligned segment: one float4 per thread and step
            const int wq = wd >> 2, total4 = rows * wq;
            for (int j = etid; j < total4; j += 128 * NESETS) {
              const int r = (int)(((unsigned long long)j * p.magic4[o]) >> 32);
              const int c = (j - r * wq) * 4;
              {
                const float4 sv = ld4(stage + r * p.OS + dd.col_begin + c);
                st4(dd.ptr + rowoff[o * DT_ROWS + r] + c, linear_act ? sv : dt_act4(p.act, sv));   // most layers (all detector heads) are linear
              }
            }
            continue;
          }
          for (int j = etid; j < total; j += 128 * NESETS) {
            const int r = (int)(((unsigned long long)j * p.magic[o]) >> 32);
            const int c = j - r * wd;
            {
              const float sv = stage[r * p.OS + dd.col_begin + c];
              dd.ptr[rowoff[o * DT_ROWS + r] + c] = linear_act ? sv : dt_act(p.act, sv);
            }
          }
        }
        if (etid == 0) stamp(i, 5);
        named_bar_sync(1, 128 * NESETS);   // the staging tile and rowoff are reused by the next tile
      }
    }
  } else if (lane_id == 0) {
    if (warp == W_ISSUE || (warp == W_ISSUE + 2 && p.niss == 2)) {
      // =============================================================== MMA issuer(s): issuer r owns tiles i = r, r + niss, ... and A ring r
      const int ring = (warp == W_ISSUE) ? 0 : 1;
      const uint32_t idesc = tc_idesc_tf32(N16);
      const uint64_t desc_fixed = tc_bdesc_fixed(N16);
      const uint32_t bhi_addr = smem_u32(s_bhi), blo_addr = smem_u32(s_blo);
      uint32_t use = 0;                                              // index of the unit within this issuer's ring
      // every instruction of this loop is on the critical path (a lone warp retires a dependent instruction every ~5 clk and
      // shares its sub-partition with four worker warps: 166 clk per MMA measured, tools/dense_trace.py): descriptors and TMEM
      // addresses advance by additions, the waits are bare try_wait loops, the trace test is hoisted
      const bool traced = p.trace != nullptr && blockIdx.x == 0;
      const uint64_t dhi0 = desc_fixed | (uint64_t)((bhi_addr >> 4) & 0x3FFF), dlo0 = desc_fixed | (uint64_t)((blo_addr >> 4) & 0x3FFF);
      const uint32_t kstep_desc = (2u * (uint32_t)N16 * 16u) >> 4;  // one k-step of W in descriptor units; no carry out of the address field
      const uint32_t stage_cols = 16u * (uint32_t)p.ku;
      const uint32_t a_ring = tmem_base + colA0 + (uint32_t)(ring * NSTG) * stage_cols;
      const int KU = p.ku, UPT = p.upt;
      for (int i = ring; i < my_tiles; i += p.niss) {
        const int d = i % ND;
        if (i >= ND) {
          mbar_wait_lean(&bar_dempty[d], ((i / ND) - 1) & 1);
          tc_fence_after();
        }
        if (traced && i > 0) stamp(i, 11);
        const uint32_t dc = tmem_base + d * N16;
        uint64_t dhi = dhi0, dlo = dlo0;
        int ks = 0;
#pragma unroll 1
        for (int u = 0; u < UPT; ++u, ++use) {
          const uint32_t s = use % NSTG;
          mbar_wait_lean(&bar_afull[ring * DT_MAXSTG + s], (use / NSTG) & 1);
          tc_fence_after();
          if (traced) {
            if (u == 0) stamp(i, 10);
            if (u == UPT - 1) stamp(i, 9);
          }
          uint32_t a = a_ring + s * stage_cols;
#pragma unroll 1
          for (int kk = 0; kk < KU && ks < KS; ++kk, ++ks, a += 16, dhi += kstep_desc, dlo += kstep_desc) {
            mma_tf32_ts(dc, a, dhi, idesc, ks > 0 ? 1u : 0u);
            mma_tf32_ts(dc, a, dlo, idesc, 1u);
            mma_tf32_ts(dc, a + 8, dhi, idesc, 1u);
          }
          tc_commit(&bar_aempty[ring * DT_MAXSTG + s]);
        }
        tc_commit(&bar_dfull[d]);
        if (traced) stamp(i, 7);
      }
    } else if (warp == W_LOAD) {
      // =============================================================== TMA loader
      int b = 0, i = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        if (i >= NBUF) mbar_wait(&bar_infree[b], ((i / NBUF) - 1) & 1);
        mbar_expect_tx(&bar_full[b], p.load_bytes);
        tma_load_4d(in_bufs + b * p.in_floats, &tm_in, &bar_full[b], 0, tile * DT_ROWS, 0, 0);
        stamp(i, 0);
        if (++b == NBUF) b = 0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ISSUE) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

}  // namespace

// shared-memory layout (floats); returns the bytes needed with `nbuf` input buffers
static size_t dense_tc_layout(int K, int N, int nbuf, bool tail, DenseTcParams* p) {
  p->K8 = round_up(K, 8); p->KS = p->K8 / 8; p->N16 = round_up(N, 16);
  p->KPAD = ((p->K8 / 4) | 1) * 4;
  p->OS = ((p->N16 / 4) | 1) * 4;
  int off = DT_BAR_FLOATS;
  p->off_b = off;
  off += 2 * p->K8 * p->N16;
  p->off_bias = off;
  off = tc_align_up(off + p->N16, 4);
  p->off_rowoff = off;
  off += 4 * DT_ROWS;                              // 2 x 128 long long
  p->off_stage = off;
  p->off_tail = off;
  off = tc_align_up(off + (tail ? 4 * p->N16 : DT_ROWS * p->OS), 256);   // the fused narrow layer needs no staging tile
  p->off_in = off;
  p->in_floats = tc_align_up(DT_ROWS * p->KPAD, 256);
  return (size_t)(off + nbuf * p->in_floats) * sizeof(float);
}

// true when the layer can run on the tensor-core kernel: forward, plain weights, 16-byte aligned rows, at least 3 k-steps
// (every gather set must own a k-step of every tile), two input buffers next to the split weights in shared memory
static bool dense_tc_ok(const float* x, int M, int K, int ldx, int N, bool tail) {
  // no lower bound on M: a layer takes the same kernel (same arithmetic, row by row) whatever the batch size, so results do
  // not depend on how a caller batches its crops
  if (M < 1 || K % 4 != 0 || K < 20 || K > 128 || N < 1 || N > 128 || ldx % 4 != 0 ||
      (((uintptr_t)x) & 15) != 0)
    return false;
  DenseTcParams p;
  return dense_tc_layout(K, N, 2, tail, &p) <= 227 * 1024 && 2 * p.N16 + DT_MAXSTG * 16 * DT_MAXKU <= 512;
}
bool hp_dense_tc_supported(const float* x, int M, int K, int ldx, int N, bool transpose_w, bool accumulate) {
  return !transpose_w && !accumulate && dense_tc_ok(x, M, K, ldx, N, false);
}
// y = act(x W + b) followed by z = act2(y W2 + b2) with n2 <= 4 outputs, y not stored
bool hp_dense_tc_tail_supported(const float* x, int M, int K, int ldx, int N, int n2) {
  return n2 >= 1 && n2 <= 4 && dense_tc_ok(x, M, K, ldx, N, true);
}

static int dense_tc_launch(hp_ctx* h, const float* x, int M, int K, int ldx, const float* W, int ldw, const float* b, int N, int act,
                           const DenseOut* outs, int n_outs, const DenseTail* tail, cudaStream_t st) {
  HP_REQUIRE(tail ? n_outs == 0 : (n_outs >= 1 && n_outs <= 2), HP_ERR_INVALID, "dense tc: 1 or 2 output segments");
  DenseTcParams p;
  p.W2 = nullptr; p.b2 = nullptr; p.n2 = 0; p.ldw2 = 0; p.act2 = 0; p.out2 = DenseOut{};
  if (tail) {
    HP_REQUIRE(tail->n2 >= 1 && tail->n2 <= 4 && tail->W2 && tail->out.ptr, HP_ERR_INVALID, "dense tc: fused layer needs 1..4 outputs");
    p.W2 = tail->W2; p.b2 = tail->b2; p.n2 = tail->n2; p.ldw2 = tail->ldw2; p.act2 = tail->act2; p.out2 = tail->out;
  }
  p.W = W; p.b = b; p.M = M; p.K = K; p.N = N; p.ldw = ldw; p.act = act;
  p.nbuf = DT_MAXB;
  while (p.nbuf > 2 && dense_tc_layout(K, N, p.nbuf, tail != nullptr, &p) > 200 * 1024) --p.nbuf;
  const size_t smem = dense_tc_layout(K, N, p.nbuf, tail != nullptr, &p);
  p.n_tiles = ceil_div(M, DT_ROWS);
  p.trace = h->tc_trace; p.trace_tiles = h->tc_trace_tiles;
  // every gather set must own a unit of every tile (it frees the input buffer after its last one): at least 3 units per tile
  p.ku = DT_MAXKU;
  while (p.ku > 1 && ceil_div(p.KS, p.ku) < 3) --p.ku;
  p.upt = ceil_div(p.KS, p.ku);
  p.nd = (tail && 4 * p.N16 + DT_MAXSTG * 16 * p.ku <= 512) ? 4 : 2;
  p.niss = 1;
  p.nstg = DT_MAXSTG;
  p.n_outs = n_outs;
  for (int i = 0; i < n_outs; ++i) {
    p.outs[i] = outs[i];
    const int wd = outs[i].col_end - outs[i].col_begin;
    HP_REQUIRE(wd >= 1 && outs[i].col_end <= N, HP_ERR_INVALID, "dense tc: bad output segment [%d, %d)", outs[i].col_begin, outs[i].col_end);
    p.magic[i] = (unsigned)((0x100000000ull + wd - 1) / wd);
    p.vec4[i] = ((wd | outs[i].col_begin | outs[i].row_stride) & 3) == 0 && (outs[i].img_stride & 3) == 0 && (((uintptr_t)outs[i].ptr) & 15) == 0;
    p.magic4[i] = p.vec4[i] ? (unsigned)((0x100000000ull + wd / 4 - 1) / (wd / 4)) : 0u;
  }
  p.load_bytes = (uint32_t)((size_t)DT_ROWS * p.KPAD * sizeof(float));
  HP_REQUIRE(smem <= 227 * 1024, HP_ERR_UNSUPPORTED, "dense tc: %zu bytes of shared memory needed for %dx%d", smem, K, N);
  CUtensorMap tin;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)M, 1, 1};
    const cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)ldx * 4 * (cuuint64_t)M, (cuuint64_t)ldx * 4 * (cuuint64_t)M};
    const cuuint32_t box[4] = {(cuuint32_t)p.KPAD, (cuuint32_t)DT_ROWS, 1, 1};
    HP_TRY(tc_make_map4(&tin, x, dims, strides, box));
  }
  long long grid = h->num_sms;
  if (grid > p.n_tiles) grid = p.n_tiles;
  // narrow layers are bound by the gather (3 sets); wide ones by the epilogue (2 sets)
  // warp sets: gather x epilogue.  Sweeps with tools/dense_trace.py (HP_DENSE_TC_SETS=<gather><epilogue>, e.g. "31", overrides the
  // defaults for such sweeps only; profiles/r01/dense_sets_sweep.log): the number of gather sets does not matter, 2 epilogue
  // sets are 20-25 % faster than 1, more than 2 change nothing
  int ns = (!tail && p.N16 <= 32) ? 3 : 2, ne = (!tail && p.N16 <= 32) ? 1 : 2;
  // measured (profiles/r01/dense_issuers_sweep.log): 2 issuers 110 -> 90 us on the 88 -> 34 detector head, 37.5 -> 32.5 us on 96 -> 64;
  // the fused narrow layer keeps one issuer and 4 accumulator buffers: alone it runs faster with two (profiles/r02/dense_nd_sweep.log:
  // 88 -> 64 -> 3 113 -> 89.5 us), inside a step slower (launch list of a bench step: 116.8 -> 126.1 us)
  int want_iss = tail ? 1 : 2, nd_env = 0;
  {
    static const char* env = getenv("HP_DENSE_TC_SETS");   // <gather sets><epilogue sets>[issuers[accumulators]]
    if (env && env[0] >= '1' && env[0] <= '3' && env[1] >= '1' && env[1] <= '2') {
      ns = env[0] - '0'; ne = env[1] - '0';
      if (env[2] == '1' || env[2] == '2') {
        want_iss = env[2] - '0';
        if (env[3] == '2' || env[3] == '4') nd_env = env[3] - '0';
      }
    }
  }
  // Two issuing threads: one thread issues a tcgen05.mma only every ~110 clk next to the worker warps and nothing else is
  // saturated (profiles/r01/ncu_heads_v6_summary.txt).  Each gets its own A ring (tiles alternate) with 2 accumulator buffers;
  // an mbarrier waiter may be one phase ahead at most, so consecutive units of a gather set within one ring must be at most
  // nstg ring slots apart: checked by walking the round-robin schedule.
  if (want_iss == 2) {
    // each issuer's tiles (i, i + 2, ...) use the accumulators i % nd: with nd = 2 an issuer owned ONE accumulator and its MMAs
    // waited for the epilogue of its previous tile (tools/dense_trace.py: 5.5 K clk of MMAs + 1.2 K of hand-off per issuer and
    // pair of tiles); 4 accumulators give every issuer two when they fit next to two rings of >= 2 stages
    for (int nd2 = 4; nd2 >= 2 && p.niss != 2; nd2 -= 2) {
      if (nd_env > 0 && nd2 != nd_env) continue;
      int stg = (512 - nd2 * p.N16) / (2 * 16 * p.ku);
      if (stg > DT_MAXSTG) stg = DT_MAXSTG;
      bool safe = stg >= 2;
      for (int x = 0; x < ns && safe; ++x) {
        long long last[2] = {-1, -1};
        for (long long g = x; g < 16ll * p.upt; g += ns) {
          const long long i = g / p.upt, u = g - i * p.upt, r = i & 1, q = (i >> 1) * p.upt + u;
          if (last[r] >= 0 && q - last[r] > stg) safe = false;
          last[r] = q;
        }
      }
      if (safe && ns <= stg * 2) { p.niss = 2; p.nstg = stg; p.nd = nd2; }
    }
  }
  HP_REQUIRE(p.nd * p.N16 + p.niss * p.nstg * 16 * p.ku <= 512 && p.upt >= 3 && p.KPAD <= 256 && ns <= p.nstg * (p.niss == 2 ? 2 : 1),
             HP_ERR_UNSUPPORTED, "dense tc: layer %dx%d too large", K, N);
#define DT_LAUNCH(NS_, NE_, TAIL_)                                                                            \
  if (ns == NS_ && ne == NE_ && (tail != nullptr) == TAIL_) {                                                  \
    auto kern = dense_tc_kernel<NS_, NE_, TAIL_>;                                                              \
    HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));              \
    kern<<<(unsigned)grid, 128 * NS_ + 128 * NE_ + 96, smem, st>>>(tin, p);                                    \
  }
  DT_LAUNCH(1, 1, false) DT_LAUNCH(1, 2, false) DT_LAUNCH(2, 1, false) DT_LAUNCH(2, 2, false) DT_LAUNCH(3, 1, false) DT_LAUNCH(3, 2, false)
  DT_LAUNCH(1, 2, true) DT_LAUNCH(2, 2, true) DT_LAUNCH(3, 2, true)
  else if (tail && ne != 2) { hp_set_error("dense tc: the fused narrow layer is built with 2 epilogue sets"); return HP_ERR_UNSUPPORTED; }
#undef DT_LAUNCH
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

int hp_launch_dense_tc(hp_ctx* h, const float* x, int M, int K, int ldx, const float* W, int ldw, const float* b, int N, int act,
                       const DenseOut* outs, int n_outs, cudaStream_t st) {
  return dense_tc_launch(h, x, M, K, ldx, W, ldw, b, N, act, outs, n_outs, nullptr, st);
}

int hp_launch_dense_tc_tail(hp_ctx* h, const float* x, int M, int K, int ldx, const float* W, int ldw, const float* b, int N, int act,
                            const DenseTail& tail, cudaStream_t st) {
  return dense_tc_launch(h, x, M, K, ldx, W, ldw, b, N, act, nullptr, 0, &tail, st);
}
