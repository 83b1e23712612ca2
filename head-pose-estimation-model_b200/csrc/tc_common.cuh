// Device-side PTX helpers shared by the tensor-core kernels (blocks_tc.cu, stem_tc.cu): mbarrier, TMA, tcgen05.
#pragma once
#include <vector>
#include <cuda.h>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------- PTX helpers
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
// Blocking wait with a watchdog.  try_wait carries a suspend-time hint, so a waiting warp sleeps in hardware until the
// phase completes instead of burning issue slots (ncu: un-hinted polling was ~25 % of all issued instructions); the clock
// is read every 256 wake-ups only.  The hint is kept at 256 ns: with 16 us, rare runs of some geometries took 1-20 ms
// instead of 0.2 ms (a completion that lands between the failed test and the sleep is only noticed at the time limit).
#ifndef HP_WAIT_HINT
#define HP_WAIT_HINT 0x100
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
#if HP_WAIT_HINT > 0
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(addr), "r"(parity), "n"(HP_WAIT_HINT) : "memory");
#else
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(addr), "r"(parity) : "memory");
#endif
  return done;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, n = 0;
  long long t0 = 0;
  while (true) {
    done = mbar_try_wait(addr, parity);
    if (done) break;
    if ((++n & 255u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();   // ~2 s: never hang the device on a lost signal
    }
  }
}
// Wait for the MMA-issuing threads: nothing but the try_wait loop.  An issuer is a lone thread whose dependent instructions
// retire every ~5 clk, so the ~25 instructions of mbar_wait (watchdog) on every k-step cost more than the MMAs they guard
// (measured in the chain kernel: 1.1-1.7K clk per k-step with nothing to wait for, tools/chain_check.py).
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "MBAR_LEAN_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra MBAR_LEAN_%=;\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
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
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// TF32 split a = hi + lo: hi keeps the top 19 bits (truncation, 1 LOP3), lo = a - hi is exact in fp32 and |lo| < 2^-10 |a|;
// the tensor core ignores the low 13 mantissa bits of lo, so the representation error is < 2^-20 |a| (cvt.rna.tf32
// would cost ~5 SASS instructions per value for one more bit).
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
// D[tmem] (+)= A[tmem] * B[smem descriptor], tf32 inputs, fp32 accumulate, M = 128
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_addr, uint32_t a_addr, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_addr), "r"(a_addr), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with fp16 inputs (two per 32-bit TMEM column), K = 16
__device__ __forceinline__ void mma_f16_ts(uint32_t d_addr, uint32_t a_addr, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_addr), "r"(a_addr), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
               "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(addr));
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// packed fp32 FMA (fma.rn.f32x2, sm_100): two channels per instruction, same rounding as two scalar FMAs; halves the FMA
// issue slots of the depthwise conv, which competes with LDS / tcgen05 traffic for the same schedulers
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  const float2 lo = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
  const float2 hi = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), make_float2(c.z, c.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 scale4(float4 a, float f) {
  const float2 lo = __fmul2_rn(make_float2(a.x, a.y), make_float2(f, f));
  const float2 hi = __fmul2_rn(make_float2(a.z, a.w), make_float2(f, f));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}


__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(addr));
}
// 8-column TMEM store, 128-bit shared-memory access helpers
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// instruction descriptor of tcgen05.mma kind::tf32: D fp32, A / B tf32, both K-major, M = 128, N = n16
__device__ __forceinline__ uint32_t tc_idesc_tf32(int n16) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// shared-memory descriptor (without the start address) of a K-major no-swizzle B slice [n16 rows][8 k] stored as
// [k / 4][n16][4]: core matrix = 8 rows x 16 B, LBO (next 4 k) = n16 * 16 B, SBO (next 8 rows) = 128 B, version 1
__device__ __forceinline__ uint64_t tc_bdesc_fixed(int n16) {
  return ((uint64_t)(((uint32_t)(n16 * 16) >> 4) & 0x3FFF) << 16) | ((uint64_t)((128u >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

typedef CUresult (*PFN_tcEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline int tc_get_encode(PFN_tcEncodeTiled* out) {
  static PFN_tcEncodeTiled fn_cached = nullptr;
  if (!fn_cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    HP_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, HP_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
    fn_cached = (PFN_tcEncodeTiled)fn;
  }
  *out = fn_cached;
  return HP_OK;
}
// fp32 tensor map of rank 4 (dims / box innermost first; strides in bytes for dims 1..3); out-of-range elements read as zero
inline int tc_make_map4(CUtensorMap* tm, const float* base, const cuuint64_t dims[4], const cuuint64_t strides[3], const cuuint32_t box[4]) {
  PFN_tcEncodeTiled enc;
  HP_TRY(tc_get_encode(&enc));
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HP_REQUIRE(r == CUDA_SUCCESS, HP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu box %u %u %u %u", (int)r,
             (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3], box[0], box[1],
             box[2], box[3]);
  return HP_OK;
}
inline int tc_align_up(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

// Deal pixel columns to lanes.  pix[i] = linear pixel index (in pixels of PS floats) of the first tap of entry i, code[i] = its
// packed coordinates.  With PS = an odd number of 16-byte chunks the bank group of a 128-bit access is pix mod 8, and the 8
// lanes of a quarter warp share the wavefronts of an LDS.128 / STS.128: every group of 8 consecutive lanes gets entries with
// different residues while the supply lasts (greedy: largest residue classes first), so that an access costs 4 wavefronts
// per warp instead of up to 8.  The groups are filled densely: lanes [0, n) are in use.
// second_bit != 0: every second lane of a residue class within its group of 8 gets that bit set (the stride-2 tail block only
// has even residues -- at least two lanes per class -- and lets those lanes walk the two 16-byte halves of a k-step, and the
// two columns of the max-pool window, in the opposite order: one chunk further, the other bank group).
inline void tc_lane_table(const std::vector<int>& pix, const std::vector<uint32_t>& code, uint32_t* tab, uint32_t second_bit = 0u) {
  std::vector<int> bucket[8];
  const int n = (int)pix.size();
  for (int i = n - 1; i >= 0; --i) bucket[((pix[i] % 8) + 8) % 8].push_back(i);   // pop_back() hands them out in natural order
  int lane = 0;
  while (lane < n) {
    int used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int want = n - lane < 8 ? n - lane : 8;
    for (int k = 0; k < want; ++k) {
      // the class with the lowest multiplicity in this group, then the fullest one
      int best = -1;
      for (int r = 0; r < 8; ++r) {
        if (bucket[r].empty()) continue;
        if (best < 0 || used[r] < used[best] || (used[r] == used[best] && bucket[r].size() > bucket[best].size())) best = r;
      }
      tab[lane++] = code[bucket[best].back()] | ((used[best] & 1) ? second_bit : 0u);
      bucket[best].pop_back();
      used[best]++;
    }
  }
  for (; lane < 128; ++lane) tab[lane] = 0u;
}
