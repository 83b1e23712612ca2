// Internal declarations shared by the translation units of libhpose.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/hpose.h"

void hp_set_error(const char* fmt, ...);

#define HP_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      hp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
      return HP_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define HP_REQUIRE(cond, code, ...)                                                           \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      hp_set_error(__VA_ARGS__);                                                              \
      return (code);                                                                          \
    }                                                                                         \
  } while (0)

#define HP_TRY(expr)                                                                          \
  do {                                                                                        \
    int rc_ = (expr);                                                                         \
    if (rc_ != HP_OK) return rc_;                                                             \
  } while (0)

// Growable device buffer owned by a handle.
extern long long g_devbuf_epoch;   // detect.cu: counts reallocations (captured CUDA graphs hold raw pointers)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t n) {
    if (n <= bytes) return HP_OK;
    ++g_devbuf_epoch;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    HP_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return HP_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  float* f() const { return (float*)p; }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// TensorFlow "SAME" padding: out = ceil(n/s), extra padding goes after.
static inline void same_pad(int n, int k, int s, int* out, int* before) {
  int o = ceil_div(n, s);
  int total = (o - 1) * s + k - n;
  if (total < 0) total = 0;
  *out = o;
  *before = total / 2;
}

// ---------------------------------------------------------------------------- backbone
struct BlockShape {
  int cin, cout, stride;
};
extern const BlockShape kBlazeBlocks[16];
static inline int chan_pad(int c) { return (c + 3) & ~3; }

struct BlockWeights {
  const float *dww, *dwb, *pww, *pwb;  // [9][CINP], [CINP], [CINP][COUTP], [COUTP] (zero padded)
  const float *bhi, *blo;              // pointwise weights split into TF32 hi / lo parts, [K8/4][N16][4] (blocks_tc.cu)
  const float *hhi, *hlo;              // the same weights times 2^h_shift split into fp16 hi / lo parts, [K16/8][N16][8 halves] (chain kernel)
  float h_unscale;                     // 2^-h_shift: applied to the accumulator in the epilogue
  const float* h_dw;                   // HOST copy of [9][CINP] dww + [CINP] dwb: passed as kernel-parameter constants
};

struct Backbone {
  bool loaded = false;
  DevBuf arena;
  std::vector<float> host_dw;     // depthwise weights + bias of the 16 blocks (BlockWeights::h_dw points into it)
  const float* stem_w = nullptr;  // [75][24]
  const float* stem_b = nullptr;
  const float *stem_bhi = nullptr, *stem_blo = nullptr;  // stem kernel split into TF32 hi / lo parts in GEMM layout (stem_tc.cu)
  BlockWeights blk[16];
  const float *det16_w = nullptr, *det16_b = nullptr;  // [88][36] = cls(2)|loc(32)|pad
  const float *det8_w = nullptr, *det8_b = nullptr;    // [96][104] = cls(6)|loc(96)|pad
  DevBuf act[2], dwtmp, feat16, feat8;
};

// ---------------------------------------------------------------------------- comm (NCCL via dlsym)
#define HP_P2P_MAX_RANKS 8
#define HP_P2P_SLICES 16          // CTAs of the fused all-reduce + optimizer kernel: each owns one slice of the flat gradient buffer
struct Comm {
  void* lib = nullptr;
  void* comm = nullptr;
  int rank = 0, nranks = 1;
  // peer-memory exchange (comm.cu): this rank's INBOX (cudaMalloc, exported through CUDA IPC) holds, per source rank and step
  // parity, a copy of that rank's flat gradient buffer plus one arrival flag per slice; peers[] are the inboxes of all ranks
  // mapped into this process (NVLink / NVSwitch peer access)
  void* inbox = nullptr;
  void* peers[HP_P2P_MAX_RANKS] = {};
  void** peers_dev = nullptr;     // device copy of peers[]
  unsigned int* p2p_seq = nullptr; // device counter of exchanges on this context: the tag of a step's slices and flags (the same on all ranks)
  int p2p_cap = 0;                // floats per slot
  bool p2p_ready = false;
};

struct hp_head;  // heads.cu

// preproc.cu: tap tables of one (input size -> output size) bicubic resize, on the device: [Hout + Wout][4] indices, then weights
struct ResizePlan {
  int hin = 0, win = 0, hout = 0, wout = 0;
  void* dev = nullptr;
};

// detect.cu: internals of hp_detect_frames (padded decode / NMS outputs before packing) and its captured graphs
struct DetectBufs {
  DevBuf x, cnt, offsets, anchor, boxes, kps, scores, poses;
  void release() { x.release(); cnt.release(); offsets.release(); anchor.release(); boxes.release(); kps.release(); scores.release(); poses.release(); }
};
struct DetectGraph {
  const void *head16 = nullptr, *head8 = nullptr, *frames = nullptr, *result = nullptr;
  int B = 0, Hin = 0, Win = 0, H = 0, W = 0, max_out = 0, cap = 0, impl = 0, chain_mode = 0;
  float logit_thr = 0.f, iou_thr = 0.f;
  long long epoch = 0;
  int launches = 0;
  cudaGraphExec_t exec = nullptr;
  bool same_key(const DetectGraph& o) const {
    return head16 == o.head16 && head8 == o.head8 && frames == o.frames && result == o.result && B == o.B && Hin == o.Hin && Win == o.Win &&
           H == o.H && W == o.W && max_out == o.max_out && cap == o.cap && impl == o.impl && chain_mode == o.chain_mode &&
           logit_thr == o.logit_thr && iou_thr == o.iou_thr;
  }
};

struct hp_ctx {
  int device = 0;
  int num_sms = 148;
  int impl = HP_IMPL_FAST;
  int64_t launches = 0;
  long long bb_generation = 0;     // successful hp_backbone_load_weights calls: lets a host model detect that another one replaced its weights
  Backbone bb;
  Comm comm;
  DevBuf pose16, pose8, cls, loc;  // unified-path internals
  DevBuf scratch;
  DevBuf status;                   // one device word of sticky flags (HP_STATUS_*), read and cleared by hp_backbone_status
  std::vector<ResizePlan> resize_plans;
  DetectBufs det;
  std::vector<DetectGraph> det_graphs;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int tile_override[16][5] = {};   // TH, TW, IMGS, nbuf, MT per block (0 = automatic)
  bool p2p_off = false;            // hp_debug_set_p2p: keep the NCCL all-reduce although the peer-memory exchange is set up (comparison runs)
  bool dense_tc = true;            // Dense / 1x1 layers may use the tensor-core kernel (cleared while a training step runs)
  int stem_tc_cfg[4] = {};         // tensor-core stem: band height, input buffers, output stages, gather sets (0 = automatic, [0] = -1: off)
  int tc_override[16][9] = {};     // tensor-core kernel: TR, NSTG, BH, npipe, nsets, nbuf per block (TR 0 = automatic, -1 = do not use)
  int chain_mode = 2;              // >= 1: blocks 6-10 / 12-15 run as fused chain kernels when the geometry allows; 2: block 11 rides on the first chain (0: one kernel per block)
  int chain_cfg[2] = {};           // chain kernel overrides: worker warp sets, MMA issuer threads (0 = default)
  int* tile_report = nullptr;      // optional int[16][8] filled by the forward pass
  long long* tc_trace = nullptr;   // optional device buffer for per-tile clock stamps of the deep tensor-core kernel
  int tc_trace_tiles = 0;
};

// generic per-token dense layer used by detector heads and regressor heads (dense.cu)
struct DenseOut {
  float* ptr;            // destination for columns [col_begin, col_end)
  int col_begin, col_end;
  int rows_per_img;      // rows are grouped by image: dst = ptr + img*img_stride + row_in_img*row_stride + (c-col_begin)
  long long img_stride;
  int row_stride;
};
// y = act(x[M,K](ld=ldx) @ W[K][ldw] + b); transpose_w: use W^T stored as [N][ldw] (for dX = dY W^T)
int hp_launch_dense(hp_ctx* h, const float* x, int M, int K, int ldx, const float* W, int ldw, const float* b,
                    int N, int act, bool transpose_w, const DenseOut* outs, int n_outs, bool accumulate,
                    cudaStream_t st);

int hp_backbone_run(hp_ctx* h, const float* x, int B, int H, int W, float* feat16, float* feat8, float* cls,
                    float* loc, int stop_after_blk, float* dbg_dst, size_t dbg_floats, float* per_layer_ms,
                    int prof_iters, cudaStream_t st);

int hp_comm_allreduce_sum(hp_ctx* h, float* buf, size_t n, cudaStream_t st);

// blocks_tma.cu: second-generation fused BlazeBlock kernel (TMA tile load/store)
struct Tile2Cfg {
  int TH, TW, IMGS, MT, nbuf, PG, TP, threads, IH, IW, tiles_y, tiles_x, n_tiles, n_runs, rpr;
  int in_tile_floats, off_w, off_tab, off_dw, off_in;
  size_t smem;
};
bool hp_tile2_fill(int B, int Hout, int Wout, int S, int CINP, int COUTP, int TH, int TW, int IMGS, int nbuf, int MT,
                   Tile2Cfg* tc);
bool hp_tile2_choose(int B, int Hout, int Wout, int S, int CINP, int COUTP, Tile2Cfg* best);
int hp_launch_block_tma(hp_ctx* h, int blk, const float* in, float* out, int B, int Hin, int Win, int Hout, int Wout,
                        int pad_t, int pad_l, const BlockWeights& w, const Tile2Cfg& tc, cudaStream_t st);

// blocks_tc.cu: third-generation fused BlazeBlock kernel (depthwise on CUDA cores -> TMEM, pointwise as 3xTF32 tcgen05 GEMM)
struct TcCfg {
  // nbuf > 0 selects the warp-specialised (deep) kernel: nbuf halo buffers, ni images per tile, npipe = epilogue warp sets
  int TR, NSTG, BH, IWB, npipe, nsets, nbuf, ni, unit, niss, place;   // niss: MMA issuer warps; place: issuers on SM sub-partition 3   // unit: depthwise work unit = 1 chunk (4 channels) or 2 (a k-step)
};
bool hp_tc_fits(int blk, int H, int W, const TcCfg& tc);
bool hp_tcd_geometry(int blk, int H, int W, int TR, int nsets, int esets, TcCfg* tc);
bool hp_tcs_geometry(int blk, int H, int W, int nsets, int esets, TcCfg* tc);
bool hp_tcs2_geometry(int blk, int Ho, int Wo, int nsets, int esets, TcCfg* tc, int force_R = 0);
int hp_launch_block_tc_s2(hp_ctx* h, int blk, const float* in, float* out, int B, int Hi, int Wi, int Ho, int Wo, int pad_t, int pad_l,
                          const BlockWeights& w, const TcCfg& tc, cudaStream_t st);
void hp_tc_split_weights(const float* pww, int cinp, int coutp, float* bhi, float* blo);
int hp_tc_weight_floats(int cinp, int coutp);
// fp16 split of the pointwise weights for the chain kernel (kind::f16 MMAs, K = 16 per instruction): w * 2^shift = hi + lo with
// the power of two chosen so that the largest weight lands in [2^12, 2^13) -- the lo parts then stay normal fp16 numbers;
// returns 2^-shift.  Storage: K16 * N16 halves per part (two per float slot).
int hp_tc_weight_floats_f16(int cinp, int coutp);
float hp_tc_split_weights_f16(const float* pww, int cinp, int coutp, float* hhi, float* hlo);
unsigned short hp_f32_to_f16_rn(float f);
float hp_f16_to_f32(unsigned short h);
bool hp_tc_choose(int blk, int H, int W, TcCfg* tc);
int hp_launch_block_tc(hp_ctx* h, int blk, const float* in, float* out, int B, int H, int W, const BlockWeights& w,
                       const TcCfg& tc, cudaStream_t st);

// blocks_chain.cu: cross-block fusion -- a chain of stride-1 BlazeBlocks on one map size as ONE persistent kernel (tile of whole
// images resident in shared memory across the blocks, pointwise weights streamed through a ring of k-step slices)
struct ChainCfg {
  int TR, NI, PS, lanes, lpi, nsets, niss;
  int ring_slots;                 // weight ring: 4 slices (split-fp16 kernels: two 16-channel slices per round) or 8 (3xTF32)
  int off_w, w_floats, off_ring, off_zero, zero_floats, off_tile, tile_floats;   // shared-memory layout (floats)
  size_t smem;
};
bool hp_chain_geometry(int first, int nblk, int chain_nblk, int H, int W, ChainCfg* cfg, int tail_blk = -1, bool f16 = true);
int hp_chain_status(unsigned int out[8]);
int hp_chain_describe(int first, int nblk, int H, int W, int tail, unsigned int* out264);
int hp_launch_chain(hp_ctx* h, int first, int nblk, const float* in, float* out, int B, int H, int W, const ChainCfg& cfg,
                    cudaStream_t st, int tail_blk = -1, float* tail_out = nullptr);

// stem_tc.cu: stem conv as an implicit 3xTF32 GEMM (tcgen05), warp-specialised
int hp_stem_tc_weight_floats();
void hp_stem_tc_split_weights(const float* w75x24, float* bhi, float* blo);
bool hp_stem_tc_supported(int H, int W);
int hp_launch_stem_tc(hp_ctx* h, const float* x, float* out, int B, int H, int W, const float* bhi, const float* blo, const float* bias,
                      const int* cfg, cudaStream_t st);

// dense_tc.cu: Dense / 1x1-conv layers of the heads as a 3xTF32 tcgen05 GEMM (forward inference)
bool hp_dense_tc_supported(const float* x, int M, int K, int ldx, int N, bool transpose_w, bool accumulate);
struct DenseTail {                 // narrow layer fused behind a dense layer: z = act2(y W2[N][ldw2] + b2), n2 <= 4
  const float *W2, *b2;
  int n2, ldw2, act2;
  DenseOut out;                    // destination of z (col_begin / col_end unused)
};
bool hp_dense_tc_tail_supported(const float* x, int M, int K, int ldx, int N, int n2);
int hp_launch_dense_tc_tail(hp_ctx* h, const float* x, int M, int K, int ldx, const float* W, int ldw, const float* b, int N, int act,
                            const DenseTail& tail, cudaStream_t st);
int hp_launch_dense_tc(hp_ctx* h, const float* x, int M, int K, int ldx, const float* W, int ldw, const float* b, int N, int act,
                       const DenseOut* outs, int n_outs, cudaStream_t st);

// ---------------------------------------------------------------------------- activations (enum hp_act), shared by heads.cu / dense_tc.cu
#ifdef __CUDACC__
#define HP_SELU_ALPHA 1.6732632423543772f
#define HP_SELU_SCALE 1.0507009873554805f
template <int ACT>
__device__ __forceinline__ float hp_act_c(float v) {
  if (ACT == HP_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == HP_ACT_TANH) return tanhf(v);
  if (ACT == HP_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  if (ACT == HP_ACT_SOFTSIGN) return v / (1.f + fabsf(v));
  if (ACT == HP_ACT_ELU) return v > 0.f ? v : expm1f(v);
  if (ACT == HP_ACT_SELU) return HP_SELU_SCALE * (v > 0.f ? v : HP_SELU_ALPHA * expm1f(v));
  if (ACT == HP_ACT_SOFTPLUS) return fmaxf(v, 0.f) + log1pf(expf(-fabsf(v)));
  if (ACT == HP_ACT_SWISH) return v / (1.f + expf(-v));
  if (ACT == HP_ACT_LEAKY_RELU) return v > 0.f ? v : 0.2f * v;
  return v;
}
__device__ __forceinline__ float hp_act_rt(int act, float v) {
  switch (act) {
    case HP_ACT_RELU: return hp_act_c<HP_ACT_RELU>(v);
    case HP_ACT_TANH: return hp_act_c<HP_ACT_TANH>(v);
    case HP_ACT_SIGMOID: return hp_act_c<HP_ACT_SIGMOID>(v);
    case HP_ACT_SOFTSIGN: return hp_act_c<HP_ACT_SOFTSIGN>(v);
    case HP_ACT_ELU: return hp_act_c<HP_ACT_ELU>(v);
    case HP_ACT_SELU: return hp_act_c<HP_ACT_SELU>(v);
    case HP_ACT_SOFTPLUS: return hp_act_c<HP_ACT_SOFTPLUS>(v);
    case HP_ACT_SWISH: return hp_act_c<HP_ACT_SWISH>(v);
    case HP_ACT_LEAKY_RELU: return hp_act_c<HP_ACT_LEAKY_RELU>(v);
    default: return v;
  }
}
#endif
