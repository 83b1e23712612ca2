// extern "C" surface of libhpose.so (declared in include/hpose.h) plus handle management and the
// fp32 FMA micro-benchmark used for roofline context.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";
void hp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// implemented in the other translation units
int hp_backbone_load_weights_impl(hp_ctx* h, const float* src, size_t n_floats, int layout_id);
int hp_head_create_impl(hp_ctx* h, const hp_head_op* ops, int n_ops, const hp_head_reg* regs, int n_regs, int out_reg,
                        int n_params, hp_head** out);
void hp_head_free_impl(hp_head* hd);
int hp_head_set_weights_impl(hp_head* hd, const float* src, int n);
int hp_head_get_weights_impl(hp_head* hd, float* dst, int n);
int hp_head_get_grads_impl(hp_head* hd, float* dst, int n);
int hp_head_in_channels(hp_head* hd);
int hp_head_out_channels(hp_head* hd);
int hp_head_forward_impl(hp_ctx* h, hp_head* hd, const float* feat, int B, int H, int W, float* out, cudaStream_t st);
int hp_head_train_step_impl(hp_ctx* h, hp_head* hd, const float* x, const float* y, int n, int H, int W, int n_global,
                            const hp_opt_config* opt, uint64_t seed, float* loss_mae_host, bool update, cudaStream_t st);
int hp_head_train_run_impl(hp_ctx* h, hp_head* hd, const float* x_all, const float* y_all, const int32_t* idx, long long n_items,
                           long long first_item, int batch_global, int n_steps, int rank, int world, int H, int W,
                           const hp_opt_config* opt, uint64_t seed, int flags, double* sums_host, cudaStream_t st);
int hp_decode_nms_impl(hp_ctx* h, const float* cls, const float* loc, const float* pose16, const float* pose8, int B,
                       int H, int W, float logit_thr, float iou_thr, int max_out, int32_t* out_cnt, int32_t* out_anchor,
                       double* boxes, double* kps, float* scores, float* poses, cudaStream_t st);
int hp_preprocess_u8_impl(hp_ctx* h, const uint8_t* bgr, int B, int H, int W, float* x, cudaStream_t st);
int hp_preprocess_resize_u8_impl(hp_ctx* h, const uint8_t* bgr, int B, int Hin, int Win, int Hout, int Wout, float* x, cudaStream_t st);
void hp_resize_plans_free(hp_ctx* h);
void hp_detect_graphs_free(hp_ctx* h);

#define HP_ENTER(h)                                                                   \
  HP_REQUIRE((h) != nullptr, HP_ERR_INVALID, "null handle");                          \
  HP_CUDA(cudaSetDevice((h)->device))

extern "C" {

const char* hp_last_error(void) { return g_err; }
int hp_version(void) { return 100; }
int hp_build_features(void) {
#ifdef HP_LEGACY_KERNELS
  return 1;
#else
  return 0;
#endif
}

int hp_create(int device, hp_handle* out) {
  HP_REQUIRE(out != nullptr, HP_ERR_INVALID, "hp_create: null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    hp_set_error("hp_create: no CUDA device available (%s); libhpose has no CPU fallback", cudaGetErrorString(e));
    return HP_ERR_CUDA;
  }
  HP_REQUIRE(device >= 0 && device < count, HP_ERR_INVALID, "hp_create: device %d out of range (0..%d)", device, count - 1);
  HP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  HP_CUDA(cudaGetDeviceProperties(&prop, device));
  HP_REQUIRE(prop.major >= 10, HP_ERR_UNSUPPORTED, "hp_create: %s is sm_%d%d; this library is built for sm_100a only",
             prop.name, prop.major, prop.minor);
  hp_ctx* h = new hp_ctx();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  if (h->status.ensure(sizeof(unsigned int)) != HP_OK || cudaMemset(h->status.p, 0, sizeof(unsigned int)) != cudaSuccess) {
    delete h;
    hp_set_error("hp_create: cannot allocate the status word");
    return HP_ERR_CUDA;
  }
  *out = h;
  return HP_OK;
}

int hp_destroy(hp_handle h) {
  if (!h) return HP_OK;
  cudaSetDevice(h->device);
  hp_comm_destroy(h);
  hp_resize_plans_free(h);
  hp_detect_graphs_free(h);
  h->det.release();
  h->bb.arena.release(); h->bb.act[0].release(); h->bb.act[1].release(); h->bb.dwtmp.release();
  h->bb.feat16.release(); h->bb.feat8.release();
  h->pose16.release(); h->pose8.release(); h->cls.release(); h->loc.release(); h->scratch.release(); h->status.release();
  if (h->ev[0]) cudaEventDestroy(h->ev[0]);
  if (h->ev[1]) cudaEventDestroy(h->ev[1]);
  delete h;
  return HP_OK;
}

int hp_set_impl(hp_handle h, int impl) {
  HP_REQUIRE(h && impl >= HP_IMPL_FAST && impl <= HP_IMPL_TMA, HP_ERR_INVALID, "hp_set_impl: bad arguments");
  h->impl = impl;
  return HP_OK;
}
int64_t hp_launch_count(hp_handle h) { return h ? h->launches : 0; }

int hp_backbone_load_weights(hp_handle h, const float* packed_host, size_t n_floats, int layout_id) {
  HP_ENTER(h);
  const int rc = hp_backbone_load_weights_impl(h, packed_host, n_floats, layout_id);
  if (rc == HP_OK) h->bb_generation++;
  return rc;
}
long long hp_backbone_generation(hp_handle h) { return h ? h->bb_generation : -1; }

int hp_num_anchors(int H, int W) { return ceil_div(H, 8) * ceil_div(W, 8) * 2 + ceil_div(H, 16) * ceil_div(W, 16) * 6; }

int hp_backbone_forward(hp_handle h, const float* x, int B, int H, int W, float* feat16, float* feat8, float* cls,
                        float* loc, void* stream) {
  HP_ENTER(h);
  return hp_backbone_run(h, x, B, H, W, feat16, feat8, cls, loc, 99, nullptr, 0, nullptr, 0, (cudaStream_t)stream);
}

int hp_backbone_read_activation(hp_handle h, const float* x, int B, int H, int W, int blk, float* dst, size_t dst_floats,
                                void* stream) {
  HP_ENTER(h);
  HP_REQUIRE(blk >= -1 && blk <= 15, HP_ERR_INVALID, "hp_backbone_read_activation: blk must be -1..15");
  return hp_backbone_run(h, x, B, H, W, nullptr, nullptr, nullptr, nullptr, blk, dst, dst_floats, nullptr, 0,
                         (cudaStream_t)stream);
}

int hp_backbone_profile(hp_handle h, const float* x, int B, int H, int W, int iters, float* per_layer_ms) {
  HP_ENTER(h);
  HP_REQUIRE(per_layer_ms != nullptr, HP_ERR_INVALID, "hp_backbone_profile: per_layer_ms is required");
  HP_TRY(h->cls.ensure((size_t)B * hp_num_anchors(H, W) * sizeof(float)));
  HP_TRY(h->loc.ensure((size_t)B * hp_num_anchors(H, W) * 16 * sizeof(float)));
  return hp_backbone_run(h, x, B, H, W, nullptr, nullptr, h->cls.f(), h->loc.f(), 99, nullptr, 0, per_layer_ms, iters,
                         (cudaStream_t)0);
}

int hp_debug_set_tile(hp_handle h, int blk, int TH, int TW, int IMGS, int nbuf, int MT) {
  HP_REQUIRE(h && blk >= 0 && blk < 16, HP_ERR_INVALID, "hp_debug_set_tile: bad arguments");
  h->tile_override[blk][0] = TH; h->tile_override[blk][1] = TW; h->tile_override[blk][2] = IMGS; h->tile_override[blk][3] = nbuf;
  h->tile_override[blk][4] = MT;
  return HP_OK;
}
int hp_debug_set_tc(hp_handle h, int blk, int TR, int NSTG, int BH, int npipe, int nsets, int nbuf) {
  HP_REQUIRE(h && blk >= 0 && blk < 16, HP_ERR_INVALID, "hp_debug_set_tc: bad arguments");
  h->tc_override[blk][0] = TR; h->tc_override[blk][1] = NSTG; h->tc_override[blk][2] = BH; h->tc_override[blk][3] = npipe; h->tc_override[blk][4] = nsets; h->tc_override[blk][5] = nbuf % 16; h->tc_override[blk][6] = (nbuf / 16) % 4; h->tc_override[blk][7] = (nbuf / 64) % 8; h->tc_override[blk][8] = nbuf / 512;
  return HP_OK;
}
int hp_debug_set_stem_tc(hp_handle h, int BH, int nbuf, int nout, int nsets) {
  HP_REQUIRE(h, HP_ERR_INVALID, "null handle");
  h->stem_tc_cfg[0] = BH; h->stem_tc_cfg[1] = nbuf; h->stem_tc_cfg[2] = nout; h->stem_tc_cfg[3] = nsets;
  return HP_OK;
}
int hp_debug_set_p2p(hp_handle h, int on) {
  HP_REQUIRE(h, HP_ERR_INVALID, "null handle");
  h->p2p_off = !on;
  return HP_OK;
}
int hp_debug_set_chain(hp_handle h, int mode, int nsets, int niss) {
  HP_REQUIRE(h, HP_ERR_INVALID, "null handle");
  h->chain_mode = mode; h->chain_cfg[0] = nsets; h->chain_cfg[1] = niss;
  return HP_OK;
}
int hp_debug_chain_status(hp_handle h, unsigned int* out8_host) {
  HP_ENTER(h);
  HP_REQUIRE(out8_host != nullptr, HP_ERR_INVALID, "hp_debug_chain_status: null pointer");
  return hp_chain_status(out8_host);
}
int hp_debug_chain_describe(int first, int nblk, int H, int W, int tail, unsigned int* out264_host) {
  HP_REQUIRE(out264_host != nullptr, HP_ERR_INVALID, "hp_debug_chain_describe: null pointer");
  return hp_chain_describe(first, nblk, H, W, tail, out264_host);
}
int hp_debug_tc_trace(hp_handle h, long long* dev_buf, int max_tiles) {
  HP_REQUIRE(h, HP_ERR_INVALID, "null handle");
  h->tc_trace = dev_buf; h->tc_trace_tiles = dev_buf ? max_tiles : 0;
  return HP_OK;
}
int hp_debug_dense(hp_handle h, const float* x, int M, int K, const float* W, const float* b, int N, int act, float* y,
                   const float* W2, const float* b2, int n2, int act2, void* stream) {
  HP_ENTER(h);
  HP_REQUIRE(x && W && y && M > 0 && K > 0 && N > 0, HP_ERR_INVALID, "hp_debug_dense: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (W2) {   // y receives z = act2(act(x W + b) W2 + b2), [M][n2]
    HP_REQUIRE(h->impl == HP_IMPL_FAST && hp_dense_tc_tail_supported(x, M, K, K, N, n2), HP_ERR_UNSUPPORTED,
               "hp_debug_dense: %d x %d -> %d -> %d is not a fused tensor-core shape", M, K, N, n2);
    DenseTail tail{W2, b2, n2, n2, act2, DenseOut{y, 0, n2, M, 0, n2}};
    return hp_launch_dense_tc_tail(h, x, M, K, K, W, N, b, N, act, tail, st);
  }
  DenseOut d{y, 0, N, M, 0, N};
  return hp_launch_dense(h, x, M, K, K, W, N, b, N, act, false, &d, 1, false, st);
}
int hp_debug_tile_report(hp_handle h, int* report16x8) {
  HP_REQUIRE(h, HP_ERR_INVALID, "null handle");
  h->tile_report = report16x8;
  return HP_OK;
}

int hp_backbone_status(hp_handle h, unsigned int* flags_host, void* stream) {
  HP_ENTER(h);
  HP_REQUIRE(flags_host != nullptr, HP_ERR_INVALID, "hp_backbone_status: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  HP_CUDA(cudaMemcpyAsync(flags_host, h->status.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  HP_CUDA(cudaMemsetAsync(h->status.p, 0, sizeof(unsigned int), st));
  HP_CUDA(cudaStreamSynchronize(st));
  return HP_OK;
}

int hp_preprocess_u8(hp_handle h, const uint8_t* bgr, int B, int H, int W, float* x, void* stream) {
  HP_ENTER(h);
  return hp_preprocess_u8_impl(h, bgr, B, H, W, x, (cudaStream_t)stream);
}

int hp_preprocess_resize_u8(hp_handle h, const uint8_t* bgr, int B, int Hin, int Win, int Hout, int Wout, float* x, void* stream) {
  HP_ENTER(h);
  if (Hin == Hout && Win == Wout) return hp_preprocess_u8_impl(h, bgr, B, Hin, Win, x, (cudaStream_t)stream);   // the resize is the identity there
  return hp_preprocess_resize_u8_impl(h, bgr, B, Hin, Win, Hout, Wout, x, (cudaStream_t)stream);
}

int hp_head_create(hp_handle h, const hp_head_op* ops, int n_ops, const hp_head_reg* regs, int n_regs, int out_reg,
                   int n_params, hp_head_t* out) {
  HP_ENTER(h);
  return hp_head_create_impl(h, ops, n_ops, regs, n_regs, out_reg, n_params, out);
}
int hp_head_destroy(hp_handle h, hp_head_t head) {
  HP_ENTER(h);
  hp_head_free_impl(head);
  return HP_OK;
}
int hp_head_set_weights(hp_handle h, hp_head_t head, const float* params_host, int n_params) {
  HP_ENTER(h);
  return hp_head_set_weights_impl(head, params_host, n_params);
}
int hp_head_get_weights(hp_handle h, hp_head_t head, float* params_host, int n_params) {
  HP_ENTER(h);
  return hp_head_get_weights_impl(head, params_host, n_params);
}
int hp_head_get_grads(hp_handle h, hp_head_t head, float* grads_host, int n_params) {
  HP_ENTER(h);
  return hp_head_get_grads_impl(head, grads_host, n_params);
}
int hp_head_forward(hp_handle h, hp_head_t head, const float* feat, int B, int H, int W, float* out, void* stream) {
  HP_ENTER(h);
  return hp_head_forward_impl(h, head, feat, B, H, W, out, (cudaStream_t)stream);
}
int hp_head_train_step(hp_handle h, hp_head_t head, const float* x, const float* y, int n, int H, int W, int n_global,
                       const hp_opt_config* opt, uint64_t seed, float* loss_mae_host, void* stream) {
  HP_ENTER(h);
  return hp_head_train_step_impl(h, head, x, y, n, H, W, n_global, opt, seed, loss_mae_host, true, (cudaStream_t)stream);
}
int hp_head_train_run(hp_handle h, hp_head_t head, const float* x_all, const float* y_all, const int32_t* idx, long long n_items,
                      long long first_item, int batch_global, int n_steps, int rank, int world, int H, int W,
                      const hp_opt_config* opt, uint64_t seed, int flags, double* sums_host, void* stream) {
  HP_ENTER(h);
  return hp_head_train_run_impl(h, head, x_all, y_all, idx, n_items, first_item, batch_global, n_steps, rank, world, H, W, opt, seed,
                                flags, sums_host, (cudaStream_t)stream);
}
int hp_head_evaluate(hp_handle h, hp_head_t head, const float* x, const float* y, int n, int H, int W,
                     float* mse_mae_host, void* stream) {
  HP_ENTER(h);
  HP_REQUIRE(mse_mae_host != nullptr, HP_ERR_INVALID, "hp_head_evaluate: result pointer required");
  return hp_head_train_step_impl(h, head, x, y, n, H, W, n, nullptr, 0, mse_mae_host, false, (cudaStream_t)stream);
}

int hp_decode_nms(hp_handle h, const float* cls, const float* loc, const float* pose16, const float* pose8, int B, int H,
                  int W, float logit_thr, float iou_thr, int max_out, int32_t* out_cnt, int32_t* out_anchor,
                  double* boxes, double* kps, float* scores, float* poses, void* stream) {
  HP_ENTER(h);
  return hp_decode_nms_impl(h, cls, loc, pose16, pose8, B, H, W, logit_thr, iou_thr, max_out, out_cnt, out_anchor, boxes,
                            kps, scores, poses, (cudaStream_t)stream);
}

int hp_unified_forward(hp_handle h, hp_head_t head16, hp_head_t head8, const float* x, int B, int H, int W,
                       float logit_thr, float iou_thr, int max_out, float* pose16, float* pose8, int32_t* out_cnt,
                       int32_t* out_anchor, double* boxes, double* kps, float* scores, float* poses, void* stream) {
  HP_ENTER(h);
  HP_REQUIRE(head16 && head8, HP_ERR_INVALID, "hp_unified_forward: both regressor heads are required");
  HP_REQUIRE(hp_head_in_channels(head16) == 88 && hp_head_in_channels(head8) == 96, HP_ERR_INVALID,
             "hp_unified_forward: head16 must take 88 channels and head8 96 (got %d, %d)", hp_head_in_channels(head16),
             hp_head_in_channels(head8));
  HP_REQUIRE(hp_head_out_channels(head16) == 3 && hp_head_out_channels(head8) == 3, HP_ERR_INVALID,
             "hp_unified_forward: heads must output 3 channels");
  cudaStream_t st = (cudaStream_t)stream;
  const int A = hp_num_anchors(H, W);
  const int H16 = ceil_div(H, 8), W16 = ceil_div(W, 8), H8 = ceil_div(H, 16), W8 = ceil_div(W, 16);
  HP_TRY(h->cls.ensure((size_t)B * A * sizeof(float)));
  HP_TRY(h->loc.ensure((size_t)B * A * 16 * sizeof(float)));
  if (!pose16) {
    HP_TRY(h->pose16.ensure((size_t)B * H16 * W16 * 3 * sizeof(float)));
    pose16 = h->pose16.f();
  }
  if (!pose8) {
    HP_TRY(h->pose8.ensure((size_t)B * H8 * W8 * 3 * sizeof(float)));
    pose8 = h->pose8.f();
  }
  HP_TRY(hp_backbone_run(h, x, B, H, W, nullptr, nullptr, h->cls.f(), h->loc.f(), 99, nullptr, 0, nullptr, 0, st));
  HP_TRY(hp_head_forward_impl(h, head16, h->bb.feat16.f(), B, H16, W16, pose16, st));
  HP_TRY(hp_head_forward_impl(h, head8, h->bb.feat8.f(), B, H8, W8, pose8, st));
  if (out_cnt && out_anchor)
    HP_TRY(hp_decode_nms_impl(h, h->cls.f(), h->loc.f(), pose16, pose8, B, H, W, logit_thr, iou_thr, max_out, out_cnt,
                              out_anchor, boxes, kps, scores, poses, st));
  return HP_OK;
}

}  // extern "C"

// ============================================================================ FMA micro-benchmark
template <int MODE>
__global__ void __launch_bounds__(256) fma_bench_kernel(float* out, int iters, float b, float c) {
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = (float)(threadIdx.x + j) * 1e-3f;
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], b, c);
    }
  } else {
    unsigned long long bb, cc, v[8];
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v[j]) : "f"(a[2 * j]), "f"(a[2 * j + 1]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[j]) : "l"(bb), "l"(cc));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * j]), "=f"(a[2 * j + 1]) : "l"(v[j]));
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += a[j];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

extern "C" int hp_fma_peak(hp_handle h, int mode, double* tflops_out) {
  HP_ENTER(h);
  HP_REQUIRE(tflops_out && (mode == 0 || mode == 1), HP_ERR_INVALID, "hp_fma_peak: bad arguments");
  const int blocks = h->num_sms * 8, iters = 8192;
  HP_TRY(h->scratch.ensure((size_t)blocks * 256 * sizeof(float)));
  cudaEvent_t e0, e1;
  HP_CUDA(cudaEventCreate(&e0));
  HP_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    HP_CUDA(cudaEventRecord(e0, 0));
    if (mode == 0) fma_bench_kernel<0><<<blocks, 256>>>(h->scratch.f(), iters, 0.999f, 1e-3f);
    else fma_bench_kernel<1><<<blocks, 256>>>(h->scratch.f(), iters, 0.999f, 1e-3f);
    h->launches++;
    HP_CUDA(cudaEventRecord(e1, 0));
    HP_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    HP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = (double)blocks * 256.0 * iters * 16.0 * 2.0;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  return HP_OK;
}
