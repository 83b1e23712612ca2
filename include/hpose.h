/*
 * hpose.h -- C ABI of libhpose.so, the B200 (sm_100a) implementation of the reference's
 * batched head-pose hot path.
 *
 * The reference (Maaz77/Head-Pose-Estimation-Model) has no FFI layer: its seam is Python
 * duck-typing on a Keras model.  Each entry point below cites the reference interface it
 * replaces; INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative hp_status; hp_last_error() returns a
 *     thread-local human readable message for the last failure;
 *   - pointers are DEVICE pointers unless the parameter name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - handles are opaque, one per GPU, not thread-safe;
 *   - all feature maps are NHWC float32, batch-major; there is NO CPU fallback anywhere.
 */
#ifndef HPOSE_H_
#define HPOSE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hp_ctx* hp_handle;

enum hp_status {
  HP_OK = 0,
  HP_ERR_INVALID = -1,   /* bad argument (mirrors the reference's ValueError sites)            */
  HP_ERR_CUDA = -2,      /* CUDA runtime failure                                               */
  HP_ERR_STATE = -3,     /* call order violated (e.g. forward before load_weights)             */
  HP_ERR_NCCL = -4,      /* NCCL missing or failed                                             */
  HP_ERR_UNSUPPORTED = -5
};

#define HP_BACKBONE_PARAMS 101390 /* stem + 16 BlazeBlocks + 4 detector heads */
#define HP_MAX_FACES 100          /* MAX_FACE_NUM, blazeFaceDetectorH5.py:9                    */
#define HP_KEYPOINTS 6            /* KEY_POINT_SIZE, blazeFaceDetectorH5.py:8                  */

/* kernel family selector for the backbone (all CUDA): FAST = fused BlazeBlock kernels with TMA tile
 * load/store (product path), NAIVE = one-thread-per-output kernels kept as an on-device cross-check,
 * CPASYNC = first-generation fused kernels (cp.async tiles), kept for A/B measurements. */
enum hp_impl { HP_IMPL_FAST = 0, HP_IMPL_NAIVE = 1, HP_IMPL_CPASYNC = 2, HP_IMPL_TMA = 3 };

const char* hp_last_error(void);
int hp_version(void);
/* bit 0: built with -DHP_LEGACY_KERNELS (the first-generation pipelined tensor-core block kernel, kept for the geometry tests
 * and sweeps only; the default build does not carry it) */
int hp_build_features(void);

int hp_create(int device, hp_handle* out);
int hp_destroy(hp_handle h);
int hp_set_impl(hp_handle h, int impl);
/* number of kernels this library has launched on the handle since creation (bench `gpu_launches`) */
int64_t hp_launch_count(hp_handle h);

/* ---------------------------------------------------------------- backbone (a1-a3)
 * Replaces the Keras model call `self.interpreter(x)` (blazeFaceDetectorH5.py:272) for the
 * detector part of the unified graph (layers conv2d .. conv2d_20).
 *
 * packed_host layout (layout_id 0, all float32, Keras memory order):
 *   stem kernel[5][5][3][24], stem bias[24],
 *   for blk in 0..15: dw kernel[3][3][Cin], dw bias[Cin], pw kernel[Cin][Cout], pw bias[Cout],
 *   cls16 kernel[88][2], bias[2], cls8 kernel[96][6], bias[6],
 *   loc16 kernel[88][32], bias[32], loc8 kernel[96][96], bias[96]          (HP_BACKBONE_PARAMS floats)
 */
int hp_backbone_load_weights(hp_handle h, const float* packed_host, size_t n_floats, int layout_id);
/* number of successful hp_backbone_load_weights calls on this context: the backbone weights belong to the context, so a host
 * object that shares it compares this counter with the value it saw after its own load to learn whether to reload */
long long hp_backbone_generation(hp_handle h);

/* anchors for an HxW input: A = ceil(H/8)*ceil(W/8)*2 + ceil(H/16)*ceil(W/16)*6 */
int hp_num_anchors(int H, int W);

/* x[B,H,W,3] in [-1,1] -> feat16[B,H/8,W/8,88] (re_lu_10), feat8[B,H/16,W/16,96] (re_lu_15),
 * cls[B,A] raw logits, loc[B,A,16] raw box/keypoint regressions (anchor order of
 * classificators_1|2 / regressors_1|2 concatenated as in blazeFaceDetectorH5.py:280-281).
 * Any of the four outputs may be NULL (feature maps are then kept in internal buffers). */
int hp_backbone_forward(hp_handle h, const float* x, int B, int H, int W,
                        float* feat16, float* feat8, float* cls, float* loc, void* stream);

/* Input range of the float entry points (hp_backbone_forward, hp_unified_forward): the tensor-core stem splits x into two
 * fp16 parts, exact to fp32 level for |x| <= 65504 -- the reference feeds normalised pixels in [-1, 1]
 * (blazeFaceDetectorH5.py:262).  A value outside that range (or inf / NaN) cannot be represented: the stem kernel then sets
 * HP_STATUS_STEM_RANGE in a sticky per-context status word instead of failing silently.  hp_backbone_status synchronises
 * `stream`, returns the word and clears it; hp_debug_set_stem_tc(h, -1, 0, 0, 0) selects the fp32 CUDA-core stem, which has
 * no such limit (the Python layer re-runs a flagged batch that way).  The uint8 entry points cannot trigger it. */
#define HP_STATUS_STEM_RANGE 1u
#define HP_STATUS_CHAIN_RANGE 4u    /* a depthwise output of blocks 6-15 does not fit fp16 (|a| >= 65520): the fused chain kernels multiply
                                     * split-fp16 operands; hp_debug_set_chain(h, 2 + 4, 0, 0) selects their 3xTF32 form, which has no such
                                     * limit (the Python layer re-runs a flagged batch that way) */
#define HP_STATUS_P2P_TIMEOUT 2u    /* a rank of the peer-memory gradient exchange waited ~2 s for a peer (see hp_p2p_open) */
int hp_backbone_status(hp_handle h, unsigned int* flags_host, void* stream);

/* debugging / parity hook: run the backbone on x up to BlazeBlock `blk` (0..15; -1 = stem only) and
 * copy that activation into dst (real channel count, NHWC). */
int hp_backbone_read_activation(hp_handle h, const float* x, int B, int H, int W, int blk, float* dst,
                                size_t dst_floats, void* stream);

/* ---------------------------------------------------------------- pre-processing
 * Replaces prepareInputForInference (blazeFaceDetectorH5.py:247-269) for images that already
 * have the network input size: BGR uint8 [B,H,W,3] -> RGB float32 ((v/255)-0.5)/0.5. */
int hp_preprocess_u8(hp_handle h, const uint8_t* bgr, int B, int H, int W, float* x, void* stream);
/* The whole of prepareInputForInference (blazeFaceDetectorH5.py:247-269) for frames of any size: BGR uint8 [B,Hin,Win,3] ->
 * RGB, /255.0, tf.image.resize(method='bicubic') to Hout x Wout (TensorFlow's ResizeBicubic with half-pixel centres: Keys
 * cubic A = -0.5 from its 1025-entry table, border taps dropped and renormalised, rows then columns, float32), (t-0.5)/0.5
 * -> float32 [B,Hout,Wout,3].  Equal sizes take the hp_preprocess_u8 path (the resize is the identity there). */
int hp_preprocess_resize_u8(hp_handle h, const uint8_t* bgr, int B, int Hin, int Win, int Hout, int Wout, float* x,
                            void* stream);

/* ---------------------------------------------------------------- regressor heads (a4-a6)
 * A head is a small program over per-token tensors ("registers"); the Python layer compiles the
 * Keras graph built by se_transformer_regr_head / create_model* (attention_model.py:16-169,
 * train_88.py:66-253, train_96.py:65-110) into this form.  Register 0 is the input feature map.
 */
enum hp_head_opcode {
  HP_OP_DENSE = 1,     /* out = act(in0 @ W[cin][cout] + b)   Conv2D 1x1 / Dense               */
  HP_OP_ACT = 2,       /* out = act(in0)                                                        */
  HP_OP_ADD = 3,       /* out = in0 + in1                                                       */
  HP_OP_MULCH = 4,     /* out[n,t,c] = in0[n,t,c] * in1[n,c]  (SE gate broadcast)               */
  HP_OP_GAP = 5,       /* out[n,c] = mean_t in0[n,t,c]        GlobalAveragePooling2D            */
  HP_OP_DROPOUT = 6,   /* SpatialDropout2D: mask (n,c), scale 1/(1-rate); identity at inference */
  HP_OP_LAYERNORM = 7, /* LayerNormalization over channels, gamma=w_off beta=b_off eps=fparam   */
  HP_OP_MHA = 8        /* self-attention over tokens: params at w_off in the order
                          Wq[C][h*d] bq[h*d] Wk bk Wv bv Wo[h*d][C] bo[C]                        */
};
/* Keras activation names; elu has alpha = 1, leaky_relu the slope 0.2 of tf.nn.leaky_relu, selu the Keras constants.  swish
 * (x sigmoid(x)) is forward-only: its derivative is not a function of its output, which is all a training step keeps. */
enum hp_act { HP_ACT_LINEAR = 0, HP_ACT_RELU = 1, HP_ACT_TANH = 2, HP_ACT_SIGMOID = 3, HP_ACT_SOFTSIGN = 4,
              HP_ACT_ELU = 5, HP_ACT_SELU = 6, HP_ACT_SOFTPLUS = 7, HP_ACT_SWISH = 8, HP_ACT_LEAKY_RELU = 9 };

typedef struct hp_head_op {
  int32_t op, in0, in1, out;
  int32_t cin, cout, act;
  int32_t w_off, b_off;      /* offsets (floats) into the head's flat parameter vector          */
  int32_t heads, key_dim;    /* HP_OP_MHA                                                       */
  int32_t op_id;             /* stable id mixed into the dropout hash                           */
  float fparam;              /* dropout rate | layer-norm epsilon                               */
  float l2_w, l2_b;          /* Keras L2 regulariser coefficients for kernel / bias             */
} hp_head_op;

typedef struct hp_head_reg {
  int32_t channels;
  int32_t per_image;         /* 1: one vector per image (after GAP), 0: one per token           */
} hp_head_reg;

typedef struct hp_head* hp_head_t;

int hp_head_create(hp_handle h, const hp_head_op* ops, int n_ops, const hp_head_reg* regs, int n_regs,
                   int out_reg, int n_params, hp_head_t* out);
int hp_head_destroy(hp_handle h, hp_head_t head);
int hp_head_set_weights(hp_handle h, hp_head_t head, const float* params_host, int n_params);
int hp_head_get_weights(hp_handle h, hp_head_t head, float* params_host, int n_params);

/* model.predict / model.__call__ (test.py:34): feat[B,H,W,Cin] -> out[B,H,W,Cout]; tokens = H*W. */
int hp_head_forward(hp_handle h, hp_head_t head, const float* feat, int B, int H, int W, float* out,
                    void* stream);

enum hp_optimizer { HP_OPT_SGD = 0, HP_OPT_ADAM = 1, HP_OPT_ADAMAX = 2 };
typedef struct hp_opt_config {
  int32_t kind;
  float lr, beta1, beta2, eps;
} hp_opt_config;

/* One optimizer step of model.fit (train_96.py:175, train_88.py:355): forward with dropout,
 * loss = mean((pred-y)^2) + sum l2*|w|^2, backward, [allreduce if hp_comm_init was called],
 * optimizer update.  x[n,H,W,Cin], y[n,H,W,3] are this rank's shard; n_global = samples summed over
 * all ranks (== n on one GPU).  loss_mae_host[2] receives {loss incl. L2, mae} of the GLOBAL batch
 * (may be NULL to avoid the device->host sync).  seed/step select the dropout masks. */
int hp_head_train_step(hp_handle h, hp_head_t head, const float* x, const float* y, int n, int H, int W,
                       int n_global, const hp_opt_config* opt, uint64_t seed, float* loss_mae_host,
                       void* stream);
/* Many steps without host round trips: the body of one epoch of model.fit (train_96.py:175-183) on a DEVICE-RESIDENT data
 * set x_all[n_items,H,W,Cin], y_all[n_items,H,W,3].  Step s (0 <= s < n_steps) trains on the global batch of items
 * idx[first_item + s * batch_global ...] (idx: device int32 permutation of the epoch, NULL = identity); this rank gathers
 * rows rank, rank + world, ... of it (a rank may get none: it then contributes zeros to the all-reduce), runs forward,
 * loss, backward, the gradient all-reduce and the optimizer exactly as hp_head_train_step does.  The step counter, the
 * Adam / Adamax step sizes and the loss / mae accumulators live in device memory, so with HP_TRAIN_GRAPH one step is
 * captured once and every further step is ONE CUDA-graph launch (needs a non-default stream).
 * sums_host (may be NULL: fully asynchronous) receives {sum over steps of loss * batch_global, of mae * batch_global}
 * after one stream synchronisation at the end of the call. */
#define HP_TRAIN_GRAPH 1
int hp_head_train_run(hp_handle h, hp_head_t head, const float* x_all, const float* y_all, const int32_t* idx,
                      long long n_items, long long first_item, int batch_global, int n_steps, int rank, int world, int H,
                      int W, const hp_opt_config* opt, uint64_t seed, int flags, double* sums_host, void* stream);
/* gradients of the last train step (after allreduce, before L2), for parity tests */
int hp_head_get_grads(hp_handle h, hp_head_t head, float* grads_host, int n_params);
/* model.evaluate (train_96.py:186): mse_mae_host[3] = {mse, mae, l2 penalty}; Keras reports
 * loss = mse + l2 penalty.  No update. */
int hp_head_evaluate(hp_handle h, hp_head_t head, const float* x, const float* y, int n, int H, int W,
                     float* mse_mae_host, void* stream);
/* SpatialDropout2D keep decision used by the train step (exposed so the oracle can replay it) */
uint32_t hp_dropout_hash(uint64_t seed, uint32_t step, uint32_t op_id, uint32_t image, uint32_t channel);

/* ---------------------------------------------------------------- decode + NMS (a7-a9)
 * Replaces filterDetections (blazeFaceDetectorH5.py:319-327), extractDetections (:284-317) and
 * filterWithNonMaxSupression (:329-357, tf.image.non_max_suppression + pose lookup) per image.
 *   cls[B,A], loc[B,A,16], pose16[B,H16,W16,3], pose8[B,H8,W8,3]
 *   logit_thr = (float)log(t/(1-t)); iou_thr as float32; max_out <= HP_MAX_FACES
 * outputs (row-major, per image, in selection order = score descending):
 *   out_cnt[B], out_anchor[B,max_out] (anchor ids), boxes[B,max_out,4] float64 [x1,y1,x2,y2],
 *   kps[B,max_out,6,2] float64, scores[B,max_out] float32, poses[B,max_out,3] float32.
 * boxes/kps/scores/poses may be NULL. */
int hp_decode_nms(hp_handle h, const float* cls, const float* loc, const float* pose16, const float* pose8,
                  int B, int H, int W, float logit_thr, float iou_thr, int max_out,
                  int32_t* out_cnt, int32_t* out_anchor, double* boxes, double* kps, float* scores,
                  float* poses, void* stream);

/* The same three steps as separate calls, one per reference method (used by the facade's
 * filterDetections / extractDetections / filterWithNonMaxSupression):
 *   hp_filter_detections: cls[B,A] -> out_idx[B,A] (ascending anchor ids, first out_cnt[b] valid),
 *                         out_scores[B,A] float32 sigmoid                       (blazeFaceDetectorH5.py:319-327)
 *   hp_extract_detections: loc[A,16] of one image + idx[n] -> boxes[n,4], kps[n,6,2] float64 (:284-317)
 *   hp_nms: boxes[n,4] float64 (cast to float32 like TensorFlow), scores[n] -> out_sel[<=max_out] indices
 *           into the candidate list in selection order, *out_cnt                (:332)            */
int hp_filter_detections(hp_handle h, const float* cls, int B, int A, float logit_thr, int32_t* out_idx,
                         float* out_scores, int32_t* out_cnt, void* stream);
int hp_extract_detections(hp_handle h, const float* loc, const int32_t* idx, int n, int H, int W,
                          double* boxes, double* kps, void* stream);
int hp_nms(hp_handle h, const double* boxes, const float* scores, int n, float iou_thr, int max_out,
           int32_t* out_sel, int32_t* out_cnt, void* stream);

/* ---------------------------------------------------------------- frames -> faces in one call
 * Replaces blazeFaceDetector.detectFaces (blazeFaceDetectorH5.py:109-126) end to end for B uint8 BGR frames [B,Hin,Win,3]
 * (device pointer): prepareInputForInference with the bicubic resize to H x W, the unified graph, filterDetections,
 * extractDetections, filterWithNonMaxSupression and the pose lookup.  The kept faces of all frames are PACKED into
 * `result` (device pointer, 8-byte aligned):
 *     int32 header[HP_RESULT_HEADER_INTS] = { total = sum(count), written = min(total, capacity), B, capacity }
 *     int32 count[B]                         faces kept per frame, frame order
 *     (padding to 8 bytes: hp_detect_result_faces_offset(B))
 *     hp_face faces[written]                 frame-major, score-descending within a frame (the reference's order)
 * so that the caller moves one buffer each way.  capacity = (result_bytes - offset) / sizeof(hp_face), at most B * max_out.
 * flags & HP_DETECT_GRAPH: the launch sequence is captured on the second call with the same arguments (pointers and
 * thresholds included) and replayed as ONE CUDA-graph launch from then on -- the per-frame latency path of the
 * reference's webcam loop (:392-444).  Needs a non-default stream. */
typedef struct hp_face {
  double box[4];          /* x1, y1, x2, y2, normalised (extractDetections :284-317, float64 like the reference) */
  double keypoints[12];   /* 6 x (x, y) */
  float score;            /* sigmoid(cls), float32 (:319-327) */
  float pose[3];          /* yaw, pitch, roll of the anchor's cell (:342-356) */
  int32_t anchor;         /* kept anchor id */
  int32_t frame;          /* index of the frame in the batch */
} hp_face;
#define HP_RESULT_HEADER_INTS 4
#define HP_DETECT_GRAPH 1
size_t hp_detect_result_faces_offset(int B);
size_t hp_detect_result_bytes(int B, int capacity);
int hp_detect_frames(hp_handle h, hp_head_t head16, hp_head_t head8, const uint8_t* frames_bgr, int B, int Hin, int Win,
                     int H, int W, float logit_thr, float iou_thr, int max_out, void* result, size_t result_bytes,
                     int flags, void* stream);

/* ---------------------------------------------------------------- unified path (config 5)
 * backbone + detector heads + two pose heads (JoinModels.py:5-90 output order) + decode + NMS.
 * pose16/pose8 outputs [B,H16,W16,3] / [B,H8,W8,3] may be NULL (internal). */
int hp_unified_forward(hp_handle h, hp_head_t head16, hp_head_t head8, const float* x, int B, int H, int W,
                       float logit_thr, float iou_thr, int max_out,
                       float* pose16, float* pose8,
                       int32_t* out_cnt, int32_t* out_anchor, double* boxes, double* kps, float* scores,
                       float* poses, void* stream);

/* ---------------------------------------------------------------- data-parallel training
 * nccl_unique_id_host: the 128-byte ncclUniqueId produced by hp_comm_unique_id on rank 0 and
 * broadcast by the caller (torch.distributed / MPI / files). */
int hp_comm_unique_id(void* id128_host);
int hp_comm_init(hp_handle h, const void* nccl_unique_id_host, int rank, int nranks);
int hp_comm_destroy(hp_handle h);
/* Peer-memory gradient exchange (the default of hp_head_train_run when it is set up): instead of ncclAllReduce followed by an
 * optimizer kernel, ONE kernel pushes every rank's gradient slices into every peer's inbox over NVLink / NVSwitch, waits for the
 * world's slices, sums them in rank order (bit-identical on all ranks) and applies the optimizer.
 *   hp_p2p_alloc : allocate this rank's inbox for buffers of up to cap_floats floats and return its 64-byte CUDA IPC handle
 *   hp_p2p_open  : all_handles_host = the nranks handles in rank order (exchanged by the caller); maps the peers' inboxes
 *   hp_p2p_close : unmap / free (also done by hp_comm_destroy / hp_destroy)
 * Needs peer access between the GPUs (one process per GPU on one node, at most 8 ranks); when it is not set up, or a head has
 * more than cap_floats - 3 parameters, the step falls back to hp_comm_init's NCCL all-reduce. */
int hp_p2p_alloc(hp_handle h, int cap_floats, int nranks, void* ipc_handle64_host);
int hp_p2p_open(hp_handle h, const void* all_handles_host, int rank, int nranks);
int hp_p2p_close(hp_handle h);
int hp_debug_set_p2p(hp_handle h, int on);

/* ---------------------------------------------------------------- measurement helpers
 * fp32 FMA micro-benchmark (SURVEY 8d asks for the measured CUDA-core peak): returns TFLOP/s.
 * mode 0 = scalar FFMA, mode 1 = packed fma.rn.f32x2 */
int hp_fma_peak(hp_handle h, int mode, double* tflops_out);
/* average device time (ms) of the backbone kernels of the last hp_backbone_forward (CUDA events on
 * the launch stream); per_layer_ms may be NULL or float[18] = stem, 16 blocks, det heads */
int hp_backbone_profile(hp_handle h, const float* x, int B, int H, int W, int iters, float* per_layer_ms);

/* tuning hooks: force the tile shape of one BlazeBlock kernel (TH=0 restores the heuristic; MT = pixels
 * per thread, 4 or 8) and have the forward pass record {TH,TW,IMGS,nbuf,threads,smem,n_tiles,MT} per block
 * into report16x8 (host int[128]) */
int hp_debug_set_tile(hp_handle h, int blk, int TH, int TW, int IMGS, int nbuf, int MT);
int hp_debug_tile_report(hp_handle h, int* report16x8);
/* tensor-core BlazeBlock kernel (HP_IMPL_FAST, stride-1 blocks): TR rows per lane, ring depth, band height, pipelines per
 * CTA, warp sets per pipeline, halo buffers of the warp-specialised variant (0 = pipelined variant; + 16 x depthwise work unit (1 or 2) + 64 x MMA issuer warps + 512 x issuer placement); TR = 0 restores the defaults, TR = -1 keeps the block on the CUDA-core kernel */
/* tensor-core stem kernel: band height, input buffers, output stages, gather warp sets (0 = automatic; BH = -1 keeps the
 * stem on the CUDA-core kernel) */
int hp_debug_set_stem_tc(hp_handle h, int BH, int nbuf, int nout, int nsets);
/* cross-block fusion (blocks 6-10 and 12-15 as one persistent kernel each): mode 0 = one kernel per block, 1 = chains, 2 = chains
 * with the stride-2 block 11 riding on the first one (default); + 4 = 3xTF32 products instead of split fp16;
 * nsets / niss override the worker warp sets / MMA issuer threads of the chain kernel (0 = default) */
int hp_debug_set_chain(hp_handle h, int mode, int nsets, int niss);
/* watchdog record of the chain kernels since the last call: out8_host[0] != 0 -> a barrier wait timed out ({1, barrier id, parity,
 * thread, CTA, step}); synchronises the device and clears the record */
int hp_debug_chain_status(hp_handle h, unsigned int* out8_host);
/* host-only (no device needed): geometry of the chain kernel for blocks [first, first + nblk) (+ the stride-2 block behind them when
 * tail != 0) on an H x W map: out[0..5] = {rows per lane, images per tile, pixel stride in floats, lanes, tail lanes, shared-memory bytes},
 * out[8..135] = lane table (image | strip << 8 | column << 16 per TMEM lane), out[136..263] = tail table (image | oy << 8 | ox << 16 |
 * swapped order << 31); HP_ERR_UNSUPPORTED when the chain kernel does not apply */
int hp_debug_chain_describe(int first, int nblk, int H, int W, int tail, unsigned int* out264_host);
/* device buffer of max_tiles x 12 clock64 stamps written by CTA 0 of the warp-specialised tensor-core kernel (NULL = off) */
int hp_debug_tc_trace(hp_handle h, long long* dev_buf, int max_tiles);
int hp_debug_set_tc(hp_handle h, int blk, int TR, int NSTG, int BH, int npipe, int nsets, int nbuf);
/* one Dense / 1x1-conv layer y[M][N] = act(x[M][K] W[K][N] + b) on device pointers, through the same dispatch as the heads
 * (tensor-core kernel when the shape allows it).  W2 != NULL: the fused pair z[M][n2] = act2(y W2[N][n2] + b2), n2 <= 4, is
 * written to y instead and the hidden activations are not stored (error if the shape is not supported).  Reference:
 * Conv2D(1x1) layers of Model-88/train_88.py:30-60 and Model-88/attention_model.py:16-169. */
int hp_debug_dense(hp_handle h, const float* x, int M, int K, const float* W, const float* b, int N, int act, float* y,
                   const float* W2, const float* b2, int n2, int act2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HPOSE_H_ */
