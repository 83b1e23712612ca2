// hp_detect_frames: the whole of blazeFaceDetector.detectFaces (BlazePoser/blazeFaceDetectorH5.py:109-126) for a batch of uint8
// BGR frames in ONE call: prepareInputForInference (:247-269, resize included) -> unified graph (:272) -> filterDetections
// (:319-327) -> extractDetections (:284-317) -> filterWithNonMaxSupression (:329-357, pose lookup included), with the kept
// faces of all frames PACKED into one result buffer (header + sum(count) records) so that a caller needs one host->device
// copy of the frames and one device->host copy of the result.  The reference pads nothing either: it returns k <= 100 faces
// per frame; the padded [B][100] form of hp_unified_forward costs 14.4 KB per frame on the way back, 12.8 KB of it padding.
//
// Latency mode (the reference's real use: one webcam frame per call, :392-444): with HP_DETECT_GRAPH the launch
// sequence (~60 kernels) is captured once per (shape, pointers, thresholds) key and replayed as one CUDA graph launch.
#include "common.cuh"

int hp_head_in_channels(hp_head* hd);
int hp_head_out_channels(hp_head* hd);
int hp_head_forward_impl(hp_ctx* h, hp_head* hd, const float* feat, int B, int H, int W, float* out, cudaStream_t st);
int hp_decode_nms_impl(hp_ctx* h, const float* cls, const float* loc, const float* pose16, const float* pose8, int B,
                       int H, int W, float logit_thr, float iou_thr, int max_out, int32_t* out_cnt, int32_t* out_anchor,
                       double* boxes, double* kps, float* scores, float* poses, cudaStream_t st);
int hp_preprocess_u8_impl(hp_ctx* h, const uint8_t* bgr, int B, int H, int W, float* x, cudaStream_t st);
int hp_preprocess_resize_u8_impl(hp_ctx* h, const uint8_t* bgr, int B, int Hin, int Win, int Hout, int Wout, float* x, cudaStream_t st);

long long g_devbuf_epoch = 0;   // bumped by every DevBuf reallocation: captured graphs hold raw pointers

namespace {

// exclusive prefix sum of the per-frame counts (one CTA; B is at most a few 10^4) -> offsets, total, records written
__global__ void __launch_bounds__(1024) pack_offsets_kernel(const int32_t* __restrict__ cnt, int B, int cap, int32_t* __restrict__ hdr,
                                                             int32_t* __restrict__ offsets) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + tid;
    const int v = b < B ? cnt[b] : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int base = s_carry + (warp ? s_warp[warp - 1] : 0);
    if (b < B) {
      offsets[b] = base + x - v;
      hdr[HP_RESULT_HEADER_INTS + b] = v;
    }
    __syncthreads();
    if (tid == 1023) s_carry = base + x;
    __syncthreads();
  }
  if (tid == 0) {
    hdr[0] = s_carry;
    hdr[1] = s_carry < cap ? s_carry : cap;
    hdr[2] = B;
    hdr[3] = cap;
  }
}

// one CTA per frame: its kept faces -> records at offsets[b]
__global__ void __launch_bounds__(128) pack_faces_kernel(const int32_t* __restrict__ cnt, const int32_t* __restrict__ offsets,
                                                         const int32_t* __restrict__ anchor, const double* __restrict__ boxes,
                                                         const double* __restrict__ kps, const float* __restrict__ scores,
                                                         const float* __restrict__ poses, int max_out, int cap, hp_face* __restrict__ faces) {
  const int b = blockIdx.x, n = cnt[b], off = offsets[b];
  // a record is 19 8-byte words: 4 box + 12 keypoints + (score, yaw) + (pitch, roll) + (anchor, frame)
  for (int i = threadIdx.x; i < n * 19; i += blockDim.x) {
    const int f = i / 19, w = i - f * 19;
    if (off + f >= cap) continue;
    const size_t src = (size_t)b * max_out + f;
    unsigned long long v;
    if (w < 4) v = (unsigned long long)__double_as_longlong(boxes[src * 4 + w]);
    else if (w < 16) v = (unsigned long long)__double_as_longlong(kps[src * 12 + (w - 4)]);
    else if (w == 16) v = (unsigned long long)__float_as_uint(scores[src]) | ((unsigned long long)__float_as_uint(poses[src * 3 + 0]) << 32);
    else if (w == 17) v = (unsigned long long)__float_as_uint(poses[src * 3 + 1]) | ((unsigned long long)__float_as_uint(poses[src * 3 + 2]) << 32);
    else v = (unsigned long long)(uint32_t)anchor[src] | ((unsigned long long)(uint32_t)b << 32);
    reinterpret_cast<unsigned long long*>(faces + off + f)[w] = v;
  }
}

int detect_launches(hp_ctx* h, hp_head* head16, hp_head* head8, const uint8_t* frames, int B, int Hin, int Win, int H, int W,
                    float logit_thr, float iou_thr, int max_out, void* result, int cap, cudaStream_t st) {
  const int A = hp_num_anchors(H, W);
  const int H16 = ceil_div(H, 8), W16 = ceil_div(W, 8), H8 = ceil_div(H, 16), W8 = ceil_div(W, 16);
  DetectBufs& d = h->det;
  HP_TRY(d.x.ensure((size_t)B * H * W * 3 * sizeof(float)));
  HP_TRY(h->cls.ensure((size_t)B * A * sizeof(float)));
  HP_TRY(h->loc.ensure((size_t)B * A * 16 * sizeof(float)));
  HP_TRY(h->pose16.ensure((size_t)B * H16 * W16 * 3 * sizeof(float)));
  HP_TRY(h->pose8.ensure((size_t)B * H8 * W8 * 3 * sizeof(float)));
  HP_TRY(d.cnt.ensure((size_t)B * sizeof(int32_t)));
  HP_TRY(d.offsets.ensure((size_t)B * sizeof(int32_t)));
  HP_TRY(d.anchor.ensure((size_t)B * max_out * sizeof(int32_t)));
  HP_TRY(d.boxes.ensure((size_t)B * max_out * 4 * sizeof(double)));
  HP_TRY(d.kps.ensure((size_t)B * max_out * 12 * sizeof(double)));
  HP_TRY(d.scores.ensure((size_t)B * max_out * sizeof(float)));
  HP_TRY(d.poses.ensure((size_t)B * max_out * 3 * sizeof(float)));
  if (Hin == H && Win == W) HP_TRY(hp_preprocess_u8_impl(h, frames, B, H, W, d.x.f(), st));
  else HP_TRY(hp_preprocess_resize_u8_impl(h, frames, B, Hin, Win, H, W, d.x.f(), st));
  HP_TRY(hp_backbone_run(h, d.x.f(), B, H, W, nullptr, nullptr, h->cls.f(), h->loc.f(), 99, nullptr, 0, nullptr, 0, st));
  HP_TRY(hp_head_forward_impl(h, head16, h->bb.feat16.f(), B, H16, W16, h->pose16.f(), st));
  HP_TRY(hp_head_forward_impl(h, head8, h->bb.feat8.f(), B, H8, W8, h->pose8.f(), st));
  HP_TRY(hp_decode_nms_impl(h, h->cls.f(), h->loc.f(), h->pose16.f(), h->pose8.f(), B, H, W, logit_thr, iou_thr, max_out,
                            (int32_t*)d.cnt.p, (int32_t*)d.anchor.p, (double*)d.boxes.p, (double*)d.kps.p, d.scores.f(), d.poses.f(), st));
  int32_t* hdr = (int32_t*)result;
  hp_face* faces = (hp_face*)((char*)result + hp_detect_result_faces_offset(B));
  pack_offsets_kernel<<<1, 1024, 0, st>>>((const int32_t*)d.cnt.p, B, cap, hdr, (int32_t*)d.offsets.p);
  pack_faces_kernel<<<B, 128, 0, st>>>((const int32_t*)d.cnt.p, (const int32_t*)d.offsets.p, (const int32_t*)d.anchor.p,
                                       (const double*)d.boxes.p, (const double*)d.kps.p, d.scores.f(), d.poses.f(), max_out, cap, faces);
  h->launches += 2;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

}  // namespace

void hp_detect_graphs_free(hp_ctx* h) {
  for (DetectGraph& g : h->det_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->det_graphs.clear();
}

extern "C" {

size_t hp_detect_result_faces_offset(int B) { return ((size_t)(HP_RESULT_HEADER_INTS + B) * sizeof(int32_t) + 7) / 8 * 8; }
size_t hp_detect_result_bytes(int B, int cap) { return hp_detect_result_faces_offset(B) + (size_t)cap * sizeof(hp_face); }

int hp_detect_frames(hp_handle h, hp_head_t head16, hp_head_t head8, const uint8_t* frames_bgr, int B, int Hin, int Win, int H, int W,
                     float logit_thr, float iou_thr, int max_out, void* result, size_t result_bytes, int flags, void* stream) {
  HP_REQUIRE(h != nullptr, HP_ERR_INVALID, "null handle");
  HP_CUDA(cudaSetDevice(h->device));
  static_assert(sizeof(hp_face) == 152, "hp_face is 19 8-byte words");
  HP_REQUIRE(head16 && head8 && frames_bgr && result, HP_ERR_INVALID, "hp_detect_frames: null argument");
  HP_REQUIRE(B > 0 && Hin > 0 && Win > 0 && H > 0 && W > 0, HP_ERR_INVALID, "hp_detect_frames: bad sizes");
  HP_REQUIRE(max_out > 0 && max_out <= HP_MAX_FACES, HP_ERR_INVALID, "hp_detect_frames: max_out must be in 1..%d", HP_MAX_FACES);
  HP_REQUIRE(hp_head_in_channels(head16) == 88 && hp_head_in_channels(head8) == 96 && hp_head_out_channels(head16) == 3 &&
                 hp_head_out_channels(head8) == 3,
             HP_ERR_INVALID, "hp_detect_frames: head16 must map 88 -> 3 channels and head8 96 -> 3");
  const size_t faces_off = hp_detect_result_faces_offset(B);
  HP_REQUIRE(result_bytes >= faces_off + sizeof(hp_face), HP_ERR_INVALID, "hp_detect_frames: result buffer of %zu bytes is too small", result_bytes);
  HP_REQUIRE(((uintptr_t)result & 7) == 0, HP_ERR_INVALID, "hp_detect_frames: result must be 8-byte aligned");
  long long cap64 = (long long)((result_bytes - faces_off) / sizeof(hp_face));
  const int cap = cap64 > (long long)B * max_out ? B * max_out : (int)cap64;
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & HP_DETECT_GRAPH))
    return detect_launches(h, head16, head8, frames_bgr, B, Hin, Win, H, W, logit_thr, iou_thr, max_out, result, cap, st);

  // ---- graph replay: key = everything the captured launches depend on
  DetectGraph key;
  key.head16 = head16; key.head8 = head8; key.frames = frames_bgr; key.result = result;
  key.B = B; key.Hin = Hin; key.Win = Win; key.H = H; key.W = W; key.max_out = max_out; key.cap = cap;
  key.logit_thr = logit_thr; key.iou_thr = iou_thr; key.impl = h->impl; key.chain_mode = h->chain_mode;
  DetectGraph* hit = nullptr;
  for (DetectGraph& g : h->det_graphs)
    if (g.same_key(key)) hit = &g;
  if (hit && hit->epoch != g_devbuf_epoch) {           // a buffer moved since the capture: the graph holds stale pointers
    hp_detect_graphs_free(h);
    hit = nullptr;
  }
  if (hit && hit->exec) {
    HP_CUDA(cudaGraphLaunch(hit->exec, st));
    h->launches += hit->launches;
    return HP_OK;
  }
  if (!hit) {
    // first call with this key: plain launches (sizes every internal buffer, sets kernel attributes, builds resize plans)
    HP_TRY(detect_launches(h, head16, head8, frames_bgr, B, Hin, Win, H, W, logit_thr, iou_thr, max_out, result, cap, st));
    if (h->det_graphs.size() >= 8) hp_detect_graphs_free(h);
    key.epoch = g_devbuf_epoch;
    h->det_graphs.push_back(key);
    return HP_OK;
  }
  // second call: capture, instantiate, launch
  HP_REQUIRE(st != nullptr, HP_ERR_INVALID, "hp_detect_frames: graph mode needs a non-default stream (the legacy stream cannot be captured)");
  const int64_t before = h->launches;
  HP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = detect_launches(h, head16, head8, frames_bgr, B, Hin, Win, H, W, logit_thr, iou_thr, max_out, result, cap, st);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != HP_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  HP_REQUIRE(ce == cudaSuccess && graph != nullptr, HP_ERR_CUDA, "hp_detect_frames: stream capture failed: %s", cudaGetErrorString(ce));
  HP_REQUIRE(hit->epoch == g_devbuf_epoch, HP_ERR_STATE, "hp_detect_frames: a buffer was reallocated during capture");
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  HP_REQUIRE(ie == cudaSuccess, HP_ERR_CUDA, "hp_detect_frames: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
  hit->exec = exec;
  hit->launches = (int)(h->launches - before);
  HP_CUDA(cudaGraphLaunch(exec, st));
  return HP_OK;
}

}  // extern "C"
