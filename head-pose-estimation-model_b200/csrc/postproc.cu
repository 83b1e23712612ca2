// Detector post-processing on device: logit threshold + sigmoid, SSD anchor decode, TensorFlow-
// compatible greedy hard NMS and pose lookup -- one CTA per image, warp ballot/shuffle selection.
//
// Reference: BlazePoser/blazeFaceDetectorH5.py:319-327 (filterDetections), :284-317
// (extractDetections, float64 arithmetic), :329-357 (tf.image.non_max_suppression(boxes, scores,
// 100, iou) + pose lookup by anchor cell) and BlazePoser/blazeFaceUtils.py:59-127 (anchor centres
// (x+0.5)/grid with fixed_anchor_size).  NMS semantics: SURVEY.md Appendix B.6.
//
// Bit-exactness: every float/double operation that NumPy/TensorFlow perform is issued with an
// explicit round-to-nearest intrinsic so that nvcc cannot contract mul+add into FMA.
#include "common.cuh"

// ---- float32 exp shared with oracle/postproc.py::exp32 (Cephes expf, rn mul/add only)
__device__ __forceinline__ float hp_exp32(float x) {
  x = fminf(fmaxf(x, -87.0f), 88.0f);
  const float k = rintf(__fmul_rn(x, 1.4426950408889634f));
  float r = __fsub_rn(x, __fmul_rn(k, 0.693359375f));
  r = __fsub_rn(r, __fmul_rn(k, -2.12194440e-4f));
  float p = 1.9875691500e-4f;
  p = __fadd_rn(__fmul_rn(p, r), 1.3981999507e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 8.3334519073e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 4.1665795894e-2f);
  p = __fadd_rn(__fmul_rn(p, r), 1.6666665459e-1f);
  p = __fadd_rn(__fmul_rn(p, r), 5.0000001201e-1f);
  const float r2 = __fmul_rn(r, r);
  float y = __fmul_rn(p, r2);
  y = __fadd_rn(y, r);
  y = __fadd_rn(y, 1.0f);
  const int ki = (int)k;
  return __fmul_rn(y, __int_as_float((ki + 127) << 23));
}
__device__ __forceinline__ float hp_sigmoid32(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, hp_exp32(-x)));
}

// TensorFlow NonMaxSuppression IOU<float>
__device__ __forceinline__ float iou32(float4 a, float4 b) {
  const float ymin_i = fminf(a.x, a.z), xmin_i = fminf(a.y, a.w), ymax_i = fmaxf(a.x, a.z), xmax_i = fmaxf(a.y, a.w);
  const float ymin_j = fminf(b.x, b.z), xmin_j = fminf(b.y, b.w), ymax_j = fmaxf(b.x, b.z), xmax_j = fmaxf(b.y, b.w);
  const float area_i = __fmul_rn(__fsub_rn(ymax_i, ymin_i), __fsub_rn(xmax_i, xmin_i));
  const float area_j = __fmul_rn(__fsub_rn(ymax_j, ymin_j), __fsub_rn(xmax_j, xmin_j));
  if (area_i <= 0.f || area_j <= 0.f) return 0.f;
  const float iy0 = fmaxf(ymin_i, ymin_j), ix0 = fmaxf(xmin_i, xmin_j);
  const float iy1 = fminf(ymax_i, ymax_j), ix1 = fminf(xmax_i, xmax_j);
  const float inter = __fmul_rn(fmaxf(__fsub_rn(iy1, iy0), 0.f), fmaxf(__fsub_rn(ix1, ix0), 0.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_i, area_j), inter));
}

// iou32(a, b) > thr for boxes already normalised to (ymin, xmin, ymax, xmax) with their areas (the same float operations
// as iou32, hoisted per box), without the division in the common cases: boxes that do not intersect give exactly 0;
// otherwise the quotient is only needed when inter is within 1e-6 (relative) of thr * union -- far more than the two
// roundings involved -- so the decision is bit-identical to comparing the correctly rounded quotient.
__device__ __forceinline__ float4 box_normalise(float4 a, float* area) {
  const float4 n = make_float4(fminf(a.x, a.z), fminf(a.y, a.w), fmaxf(a.x, a.z), fmaxf(a.y, a.w));
  *area = __fmul_rn(__fsub_rn(n.z, n.x), __fsub_rn(n.w, n.y));
  return n;
}
__device__ __forceinline__ bool iou32_gt(float4 a, float area_a, float4 b, float area_b, float thr) {
  if (area_a <= 0.f || area_b <= 0.f) return 0.f > thr;
  const float iy0 = fmaxf(a.x, b.x), ix0 = fmaxf(a.y, b.y);
  const float iy1 = fminf(a.z, b.z), ix1 = fminf(a.w, b.w);
  const float inter = __fmul_rn(fmaxf(__fsub_rn(iy1, iy0), 0.f), fmaxf(__fsub_rn(ix1, ix0), 0.f));
  if (inter <= 0.f) return 0.f > thr;                          // 0 / union
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  if (thr < 0.f || !(uni > 0.f)) return __fdiv_rn(inter, uni) > thr;
  const float t = __fmul_rn(thr, uni);
  if (inter > __fmul_rn(t, 1.000001f)) return true;
  if (inter < __fmul_rn(t, 0.999999f)) return false;
  return __fdiv_rn(inter, uni) > thr;
}

struct NmsParams {
  const float* cls;
  const float* loc;
  const float* pose16;
  const float* pose8;
  int B, A, A16, H16, W16, H8, W8, in_h, in_w;
  float logit_thr, iou_thr;
  int max_out, n2;
  int32_t* out_cnt;
  int32_t* out_anchor;
  double* boxes;
  double* kps;
  float* scores;
  float* poses;
};

__device__ __forceinline__ void anchor_centre(const NmsParams& p, int a, double* ax, double* ay) {
  int cell, gw, gh;
  if (a < p.A16) { cell = a >> 1; gw = p.W16; gh = p.H16; }
  else { cell = (a - p.A16) / 6; gw = p.W8; gh = p.H8; }
  const int x = cell % gw, y = cell / gw;
  *ax = __ddiv_rn(__dmul_rn((double)x + 0.5, 1.0), (double)gw);
  *ay = __ddiv_rn(__dmul_rn((double)y + 0.5, 1.0), (double)gh);
}
// (v + a*size)/size in float64, exactly as NumPy evaluates it
__device__ __forceinline__ double decode_centre(float v, double a, double size) {
  return __ddiv_rn(__dadd_rn((double)v, __dmul_rn(a, size)), size);
}
__device__ __forceinline__ void decode_box64(const NmsParams& p, const float* l, int a, double* b4) {
  double ax, ay;
  anchor_centre(p, a, &ax, &ay);
  const double cx = decode_centre(l[0], ax, (double)p.in_w), cy = decode_centre(l[1], ay, (double)p.in_h);
  const double w = __ddiv_rn((double)l[2], (double)p.in_w), hh = __ddiv_rn((double)l[3], (double)p.in_h);
  const double hw = __dmul_rn(w, 0.5), hhh = __dmul_rn(hh, 0.5);
  b4[0] = __dsub_rn(cx, hw);
  b4[1] = __dsub_rn(cy, hhh);
  b4[2] = __dadd_rn(cx, hw);
  b4[3] = __dadd_rn(cy, hhh);
}

__global__ void __launch_bounds__(128) decode_nms_kernel(NmsParams p) {
  extern __shared__ unsigned long long keys[];  // [n2]
  __shared__ int s_count;
  __shared__ int s_nsel;
  __shared__ int s_sel_anchor[HP_MAX_FACES];
  __shared__ float s_sel_score[HP_MAX_FACES];
  __shared__ float4 s_sel_box[HP_MAX_FACES];
  __shared__ float s_sel_area[HP_MAX_FACES];
  __shared__ float s_carea[128];
  __shared__ float4 s_cbox[128];          // the chunk of candidates being resolved: normalised box (+ area), anchor, score
  __shared__ int s_canchor[128];
  __shared__ float s_cscore[128];
  __shared__ unsigned s_mask[128][4];     // row i: later candidates of the chunk that box i suppresses
  __shared__ unsigned s_alive[4];         // candidates of the chunk that no earlier selection suppresses
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* cls = p.cls + (long long)b * p.A;
  const float* loc = p.loc + (long long)b * p.A * 16;
  if (tid == 0) { s_count = 0; s_nsel = 0; }
  for (int i = tid; i < p.n2; i += 128) keys[i] = 0ull;
  __syncthreads();
  // ---- filterDetections: logit > threshold, score = sigmoid (float32); unordered compaction, the
  //      sort key carries the anchor id for TensorFlow's lower-index-first tie break
  for (int a = tid; a < p.A; a += 128) {
    const float v = cls[a];
    if (v > p.logit_thr) {
      const float s = hp_sigmoid32(v);
      const int slot = atomicAdd(&s_count, 1);
      keys[slot] = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a);
    }
  }
  __syncthreads();
  const int n = s_count;
  // ---- bitonic sort, descending
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  // one compare-exchange PAIR per thread and step (i = the pair's lower index, partner i | j): no idle half of the threads
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n2 >> 1); t += 128) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const unsigned long long x = keys[i], y = keys[ixj];
        const bool desc = ((i & k) == 0);
        if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[ixj] = x; }
      }
      __syncthreads();
    }
  }
  // ---- greedy selection (tf.image.non_max_suppression order), 128 sorted candidates at a time:
  //  1. every thread decodes one candidate and tests it against the boxes selected in earlier chunks      (parallel)
  //  2. every surviving candidate i builds the bit row "candidate j > i of this chunk overlaps me"          (parallel)
  //  3. warp 0 walks the chunk in order with the removed-bits in lanes 0..3: a surviving candidate is selected and ORs its row
  //     in -- a few instructions per candidate instead of an IoU loop per candidate on one warp
  // A candidate is selected iff no earlier SELECTED box overlaps it: the same decisions, IoU arguments (candidate, selected).
  {
    const int lane = tid & 31, warp = tid >> 5;
    for (int base = 0; base < n && s_nsel < p.max_out; base += 128) {
      const int nsel0 = s_nsel;
      const int pos = base + tid;
      const int cn = (n - base < 128) ? n - base : 128;          // candidates in this chunk
      bool alive = tid < cn;
      float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
      float area = 0.f;
      if (alive) {
        const unsigned long long key = keys[pos];
        const int a = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        double b4[4];
        decode_box64(p, loc + (long long)a * 16, a, b4);
        bx = box_normalise(make_float4((float)b4[0], (float)b4[1], (float)b4[2], (float)b4[3]), &area);  // tf casts boxes to float32
        s_canchor[tid] = a;
        s_cscore[tid] = __uint_as_float((unsigned)(key >> 32));
        for (int j = 0; j < nsel0; ++j)
          if (iou32_gt(bx, area, s_sel_box[j], s_sel_area[j], p.iou_thr)) { alive = false; break; }
      }
      s_cbox[tid] = bx;
      s_carea[tid] = area;
      const unsigned am = __ballot_sync(0xffffffffu, alive);
      if (lane == 0) s_alive[warp] = am;
      __syncthreads();
      {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          unsigned word = 0u;
          if (alive && w * 32 + 31 > tid) {
            const unsigned al = s_alive[w];
            for (int jj = 0; jj < 32; ++jj) {
              const int j = w * 32 + jj;
              if (j > tid && j < cn && ((al >> jj) & 1u) && iou32_gt(s_cbox[j], s_carea[j], bx, area, p.iou_thr)) word |= 1u << jj;
            }
          }
          s_mask[tid][w] = word;
        }
      }
      __syncthreads();
      if (warp == 0) {
        unsigned rem = lane < 4 ? ~s_alive[lane] : 0xffffffffu;
        int cnt = nsel0;
        for (int i = 0; i < cn && cnt < p.max_out; ++i) {
          const unsigned w = __shfl_sync(0xffffffffu, rem, i >> 5);
          if (!((w >> (i & 31)) & 1u)) {
            if (lane == 0) {
              s_sel_anchor[cnt] = s_canchor[i];
              s_sel_score[cnt] = s_cscore[i];
              s_sel_box[cnt] = s_cbox[i];
              s_sel_area[cnt] = s_carea[i];
            }
            if (lane < 4) rem |= s_mask[i][lane];
            ++cnt;
          }
        }
        if (lane == 0) s_nsel = cnt;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  const int nsel = s_nsel;
  if (tid == 0) p.out_cnt[b] = nsel;
  // ---- gather outputs (selection order)
  for (int k = tid; k < p.max_out; k += 128) {
    const long long o = (long long)b * p.max_out + k;
    if (k >= nsel) { p.out_anchor[o] = -1; continue; }
    const int a = s_sel_anchor[k];
    p.out_anchor[o] = a;
    if (p.scores) p.scores[o] = s_sel_score[k];
    const float* l = loc + (long long)a * 16;
    if (p.boxes) {
      double b4[4];
      decode_box64(p, l, a, b4);
      for (int j = 0; j < 4; ++j) p.boxes[o * 4 + j] = b4[j];
    }
    if (p.kps) {
      double ax, ay;
      anchor_centre(p, a, &ax, &ay);
      for (int j = 0; j < HP_KEYPOINTS; ++j) {
        p.kps[o * 12 + 2 * j + 0] = decode_centre(l[4 + 2 * j], ax, (double)p.in_w);
        p.kps[o * 12 + 2 * j + 1] = decode_centre(l[5 + 2 * j], ay, (double)p.in_h);
      }
    }
    if (p.poses) {
      const float* src;
      if (a < p.A16) src = p.pose16 + ((long long)b * p.H16 * p.W16 + (a >> 1)) * 3;
      else src = p.pose8 + ((long long)b * p.H8 * p.W8 + (a - p.A16) / 6) * 3;
      p.poses[o * 3 + 0] = src[0];
      p.poses[o * 3 + 1] = src[1];
      p.poses[o * 3 + 2] = src[2];
    }
  }
}

int hp_decode_nms_impl(hp_ctx* h, const float* cls, const float* loc, const float* pose16, const float* pose8, int B,
                       int H, int W, float logit_thr, float iou_thr, int max_out, int32_t* out_cnt,
                       int32_t* out_anchor, double* boxes, double* kps, float* scores, float* poses,
                       cudaStream_t st) {
  HP_REQUIRE(cls && loc && out_cnt && out_anchor && B > 0, HP_ERR_INVALID, "hp_decode_nms: bad arguments");
  HP_REQUIRE(max_out > 0 && max_out <= HP_MAX_FACES, HP_ERR_INVALID, "hp_decode_nms: max_out must be in 1..%d", HP_MAX_FACES);
  HP_REQUIRE(!poses || (pose16 && pose8), HP_ERR_INVALID, "hp_decode_nms: poses requested without pose maps");
  NmsParams p;
  p.cls = cls; p.loc = loc; p.pose16 = pose16; p.pose8 = pose8;
  p.B = B; p.in_h = H; p.in_w = W;
  p.H16 = ceil_div(H, 8); p.W16 = ceil_div(W, 8); p.H8 = ceil_div(H, 16); p.W8 = ceil_div(W, 16);
  p.A16 = p.H16 * p.W16 * 2; p.A = p.A16 + p.H8 * p.W8 * 6;
  p.logit_thr = logit_thr; p.iou_thr = iou_thr; p.max_out = max_out;
  int n2 = 1;
  while (n2 < p.A) n2 <<= 1;
  p.n2 = n2;
  HP_REQUIRE(n2 <= 16384, HP_ERR_UNSUPPORTED, "hp_decode_nms: %d anchors per image exceed the kernel limit", p.A);
  p.out_cnt = out_cnt; p.out_anchor = out_anchor; p.boxes = boxes; p.kps = kps; p.scores = scores; p.poses = poses;
  const size_t smem = (size_t)n2 * sizeof(unsigned long long);
  HP_CUDA(cudaFuncSetAttribute(decode_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  decode_nms_kernel<<<B, 128, smem, st>>>(p);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

// ============================================================================ step-wise entry points
// The reference exposes the three post-processing steps as separate methods; these kernels back them
// one to one so that each step can be checked on its own.

// filterDetections: ascending anchor order (np.where), float32 sigmoid.  One CTA per image.
__global__ void __launch_bounds__(128) filter_detections_kernel(const float* cls, int A, float logit_thr,
                                                                int32_t* out_idx, float* out_scores, int32_t* out_cnt) {
  __shared__ int s_warp[4];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* c = cls + (long long)b * A;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int a0 = 0; a0 < A; a0 += 128) {
    const int a = a0 + tid;
    const float v = (a < A) ? c[a] : 0.f;
    const bool good = (a < A) && (v > logit_thr);
    const unsigned m = __ballot_sync(0xffffffffu, good);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (good) {
      const int slot = off + __popc(m & ((1u << lane) - 1u));
      out_idx[(long long)b * A + slot] = a;
      out_scores[(long long)b * A + slot] = hp_sigmoid32(v);
    }
    __syncthreads();
    if (tid == 0) s_base += s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
    __syncthreads();
  }
  if (tid == 0) out_cnt[b] = s_base;
}

// extractDetections for a list of anchor ids of ONE image grid: boxes (n,4) / keypoints (n,6,2) float64
__global__ void extract_detections_kernel(NmsParams p, const float* loc, const int32_t* idx, int n, double* boxes,
                                          double* kps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = idx[i];
  const float* l = loc + (long long)a * 16;
  double b4[4];
  decode_box64(p, l, a, b4);
  for (int j = 0; j < 4; ++j) boxes[(long long)i * 4 + j] = b4[j];
  double ax, ay;
  anchor_centre(p, a, &ax, &ay);
  for (int j = 0; j < HP_KEYPOINTS; ++j) {
    kps[(long long)i * 12 + 2 * j + 0] = decode_centre(l[4 + 2 * j], ax, (double)p.in_w);
    kps[(long long)i * 12 + 2 * j + 1] = decode_centre(l[5 + 2 * j], ay, (double)p.in_h);
  }
}

// tf.image.non_max_suppression(boxes, scores, max_out, iou) on explicit boxes (float64 in, cast to
// float32 as TensorFlow does).  One CTA; returns indices into the candidate list in selection order.
__global__ void __launch_bounds__(128) nms_boxes_kernel(const double* boxes, const float* scores, int n, int n2,
                                                        float iou_thr, int max_out, int32_t* out_sel, int32_t* out_cnt) {
  extern __shared__ unsigned long long keys[];
  __shared__ float4 s_sel_box[HP_MAX_FACES];
  const int tid = threadIdx.x;
  for (int i = tid; i < n2; i += 128) {
    unsigned long long k = 0ull;
    if (i < n) k = ((unsigned long long)__float_as_uint(scores[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    keys[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += 128) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = keys[i], y = keys[ixj];
          const bool desc = ((i & k) == 0);
          if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
  if (tid < 32) {
    const int lane = tid;
    int nsel = 0;
    for (int base = 0; base < n && nsel < max_out; base += 32) {
      const int pos = base + lane;
      bool alive = pos < n;
      int ci = 0;
      float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
      if (alive) {
        ci = (int)(0xFFFFFFFFu - (unsigned)(keys[pos] & 0xFFFFFFFFull));
        const double* bp = boxes + (long long)ci * 4;
        bx = make_float4((float)bp[0], (float)bp[1], (float)bp[2], (float)bp[3]);
        for (int j = 0; j < nsel; ++j)
          if (iou32(bx, s_sel_box[j]) > iou_thr) { alive = false; break; }
      }
      unsigned mask = __ballot_sync(0xffffffffu, alive);
      while (mask != 0u && nsel < max_out) {
        const int leader = __ffs(mask) - 1;
        const float4 lb = make_float4(__shfl_sync(0xffffffffu, bx.x, leader), __shfl_sync(0xffffffffu, bx.y, leader),
                                      __shfl_sync(0xffffffffu, bx.z, leader), __shfl_sync(0xffffffffu, bx.w, leader));
        if (lane == leader) {
          out_sel[nsel] = ci;
          s_sel_box[nsel] = bx;
          alive = false;
        }
        nsel++;
        if (alive && lane > leader && iou32(bx, lb) > iou_thr) alive = false;
        mask = __ballot_sync(0xffffffffu, alive);
      }
      __syncwarp();
    }
    if (lane == 0) *out_cnt = nsel;
  }
}

static void fill_grid(NmsParams* p, int H, int W) {
  p->in_h = H; p->in_w = W;
  p->H16 = ceil_div(H, 8); p->W16 = ceil_div(W, 8); p->H8 = ceil_div(H, 16); p->W8 = ceil_div(W, 16);
  p->A16 = p->H16 * p->W16 * 2; p->A = p->A16 + p->H8 * p->W8 * 6;
}

extern "C" int hp_filter_detections(hp_handle h, const float* cls, int B, int A, float logit_thr, int32_t* out_idx,
                                    float* out_scores, int32_t* out_cnt, void* stream) {
  HP_REQUIRE(h && cls && out_idx && out_scores && out_cnt && B > 0 && A > 0, HP_ERR_INVALID, "hp_filter_detections: bad arguments");
  HP_CUDA(cudaSetDevice(h->device));
  filter_detections_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(cls, A, logit_thr, out_idx, out_scores, out_cnt);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

extern "C" int hp_extract_detections(hp_handle h, const float* loc, const int32_t* idx, int n, int H, int W,
                                     double* boxes, double* kps, void* stream) {
  HP_REQUIRE(h && n >= 0, HP_ERR_INVALID, "hp_extract_detections: bad arguments");
  if (n == 0) return HP_OK;
  HP_REQUIRE(loc && boxes && kps && idx, HP_ERR_INVALID, "hp_extract_detections: null pointer");
  HP_CUDA(cudaSetDevice(h->device));
  NmsParams p;
  memset(&p, 0, sizeof(p));
  fill_grid(&p, H, W);
  extract_detections_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(p, loc, idx, n, boxes, kps);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

extern "C" int hp_nms(hp_handle h, const double* boxes, const float* scores, int n, float iou_thr, int max_out,
                      int32_t* out_sel, int32_t* out_cnt, void* stream) {
  HP_REQUIRE(h && out_sel && out_cnt && n >= 0 && (n == 0 || (boxes && scores)), HP_ERR_INVALID, "hp_nms: bad arguments");
  HP_REQUIRE(max_out > 0 && max_out <= HP_MAX_FACES, HP_ERR_INVALID, "hp_nms: max_out must be in 1..%d", HP_MAX_FACES);
  HP_CUDA(cudaSetDevice(h->device));
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  HP_REQUIRE(n2 <= 16384, HP_ERR_UNSUPPORTED, "hp_nms: at most 16384 candidates");
  HP_CUDA(cudaFuncSetAttribute(nms_boxes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  nms_boxes_kernel<<<1, 128, (size_t)n2 * sizeof(unsigned long long), (cudaStream_t)stream>>>(boxes, scores, n, n2, iou_thr,
                                                                                           max_out, out_sel, out_cnt);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}

// ============================================================================ pre-processing
// prepareInputForInference (blazeFaceDetectorH5.py:247-269) without the resize: BGR->RGB,
// v/255.0 in float64 -> float32 (tf.image.resize output dtype) -> (t-0.5)/0.5 in float32.
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ bgr, float* __restrict__ x, long long npix) {
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const float t = (float)__ddiv_rn((double)i, 255.0);
    lut[i] = __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);
  }
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t b = bgr[i * 3 + 0], g = bgr[i * 3 + 1], r = bgr[i * 3 + 2];
    x[i * 3 + 0] = lut[r];
    x[i * 3 + 1] = lut[g];
    x[i * 3 + 2] = lut[b];
  }
}

int hp_preprocess_u8_impl(hp_ctx* h, const uint8_t* bgr, int B, int H, int W, float* x, cudaStream_t st) {
  HP_REQUIRE(bgr && x && B > 0 && H > 0 && W > 0, HP_ERR_INVALID, "hp_preprocess_u8: bad arguments");
  const long long npix = (long long)B * H * W;
  long long grid = (npix + 255) / 256;
  if (grid > (long long)h->num_sms * 16) grid = (long long)h->num_sms * 16;
  preprocess_u8_kernel<<<(unsigned)grid, 256, 0, st>>>(bgr, x, npix);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}
