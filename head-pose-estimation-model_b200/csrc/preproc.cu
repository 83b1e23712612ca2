// Pre-processing with resize: prepareInputForInference (BlazePoser/blazeFaceDetectorH5.py:247-269) for uint8 BGR
// frames of ANY size: BGR -> RGB, v / 255.0 (double, rounded to float where TensorFlow's kernel casts its input), bicubic resize
// to the network input size, (t - 0.5) / 0.5 in float32.
//
// The resize is TensorFlow's ResizeBicubic CPU kernel with half_pixel_centers = true (tf.image.resize(method='bicubic'),
// antialias off; un-vendored dependency tensorflow>=2.8.0): Keys cubic A = -0.5 read from a 1025-entry coefficient table, taps
// outside the image dropped and the rest renormalised, the 4 rows combined first and then the 4 columns, all in float32 with
// separate multiplies and adds.  The tap tables (4 clamped indices + 4 weights per output row / column) are computed on the
// host in exactly that arithmetic, once per (input size, output size), and kept on the device; oracle/preprocess.py restates
// the same algorithm in numpy and the two agree bit for bit.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kTableSize = 1 << 10;

// the kernel's InitCoeffsTable(const double a): polynomials in double, stored as float
const float* coeff_table() {
  static float table[(kTableSize + 1) * 2];
  static bool ready = false;
  if (!ready) {
    const double a = -0.5;
    for (int i = 0; i <= kTableSize; ++i) {
      float x = (float)(i * 1.0 / kTableSize);
      table[i * 2] = (float)(((a + 2) * x - (a + 3)) * x * x + 1);
      x += 1.0f;
      table[i * 2 + 1] = (float)(((a * x - 5 * a) * x + 8 * a) * x - 4 * a);
    }
    ready = true;
  }
  return table;
}

// GetWeightsAndIndices<HalfPixelScaler, use_keys_cubic = true> for every output index; every intermediate is a separate float
// (no contraction: the reference binary has none either)
void make_taps(int in_size, int out_size, int* idx, float* wgt) {
  const float* tab = coeff_table();
  const float scale = (float)in_size / (float)out_size;
  for (int o = 0; o < out_size; ++o) {
    const float shifted = (float)o + 0.5f;
    const float scaled = shifted * scale;
    const float loc_f = scaled - 0.5f;
    const long loc = (long)std::floor(loc_f);
    const float delta = loc_f - (float)loc;
    const float scaled_delta = delta * (float)kTableSize;
    const long off = lrintf(scaled_delta);
    const long cand[4] = {loc - 1, loc, loc + 1, loc + 2};
    const float raw[4] = {tab[off * 2 + 1], tab[off * 2], tab[(kTableSize - off) * 2], tab[(kTableSize - off) * 2 + 1]};
    float w[4];
    for (int k = 0; k < 4; ++k) {
      long b = cand[k] < 0 ? 0 : (cand[k] > in_size - 1 ? in_size - 1 : cand[k]);
      idx[o * 4 + k] = (int)b;
      w[k] = (b == cand[k]) ? raw[k] : 0.0f;
    }
    const float s01 = w[0] + w[1];
    const float s012 = s01 + w[2];
    const float sum = s012 + w[3];
    if (std::fabs(sum) >= 1000.0f * 1.17549435e-38f) {
      const float inv = 1.0f / sum;
      for (int k = 0; k < 4; ++k) w[k] = w[k] * inv;
    }
    for (int k = 0; k < 4; ++k) wgt[o * 4 + k] = w[k];
  }
}

// one thread per output pixel; blockIdx.y = output row, blockIdx.z = image
__global__ void __launch_bounds__(128)
resize_bicubic_u8_kernel(const uint8_t* __restrict__ bgr, float* __restrict__ x, int Hin, int Win, int Hout, int Wout,
                         const int* __restrict__ yidx, const float* __restrict__ ywgt, const int* __restrict__ xidx,
                         const float* __restrict__ xwgt) {
  __shared__ float lut[256];                    // (float)(v / 255.0): the double quotient of :254, cast where the kernel reads it
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = (float)__ddiv_rn((double)i, 255.0);
  __syncthreads();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (ox >= Wout) return;
  const size_t img = blockIdx.z;
  const uint8_t* src = bgr + img * (size_t)Hin * Win * 3;
  const int4 yi = *reinterpret_cast<const int4*>(yidx + oy * 4);
  const float4 yw = *reinterpret_cast<const float4*>(ywgt + oy * 4);
  const int4 xi = *reinterpret_cast<const int4*>(xidx + ox * 4);
  const float4 xw = *reinterpret_cast<const float4*>(xwgt + ox * 4);
  const int xs[4] = {xi.x, xi.y, xi.z, xi.w};
  const uint8_t* r0 = src + (size_t)yi.x * Win * 3;
  const uint8_t* r1 = src + (size_t)yi.y * Win * 3;
  const uint8_t* r2 = src + (size_t)yi.z * Win * 3;
  const uint8_t* r3 = src + (size_t)yi.w * Win * 3;
  float col[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int o = xs[k] * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // Interpolate1D over the 4 rows: v0 w0 + v1 w1 + v2 w2 + v3 w3, left to right
      float acc = __fmul_rn(lut[r0[o + c]], yw.x);
      acc = __fadd_rn(acc, __fmul_rn(lut[r1[o + c]], yw.y));
      acc = __fadd_rn(acc, __fmul_rn(lut[r2[o + c]], yw.z));
      acc = __fadd_rn(acc, __fmul_rn(lut[r3[o + c]], yw.w));
      col[k][c] = acc;
    }
  }
  float* dst = x + ((img * Hout + oy) * (size_t)Wout + ox) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = __fmul_rn(col[0][c], xw.x);
    acc = __fadd_rn(acc, __fmul_rn(col[1][c], xw.y));
    acc = __fadd_rn(acc, __fmul_rn(col[2][c], xw.z));
    acc = __fadd_rn(acc, __fmul_rn(col[3][c], xw.w));
    dst[2 - c] = __fdiv_rn(__fsub_rn(acc, 0.5f), 0.5f);          // BGR -> RGB, (t - 0.5) / 0.5
  }
}

}  // namespace

void hp_resize_plans_free(hp_ctx* h) {
  for (ResizePlan& p : h->resize_plans) cudaFree(p.dev);
  h->resize_plans.clear();
}

int hp_preprocess_resize_u8_impl(hp_ctx* h, const uint8_t* bgr, int B, int Hin, int Win, int Hout, int Wout, float* x, cudaStream_t st) {
  HP_REQUIRE(bgr && x && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, HP_ERR_INVALID, "hp_preprocess_resize_u8: bad arguments");
  HP_REQUIRE(B <= 65535 && Hout <= 65535, HP_ERR_INVALID, "hp_preprocess_resize_u8: at most 65535 images / output rows per call");
  const ResizePlan* plan = nullptr;
  for (const ResizePlan& p : h->resize_plans)
    if (p.hin == Hin && p.win == Win && p.hout == Hout && p.wout == Wout) plan = &p;
  if (!plan) {
    if (h->resize_plans.size() >= 16) hp_resize_plans_free(h);   // frames of a stream keep their size: a tiny cache is enough
    std::vector<int> idx((size_t)(Hout + Wout) * 4);
    std::vector<float> wgt((size_t)(Hout + Wout) * 4);
    make_taps(Hin, Hout, idx.data(), wgt.data());
    make_taps(Win, Wout, idx.data() + (size_t)Hout * 4, wgt.data() + (size_t)Hout * 4);
    ResizePlan p;
    p.hin = Hin; p.win = Win; p.hout = Hout; p.wout = Wout;
    const size_t n = (size_t)(Hout + Wout) * 4;
    HP_CUDA(cudaStreamSynchronize(st));                            // an earlier launch may still read a plan that was just evicted
    HP_CUDA(cudaMalloc(&p.dev, n * 8));
    HP_CUDA(cudaMemcpy(p.dev, idx.data(), n * 4, cudaMemcpyHostToDevice));
    HP_CUDA(cudaMemcpy((char*)p.dev + n * 4, wgt.data(), n * 4, cudaMemcpyHostToDevice));
    h->resize_plans.push_back(p);
    plan = &h->resize_plans.back();
  }
  const size_t n = (size_t)(Hout + Wout) * 4;
  const int* yidx = (const int*)plan->dev;
  const int* xidx = yidx + (size_t)Hout * 4;
  const float* ywgt = (const float*)((const char*)plan->dev + n * 4);
  const float* xwgt = ywgt + (size_t)Hout * 4;
  dim3 grid((unsigned)((Wout + 127) / 128), (unsigned)Hout, (unsigned)B);
  resize_bicubic_u8_kernel<<<grid, 128, 0, st>>>(bgr, x, Hin, Win, Hout, Wout, yidx, ywgt, xidx, xwgt);
  h->launches++;
  HP_CUDA(cudaGetLastError());
  return HP_OK;
}
