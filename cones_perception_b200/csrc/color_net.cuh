// color_net.cuh — the reference's cone-colour classifier on the device.
//
// The reference evaluates models/dam_net/dam_net.tflite with the TFLite interpreter inside a Python ROS
// service (scripts/color_classifier_server.py:66-71 loads it, :108-120 runs it per cone): the 15x12x1 uint8
// range image is cast to float32 without scaling (:108), and the answer is argmax + 1 when the largest softmax
// output is >= 0.8, else 0 = unknown (:116-120).  The graph in that file is
//     CONV_2D 3x3 (1 -> 16, VALID, ReLU) -> MAX_POOL 2x2/2 -> CONV_2D 3x3 (16 -> 32, VALID, ReLU) -> MAX_POOL 2x2/2
//     -> MUL, ADD (folded batch normalisation, per channel) -> RESHAPE (flatten, NHWC) -> FULLY_CONNECTED (64 -> 3)
//     -> SOFTMAX
// cone_color_kernel evaluates exactly this architecture (channel counts are parameters up to the shared-memory
// budget below), one CTA per cone, chained behind cone_raster_kernel so that only one byte per cone travels back.
//
// Arithmetic: fp32, every accumulation in the order of TFLite's reference kernels (filter_y, filter_x,
// in_channel; then the bias), multiply and add rounded separately (no FMA contraction), which is also the order
// of the numpy checker (oracle/dam_net_ref.py) — logits are bit-identical to it; the softmax differs from numpy
// only by expf's last ulp.  75 k multiply-adds per cone: latency-bound, no tensor-core shape (K = 9 and 144).
#pragma once
#include "color_kernels.cuh"

namespace cp {

constexpr int kNetThreads = 128;
constexpr int kNetMaxC1 = 16, kNetMaxC2 = 32, kNetMaxClasses = 4;
constexpr int kConv1H = kImgRows - 2, kConv1W = kImgCols - 2;   // 13 x 10
constexpr int kPool1H = kConv1H / 2, kPool1W = kConv1W / 2;     // 6 x 5
constexpr int kConv2H = kPool1H - 2, kConv2W = kPool1W - 2;     // 4 x 3
constexpr int kPool2H = kConv2H / 2, kPool2W = kConv2W / 2;     // 2 x 1
constexpr u32 kConeLowConfidence = 16u;  // flag: the decision sits within 2e-6 of the 0.8 threshold

// Weights as laid out by cp_color_net_load (transposed so that consecutive threads read consecutive words):
//   w1t[9][c1]  b1[c1]  w2t[9 * c1][c2]  b2[c2]  scale[c2]  shift[c2]  wd[classes][kPool2H * kPool2W * c2]  bd[classes]
struct ColorNetDev {
  const float* w1t;
  const float* b1;
  const float* w2t;
  const float* b2;
  const float* scale;
  const float* shift;
  const float* wd;
  const float* bd;
  u32 c1, c2, classes;
  float threshold;
};

// grid = cones.  images [n][15][12] uint8; raster_flags (optional) = cone_raster_kernel's flags: a cone whose image
// could not be drawn (empty crop, numpy would raise) gets colour 255 and zero probabilities.
__global__ void __launch_bounds__(kNetThreads) cone_color_kernel(const uint8_t* __restrict__ images,
                                                                 const u32* __restrict__ raster_flags,
                                                                 const __grid_constant__ ColorNetDev net,
                                                                 uint8_t* __restrict__ colors,
                                                                 float* __restrict__ probs, float* __restrict__ logits_out,
                                                                 u32* __restrict__ flags_out) {
  __shared__ float s_in[kImgPix];
  __shared__ float s_a[kConv1H * kConv1W * kNetMaxC1];   // conv1 output, later conv2 output
  __shared__ float s_b[kPool1H * kPool1W * kNetMaxC1];   // pool1 output, later the flattened features
  __shared__ float s_logit[kNetMaxClasses];
  const u32 cone = blockIdx.x, tid = threadIdx.x;
  const u32 C1 = net.c1, C2 = net.c2, NC = net.classes;
  const u32 rf = raster_flags ? raster_flags[cone] : 0u;
  if (rf & ~kConeAmbiguous) {   // no image: the service skips (empty) or raises (bad index / intensity)
    if (tid == 0) {
      colors[cone] = 255;
      if (flags_out) flags_out[cone] = rf;
    }
    if (probs && tid < NC) probs[(size_t)cone * NC + tid] = 0.f;
    if (logits_out && tid < NC) logits_out[(size_t)cone * NC + tid] = 0.f;
    return;
  }
  for (u32 i = tid; i < kImgPix; i += kNetThreads) s_in[i] = (float)images[(size_t)cone * kImgPix + i];
  __syncthreads();
  // conv1 + ReLU: out[y][x][c], c fastest
  for (u32 o = tid; o < kConv1H * kConv1W * C1; o += kNetThreads) {
    const u32 c = o % C1, x = (o / C1) % kConv1W, y = o / (C1 * kConv1W);
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
        acc = __fadd_rn(acc, __fmul_rn(s_in[(y + ky) * kImgCols + x + kx], __ldg(net.w1t + (ky * 3 + kx) * C1 + c)));
    acc = __fadd_rn(acc, __ldg(net.b1 + c));
    s_a[o] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  // pool1 2x2 stride 2 (VALID: the odd last row of 13 is dropped)
  for (u32 o = tid; o < kPool1H * kPool1W * C1; o += kNetThreads) {
    const u32 c = o % C1, x = (o / C1) % kPool1W, y = o / (C1 * kPool1W);
    const float* p = s_a + ((2 * y) * kConv1W + 2 * x) * C1 + c;
    s_b[o] = fmaxf(fmaxf(p[0], p[C1]), fmaxf(p[kConv1W * C1], p[kConv1W * C1 + C1]));
  }
  __syncthreads();
  // conv2 + ReLU into s_a
  for (u32 o = tid; o < kConv2H * kConv2W * C2; o += kNetThreads) {
    const u32 c = o % C2, x = (o / C2) % kConv2W, y = o / (C2 * kConv2W);
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const float* in = s_b + ((y + ky) * kPool1W + x + kx) * C1;
        const float* w = net.w2t + (size_t)((ky * 3 + kx) * C1) * C2 + c;
        for (u32 ci = 0; ci < C1; ++ci) acc = __fadd_rn(acc, __fmul_rn(in[ci], __ldg(w + (size_t)ci * C2)));
      }
    acc = __fadd_rn(acc, __ldg(net.b2 + c));
    s_a[o] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  // pool2 (4x3 -> 2x1: the third column is dropped) + per-channel scale, shift; flatten in NHWC order
  for (u32 o = tid; o < kPool2H * kPool2W * C2; o += kNetThreads) {
    const u32 c = o % C2, y = o / C2;
    const float* p = s_a + ((2 * y) * kConv2W) * C2 + c;
    const float m = fmaxf(fmaxf(p[0], p[C2]), fmaxf(p[kConv2W * C2], p[kConv2W * C2 + C2]));
    s_b[o] = __fadd_rn(__fmul_rn(m, __ldg(net.scale + c)), __ldg(net.shift + c));
  }
  __syncthreads();
  if (tid < NC) {
    const u32 K = kPool2H * kPool2W * C2;
    float acc = 0.f;
    for (u32 k = 0; k < K; ++k) acc = __fadd_rn(acc, __fmul_rn(s_b[k], __ldg(net.wd + (size_t)tid * K + k)));
    s_logit[tid] = __fadd_rn(acc, __ldg(net.bd + tid));
  }
  __syncthreads();
  if (tid == 0) {
    float m = s_logit[0];
    for (u32 j = 1; j < NC; ++j) m = fmaxf(m, s_logit[j]);
    float e[kNetMaxClasses], s = 0.f;
    for (u32 j = 0; j < NC; ++j) {
      e[j] = expf(__fsub_rn(s_logit[j], m));
      s = __fadd_rn(s, e[j]);
    }
    float best = -1.f;
    u32 arg = 0;
    for (u32 j = 0; j < NC; ++j) {
      const float p = __fdiv_rn(e[j], s);
      if (probs) probs[(size_t)cone * NC + j] = p;
      if (logits_out) logits_out[(size_t)cone * NC + j] = s_logit[j];
      if (p > best) {   // np.argmax: first maximum
        best = p;
        arg = j;
      }
    }
    colors[cone] = best >= net.threshold ? (uint8_t)(arg + 1) : (uint8_t)0;   // :116-120
    if (flags_out) flags_out[cone] = rf | (fabsf(best - net.threshold) < 2e-6f ? kConeLowConfidence : 0u);
  }
}

}  // namespace cp
