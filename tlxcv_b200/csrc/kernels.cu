// Memory-bound CUDA-core kernels of the CNN-backbone forward path (NHWC, 128-bit vectorised along
// C), the one-time weight/BN preparation kernels, and the fp32 direct convolution used by the
// fp32 validation mode.  Each replaces one tensorlayerx layer of the reference hot path:
//   import/export_nchw  the NCHW fp32 tensors the reference passes in / gets back (layout pass, K7)
//   maxpool_nhwc        nn.MaxPool2d(3,2,1)       classification/resnet.py:213-218, resnext.py:159-164
//   gap_nhwc            nn.AdaptiveAvgPool2d(1)   resnet.py:227-231
//   dwconv_nhwc         depthwise GroupConv2d + BN + ReLU6/ReLU  ops/ops_fusion.py:39-48, mobilenetv1.py:81-89
//   argmax_rows         tlx.argmax(axis=-1)       tasks/image_classification.py:23
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace tlxcv {

namespace {

constexpr int kThreads = 256;

// forward kernels are launched with the programmatic-dependent-launch attribute (common.cuh); the macro keeps
// template commas inside the kernel name out of the argument list
#define TLXCV_LAUNCH(kernel, grid, block, smem, st, ...)                                              \
  do {                                                                                              \
    cudaError_t _le = launch_pdl((kernel), dim3(grid), dim3(block), (smem), (st), __VA_ARGS__);     \
    if (_le != cudaSuccess) return _le;                                                             \
  } while (0)

inline int blocks_for(size_t work, int threads = kThreads) {
  size_t b = (work + threads - 1) / threads;
  return static_cast<int>(b < 1 ? 1 : (b > 0x7fffffffull ? 0x7fffffffull : b));
}

// ------------------------------------------------------------------------------------------------
// weight / BN preparation (run once per plan)
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout,
                                         int Cout_pad, int Cin, int R, int S, int groups, int mode, int Ktot) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<size_t>(Cout_pad) * Ktot) return;
  const int o = static_cast<int>(idx / Ktot), k = static_cast<int>(idx % Ktot);
  const int Cg = Cin / groups;
  float v = 0.0f;
  if (o < Cout) {
    if (mode == kModeGatherC4) {
      const int KR = (S * 4 <= 16) ? 16 : 32;
      const int r_per_kb = 64 / KR;
      const int kb = k / 64, within = k % 64;
      const int r = kb * r_per_kb + within / KR, e = within % KR;
      const int s = e / 4, c = e % 4;
      if (r < R && s < S && c < Cin) v = w[((static_cast<size_t>(o) * Cin + c) * R + r) * S + s];
    } else if (mode == kModePixelPairs) {
      // K block j < 3: filter row {1, 2, 0}[j], taps s = 1 | 2 in the two halves; j >= 3: same rows, tap s = 0 in the UPPER half
      // (the pair to the left holds pixel 2*ox - 1 in its upper 32 channels), lower half zero
      const int j = k / 64, half = (k % 64) / 32, c = k % 32;
      const int r = (j % 3 == 0) ? 1 : (j % 3 == 1 ? 2 : 0);
      if (j < 3)
        v = w[((static_cast<size_t>(o) * Cin + c) * R + r) * S + 1 + half];
      else if (half == 1)
        v = w[((static_cast<size_t>(o) * Cin + c) * R + r) * S + 0];
    } else if (groups > 1) {
      // 64-channel block-diagonal expansion: K slot (tap, cl) holds input channel 64*(o/64)+cl
      const int tap = k / 64, cl = k % 64;
      const int r = tap / S, s = tap % S;
      const int c = (o / 64) * 64 + cl;
      const int cpg_out = Cout / groups;
      if (c < Cin && c / Cg == o / cpg_out) v = w[((static_cast<size_t>(o) * Cg + (c % Cg)) * R + r) * S + s];
    } else {
      const int kpt = mode == kModeSlabDense ? Cin : Ktot / (R * S);  // channels reserved per tap (a multiple of the K block)
      const int tap = k / kpt, c = k % kpt;
      const int r = tap / S, s = tap % S;
      if (c < Cin) v = w[((static_cast<size_t>(o) * Cin + c) * R + r) * S + s];
    }
  }
  dst[idx] = __float2bfloat16_rn(v);
}

__global__ void pack_linear_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int F, int Kout,
                                           int Kout_pad) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<size_t>(Kout_pad) * F) return;
  const int o = static_cast<int>(idx / F), f = static_cast<int>(idx % F);
  dst[idx] = __float2bfloat16_rn(o < Kout ? w[static_cast<size_t>(f) * Kout + o] : 0.0f);
}

__global__ void pack_conv_weights_f32_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cg,
                                             int R, int S) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t total = static_cast<size_t>(Cout) * Cg * R * S;
  if (idx >= total) return;
  // dst [R][S][Cg][Cout]
  const int o = static_cast<int>(idx % Cout);
  size_t t = idx / Cout;
  const int c = static_cast<int>(t % Cg);
  t /= Cg;
  const int s = static_cast<int>(t % S), r = static_cast<int>(t / S);
  dst[idx] = w[((static_cast<size_t>(o) * Cg + c) * R + r) * S + s];
}

template <typename T>
__global__ void pack_dw_weights_kernel(const float* __restrict__ w, T* __restrict__ dst, int C, int RS) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * RS) return;
  const int c = idx % C, t = idx / C;
  dst[idx] = from_f32<T>(w[static_cast<size_t>(c) * RS + t]);
}

__global__ void fold_bn_kernel(float* scale, float* shift, const float* gamma, const float* beta, const float* mean,
                               const float* var, const float* bias, float eps, int K, int K_pad) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K_pad) return;
  float sc = 0.0f, sh = 0.0f;
  if (k < K) {
    const float b = bias ? bias[k] : 0.0f;
    if (gamma) {
      // same operation order as F.batch_norm's eval formula: (x - mean) * rsqrt(var + eps) * gamma + beta
      sc = gamma[k] / sqrtf(var[k] + eps);
      sh = beta[k] + (b - mean[k]) * sc;
    } else {
      sc = 1.0f;
      sh = b;
    }
  }
  scale[k] = sc;
  shift[k] = sh;
}

// ------------------------------------------------------------------------------------------------
// layout passes
// ------------------------------------------------------------------------------------------------
// small-C import (C <= 4): one thread per pixel, planes read coalesced, one 8 B / 16 B store
template <typename T>
__global__ void import_nchw_smallc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, size_t HW,
                                          size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const size_t n = idx / HW, hw = idx % HW;
  const float* s = src + n * C * HW + hw;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) v[c] = __ldg(s + c * HW);
  if constexpr (sizeof(T) == 2) {
    uint2 o = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    reinterpret_cast<uint2*>(dst)[idx] = o;
  } else {
    reinterpret_cast<float4*>(dst)[idx] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// uint8 NHWC image batch -> normalised NHWC4 activations (optionally with zero pad columns): one thread per stored pixel
template <typename T>
__global__ void import_u8_nhwc_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, const float* __restrict__ mean,
                                      const float* __restrict__ stdv, int C, int H, int W, int Wp, int pad_l, size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int wp = static_cast<int>(idx % Wp);
  const size_t nh = idx / Wp;
  const int w = wp - pad_l;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (w >= 0 && w < W) {
    const uint8_t* s = src + (nh * W + w) * C;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < C) v[c] = __fdiv_rn(static_cast<float>(s[c]) - __ldg(mean + c), __ldg(stdv + c));
  }
  if constexpr (sizeof(T) == 2)
    reinterpret_cast<uint2*>(dst)[idx] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  else
    reinterpret_cast<float4*>(dst)[idx] = make_float4(v[0], v[1], v[2], v[3]);
}

// RGB fast path: four stored pixels per thread = three aligned 4-byte loads (12 bytes) and two 16-byte stores;
// needs C == 3 and W, pad_l, Wp multiples of 4 (no group of four straddles the image edge)
__global__ void import_u8_rgb_x4_kernel(const uint8_t* __restrict__ src, uint4* __restrict__ dst, const float* __restrict__ mean,
                                        const float* __restrict__ stdv, int H, int W, int Wp4, int pad_l, size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int wq = static_cast<int>(idx % Wp4);
  const size_t nh = idx / Wp4;
  const int w = wq * 4 - pad_l;
  uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
  if (w >= 0 && w < W) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src + (nh * W + w) * 3);
    const uint32_t b0 = __ldg(s), b1 = __ldg(s + 1), b2 = __ldg(s + 2);
    const float m0 = __ldg(mean), m1 = __ldg(mean + 1), m2 = __ldg(mean + 2);
    const float d0 = __ldg(stdv), d1 = __ldg(stdv + 1), d2 = __ldg(stdv + 2);
    auto byte = [&](int k) { return static_cast<float>(((k < 4 ? b0 : (k < 8 ? b1 : b2)) >> (8 * (k & 3))) & 0xffu); };
    auto px = [&](int j, uint32_t& lo, uint32_t& hi) {
      lo = pack_bf16x2(__fdiv_rn(byte(3 * j) - m0, d0), __fdiv_rn(byte(3 * j + 1) - m1, d1));
      hi = pack_bf16x2(__fdiv_rn(byte(3 * j + 2) - m2, d2), 0.0f);
    };
    px(0, o0.x, o0.y), px(1, o0.z, o0.w), px(2, o1.x, o1.y), px(3, o1.z, o1.w);
  }
  dst[idx * 2] = o0;
  dst[idx * 2 + 1] = o1;
}

// uint8 NHWC (N, Hs, Ws, C <= 4) -> bilinear resize to (H, W) -> (x - mean) / std -> [N][H][Wp][4] activations: the reference's
// Resize + Normalize + ToTensor (demo/image_classification/predict-resnet.py:50-54) in the plan's input pass.  The resize
// is OpenCV's INTER_LINEAR for 8-bit images bit for bit (what tensorlayerx's Resize runs on a numpy image): 11-bit
// fixed-point weights from the host-built tables `tx` / `ty` = {i0, i1, c0, c1} per destination column / row,
//   D = S[i0] * c0 + S[i1] * c1  (horizontal, 32 bit),  dst = ((b0 * (D0 >> 4) >> 16) + (b1 * (D1 >> 4) >> 16) + 2) >> 2.
template <typename T>
__global__ void import_u8_resize_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, const float* __restrict__ mean,
                                        const float* __restrict__ stdv, const int4* __restrict__ tx, const int4* __restrict__ ty,
                                        int C, int Hs, int Ws, int H, int W, int Wp, int pad_l, size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int xp = static_cast<int>(idx % Wp);
  const size_t row = idx / Wp;
  const int y = static_cast<int>(row % H);
  const size_t n = row / H;
  const int x = xp - pad_l;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (x >= 0 && x < W) {
    const int4 cx = __ldg(tx + x), cy = __ldg(ty + y);
    const uint8_t* r0 = src + (n * Hs + cy.x) * static_cast<size_t>(Ws) * C;
    const uint8_t* r1 = src + (n * Hs + cy.y) * static_cast<size_t>(Ws) * C;
    for (int c = 0; c < C; ++c) {
      const int d0 = r0[cx.x * C + c] * cx.z + r0[cx.y * C + c] * cx.w;
      const int d1 = r1[cx.x * C + c] * cx.z + r1[cx.y * C + c] * cx.w;
      int o = (((cy.z * (d0 >> 4)) >> 16) + ((cy.w * (d1 >> 4)) >> 16) + 2) >> 2;
      o = min(max(o, 0), 255);
      v[c] = (static_cast<float>(o) - __ldg(mean + c)) / __ldg(stdv + c);
    }
  }
  T* d = dst + idx * 4;
  if constexpr (sizeof(T) == 2) {
    *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  } else {
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// general transpose [N][C][HW] fp32 -> [N][HW][C] T through a 32x32 smem tile
template <typename T>
__global__ void import_nchw_tile_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int HW) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* s = src + static_cast<size_t>(n) * C * HW;
  T* d = dst + static_cast<size_t>(n) * HW * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? __ldg(s + static_cast<size_t>(c) * HW + p) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) d[static_cast<size_t>(p) * C + c] = from_f32<T>(tile[threadIdx.x][i]);
  }
}

template <typename T>
__global__ void export_nchw_tile_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int Cs, int HW) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const T* s = src + static_cast<size_t>(n) * HW * Cs;
  float* d = dst + static_cast<size_t>(n) * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? to_f32(s[static_cast<size_t>(p) * Cs + c]) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) d[static_cast<size_t>(c) * HW + p] = tile[threadIdx.x][i];
  }
}

// bf16 NHWC -> fp32 NCHW, 64 channels x 64 pixels per block: 16-byte loads along C, smem transpose, 16-byte (or,
// when HW is not a multiple of 4, 4-byte) stores along the pixels of one channel plane
template <bool VEC4>
__global__ void __launch_bounds__(256) export_nchw_bf16_64_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst,
                                                                   int C, int HW) {
  pdl_wait();
  __shared__ float tile[64][65];  // [channel][pixel]
  const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const __nv_bfloat16* s = src + static_cast<size_t>(n) * HW * C;
  float* d = dst + static_cast<size_t>(n) * C * HW;
  const int t = threadIdx.x;
  {
    const int cg = t & 7, pr = t >> 3;  // 8 channel groups of 8, 32 pixel rows per pass
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int pl = pr + pass * 32, p = p0 + pl, c = c0 + cg * 8;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p < HW && c < C) Vec8<__nv_bfloat16>::load(s + static_cast<size_t>(p) * C + c, v);  // C is a multiple of 8
#pragma unroll
      for (int j = 0; j < 8; ++j) tile[cg * 8 + j][pl] = v[j];
    }
  }
  __syncthreads();
  if (VEC4) {
    const int pq = t & 15, cr = t >> 4;  // 16 pixel quads, 16 channel rows per pass
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int cl = cr + pass * 16, c = c0 + cl, p = p0 + pq * 4;
      if (c < C && p < HW)  // HW % 4 == 0: a quad is entirely inside or outside
        *reinterpret_cast<float4*>(d + static_cast<size_t>(c) * HW + p) =
            make_float4(tile[cl][pq * 4], tile[cl][pq * 4 + 1], tile[cl][pq * 4 + 2], tile[cl][pq * 4 + 3]);
    }
  } else {
    const int pl = t & 63, cr = t >> 6;  // 64 pixels, 4 channel rows per pass
#pragma unroll
    for (int pass = 0; pass < 16; ++pass) {
      const int cl = cr + pass * 4, c = c0 + cl, p = p0 + pl;
      if (c < C && p < HW) d[static_cast<size_t>(c) * HW + p] = tile[cl][pl];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pooling
// ------------------------------------------------------------------------------------------------
template <typename T, int KS>  // KS = compile-time window (3) or 0 for a runtime window
__global__ void maxpool_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int H, int W, int C8, int P, int Q,
                                    int k_rt, int stride, int pad, size_t total) {
  pdl_wait();
  const int k = KS > 0 ? KS : k_rt;
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % C8);
  size_t t = idx / C8;
  const int q = static_cast<int>(t % Q);
  t /= Q;
  const int pp = static_cast<int>(t % P);
  const size_t n = t / P;
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -FLT_MAX;  // padding behaves as -inf (F.max_pool2d)
  const int h0 = pp * stride - pad, w0 = q * stride - pad;
#pragma unroll
  for (int r = 0; r < k; ++r) {
    const int h = h0 + r;
    if (h < 0 || h >= H) continue;
#pragma unroll
    for (int s = 0; s < k; ++s) {
      const int w = w0 + s;
      if (w < 0 || w >= W) continue;
      float v[8];
      Vec8<T>::load(src + ((n * H + h) * W + w) * static_cast<size_t>(C8) * 8 + c8 * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
    }
  }
  Vec8<T>::store(dst + idx * 8, m);
}

// AvgPool2d(k, stride, pad) (the 2x2 / stride-2 pool in front of the shortcut conv of a ResNet_vd / ResNeSt block,
// segmentation/backbones/resnet_vd.py:25-27,45-46; ResNeSt's 3x3 / stride-2 / pad-1 "avd" pool, classification/resnest.py:245-250):
// fp32 sum of the window in (r, s) order, times 1 / k^2 (padding counts as zeros: torch's count_include_pad default), one rounding
template <typename T>
__global__ void avgpool_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int H, int W, int C8, int P, int Q, int k,
                                    int stride, int pad, size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % C8);
  size_t t = idx / C8;
  const int q = static_cast<int>(t % Q);
  t /= Q;
  const int pp = static_cast<int>(t % P);
  const size_t n = t / P;
  float m[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < k; ++r)
    for (int s2 = 0; s2 < k; ++s2) {
      const int h = pp * stride - pad + r, w = q * stride - pad + s2;
      if (h < 0 || h >= H || w < 0 || w >= W) continue;
      float v[8];
      Vec8<T>::load(src + ((n * H + h) * W + w) * static_cast<size_t>(C8) * 8 + c8 * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] += v[i];
    }
  const float inv = 1.0f / static_cast<float>(k * k);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] *= inv;
  Vec8<T>::store(dst + idx * 8, m);
}

// Split attention (ResNeSt SplatConv, classification/resnest.py:53-82,146-166): x holds `radix` channel groups of C = G * cpc
// channels ([r][g][j] order: `tlx.split(x, radix)`), att the attention logits of the block in the conv's [g][r][j] order
// (rSoftmax reshapes them to (batch, G, radix, cpc), soft-maxes over radix and flattens to [r][g][j]):
//   out[n, p, c] = sum_r softmax_r(att[n, g, :, j])[r] * x[n, p, r * C + c],   c = g * cpc + j
// Block = (image, slice of the pixels): the radix x C probabilities are computed once into shared memory, then every
// thread streams 8 channels of a pixel per step.
__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__global__ void splat_apply_kernel(const T* __restrict__ x, const T* __restrict__ att, T* __restrict__ dst, int HW, int C, int radix,
                                   int cpc, int slices) {
  extern __shared__ float prob[];  // [radix][C]
  pdl_wait();
  const int n = blockIdx.x / slices, slice = blockIdx.x % slices;
  const T* a = att + static_cast<size_t>(n) * radix * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpc, j = c - g * cpc;
    float mx = -3.402823466e38f;
    for (int r = 0; r < radix; ++r) mx = fmaxf(mx, to_float(a[(g * radix + r) * cpc + j]));
    float sum = 0.0f;
    for (int r = 0; r < radix; ++r) {
      const float e = expf(to_float(a[(g * radix + r) * cpc + j]) - mx);
      prob[r * C + c] = e;
      sum += e;
    }
    const float inv = 1.0f / sum;
    for (int r = 0; r < radix; ++r) prob[r * C + c] *= inv;
  }
  __syncthreads();
  const int C8 = C / 8;
  const int per = (HW + slices - 1) / slices;
  const int p0 = slice * per, p1 = min(HW, p0 + per);
  const size_t total = static_cast<size_t>(p1 - p0) * C8;
  for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    const size_t pix = static_cast<size_t>(n) * HW + p0 + i / C8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < radix; ++r) {
      float v[8];
      Vec8<T>::load(x + (pix * radix + r) * static_cast<size_t>(C) + c8 * 8, v);
      const float* pr = prob + r * C + c8 * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(pr[k], v[k], acc[k]);
    }
    Vec8<T>::store(dst + pix * static_cast<size_t>(C) + c8 * 8, acc);
  }
}

// global average pool: block per image, blockDim = (8-channel groups, kGapSlices pixel slices).  Each thread sums its
// slice of the pixels in pixel order (fp32), the slices are then added in slice order through shared memory: four times
// the loads in flight of a thread-per-group kernel, which was latency-bound (49 dependent-issue loads per thread).
// (blockDim.y = 4 slices for 2048 channels up to 32 for 128: narrow maps - the pooled radix groups of a ResNeSt block - would
// otherwise leave most of the block idle)
template <typename T>
__global__ void gap_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int HW, int C8) {
  extern __shared__ float gap_part[];  // [slices - 1][C8 * 8]
  const int kGapSlices = blockDim.y;
  pdl_wait();
  const int n = blockIdx.x;
  const T* s = src + static_cast<size_t>(n) * HW * C8 * 8;
  const float inv = 1.0f / static_cast<float>(HW);
  const int slice = threadIdx.y;
  const int per = (HW + kGapSlices - 1) / kGapSlices;
  const int p0 = slice * per, p1 = min(HW, p0 + per);
  for (int g0 = 0; g0 < C8; g0 += blockDim.x) {
    const int g = g0 + threadIdx.x;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (g < C8) {
#pragma unroll 4
      for (int p = p0; p < p1; ++p) {
        float v[8];
        Vec8<T>::load(s + (static_cast<size_t>(p) * C8 + g) * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
      if (slice > 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) gap_part[(static_cast<size_t>(slice - 1) * C8 + g) * 8 + i] = acc[i];
      }
    }
    __syncthreads();
    if (slice == 0 && g < C8) {
      for (int k = 0; k < kGapSlices - 1; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += gap_part[(static_cast<size_t>(k) * C8 + g) * 8 + i];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= inv;
      Vec8<T>::store(dst + (static_cast<size_t>(n) * C8 + g) * 8, acc);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise conv + folded BN + activation (+ residual), NHWC, 8 channels per thread
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void dwconv_nhwc_kernel(const T* __restrict__ src, const T* __restrict__ w_rsc, T* __restrict__ dst,
                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                   const T* __restrict__ residual, int H, int W, int C8, int P, int Q, int R, int S,
                                   int stride, int pad, int act1, float alpha1, int act2, float alpha2, size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % C8);
  size_t t = idx / C8;
  const int q = static_cast<int>(t % Q);
  t /= Q;
  const int pp = static_cast<int>(t % P);
  const size_t n = t / P;
  const int C = C8 * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int h0 = pp * stride - pad, w0 = q * stride - pad;
  for (int r = 0; r < R; ++r) {
    const int h = h0 + r;
    if (h < 0 || h >= H) continue;
    for (int s = 0; s < S; ++s) {
      const int w = w0 + s;
      if (w < 0 || w >= W) continue;
      float v[8], k[8];
      Vec8<T>::load(src + ((n * H + h) * W + w) * static_cast<size_t>(C) + c8 * 8, v);
      Vec8<T>::load(w_rsc + static_cast<size_t>(r * S + s) * C + c8 * 8, k);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], k[i], acc[i]);
    }
  }
  float res[8];
  if (residual) Vec8<T>::load(residual + idx * 8, res);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float y = apply_act(fmaf(acc[i], __ldg(scale + c8 * 8 + i), __ldg(shift + c8 * 8 + i)), act1, alpha1);
    if (residual) y += res[i];
    acc[i] = apply_act(y, act2, alpha2);
  }
  Vec8<T>::store(dst + idx * 8, acc);
}

// 3x3 depthwise, pad 1, stride 1 or 2: the HBM-bound workhorse of MobileNet (bf16).
// At 9 MACs per 4 bytes of traffic this kernel is closer to the issue limit than to the HBM limit,
// so it is written for instruction count:
//   * thread = 4 channels x (TH x TW) output pixels; the 9 x 4 filter taps are unpacked to fp32 ONCE;
//   * every input vector (8 B = 4 channels) is loaded and unpacked once and feeds all outputs it touches
//     ((TH*S+2)(TW*S+2) loads for TH*TW outputs: 3 per output at stride 1 instead of 9);
//   * math is packed fp32x2 FMA (sm_100 `fma.rn.f32x2`), two channels per instruction;
//   * the activation switch sits outside the element loops.
// Adjacent lanes own adjacent channel quads, so each load instruction reads contiguous 256 B runs.
// ffma2 / bf16x2_to_f32x2: common.cuh
__device__ __forceinline__ float2 unpack_f32x2(unsigned long long v) {
  return make_float2(__uint_as_float(static_cast<uint32_t>(v)), __uint_as_float(static_cast<uint32_t>(v >> 32)));
}

template <int STRIDE, int TH, int TW>
__global__ void __launch_bounds__(128)
dwconv3x3_nhwc_kernel(const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ w_rsc,
                      __nv_bfloat16* __restrict__ dst, const float* __restrict__ scale,
                      const float* __restrict__ shift, const __nv_bfloat16* __restrict__ residual, int H, int W,
                      int C4, int P, int Q, int PT, int QT, int act1, float alpha1, int act2, float alpha2,
                      size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c4 = static_cast<int>(idx % C4);
  size_t t = idx / C4;
  const int qt = static_cast<int>(t % QT);
  t /= QT;
  const int pt = static_cast<int>(t % PT);
  const size_t n = t / PT;
  const int C = C4 * 4;
  const int p0 = pt * TH, q0 = qt * TW;
  constexpr int ROWS = (TH - 1) * STRIDE + 3, COLS = (TW - 1) * STRIDE + 3;

  unsigned long long wv[9][2];  // 9 taps x 4 channels, fp32, packed in pairs
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(w_rsc + static_cast<size_t>(k) * C + c4 * 4));
    wv[k][0] = bf16x2_to_f32x2(u.x), wv[k][1] = bf16x2_to_f32x2(u.y);
  }
  unsigned long long acc[TH][TW][2];
#pragma unroll
  for (int i = 0; i < TH; ++i)
#pragma unroll
    for (int j = 0; j < TW; ++j) acc[i][j][0] = 0ull, acc[i][j][1] = 0ull;

  const int ih0 = p0 * STRIDE - 1, iw0 = q0 * STRIDE - 1;
  const __nv_bfloat16* img = src + n * static_cast<size_t>(H) * W * C + c4 * 4;
#pragma unroll
  for (int ri = 0; ri < ROWS; ++ri) {
    const int ih = ih0 + ri;
    const bool rok = ih >= 0 && ih < H;
    const __nv_bfloat16* rowp = img + static_cast<size_t>(rok ? ih : 0) * W * C;
    uint2 raw[COLS];
#pragma unroll
    for (int ci = 0; ci < COLS; ++ci) {
      const int iw = iw0 + ci;
      raw[ci] = (rok && iw >= 0 && iw < W) ? __ldg(reinterpret_cast<const uint2*>(rowp + static_cast<size_t>(iw) * C))
                                           : make_uint2(0, 0);
    }
#pragma unroll
    for (int ci = 0; ci < COLS; ++ci) {
      const unsigned long long x0 = bf16x2_to_f32x2(raw[ci].x), x1 = bf16x2_to_f32x2(raw[ci].y);
#pragma unroll
      for (int to = 0; to < TH; ++to) {
        const int r = ri - to * STRIDE;  // compile time after unrolling
        if (r < 0 || r > 2) continue;
#pragma unroll
        for (int tq = 0; tq < TW; ++tq) {
          const int s3 = ci - tq * STRIDE;
          if (s3 < 0 || s3 > 2) continue;
          acc[to][tq][0] = ffma2(x0, wv[r * 3 + s3][0], acc[to][tq][0]);
          acc[to][tq][1] = ffma2(x1, wv[r * 3 + s3][1], acc[to][tq][1]);
        }
      }
    }
  }
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c4 * 4));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c4 * 4));
  float y[TH * TW][4];
#pragma unroll
  for (int to = 0; to < TH; ++to)
#pragma unroll
    for (int tq = 0; tq < TW; ++tq) {
      const float2 a = unpack_f32x2(acc[to][tq][0]), b = unpack_f32x2(acc[to][tq][1]);
      float* o = y[to * TW + tq];
      o[0] = fmaf(a.x, sc.x, sh.x), o[1] = fmaf(a.y, sc.y, sh.y), o[2] = fmaf(b.x, sc.z, sh.z), o[3] = fmaf(b.y, sc.w, sh.w);
    }
  // activation switches outside the element loops
  if (act1 == TLXCV_ACT_RELU6) {
#pragma unroll
    for (int i = 0; i < TH * TW; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) y[i][j] = fminf(fmaxf(y[i][j], 0.0f), 6.0f);
  } else if (act1 == TLXCV_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < TH * TW; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) y[i][j] = fmaxf(y[i][j], 0.0f);
  } else if (act1 == TLXCV_ACT_LEAKY) {
#pragma unroll
    for (int i = 0; i < TH * TW; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) y[i][j] = y[i][j] > 0.0f ? y[i][j] : y[i][j] * alpha1;
  }
#pragma unroll
  for (int to = 0; to < TH; ++to)
#pragma unroll
    for (int tq = 0; tq < TW; ++tq) {
      const int pp = p0 + to, qq = q0 + tq;
      if (pp >= P || qq >= Q) continue;
      const size_t o = ((n * P + pp) * Q + qq) * static_cast<size_t>(C) + c4 * 4;
      float* v = y[to * TW + tq];
      if (residual != nullptr) {  // rare (no hot-path model adds to a depthwise output): kept simple
        const uint2 ru = __ldg(reinterpret_cast<const uint2*>(residual + o));
        const float2 r0 = unpack_f32x2(bf16x2_to_f32x2(ru.x)), r1 = unpack_f32x2(bf16x2_to_f32x2(ru.y));
        v[0] = apply_act(v[0] + r0.x, act2, alpha2), v[1] = apply_act(v[1] + r0.y, act2, alpha2);
        v[2] = apply_act(v[2] + r1.x, act2, alpha2), v[3] = apply_act(v[3] + r1.y, act2, alpha2);
      }
      *reinterpret_cast<uint2*>(dst + o) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
}

template <typename T>
__global__ void add_act_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ dst, size_t n8, int act,
                               float alpha) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= n8) return;
  float x[8], y[8];
  Vec8<T>::load(a + idx * 8, x);
  if (b) {
    Vec8<T>::load(b + idx * 8, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] += y[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = apply_act(x[i], act, alpha);
  Vec8<T>::store(dst + idx * 8, x);
}

// out[n][y][x][0:C0] = a[n][y / ra][x / ra][:],  out[n][y][x][C0:C0+C1] = b[n][y / rb][x / rb][:]   (nearest up-sampling by
// integer factors + channel concat in ONE pass: Interpolater + tlx.concat of YOLOv3FPN.forward, detection/yolov3.py:244-253).
// One thread per 8 output channels (16 B for bf16); b == nullptr: up-sampling alone.
template <typename T>
__global__ void upsample_concat_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ dst, int H, int W,
                                       int C0_8, int C1_8, int ra, int rb, size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int C8 = C0_8 + C1_8;
  const int c8 = static_cast<int>(idx % C8);
  size_t pix = idx / C8;
  const int x = static_cast<int>(pix % W);
  pix /= W;
  const int y = static_cast<int>(pix % H);
  const size_t n = pix / H;
  float v[8];
  if (c8 < C0_8) {
    const int Ha = H / ra, Wa = W / ra;
    Vec8<T>::load(a + ((n * Ha + y / ra) * Wa + x / ra) * (static_cast<size_t>(C0_8) * 8) + c8 * 8, v);
  } else {
    const int Hb = H / rb, Wb = W / rb;
    Vec8<T>::load(b + ((n * Hb + y / rb) * Wb + x / rb) * (static_cast<size_t>(C1_8) * 8) + (c8 - C0_8) * 8, v);
  }
  Vec8<T>::store(dst + idx * 8, v);
}

// one warp per row; first maximal index wins (torch.argmax on ties returns the first occurrence)
__global__ void argmax_rows_kernel(const float* __restrict__ logits, long long* __restrict__ dst, int N, int K) {
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* p = logits + static_cast<size_t>(row) * K;
  float best = -FLT_MAX;
  int besti = 0x7fffffff;
  for (int i = lane; i < K; i += 32) {
    const float v = p[i];
    if (v > best) best = v, besti = i;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
    if (ob > best || (ob == best && oi < besti)) best = ob, besti = oi;
  }
  if (lane == 0) dst[row] = besti;
}

// Row-wise softmax of fp32 logits, one warp per row: max and sum by warp shuffles, exp of the shifted logits
// (tlx.softmax(logits, axis=-1); the "final FC + softmax" of the north star).
__global__ void softmax_rows_kernel(const float* __restrict__ logits, float* __restrict__ dst, int N, int K) {
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* p = logits + static_cast<size_t>(row) * K;
  float mx = -FLT_MAX;
  for (int i = lane; i < K; i += 32) mx = fmaxf(mx, p[i]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float sum = 0.0f;
  for (int i = lane; i < K; i += 32) sum += expf(p[i] - mx);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float inv = 1.0f / sum;
  float* d = dst + static_cast<size_t>(row) * K;
  for (int i = lane; i < K; i += 32) d[i] = expf(p[i] - mx) * inv;
}

// Mean softmax cross-entropy of fp32 logits against int64 class labels (tlx.losses.softmax_cross_entropy_with_logits,
// tasks/image_classification.py:10-15): one warp per row writes  log(sum exp(l - max)) + max - l[target]  into a scratch
// row, the LAST block to finish (ticket counter, reset for the next launch) sums the rows in index order - so the result
// does not depend on block scheduling - and writes the mean.  A label outside [0, K) gives NaN.
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const long long* __restrict__ target, float* __restrict__ row_loss,
                                  unsigned int* __restrict__ ticket, float* __restrict__ dst, int N, int K) {
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row < N) {
    const float* p = logits + static_cast<size_t>(row) * K;
    float mx = -FLT_MAX;
    for (int i = lane; i < K; i += 32) mx = fmaxf(mx, p[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.0f;
    for (int i = lane; i < K; i += 32) sum += expf(p[i] - mx);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) {
      const long long t = target[row];
      row_loss[row] = (t >= 0 && t < K) ? logf(sum) + mx - p[t] : __int_as_float(0x7fc00000);
    }
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    float acc = 0.0f;
    for (int base = 0; base < N; base += 32) {  // fixed order: lanes over 32 consecutive rows, then a shuffle tree
      float v = base + lane < N ? __ldcg(row_loss + base + lane) : 0.0f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      acc += v;
    }
    if (lane == 0) {
      dst[0] = acc / static_cast<float>(N);
      *ticket = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 direct convolution (validation mode): thread = (pixel, 4 output channels)
// weights [R][S][C/g][K]; lanes run along K so weight loads coalesce and input loads broadcast
// ------------------------------------------------------------------------------------------------
__global__ void conv_direct_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                       const float* __restrict__ residual, int H, int W, int C, int Cs, int P, int Q,
                                       int K, int R, int S, int stride, int pad, int dil, int groups, int act1,
                                       float alpha1, int act2, float alpha2, size_t total) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int k = static_cast<int>(idx % K);
  size_t t = idx / K;
  const int q = static_cast<int>(t % Q);
  t /= Q;
  const int pp = static_cast<int>(t % P);
  const size_t n = t / P;
  const int Cg = C / groups, Kg = K / groups;
  const int cbeg = (k / Kg) * Cg;
  float acc = 0.0f;
  const int h0 = pp * stride - pad, w0 = q * stride - pad;
  for (int r = 0; r < R; ++r) {
    const int h = h0 + r * dil;
    if (h < 0 || h >= H) continue;
    for (int s = 0; s < S; ++s) {
      const int ww = w0 + s * dil;
      if (ww < 0 || ww >= W) continue;
      const float* ip = in + ((n * H + h) * W + ww) * static_cast<size_t>(Cs) + cbeg;
      const float* wp = w + (static_cast<size_t>(r * S + s) * Cg) * K + k;
      for (int c = 0; c < Cg; ++c) acc = fmaf(__ldg(ip + c), __ldg(wp + static_cast<size_t>(c) * K), acc);
    }
  }
  float y = apply_act(fmaf(acc, scale[k], shift[k]), act1, alpha1);
  if (residual) y += residual[idx];
  out[idx] = apply_act(y, act2, alpha2);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t pack_conv_weights(const float* oihw, __nv_bfloat16* dst, int Cout, int Cout_pad, int Cin, int R, int S,
                              int groups, int mode, int Ktot, cudaStream_t st) {
  const size_t total = static_cast<size_t>(Cout_pad) * Ktot;
  pack_conv_weights_kernel<<<blocks_for(total), kThreads, 0, st>>>(oihw, dst, Cout, Cout_pad, Cin, R, S, groups, mode, Ktot);
  return cudaGetLastError();
}

cudaError_t pack_linear_weights(const float* w, __nv_bfloat16* dst, int F, int Kout, int Kout_pad, cudaStream_t st) {
  const size_t total = static_cast<size_t>(Kout_pad) * F;
  pack_linear_weights_kernel<<<blocks_for(total), kThreads, 0, st>>>(w, dst, F, Kout, Kout_pad);
  return cudaGetLastError();
}

cudaError_t pack_conv_weights_f32(const float* oihw, float* dst, int Cout, int Cg, int R, int S, cudaStream_t st) {
  const size_t total = static_cast<size_t>(Cout) * Cg * R * S;
  pack_conv_weights_f32_kernel<<<blocks_for(total), kThreads, 0, st>>>(oihw, dst, Cout, Cg, R, S);
  return cudaGetLastError();
}

cudaError_t pack_dw_weights(const float* oihw, void* dst, int C, int RS, int is_f32, cudaStream_t st) {
  const int total = C * RS;
  if (is_f32)
    pack_dw_weights_kernel<float><<<blocks_for(total), kThreads, 0, st>>>(oihw, static_cast<float*>(dst), C, RS);
  else
    pack_dw_weights_kernel<__nv_bfloat16><<<blocks_for(total), kThreads, 0, st>>>(oihw, static_cast<__nv_bfloat16*>(dst), C, RS);
  return cudaGetLastError();
}

cudaError_t fold_bn(float* scale, float* shift, const float* gamma, const float* beta, const float* mean,
                    const float* var, const float* bias, float eps, int K, int K_pad, cudaStream_t st) {
  fold_bn_kernel<<<blocks_for(K_pad), kThreads, 0, st>>>(scale, shift, gamma, beta, mean, var, bias, eps, K, K_pad);
  return cudaGetLastError();
}

cudaError_t import_nchw(const float* src, void* dst, int N, int C, int H, int W, int Cs, int is_f32, cudaStream_t st) {
  const size_t HW = static_cast<size_t>(H) * W;
  if (C <= 4 && Cs == 4) {
    const size_t total = static_cast<size_t>(N) * HW;
    if (is_f32)
      TLXCV_LAUNCH(import_nchw_smallc_kernel<float>, blocks_for(total), kThreads, 0, st, src, static_cast<float*>(dst), C, HW, total);
    else
      TLXCV_LAUNCH(import_nchw_smallc_kernel<__nv_bfloat16>, blocks_for(total), kThreads, 0, st, src, static_cast<__nv_bfloat16*>(dst), C, HW, total);
    return cudaGetLastError();
  }
  if (Cs != C) return cudaErrorInvalidValue;
  dim3 grid(static_cast<unsigned>((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
  if (is_f32)
    TLXCV_LAUNCH(import_nchw_tile_kernel<float>, grid, block, 0, st, src, static_cast<float*>(dst), C, static_cast<int>(HW));
  else
    TLXCV_LAUNCH(import_nchw_tile_kernel<__nv_bfloat16>, grid, block, 0, st, src, static_cast<__nv_bfloat16*>(dst), C, static_cast<int>(HW));
  return cudaGetLastError();
}

cudaError_t import_u8_nhwc(const uint8_t* src, void* dst, const float* mean, const float* stdv, int N, int C, int H, int W,
                           int Wp, int pad_l, int is_f32, cudaStream_t st) {
  if (C > 4) return cudaErrorInvalidValue;
  if (!is_f32 && C == 3 && W % 4 == 0 && pad_l % 4 == 0 && Wp % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
    const size_t total4 = static_cast<size_t>(N) * H * (Wp / 4);
    TLXCV_LAUNCH(import_u8_rgb_x4_kernel, blocks_for(total4), kThreads, 0, st, src, static_cast<uint4*>(dst), mean, stdv, H, W, Wp / 4,
                 pad_l, total4);
    return cudaGetLastError();
  }
  const size_t total = static_cast<size_t>(N) * H * Wp;
  if (is_f32)
    TLXCV_LAUNCH(import_u8_nhwc_kernel<float>, blocks_for(total), kThreads, 0, st, src, static_cast<float*>(dst), mean, stdv, C, H, W, Wp, pad_l, total);
  else
    TLXCV_LAUNCH(import_u8_nhwc_kernel<__nv_bfloat16>, blocks_for(total), kThreads, 0, st, src, static_cast<__nv_bfloat16*>(dst), mean, stdv, C, H, W, Wp, pad_l, total);
  return cudaGetLastError();
}

cudaError_t import_u8_resize(const uint8_t* src, void* dst, const float* mean, const float* stdv, const int4* tx, const int4* ty,
                             int N, int C, int Hs, int Ws, int H, int W, int Wp, int pad_l, int is_f32, cudaStream_t st) {
  if (C > 4) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(N) * H * Wp;
  if (is_f32)
    TLXCV_LAUNCH(import_u8_resize_kernel<float>, blocks_for(total), kThreads, 0, st, src, static_cast<float*>(dst), mean, stdv, tx, ty, C, Hs, Ws, H, W, Wp, pad_l, total);
  else
    TLXCV_LAUNCH(import_u8_resize_kernel<__nv_bfloat16>, blocks_for(total), kThreads, 0, st, src, static_cast<__nv_bfloat16*>(dst), mean, stdv, tx, ty, C, Hs, Ws, H, W, Wp, pad_l, total);
  return cudaGetLastError();
}

cudaError_t export_nchw(const void* src, float* dst, int N, int C, int Cs, int H, int W, int is_f32, cudaStream_t st) {
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  if (is_f32) {
    TLXCV_LAUNCH(export_nchw_tile_kernel<float>, grid, block, 0, st, static_cast<const float*>(src), dst, C, Cs, HW);
  } else if (C % 8 == 0 && Cs == C) {
    dim3 grid64((HW + 63) / 64, (C + 63) / 64, N);
    if (HW % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0)
      TLXCV_LAUNCH(export_nchw_bf16_64_kernel<true>, grid64, 256, 0, st, static_cast<const __nv_bfloat16*>(src), dst, C, HW);
    else
      TLXCV_LAUNCH(export_nchw_bf16_64_kernel<false>, grid64, 256, 0, st, static_cast<const __nv_bfloat16*>(src), dst, C, HW);
  } else {
    TLXCV_LAUNCH(export_nchw_tile_kernel<__nv_bfloat16>, grid, block, 0, st, static_cast<const __nv_bfloat16*>(src), dst, C, Cs, HW);
  }
  return cudaGetLastError();
}

cudaError_t maxpool_nhwc(const void* src, void* dst, int N, int H, int W, int C, int P, int Q, int k, int stride, int pad,
                         int is_f32, cudaStream_t st) {
  if (C % 8) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(N) * P * Q * (C / 8);
  if (is_f32)
    TLXCV_LAUNCH((maxpool_nhwc_kernel<float, 0>), blocks_for(total), kThreads, 0, st, static_cast<const float*>(src), static_cast<float*>(dst), H, W, C / 8, P, Q, k, stride, pad, total);
  else if (k == 3)
    TLXCV_LAUNCH((maxpool_nhwc_kernel<__nv_bfloat16, 3>), blocks_for(total), kThreads, 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), H, W, C / 8, P, Q, k, stride, pad, total);
  else
    TLXCV_LAUNCH((maxpool_nhwc_kernel<__nv_bfloat16, 0>), blocks_for(total), kThreads, 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), H, W, C / 8, P, Q, k, stride, pad, total);
  return cudaGetLastError();
}

cudaError_t splat_apply(const void* x, const void* att, void* dst, int N, int HW, int C, int radix, int cardinality, int is_f32,
                        cudaStream_t st) {
  if (C % 8 || radix < 1 || cardinality < 1 || C % cardinality) return cudaErrorInvalidValue;
  // enough blocks to fill the chip: images x pixel slices
  int slices = 1;
  while (N * slices < 2 * 148 && slices * 64 < HW) slices *= 2;
  const size_t smem = static_cast<size_t>(radix) * C * sizeof(float);
  if (is_f32)
    TLXCV_LAUNCH(splat_apply_kernel<float>, N * slices, kThreads, smem, st, static_cast<const float*>(x), static_cast<const float*>(att),
                 static_cast<float*>(dst), HW, C, radix, C / cardinality, slices);
  else
    TLXCV_LAUNCH(splat_apply_kernel<__nv_bfloat16>, N * slices, kThreads, smem, st, static_cast<const __nv_bfloat16*>(x),
                 static_cast<const __nv_bfloat16*>(att), static_cast<__nv_bfloat16*>(dst), HW, C, radix, C / cardinality, slices);
  return cudaSuccess;
}

cudaError_t avgpool_nhwc(const void* src, void* dst, int N, int H, int W, int C, int P, int Q, int k, int stride, int pad, int is_f32,
                         cudaStream_t st) {
  if (C % 8 || pad < 0 || (P - 1) * stride + k > H + 2 * pad || (Q - 1) * stride + k > W + 2 * pad) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(N) * P * Q * (C / 8);
  if (is_f32)
    TLXCV_LAUNCH(avgpool_nhwc_kernel<float>, blocks_for(total), kThreads, 0, st, static_cast<const float*>(src), static_cast<float*>(dst), H, W, C / 8, P, Q, k, stride, pad, total);
  else
    TLXCV_LAUNCH(avgpool_nhwc_kernel<__nv_bfloat16>, blocks_for(total), kThreads, 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), H, W, C / 8, P, Q, k, stride, pad, total);
  return cudaGetLastError();
}

cudaError_t gap_nhwc(const void* src, void* dst, int N, int HW, int C, int is_f32, cudaStream_t st) {
  if (C % 8) return cudaErrorInvalidValue;
  int tx = 16;
  while (tx < C / 8 && tx < 256) tx *= 2;
  const int slices = std::max(4, std::min(32, 1024 / tx));
  const dim3 threads(tx, slices);
  const int smem = (slices - 1) * C * static_cast<int>(sizeof(float));
  if (smem > 48 * 1024) return cudaErrorInvalidValue;  // C > 4096: not a shape of this path
  if (is_f32)
    TLXCV_LAUNCH(gap_nhwc_kernel<float>, N, threads, smem, st, static_cast<const float*>(src), static_cast<float*>(dst), HW, C / 8);
  else
    TLXCV_LAUNCH(gap_nhwc_kernel<__nv_bfloat16>, N, threads, smem, st, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), HW, C / 8);
  return cudaGetLastError();
}

// Depthwise 3x3 with the channel count as a template parameter: the generic kernel above spends ~1280 instructions per
// 32 outputs, two thirds of them integer work (64-bit address arithmetic, bound checks and index divisions per load).
// With C known at compile time every pixel / tap offset is an immediate, the thread -> (channel quad, tile) map needs one
// constant division, validity is one predicate per input row and column, and ReLU / ReLU6 run on the packed bf16 pairs
// (clamping commutes with the rounding: 0 and 6 are exact in bf16).
template <int STRIDE, int TH, int TW, int C>
__global__ void __launch_bounds__(128)
dwconv3x3_nhwc_c_kernel(const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ w_rsc,
                        __nv_bfloat16* __restrict__ dst, const float* __restrict__ scale, const float* __restrict__ shift,
                        int H, int W, int P, int Q, int QT, int act1, float alpha1) {
  pdl_wait();
  constexpr int C4 = C / 4;
  constexpr int ROWS = (TH - 1) * STRIDE + 3, COLS = (TW - 1) * STRIDE + 3;
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= QT * C4) return;
  const int c4 = i % C4, qt = i / C4;
  const int pt = blockIdx.y, n = blockIdx.z;
  const int p0 = pt * TH, q0 = qt * TW;
  const int ih0 = p0 * STRIDE - 1, iw0 = q0 * STRIDE - 1;

  unsigned long long wv[9][2];  // 9 taps x 4 channels, fp32, packed in pairs
  {
    const __nv_bfloat16* wp = w_rsc + c4 * 4;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(wp + k * C));
      wv[k][0] = bf16x2_to_f32x2(u.x), wv[k][1] = bf16x2_to_f32x2(u.y);
    }
  }
  unsigned long long acc[TH][TW][2];
#pragma unroll
  for (int a = 0; a < TH; ++a)
#pragma unroll
    for (int b = 0; b < TW; ++b) acc[a][b][0] = 0ull, acc[a][b][1] = 0ull;

  bool cok[COLS];
#pragma unroll
  for (int ci = 0; ci < COLS; ++ci) cok[ci] = static_cast<unsigned>(iw0 + ci) < static_cast<unsigned>(W);
  // pointer to the (possibly virtual) top-left input pixel of the tile; only valid positions are dereferenced
  const __nv_bfloat16* rowp = src + ((static_cast<long long>(n) * H + ih0) * W + iw0) * C + c4 * 4;
  const long long row_stride = static_cast<long long>(W) * C;
#pragma unroll
  for (int ri = 0; ri < ROWS; ++ri, rowp += row_stride) {
    const bool rok = static_cast<unsigned>(ih0 + ri) < static_cast<unsigned>(H);
    uint2 raw[COLS];
#pragma unroll
    for (int ci = 0; ci < COLS; ++ci)
      raw[ci] = (rok && cok[ci]) ? __ldg(reinterpret_cast<const uint2*>(rowp + ci * C)) : make_uint2(0, 0);
#pragma unroll
    for (int ci = 0; ci < COLS; ++ci) {
      const unsigned long long x0 = bf16x2_to_f32x2(raw[ci].x), x1 = bf16x2_to_f32x2(raw[ci].y);
#pragma unroll
      for (int to = 0; to < TH; ++to) {
        const int r = ri - to * STRIDE;  // compile time after unrolling
        if (r < 0 || r > 2) continue;
#pragma unroll
        for (int tq = 0; tq < TW; ++tq) {
          const int s3 = ci - tq * STRIDE;
          if (s3 < 0 || s3 > 2) continue;
          acc[to][tq][0] = ffma2(x0, wv[r * 3 + s3][0], acc[to][tq][0]);
          acc[to][tq][1] = ffma2(x1, wv[r * 3 + s3][1], acc[to][tq][1]);
        }
      }
    }
  }
  // folded BN on the packed pairs, then bf16, then the activation on bf16 pairs
  const uint4 scu = __ldg(reinterpret_cast<const uint4*>(scale + c4 * 4));
  const uint4 shu = __ldg(reinterpret_cast<const uint4*>(shift + c4 * 4));
  const unsigned long long sc0 = (static_cast<unsigned long long>(scu.y) << 32) | scu.x, sc1 = (static_cast<unsigned long long>(scu.w) << 32) | scu.z;
  const unsigned long long sh0 = (static_cast<unsigned long long>(shu.y) << 32) | shu.x, sh1 = (static_cast<unsigned long long>(shu.w) << 32) | shu.z;
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f), six2 = __floats2bfloat162_rn(6.0f, 6.0f);
  __nv_bfloat16* outp = dst + ((static_cast<long long>(n) * P + p0) * Q + q0) * C + c4 * 4;
  const long long orow = static_cast<long long>(Q) * C;
#pragma unroll
  for (int to = 0; to < TH; ++to)
#pragma unroll
    for (int tq = 0; tq < TW; ++tq) {
      const float2 a = unpack_f32x2(ffma2(acc[to][tq][0], sc0, sh0)), b = unpack_f32x2(ffma2(acc[to][tq][1], sc1, sh1));
      __nv_bfloat162 y0, y1;
      if (act1 == TLXCV_ACT_LEAKY) {
        y0 = __floats2bfloat162_rn(a.x > 0.f ? a.x : a.x * alpha1, a.y > 0.f ? a.y : a.y * alpha1);
        y1 = __floats2bfloat162_rn(b.x > 0.f ? b.x : b.x * alpha1, b.y > 0.f ? b.y : b.y * alpha1);
      } else {
        y0 = __floats2bfloat162_rn(a.x, a.y), y1 = __floats2bfloat162_rn(b.x, b.y);
        if (act1 == TLXCV_ACT_RELU || act1 == TLXCV_ACT_RELU6) y0 = __hmax2(y0, zero2), y1 = __hmax2(y1, zero2);
        if (act1 == TLXCV_ACT_RELU6) y0 = __hmin2(y0, six2), y1 = __hmin2(y1, six2);
      }
      if (p0 + to < P && q0 + tq < Q)
        *reinterpret_cast<uint2*>(outp + to * orow + tq * C) =
            make_uint2(*reinterpret_cast<uint32_t*>(&y0), *reinterpret_cast<uint32_t*>(&y1));
    }
}

template <int STRIDE, int C>
cudaError_t launch_dw_c(const __nv_bfloat16* x, const __nv_bfloat16* wp, __nv_bfloat16* y, const float* scale, const float* shift,
                        int N, int H, int W, int P, int Q, int act1, float alpha1, cudaStream_t st) {
  constexpr int TH = STRIDE == 1 ? 4 : 2, TW = 2;
  const int PT = (P + TH - 1) / TH, QT = (Q + TW - 1) / TW;
  dim3 grid((QT * (C / 4) + 127) / 128, PT, N);
  TLXCV_LAUNCH((dwconv3x3_nhwc_c_kernel<STRIDE, TH, TW, C>), grid, 128, 0, st, x, wp, y, scale, shift, H, W, P, Q, QT, act1, alpha1);
  return cudaSuccess;
}

// channel counts of the depthwise layers of MobileNetV1 / V2 (classification/mobilenetv1.py:134-243, mobilenetv2.py:76-78)
#define TLXCV_DW_CHANNELS(X) X(32) X(64) X(96) X(128) X(144) X(192) X(256) X(384) X(512) X(576) X(960) X(1024)

cudaError_t dwconv_nhwc(const void* src, const void* w_rsc, void* dst, const float* scale, const float* shift,
                        const void* residual, int N, int H, int W, int C, int P, int Q, int R, int S, int stride, int pad,
                        int act1, float alpha1, int act2, float alpha2, int is_f32, cudaStream_t st) {
  if (C % 8) return cudaErrorInvalidValue;
  if (!is_f32 && R == 3 && S == 3 && pad == 1 && (stride == 1 || stride == 2)) {
    const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(src);
    const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(w_rsc);
    const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(residual);
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(dst);
    if (rp == nullptr && N <= 65535 && !tuning_env("TLXCV_NO_DW_FAST")) {
#define TLXCV_X(CC)                                                                                                   \
  if (C == CC)                                                                                                        \
    return stride == 1 ? launch_dw_c<1, CC>(x, wp, y, scale, shift, N, H, W, P, Q, act1, alpha1, st)                  \
                       : launch_dw_c<2, CC>(x, wp, y, scale, shift, N, H, W, P, Q, act1, alpha1, st);
      TLXCV_DW_CHANNELS(TLXCV_X)
#undef TLXCV_X
    }
    if (stride == 1) {
      constexpr int TH = 4, TW = 2;
      const int PT = (P + TH - 1) / TH, QT = (Q + TW - 1) / TW;
      const size_t total = static_cast<size_t>(N) * PT * QT * (C / 4);
      TLXCV_LAUNCH((dwconv3x3_nhwc_kernel<1, TH, TW>), blocks_for(total, 128), 128, 0, st, 
          x, wp, y, scale, shift, rp, H, W, C / 4, P, Q, PT, QT, act1, alpha1, act2, alpha2, total);
    } else {
      constexpr int TH = 2, TW = 2;
      const int PT = (P + TH - 1) / TH, QT = (Q + TW - 1) / TW;
      const size_t total = static_cast<size_t>(N) * PT * QT * (C / 4);
      TLXCV_LAUNCH((dwconv3x3_nhwc_kernel<2, TH, TW>), blocks_for(total, 128), 128, 0, st, 
          x, wp, y, scale, shift, rp, H, W, C / 4, P, Q, PT, QT, act1, alpha1, act2, alpha2, total);
    }
    return cudaGetLastError();
  }
  const size_t total = static_cast<size_t>(N) * P * Q * (C / 8);
  if (is_f32)
    TLXCV_LAUNCH(dwconv_nhwc_kernel<float>, blocks_for(total), kThreads, 0, st, 
        static_cast<const float*>(src), static_cast<const float*>(w_rsc), static_cast<float*>(dst), scale, shift,
        static_cast<const float*>(residual), H, W, C / 8, P, Q, R, S, stride, pad, act1, alpha1, act2, alpha2, total);
  else
    TLXCV_LAUNCH(dwconv_nhwc_kernel<__nv_bfloat16>, blocks_for(total), kThreads, 0, st, 
        static_cast<const __nv_bfloat16*>(src), static_cast<const __nv_bfloat16*>(w_rsc),
        static_cast<__nv_bfloat16*>(dst), scale, shift, static_cast<const __nv_bfloat16*>(residual), H, W, C / 8, P, Q,
        R, S, stride, pad, act1, alpha1, act2, alpha2, total);
  return cudaGetLastError();
}

cudaError_t add_act(const void* a, const void* b, void* dst, size_t n, int act, float alpha, int is_f32, cudaStream_t st) {
  if (n % 8) return cudaErrorInvalidValue;
  const size_t n8 = n / 8;
  if (is_f32)
    TLXCV_LAUNCH(add_act_kernel<float>, blocks_for(n8), kThreads, 0, st, static_cast<const float*>(a), static_cast<const float*>(b), static_cast<float*>(dst), n8, act, alpha);
  else
    TLXCV_LAUNCH(add_act_kernel<__nv_bfloat16>, blocks_for(n8), kThreads, 0, st, static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(dst), n8, act, alpha);
  return cudaGetLastError();
}

cudaError_t upsample_concat(const void* a, const void* b, void* dst, int N, int H, int W, int C0, int C1, int ra, int rb,
                            int is_f32, cudaStream_t st) {
  if (C0 % 8 || C1 % 8 || ra < 1 || rb < 1 || H % ra || W % ra || H % rb || W % rb || (b == nullptr && C1 != 0)) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(N) * H * W * ((C0 + C1) / 8);
  if (is_f32)
    TLXCV_LAUNCH(upsample_concat_kernel<float>, blocks_for(total), kThreads, 0, st, static_cast<const float*>(a), static_cast<const float*>(b), static_cast<float*>(dst), H, W, C0 / 8, C1 / 8, ra, rb, total);
  else
    TLXCV_LAUNCH(upsample_concat_kernel<__nv_bfloat16>, blocks_for(total), kThreads, 0, st, static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(dst), H, W, C0 / 8, C1 / 8, ra, rb, total);
  return cudaGetLastError();
}

cudaError_t argmax_rows(const float* logits, long long* dst, int N, int K, cudaStream_t st) {
  const int warps_per_block = 8;
  TLXCV_LAUNCH(argmax_rows_kernel, (N + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st, logits, dst, N, K);
  return cudaGetLastError();
}

cudaError_t softmax_rows(const float* logits, float* dst, int N, int K, cudaStream_t st) {
  const int warps_per_block = 8;
  TLXCV_LAUNCH(softmax_rows_kernel, (N + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st, logits, dst, N, K);
  return cudaGetLastError();
}

cudaError_t softmax_ce(const float* logits, const long long* target, float* row_loss, unsigned int* ticket, float* dst, int N, int K,
                       cudaStream_t st) {
  const int warps_per_block = 8;
  TLXCV_LAUNCH(softmax_ce_kernel, (N + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st, logits, target, row_loss,
               ticket, dst, N, K);
  return cudaGetLastError();
}

cudaError_t conv_direct_f32(const float* in, const float* w_rsck, float* out, const float* scale, const float* shift,
                            const float* residual, int N, int H, int W, int C, int P, int Q, int K, int R, int S,
                            int stride, int pad, int dil, int groups, int act1, float alpha1, int act2, float alpha2,
                            cudaStream_t st) {
  const size_t total = static_cast<size_t>(N) * P * Q * K;
  const int Cs = C <= 4 ? 4 : C;  // stems read the NHWC4 import
  conv_direct_f32_kernel<<<blocks_for(total), kThreads, 0, st>>>(in, w_rsck, out, scale, shift, residual, H, W, C, Cs, P,
                                                                  Q, K, R, S, stride, pad, dil, groups, act1, alpha1,
                                                                  act2, alpha2, total);
  return cudaGetLastError();
}

}  // namespace tlxcv
