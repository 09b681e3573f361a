// Host-side declarations of the kernel launchers (internal to libtlxcv_b200.so).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/tlxcv_b200.h"
#include "env.h"

namespace tlxcv {

// ---- implicit-GEMM conv on tcgen05 -------------------------------------------------------------
enum ConvMode : int {
  kModeTiled = 0,   // 1x1 stride-1: A = [M][C] plain 2-D TMA tiles
  kModeIm2col = 1,  // general RxS / stride / pad: A tiles by im2col-mode TMA from NHWC
  kModeGatherC4 = 2, // C_in <= 4 stems: producer warps gather from NHWC4 into the swizzled A tile
  kModeSlabDense = 3, // weight packing only: K = taps x exactly C_in channels (conv3x3_slab.cu, dense layers)
  kModePixelPairs = 4 // weight packing only: 3x3 / stride-2 conv over 32 channels as 6 K blocks of pixel pairs (kLayoutPixelPairs)
};

// K layout of a dense im2col conv (tc_conv_layout): how the filter taps and channels map onto K blocks
enum TcConvLayout {
  kLayoutPlain = 0,      // K block = 64 channels of one tap (a tap with fewer channels is zero padded)
  kLayoutKb32 = 1,       // C_in == 32, C_out <= 64: K block = the 32 channels of one tap (64-byte rows, SWIZZLE_64B)
  kLayoutPixelPairs = 2  // C_in == 32, 3x3 / stride 2 / pad 1 on even H, W: the input viewed as (N, H/2, [row parity], W/2, 64):
                         // two horizontally adjacent pixels form one 64-channel "pixel pair", so the taps s = 1, 2 of a filter
                         // row are ONE 128-byte-row K block (stride 1 in pair space) and s = 0 is the upper half of the
                         // pair to the left: 6 K blocks of full 128-byte rows instead of 9 half-empty ones
};

struct ConvKernelParams {
  int M, Cout;
  int num_kb;      // 64-wide K blocks per output tile
  int kb_per_tap;  // K blocks per filter tap (ceil(Cin/64)); taps = num_kb / kb_per_tap
  int S;           // filter width (tap -> (r, s))
  int P, Q, H, W;  // output / input spatial size
  int stride, pad, dil;
  int m_tiles, n_tiles;
  int a_chan_from_n;  // grouped conv with 64-channel block-diagonal weights: A channel base = n_tile * 64
  unsigned pair_taps; // kLayoutPixelPairs: per K block a nibble {8 valid | 4 row offset | 2 pair offset | 1 odd-row map}; 0 otherwise
  // epilogue: y = act2(act1(acc * scale + shift) + residual)
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  void* out;
  int act1;
  float alpha1;
  int act2;
  float alpha2;
  int out_f32;
  // gather mode
  const __nv_bfloat16* in_c4;
  int R, KR;  // filter height; K elements reserved per filter row (16 or 32)
  // shared-memory configuration chosen per layer
  int ablate;  // debug: TLXCV_DEBUG_ABLATE bit mask (timing experiments; 0 in normal operation)
  int stages, ring;  // operand pipeline slots (barrier pairs); epilogue store/residual ring depth per warp (2 or 4)
  int kgroup;        // K blocks per pipeline slot: 1, or 2 for 64-wide tiles with >= 4 K blocks (0 = 1)
  int sc_bufs;       // scale/shift smem buffers: 1 (filled once, or unused) or 2 (refreshed per tile)
  int epi_warps;     // epilogue warps taking part: 8 or 16
  // dual-accumulator launches (conv3 + downsample conv of a stage's first block in one kernel)
  int num_kb1;       // K blocks of the first GEMM (the remaining num_kb - num_kb1 belong to the second)
  int a2_im2col;     // second A operand: 0 = plain [M][C2] matrix, 1 = im2col-mode TMA (strided 1x1)
  const float* scale2;
  const float* shift2;
  unsigned long long* trace;  // debug: TLXCV_DEBUG_TRACE_CONV timeline buffer (NULL in normal operation)
  // argmax over the output channels fused into an fp32-output (Linear) launch: every epilogue thread folds its 32 columns
  // into amax_keys[row] with one 64-bit atomicMax of (order-preserving float bits << 32 | ~column), the last CTA to finish
  // decodes the keys into amax_out[row] (first maximal index, like torch.argmax) and resets keys and ticket for the next launch
  unsigned long long* amax_keys;
  long long* amax_out;
  unsigned int* amax_ticket;
};

struct TcConvLaunch {
  CUtensorMap tmapA, tmapB, tmapOut, tmapRes, tmapA2, tmapB2;
  ConvKernelParams p;
  int mode, block_n, grid, threads, smem, dual;
  int two;  // launched as CTA pairs (cluster of 2, tcgen05.mma.cta_group::2)
  int kblock;    // channels per K block: 64, or 32 for 32-input-channel layers (0 = 64)
  int chain_n1;  // > 0: chain launch (conv -> 1x1 conv in one kernel), width of the first conv (64 or 128)
  int chain_w1res;  // chain launch keeps the first conv's weights resident in shared memory
};

// Encodes the TMA descriptors and picks tile shape; returns an empty string or an error message.
std::string tc_conv_prepare(TcConvLaunch& L, int sm_count, const __nv_bfloat16* act_in, int N, int H, int W, int Cin,
                            int Cin_storage, const __nv_bfloat16* packed_w, int Ktot, int Cout, int R, int S, int stride,
                            int pad, int dil, int groups, int force_block_n, void* out_bf16, int Cout_storage, const void* residual_bf16);
// out = act( A1[M][K1] * W1^T * scale + shift  +  conv1x1_stride(A2) * W2^T * scale2 + shift2 ): two accumulators per tile
std::string tc_conv_prepare_dual(TcConvLaunch& L, int sm_count, const __nv_bfloat16* a1, int M, int K1,
                                 const __nv_bfloat16* w1, const __nv_bfloat16* a2, int N, int H2, int W2, int C2, int stride2,
                                 const __nv_bfloat16* w2, int Cout, void* out_bf16);
// out = act2( bn2(conv1x1( relu(bn1(conv_RxS(in))) )) + residual ): first conv N1 = 64 / 128 channels wide, its output never leaves
// the SM (conv_chain_kernel).  w1: im2col packing [256][K1tot]; w2: [N2 padded to 256][N1].
std::string tc_chain_prepare(TcConvLaunch& L, int sm_count, const __nv_bfloat16* act_in, int N, int H, int W, int Cin,
                             const __nv_bfloat16* w1, int K1tot, int N1, int R, int S, int stride, int pad, int dil,
                             const __nv_bfloat16* w2, int N2, void* out_bf16, const void* residual_bf16);
bool tc_chain_supported(int Cin, int N1, int N2);
cudaError_t tc_conv_launch(const TcConvLaunch& L, cudaStream_t stream);
cudaError_t tc_conv_set_attributes();
// number of K elements per output channel in the packed weight matrix for this geometry
int tc_conv_packed_k(int Cin, int R, int S, int groups, int mode, int layout = kLayoutPlain);
// K layout for a dense conv of this geometry (kLayoutPlain for everything but 32-input-channel im2col layers)
int tc_conv_layout(int Cin, int Cout, int R, int S, int stride, int pad, int dil, int groups, int H, int W);
int tc_conv_mode(int Cin, int R, int S, int stride, int pad, int groups);

// ---- stem conv (C_in <= 4) as a row-ring implicit GEMM, optional fused 3x3/s2/p1 max-pool ------------
struct StemGeometry {
  int pairs;    // stride-1 stems run over pairs of output pixels (GEMM N = 2 * C_out)
  int block_n;  // GEMM N: 32 or 64
  int P, Q;     // conv output size
  int Qw;       // A rows (windows) per output row: Q, or Q/2 in pair mode
  int pad_l;    // zero columns stored left of each input row (even, >= pad)
  int xoff;     // window position of filter column 0 (pad_l - pad)
  int Wp;       // stored row pitch in pixels (input is [N][H][Wp][4] bf16)
};

struct StemParams {
  int N, H, P, Qw;
  int R, sv, pad;
  int bands, band_rows, q_tiles;
  int pool, Pp, Qp;
  // step = 2 output rows; ring of NG groups of G = 2*sv input-row slices; a step's window spans n_in
  // slices = ng groups; stationary weight chain of nblk blocks, bblk[i] = first block of the stacked
  // B operand of input-row position i
  int G, n_in, ng, NG, slice_bytes, nblk;
  int bblk[9];
  const float* scale;  // [block_n] (pair mode: the per-channel values twice)
  const float* shift;
  int act;
  float alpha;
  const __nv_bfloat16* in;  // padded input [N][H][Wp][4] (L2 prefetch addresses; the operand loads go through tmapA)
  int Wp;
  __nv_bfloat16* out;  // [N][P][Qw][block_n] == NHWC, or the pooled map [N][Pp][Qp][block_n]
  int ablate;          // debug: TLXCV_DEBUG_ABLATE_STEM bit mask (timing experiments; 0 in normal operation)
};

struct StemLaunch {
  CUtensorMap tmapA, tmapB, tmapOut;
  StemParams p;
  int block_n, grid, threads, smem;
};

// false when the geometry is not one the row-ring kernel covers (the gather path of conv_tcgen05 is used instead)
bool stem_rowring_geometry(StemGeometry& g, int Cin, int Cout, int H, int W, int R, int S, int stride, int pad, int dil,
                           int groups);
std::string stem_rowring_prepare(StemLaunch& L, int sm_count, const StemGeometry& g, const __nv_bfloat16* in_padded, int N,
                                 int H, int R, int stride, int pad, const __nv_bfloat16* packed_w, void* out, int pool,
                                 int Pp, int Qp);
cudaError_t stem_rowring_launch(const StemLaunch& L, cudaStream_t st);
cudaError_t stem_rowring_set_attributes();
// number of bf16 elements of the packed (chained) stem weights
int stem_rowring_weight_elems(const StemGeometry& g, int R, int stride);
cudaError_t pack_stem_weights(const float* oihw, __nv_bfloat16* dst, int Cout, int Cin, int R, int S, int stride,
                              const StemGeometry& g, cudaStream_t st);
// NCHW fp32 (C <= 4) -> [N][H][Wp][4] bf16, pixel w at column w + pad_l, zero elsewhere
cudaError_t import_nchw_c4_padded(const float* src, void* dst, int N, int C, int H, int W, int Wp, int pad_l, cudaStream_t st);

// ---- 3x3 stride-1 conv with 64 input channels as a slab implicit GEMM (conv3x3_slab.cu) -----------
struct SlabParams {
  int N, H, W, Cout;
  int cblocks;           // 64-channel blocks worked on independently (grouped conv), 1 for a dense layer
  int segs, Ws;          // column segments per row and their width (Ws + 2 <= 128 slab pixels with the halo)
  int T;                 // output rows per step = floor(128 / (Ws + 2))
  int ng;                // row groups a step's window spans: 1 + ceil(2 / T)
  int NG;                // ring depth in groups of T input rows
  int bands, band_rows;  // bands per image; output rows per band
  __nv_bfloat16* out;       // NHWC output (all Cout channels; an item writes its 64-channel block)
  const __nv_bfloat16* residual;  // NHWC, Cout channels, or NULL:  y = act2(act(acc * scale + shift) + residual)
  const float* scale;
  const float* shift;
  int act;
  float alpha;
  int act2;
  float alpha2;
  int ablate;  // debug: TLXCV_DEBUG_ABLATE_SLAB bit mask (timing experiments; 0 in normal operation)
  unsigned long long* trace;  // debug: TLXCV_DEBUG_TRACE_SLAB timeline buffer (NULL in normal operation)
};

struct SlabLaunch {
  CUtensorMap tmapA, tmapB;
  SlabParams p;
  int block_n, cb, grid, threads, smem;
};

bool conv3x3_slab_supported(int Cin, int Cout, int H, int W, int R, int S, int stride, int pad, int dil, int groups);
// packed_w: [Cout_pad][9 x CB] with K order (tap, channel of the block): dense layers CB = Cin (kModeSlabDense packing),
// grouped layers CB = 64 with the block-diagonal expansion of the grouped packing
std::string conv3x3_slab_prepare(SlabLaunch& L, int sm_count, const __nv_bfloat16* in, int N, int H, int W, int Cin, int Cout,
                                 int groups, const __nv_bfloat16* packed_w, int Ktot, void* out, const void* residual);
cudaError_t conv3x3_slab_launch(const SlabLaunch& L, cudaStream_t st);
cudaError_t conv3x3_slab_set_attributes();

// ---- weight / BN preparation -------------------------------------------------------------------
// OIHW fp32 -> [Cout_pad][Ktot] bf16 in the K order the conv kernel consumes.
cudaError_t pack_conv_weights(const float* oihw, __nv_bfloat16* dst, int Cout, int Cout_pad, int Cin, int R, int S,
                              int groups, int mode, int Ktot, cudaStream_t st);
// (in,out) fp32 -> [Kout_pad][F] bf16
cudaError_t pack_linear_weights(const float* w_in_out, __nv_bfloat16* dst, int F, int Kout, int Kout_pad, cudaStream_t st);
// OIHW fp32 -> [R][S][C/g][K] fp32 (validation path) ; (in,out) linear is already [F][K]
cudaError_t pack_conv_weights_f32(const float* oihw, float* dst, int Cout, int Cg, int R, int S, cudaStream_t st);
// depthwise: OIHW [C][1][R][S] -> [R*S][C] (T = bf16 or fp32)
cudaError_t pack_dw_weights(const float* oihw, void* dst, int C, int RS, int is_f32, cudaStream_t st);
// scale = gamma / sqrt(var + eps), shift = beta + (bias - mean) * scale   (any pointer may be NULL)
cudaError_t fold_bn(float* scale, float* shift, const float* gamma, const float* beta, const float* mean,
                    const float* var, const float* bias, float eps, int K, int K_pad, cudaStream_t st);

// ---- memory-bound kernels (T = __nv_bfloat16 or float, is_f32 selects) -------------------------
cudaError_t import_nchw(const float* src, void* dst, int N, int C, int H, int W, int Cs, int is_f32, cudaStream_t st);
// src rows hold Cs >= C channels (Cs == C except for maps whose channel count is not a multiple of 8)
cudaError_t export_nchw(const void* src, float* dst, int N, int C, int Cs, int H, int W, int is_f32, cudaStream_t st);
// NHWC uint8 (C <= 4) -> [N][H][Wp][4] activations, (x - mean[c]) / std[c]; Wp == W, pad_l == 0 for the dense layout
cudaError_t import_u8_nhwc(const uint8_t* src, void* dst, const float* mean, const float* stdv, int N, int C, int H, int W,
                           int Wp, int pad_l, int is_f32, cudaStream_t st);
// the same with OpenCV's 8-bit INTER_LINEAR resize (Hs, Ws) -> (H, W) in front; tx[W] / ty[H] = {i0, i1, c0, c1} (resize_tables)
cudaError_t import_u8_resize(const uint8_t* src, void* dst, const float* mean, const float* stdv, const int4* tx, const int4* ty,
                             int N, int C, int Hs, int Ws, int H, int W, int Wp, int pad_l, int is_f32, cudaStream_t st);
cudaError_t maxpool_nhwc(const void* src, void* dst, int N, int H, int W, int C, int P, int Q, int k, int stride, int pad,
                         int is_f32, cudaStream_t st);
// AvgPool2d(k, stride), no padding (windows inside the map)
cudaError_t splat_apply(const void* x, const void* att, void* dst, int N, int HW, int C, int radix, int cardinality, int is_f32,
                        cudaStream_t st);
cudaError_t avgpool_nhwc(const void* src, void* dst, int N, int H, int W, int C, int P, int Q, int k, int stride, int pad, int is_f32,
                         cudaStream_t st);
cudaError_t gap_nhwc(const void* src, void* dst, int N, int HW, int C, int is_f32, cudaStream_t st);
cudaError_t dwconv_nhwc(const void* src, const void* w_rsc, void* dst, const float* scale, const float* shift,
                        const void* residual, int N, int H, int W, int C, int P, int Q, int R, int S, int stride, int pad,
                        int act1, float alpha1, int act2, float alpha2, int is_f32, cudaStream_t st);
cudaError_t add_act(const void* a, const void* b, void* dst, size_t n, int act, float alpha, int is_f32, cudaStream_t st);
cudaError_t argmax_rows(const float* logits, long long* dst, int N, int K, cudaStream_t st);
// probabilities of fp32 logits (N, K), row-wise
cudaError_t softmax_rows(const float* logits, float* dst, int N, int K, cudaStream_t st);
// mean softmax cross-entropy of (N, K) fp32 logits against int64 labels -> dst[0]; row_loss: N floats of scratch, ticket: one
// zero-initialised counter (the kernel leaves it zero)
cudaError_t softmax_ce(const float* logits, const long long* target, float* row_loss, unsigned int* ticket, float* dst, int N, int K,
                       cudaStream_t st);
// dst[N][H][W][C0 + C1]: channels [0, C0) = a nearest-up-sampled ra times, [C0, C0 + C1) = b up-sampled rb times (b may be NULL)
cudaError_t upsample_concat(const void* a, const void* b, void* dst, int N, int H, int W, int C0, int C1, int ra, int rb,
                            int is_f32, cudaStream_t st);
// fp32 direct conv on CUDA cores (validation mode; dense, grouped and depthwise)
cudaError_t conv_direct_f32(const float* in, const float* w_rsck, float* out, const float* scale, const float* shift,
                            const float* residual, int N, int H, int W, int C, int P, int Q, int K, int R, int S,
                            int stride, int pad, int dil, int groups, int act1, float alpha1, int act2, float alpha2,
                            cudaStream_t st);

}  // namespace tlxcv
