// Stem convolution (C_in <= 4) for sm_100a as a ROW-RING implicit GEMM, with the 3x3/s2 max-pool
// of the ResNet / ResNeXt stem fused into its epilogue.
//
// Replaces  nn.GroupConv2d(7,2,3)/(3,2,1)/(3,1,1) on the 3-channel input + BatchNorm2d + ReLU/ReLU6/
// LeakyReLU (+ nn.MaxPool2d(3,2,1))  of  classification/resnet.py:199-218,287-290,
// resnext.py:151-164,201-203, mobilenetv2.py:82-84 (ops_fusion.py:39-48), mobilenetv1.py:124-132,
// detection/backbones/darknet.py:250-260, classification/darknet53.py:64-66.
//
// Why a separate kernel: with C_in = 3 the K dimension of the implicit GEMM is tiny per tap, and
// building 128 x 64 im2col tiles element-wise (the gather path in conv_tcgen05.cu) is bound by
// address arithmetic.  Here the input is stored NHWC4 with zero-padded columns ([N][H][Wp][4] bf16,
// pixel w at column w + pad_l) so that
//   * the 8 consecutive pixels (32 bf16 = 64 B) an output pixel needs from ONE input row start at a
//     16-byte aligned address 16*q  ->  a TMA tensor map with an OVERLAPPING dimension
//     (dim0 = 32 elements, dim1 = q with stride 16 B) delivers, for one input row h, the complete
//     [128 windows][64 B] A operand slice in ONE bulk copy, already in the SWIZZLE_64B K-major layout
//     tcgen05.mma consumes;
//   * that slice is the A operand of filter row r for EVERY output row p with p*sv - pad + r == h, so
//     a persistent CTA that walks down a band of output rows keeps a ring of input-row slices in
//     shared memory and fetches each input row once (sv new rows per output row instead of R).
// Stride-1 stems (DarkNet) are run as stride-2 over PAIRS of output pixels: the GEMM N dimension
// holds (pixel parity e, output channel), the weights of parity e are shifted by e pixels inside the
// 8-pixel window.  Their output row [Q/2 pairs][2*C_out] is byte-identical to NHWC [Q][C_out].
//
// GEMM per output row:  D[128 windows][BLOCK_N] = sum_r  A_h(r)[128][32] * B_r[BLOCK_N][32]^T
// (two K=16 MMAs per filter row), weights B stationary in shared memory for the whole kernel,
// D in TMEM (4 accumulators in flight).
// Epilogue (8 warps): tcgen05.ld -> fp32 scale/shift -> activation -> bf16 -> swizzled row buffer in
// shared memory (ring of 4 rows) -> either a coalesced 16 B/thread copy of the row to HBM, or (pool)
// the 3x3/s2/p1 maximum over the last three conv rows, written as the pooled row.  The conv
// activation map never reaches HBM in the pooled variant.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace tlxcv {

namespace {

constexpr int kTileM = 128;                 // windows (A rows) per MMA
constexpr int kSliceBytes = kTileM * 64;    // one input-row slice: 128 windows x 64 B
constexpr int kRingSlots = 14;              // input-row slices resident per CTA
constexpr int kMaxR = 7;
constexpr int kAccBufs = 4;
constexpr int kRowSlots = 4;                // conv-row buffers (bf16) for the epilogue / pool
constexpr int kEpiWarpsS = 8;
constexpr int kThreadsS = (2 + kEpiWarpsS) * 32;  // 320

template <int BLOCK_N>
struct SCfg {
  static constexpr int kRowBytes = BLOCK_N * 2;                 // bytes per window in the output row
  static constexpr int kBBytes = kMaxR * BLOCK_N * 64;          // stationary weights
  static constexpr int kRowBuf = kTileM * kRowBytes;            // one conv-row buffer
  static constexpr int kSmem = kRingSlots * kSliceBytes + kBBytes + kRowSlots * kRowBuf + 2 * BLOCK_N * 4 + 512;
};

__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;           // LBO: ignored for swizzled K-major
  d |= static_cast<uint64_t>(512 >> 4) << 32;    // SBO: 8 rows x 64 B
  d |= static_cast<uint64_t>(1) << 46;           // descriptor version 1 (sm_100)
  d |= static_cast<uint64_t>(4) << 61;           // SWIZZLE_64B
  return d;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// One work item = (image, band of output rows, 128-window column tile).
struct Item {
  int n, qt;
  int c0, c1;  // conv rows [c0, c1)
  int j0, j1;  // pooled rows [j0, j1) (pool only)
  int h0, h1;  // input rows [h0, h1) fetched for the band
};

__device__ __forceinline__ Item decode_item(const StemParams& p, int item) {
  Item it;
  it.qt = item % p.q_tiles;
  int t = item / p.q_tiles;
  const int b = t % p.bands;
  it.n = t / p.bands;
  if (p.pool) {
    it.j0 = b * p.band_rows;
    it.j1 = min(p.Pp, it.j0 + p.band_rows);
    it.c0 = max(0, 2 * it.j0 - 1);
    it.c1 = min(p.P, 2 * it.j1);
  } else {
    it.j0 = it.j1 = 0;
    it.c0 = b * p.band_rows;
    it.c1 = min(p.P, it.c0 + p.band_rows);
  }
  it.h0 = max(0, it.c0 * p.sv - p.pad);
  it.h1 = min(p.H, (it.c1 - 1) * p.sv - p.pad + p.R);
  return it;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreadsS, 1)
stem_rowring_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const StemParams p) {
  using C = SCfg<BLOCK_N>;
  constexpr int kRowBytes = C::kRowBytes;
  constexpr int kChunks = kRowBytes / 16;  // 16-byte chunks per window in the output row (4 or 8)
  constexpr int kNc = BLOCK_N / 2;         // accumulator columns per epilogue warp
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* ring = smem;
  uint8_t* bsm = ring + kRingSlots * kSliceBytes;
  uint8_t* rows = bsm + C::kBBytes;
  float* sc_s = reinterpret_cast<float*>(rows + kRowSlots * C::kRowBuf);
  float* sh_s = sc_s + BLOCK_N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sh_s + BLOCK_N);
  uint64_t* full_bar = bars;                      // [kRingSlots] slice landed
  uint64_t* empty_bar = bars + kRingSlots;        // [kRingSlots] MMAs reading the slice retired
  uint64_t* tfull_bar = bars + 2 * kRingSlots;    // [kAccBufs]
  uint64_t* tempty_bar = tfull_bar + kAccBufs;    // [kAccBufs]
  uint64_t* b_bar = tempty_bar + kAccBufs;        // weights landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.N * p.bands * p.q_tiles;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    for (int i = 0; i < kRingSlots; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < kAccBufs; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), kEpiWarpsS);
    }
    mbar_init(smem_u32(b_bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<kAccBufs * BLOCK_N>(smem_u32(tmem_ptr_smem));
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BLOCK_N; i += kEpiWarpsS * 32) {
      sc_s[i] = p.scale[i];
      sh_s[i] = p.shift[i];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one slice per input row =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(b_bar), p.R * BLOCK_N * 64);
      for (int r = 0; r < p.R; ++r) tma_load_2d(smem_u32(bsm + r * BLOCK_N * 64), &tmapB, smem_u32(b_bar), r * 32, 0);
      uint32_t slot = 0, phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        for (int h = it.h0; h < it.h1; ++h) {
          mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1);
          const uint32_t bar = smem_u32(&full_bar[slot]);
          if (p.ablate & 1) {
            mbar_arrive(bar);
          } else {
            mbar_arrive_expect_tx(bar, kSliceBytes);
            tma_load_4d(smem_u32(ring + slot * kSliceBytes), &tmapA, bar, 0, it.qt * kTileM, h, it.n);
          }
          if (++slot == kRingSlots) slot = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N);
      mbar_wait(smem_u32(b_bar), 0);
      tcgen05_fence_after();
      uint32_t slot0 = 0;                     // ring slot of input row it.h0
      uint32_t wslot = 0, wphase = 0;         // next slice to wait for
      uint32_t acc = 0, acc_phase = 0;
      const uint32_t ring_addr = smem_u32(ring), b_addr = smem_u32(bsm);
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        int landed = it.h0;  // input rows [h0, landed) are known to be in the ring
        int freed = it.h0;   // input rows [h0, freed) have been handed back to the producer
        for (int c = it.c0; c < it.c1; ++c) {
          mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
          const int top = c * p.sv - p.pad;
          const int need = min(it.h1, top + p.R);
          while (landed < need) {
            mbar_wait(smem_u32(&full_bar[wslot]), wphase);
            if (++wslot == kRingSlots) wslot = 0, wphase ^= 1;
            ++landed;
          }
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
          uint32_t accumulate = 0;
          for (int r = 0; r < p.R; ++r) {
            const int h = top + r;
            if (h < 0 || h >= p.H) continue;  // zero padding rows contribute nothing
            uint32_t s = slot0 + static_cast<uint32_t>(h - it.h0);
            s -= (s / kRingSlots) * kRingSlots;
            const uint64_t adesc = make_kmajor_sw64_desc(ring_addr + s * kSliceBytes);
            const uint64_t bdesc = make_kmajor_sw64_desc(b_addr + r * BLOCK_N * 64);
            if (p.ablate & 2) continue;
            umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
            umma_bf16(tmem_d, adesc + 2, bdesc + 2, idesc, 1);  // second K=16 step: +32 B inside the swizzle atom
            accumulate = 1;
          }
          umma_commit(smem_u32(&tfull_bar[acc]));
          if (++acc == kAccBufs) acc = 0, acc_phase ^= 1;
          // input rows above the next output row's window are dead; the last output row frees the band
          const int free_to = (c == it.c1 - 1) ? it.h1 : min(it.h1, max(it.h0, (c + 1) * p.sv - p.pad));
          while (freed < free_to) {
            uint32_t s = slot0 + static_cast<uint32_t>(freed - it.h0);
            s -= (s / kRingSlots) * kRingSlots;
            umma_commit(smem_u32(&empty_bar[s]));
            ++freed;
          }
        }
        slot0 = (slot0 + static_cast<uint32_t>(it.h1 - it.h0)) % kRingSlots;
      }
    }
  } else {
    // ===================== epilogue: 8 warps =====================
    const int lg = warp & 3, ch = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;              // 0..255
    const int px = lg * 32 + lane;                // window (TMEM lane) this thread owns
    const uint32_t swz = kChunks == 8 ? (px & 7) : ((px >> 1) & 3);
    const uint32_t rows_addr = smem_u32(rows);
    const float alpha = p.alpha;
    uint32_t acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item it = decode_item(p, item);
      const int valid_w = min(kTileM, p.Qw - it.qt * kTileM);  // windows of this tile that exist
      epi_bar_sync();  // the previous band's last row buffers are no longer being read
      for (int c = it.c0; c < it.c1; ++c) {
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tcgen05_fence_after();
        uint32_t v[kNc];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + acc * BLOCK_N + ch * kNc;
        if constexpr (kNc == 32)
          tmem_ld_32x32b_x32(taddr, v);
        else
          tmem_ld_32x32b_x16(taddr, v);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
        if (++acc == kAccBufs) acc = 0, acc_phase ^= 1;
        if (p.ablate & 4) continue;

        const uint32_t slot_addr = rows_addr + static_cast<uint32_t>((c + 1) & (kRowSlots - 1)) * C::kRowBuf;
        const uint32_t my_row = slot_addr + px * kRowBytes;
#pragma unroll
        for (int j = 0; j < kNc / 8; ++j) {
          float f[8];
          const float4 s0 = *reinterpret_cast<const float4*>(sc_s + ch * kNc + 8 * j);
          const float4 s1 = *reinterpret_cast<const float4*>(sc_s + ch * kNc + 8 * j + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(sh_s + ch * kNc + 8 * j);
          const float4 h1 = *reinterpret_cast<const float4*>(sh_s + ch * kNc + 8 * j + 4);
          f[0] = fmaf(__uint_as_float(v[8 * j + 0]), s0.x, h0.x);
          f[1] = fmaf(__uint_as_float(v[8 * j + 1]), s0.y, h0.y);
          f[2] = fmaf(__uint_as_float(v[8 * j + 2]), s0.z, h0.z);
          f[3] = fmaf(__uint_as_float(v[8 * j + 3]), s0.w, h0.w);
          f[4] = fmaf(__uint_as_float(v[8 * j + 4]), s1.x, h1.x);
          f[5] = fmaf(__uint_as_float(v[8 * j + 5]), s1.y, h1.y);
          f[6] = fmaf(__uint_as_float(v[8 * j + 6]), s1.z, h1.z);
          f[7] = fmaf(__uint_as_float(v[8 * j + 7]), s1.w, h1.w);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = apply_act(f[e], p.act, alpha);
          const uint32_t cidx = static_cast<uint32_t>(ch * (kNc / 8) + j);
          const uint32_t addr = my_row + ((cidx ^ swz) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(f[0], f[1])),
                       "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])), "r"(pack_bf16x2(f[6], f[7]))
                       : "memory");
        }
        epi_bar_sync();  // conv row c is complete in its buffer
        if (p.ablate & 8) continue;
        if (!p.pool) {
          // copy the row out: consecutive threads -> consecutive 16 B of the NHWC row
          uint8_t* grow = reinterpret_cast<uint8_t*>(p.out) +
                          ((static_cast<size_t>(it.n) * p.P + c) * p.Qw + static_cast<size_t>(it.qt) * kTileM) * kRowBytes;
          for (int i = et; i < valid_w * kChunks; i += kEpiWarpsS * 32) {
            const uint32_t w = static_cast<uint32_t>(i) / kChunks, cidx = static_cast<uint32_t>(i) % kChunks;
            const uint32_t wswz = kChunks == 8 ? (w & 7) : ((w >> 1) & 3);
            uint4 val;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                         : "r"(slot_addr + w * kRowBytes + ((cidx ^ wswz) << 4)));
            *reinterpret_cast<uint4*>(grow + static_cast<size_t>(i) * 16) = val;
          }
        } else {
          // pooled row j = max over conv rows 2j-1, 2j, 2j+1 and columns 2i-1, 2i, 2i+1 (padding = -inf:
          // positions outside the map are skipped; the centre (2j, 2i) always exists)
          const bool completes = (c & 1) ? true : (c == p.P - 1);
          const int j = c >> 1;
          if (completes && j >= it.j0 && j < it.j1) {
            uint8_t* grow = reinterpret_cast<uint8_t*>(p.out) + (static_cast<size_t>(it.n) * p.Pp + j) * p.Qp * kRowBytes;
            for (int i = et; i < p.Qp * kChunks; i += kEpiWarpsS * 32) {
              const int qo = i / kChunks;
              const uint32_t cidx = static_cast<uint32_t>(i % kChunks);
              __nv_bfloat162 m[4];
              bool have = false;
#pragma unroll
              for (int dr = -1; dr <= 1; ++dr) {
                const int cr = 2 * j + dr;
                if (cr < 0 || cr >= p.P) continue;
                const uint32_t sa = rows_addr + static_cast<uint32_t>((cr + 1) & (kRowSlots - 1)) * C::kRowBuf;
#pragma unroll
                for (int dc = -1; dc <= 1; ++dc) {
                  const int w = 2 * qo + dc;
                  if (w < 0 || w >= p.Qw) continue;
                  const uint32_t wswz = kChunks == 8 ? (w & 7) : ((w >> 1) & 3);
                  uint4 val;
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                               : "r"(sa + w * kRowBytes + ((cidx ^ wswz) << 4)));
                  const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&val);
                  if (!have) {
                    m[0] = hv[0], m[1] = hv[1], m[2] = hv[2], m[3] = hv[3];
                    have = true;
                  } else {
                    m[0] = __hmax2(m[0], hv[0]), m[1] = __hmax2(m[1], hv[1]);
                    m[2] = __hmax2(m[2], hv[2]), m[3] = __hmax2(m[3], hv[3]);
                  }
                }
              }
              *reinterpret_cast<uint4*>(grow + static_cast<size_t>(i) * 16) = *reinterpret_cast<const uint4*>(m);
            }
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<kAccBufs * BLOCK_N>(tmem_base);
  }
}

// weights OIHW fp32 -> [BLOCK_N][R][8 window pixels][4] bf16; row o' = (parity e, output channel) for
// pair mode; window position x holds filter column s = x - xoff - e
__global__ void pack_stem_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin,
                                         int R, int S, int block_n, int pairs, int xoff) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int Ktot = R * 32;
  if (idx >= block_n * Ktot) return;
  const int o = idx / Ktot, k = idx % Ktot;
  const int r = k / 32, x = (k % 32) / 4, c = k % 4;
  const int e = pairs ? o / Cout : 0, co = pairs ? o % Cout : o;
  const int s = x - xoff - e;
  float v = 0.0f;
  if (co < Cout && e < 2 && c < Cin && s >= 0 && s < S) v = w[((static_cast<size_t>(co) * Cin + c) * R + r) * S + s];
  dst[idx] = __float2bfloat16_rn(v);
}

// NCHW fp32 (C <= 4) -> [N][H][Wp][4] bf16 with zero pad columns
__global__ void import_nchw_c4_padded_kernel(const float* __restrict__ src, uint2* __restrict__ dst, int C, int H, int W,
                                             int Wp, int pad_l, size_t total) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int wp = static_cast<int>(idx % Wp);
  const size_t nh = idx / Wp;
  const int w = wp - pad_l;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (w >= 0 && w < W) {
    const size_t n = nh / H, h = nh % H;
    const size_t HW = static_cast<size_t>(H) * W;
    const float* s = src + n * C * HW + h * W + w;
    for (int c = 0; c < C; ++c) v[c] = __ldg(s + c * HW);
  }
  dst[idx] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

using EncodeTiledFnS = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnS g_encode = nullptr;

}  // namespace

bool stem_rowring_geometry(StemGeometry& g, int Cin, int Cout, int H, int W, int R, int S, int stride, int pad, int dil,
                           int groups) {
  memset(&g, 0, sizeof g);
  if (Cin > 4 || groups != 1 || dil != 1 || R > kMaxR || R != S || (stride != 1 && stride != 2) || pad >= R) return false;
  g.pairs = stride == 1 ? 1 : 0;
  g.block_n = Cout * (g.pairs ? 2 : 1);
  if (g.block_n != 32 && g.block_n != 64) return false;
  g.P = (H + 2 * pad - R) / stride + 1;
  g.Q = (W + 2 * pad - S) / stride + 1;
  if (g.pairs && (g.Q & 1)) return false;
  g.Qw = g.pairs ? g.Q / 2 : g.Q;
  g.pad_l = pad + (pad & 1);
  g.xoff = g.pad_l - pad;
  if (S + g.xoff + g.pairs > 8) return false;
  g.Wp = std::max(2 * (g.Qw - 1) + 8, W + g.pad_l);
  g.Wp += g.Wp & 1;
  return true;
}

cudaError_t stem_rowring_set_attributes() {
  cudaError_t e = cudaFuncSetAttribute(stem_rowring_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCfg<64>::kSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(stem_rowring_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCfg<32>::kSmem);
}

cudaError_t pack_stem_weights(const float* oihw, __nv_bfloat16* dst, int Cout, int Cin, int R, int S, const StemGeometry& g,
                              cudaStream_t st) {
  const int total = g.block_n * R * 32;
  pack_stem_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(oihw, dst, Cout, Cin, R, S, g.block_n, g.pairs, g.xoff);
  return cudaGetLastError();
}

cudaError_t import_nchw_c4_padded(const float* src, void* dst, int N, int C, int H, int W, int Wp, int pad_l, cudaStream_t st) {
  const size_t total = static_cast<size_t>(N) * H * Wp;
  const size_t blocks = (total + 255) / 256;
  import_nchw_c4_padded_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, static_cast<uint2*>(dst), C, H, W, Wp, pad_l, total);
  return cudaGetLastError();
}

std::string stem_rowring_prepare(StemLaunch& L, int sm_count, const StemGeometry& g, const __nv_bfloat16* in_padded, int N,
                                 int H, int R, int stride, int pad, const __nv_bfloat16* packed_w, void* out, int pool,
                                 int Pp, int Qp) {
  if (!g_encode) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return "cuTensorMapEncodeTiled is not available from the driver";
    g_encode = reinterpret_cast<EncodeTiledFnS>(fn);
  }
  memset(&L, 0, sizeof L);
  StemParams& p = L.p;
  p.N = N, p.H = H, p.P = g.P, p.Qw = g.Qw, p.R = R, p.sv = stride, p.pad = pad;
  p.q_tiles = (g.Qw + kTileM - 1) / kTileM;
  p.pool = pool, p.Pp = Pp, p.Qp = Qp;
  p.out = static_cast<__nv_bfloat16*>(out);
  if (const char* e = getenv("TLXCV_DEBUG_ABLATE_STEM")) p.ablate = atoi(e);  // timing experiments only: results are wrong
  if (pool && (p.q_tiles != 1 || g.pairs)) return "stem: the fused max-pool needs a single column tile";
  // band height: minimise the rows the busiest CTA walks (including the R - sv warm-up rows per band)
  const int units = pool ? Pp : g.P;  // rows a band is measured in
  long long best = -1;
  for (int bands = 1; bands <= units; ++bands) {
    const int rows = (units + bands - 1) / bands;
    const int real_bands = (units + rows - 1) / rows;
    if (real_bands != bands) continue;
    const long long items = static_cast<long long>(N) * bands * p.q_tiles;
    const long long per_cta = (items + sm_count - 1) / sm_count;
    const long long conv_rows = pool ? 2 * rows + 1 : rows;
    const long long cost = per_cta * (conv_rows * 10 + (R - stride) * 4 + 6);  // row work + ring warm-up + band overhead
    if (best < 0 || cost < best) best = cost, p.bands = bands, p.band_rows = rows;
  }
  const long long items = static_cast<long long>(N) * p.bands * p.q_tiles;
  L.grid = static_cast<int>(std::min<long long>(items, sm_count));
  L.block_n = g.block_n;
  L.threads = kThreadsS;
  L.smem = g.block_n == 64 ? SCfg<64>::kSmem : SCfg<32>::kSmem;

  // A: overlapping windows. dim0 = 32 bf16 (8 pixels x 4 channels), dim1 = window q (stride 2 pixels = 16 B),
  // dim2 = input row, dim3 = image.  Out-of-range rows / windows are zero-filled by the hardware.
  {
    cuuint64_t dims[4] = {32, (cuuint64_t)g.Qw, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {16, (cuuint64_t)g.Wp * 8, (cuuint64_t)H * g.Wp * 8};
    cuuint32_t box[4] = {32, kTileM, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&L.tmapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(in_padded), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      char buf[160];
      snprintf(buf, sizeof buf, "stem: cuTensorMapEncodeTiled (overlapping windows) failed (%d)", int(r));
      return buf;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)R * 32, (cuuint64_t)g.block_n};
    cuuint64_t strides[1] = {(cuuint64_t)R * 64};
    cuuint32_t box[2] = {32, (cuuint32_t)g.block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&L.tmapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(packed_w), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "stem: cuTensorMapEncodeTiled (weights) failed";
  }
  return "";
}

cudaError_t stem_rowring_launch(const StemLaunch& L, cudaStream_t st) {
  if (L.block_n == 64)
    stem_rowring_kernel<64><<<L.grid, L.threads, L.smem, st>>>(L.tmapA, L.tmapB, L.p);
  else
    stem_rowring_kernel<32><<<L.grid, L.threads, L.smem, st>>>(L.tmapA, L.tmapB, L.p);
  return cudaGetLastError();
}

}  // namespace tlxcv
