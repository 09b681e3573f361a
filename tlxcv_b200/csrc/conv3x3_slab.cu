// 3x3 / stride 1 / pad 1 convolution (64 input channels) for sm_100a as a SLAB implicit GEMM.
//
// Replaces  nn.GroupConv2d(3,1,1) + BatchNorm2d + ReLU/LeakyReLU  of the wide early layers
// (classification/resnet.py:111-121 `conv2` of layer1, resnext / darknet equivalents) where the
// im2col-mode TMA path of conv_tcgen05.cu is bound by L2 -> SM operand traffic: it fetches every
// input pixel nine times (once per filter tap) plus the weights of every tap for every tile.
//
// Here a persistent CTA walks down a band of output rows and keeps
//   * the input rows it needs in a shared-memory SLAB: each row stored as [zero][W pixels][zero],
//     128 B (64 channels, SWIZZLE_128B) per pixel, consecutive rows contiguous (pitch Wp = W + 2).
//     Every input row is fetched ONCE by TMA (box 64 ch x W px) into its slot;
//   * all nine weight taps (9 x [C_out][64] K-major tiles, 72 KB) stationary for the whole kernel.
// The A operand of tap (r, s) for the T = floor(128 / Wp) output rows of a step is then simply the
// slab read at a ROW-SHIFTED start address  slab + ((row0 + r) * Wp + s) * 128 B : the hardware
// swizzle of both TMA and tcgen05.mma is a function of the absolute shared-memory address, so a
// K-major SWIZZLE_128B operand may start at any 128-byte row (tools/micro/umma_shift.cu verifies
// this on B200).  The zero columns between rows provide the left/right padding, rows outside the
// image are zero-filled by TMA, and the (128 - T*W) accumulator rows that fall on pad positions are
// simply not written out.
//
// Per step: 9 taps x 4 K-steps = 36 tcgen05.mma (128 x C_out x 16) into one TMEM accumulator
// (4 in flight); epilogue warps apply scale/shift/activation, stage the T output rows in shared
// memory and copy them to HBM as one contiguous, fully coalesced run (T consecutive NHWC rows).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace tlxcv {

namespace {

constexpr int kTileM = 128;
constexpr int kEpiWarpsB = 8;
constexpr int kThreadsB = (2 + kEpiWarpsB) * 32;  // 320
constexpr int kMaxGroupsB = 8;
constexpr int kSmemLimitB = 232448;
constexpr int kMiscB = 1024;

__device__ __forceinline__ uint32_t sw128_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t kSw128DescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B

template <bool kAccumulate>
__device__ __forceinline__ void umma_sw128(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(kSw128DescHi), "r"(idesc), "n"(kAccumulate ? 1 : 0)
      : "memory");
}

__device__ __forceinline__ bool elect_one_b() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_4d_b(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// debug timeline (TLXCV_DEBUG_TRACE_SLAB=1): CTA 0 records %clock64 at pipeline events, role-major
constexpr int kTraceLen = 4096;
__device__ __forceinline__ void trace(unsigned long long* buf, int role, int& idx) {
  if (buf != nullptr && blockIdx.x == 0 && idx < kTraceLen) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    buf[role * kTraceLen + idx++] = t;
  }
}

__device__ __forceinline__ void epi_bar_sync_b() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct BandItem {
  int n;
  int p0, p1;  // output rows [p0, p1)
  int steps;   // ceil((p1 - p0) / T)
  int groups;  // row groups the band loads: steps + 1 (group g = input rows p0 - 1 + g*T ... + T)
};

__device__ __forceinline__ BandItem decode_band(const SlabParams& p, int item) {
  BandItem it;
  const int b = item % p.bands;
  it.n = item / p.bands;
  it.p0 = b * p.band_rows;
  it.p1 = min(p.H, it.p0 + p.band_rows);
  it.steps = (it.p1 - it.p0 + p.T - 1) / p.T;
  it.groups = it.steps + 1;
  return it;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreadsB, 1)
conv3x3_slab_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const SlabParams p) {
  constexpr int kAccBufs = 512 / BLOCK_N > 4 ? 4 : 512 / BLOCK_N;
  constexpr int kRowBytes = BLOCK_N * 2;   // bytes per output pixel
  constexpr int kChunks = kRowBytes / 16;  // 16-byte chunks per output pixel
  constexpr int kTapBytes = BLOCK_N * 128; // one weight tap: BLOCK_N rows x 64 K (SWIZZLE_128B)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int T = p.T, W = p.W, Wp = p.W + 2;
  const int row_bytes = Wp * 128;               // one slab row (64 channels)
  const int slots = p.NG * T + 2;               // ring rows + two mirror rows
  const int slab_bytes = (slots * row_bytes + 1023) / 1024 * 1024;
  const int stage_bytes = T * W * kRowBytes;    // one staging buffer: T output rows
  uint8_t* slab = smem;
  uint8_t* bsm = slab + slab_bytes;
  uint8_t* stage = bsm + 9 * kTapBytes;
  float* sc_s = reinterpret_cast<float*>(stage + 2 * ((stage_bytes + 127) / 128 * 128));
  float* sh_s = sc_s + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sh_s + 64);
  uint64_t* full_bar = bars;                      // [kMaxGroupsB] row group landed
  uint64_t* empty_bar = bars + kMaxGroupsB;       // [kMaxGroupsB] MMAs reading the group retired
  uint64_t* tfull_bar = bars + 2 * kMaxGroupsB;   // [4]
  uint64_t* tempty_bar = tfull_bar + 4;           // [4]
  uint64_t* b_bar = tempty_bar + 4;               // weights landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.N * p.bands;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    for (int i = 0; i < kMaxGroupsB; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), kEpiWarpsB);
    }
    mbar_init(smem_u32(b_bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<kAccBufs * BLOCK_N>(smem_u32(tmem_ptr_smem));
  if (warp >= 2) {
    const int et = threadIdx.x - 64;
    for (int i = et; i < BLOCK_N; i += kEpiWarpsB * 32) {
      sc_s[i] = p.scale[i];
      sh_s[i] = p.shift[i];
    }
    // the pad columns of every slab row (first and last 128 B) stay zero for the whole kernel
    for (int i = et; i < slots * 2 * 8; i += kEpiWarpsB * 32) {
      const int s = i / 16, side = (i / 8) & 1, c = i & 7;
      *reinterpret_cast<uint4*>(slab + s * row_bytes + (side ? (W + 1) * 128 : 0) + c * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();  // generic-proxy zeros -> visible to tcgen05.mma operand reads
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t NG = static_cast<uint32_t>(p.NG);
  pdl_wait();               // the previous kernel's activations are complete and visible from here on
  pdl_launch_dependents();  // the next kernel may take this SM as soon as this CTA exits

  if (warp == 0) {
    // ===================== TMA producer: weights once, then T input rows per step =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(b_bar), 9 * kTapBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(bsm + t * kTapBytes), &tmapB, smem_u32(b_bar), t * 64, 0);
    }
    uint32_t slot = 0, phase = 0;
    int tr = 0;
    const int ahead = p.NG + 1;
    const size_t in_row_bytes = static_cast<size_t>(W) * 128;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const BandItem it = decode_band(p, item);
      const uint8_t* img = reinterpret_cast<const uint8_t*>(p.in) + static_cast<size_t>(it.n) * p.H * in_row_bytes;
      for (int g = -ahead; g < it.groups; ++g) {
        const int gp = g + ahead;
        if (gp < it.groups && lane == 0 && !(p.ablate & 16)) {
          // L2 prefetch of a group further down the band: its rows are one contiguous run of the NHWC input
          const int h_lo = max(0, it.p0 - 1 + gp * T), h_hi = min(p.H, it.p0 - 1 + gp * T + T);
          if (h_hi > h_lo)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(img + static_cast<size_t>(h_lo) * in_row_bytes),
                         "r"(static_cast<uint32_t>((h_hi - h_lo) * in_row_bytes))
                         : "memory");
        }
        if (g < 0) continue;
        if (lane == 0) {
          trace(p.trace, 0, tr);  // [3k] before the empty wait
          mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1);
          trace(p.trace, 0, tr);  // [3k+1] slot free
          const uint32_t bar = smem_u32(&full_bar[slot]);
          const int h0 = it.p0 - 1 + g * T;
          const int mirror = slot == 0 ? 2 : 0;  // the ring's first two rows are mirrored behind its end
          if (p.ablate & 1) {
            mbar_arrive(bar);
          } else {
            mbar_arrive_expect_tx(bar, (T + mirror) * W * 128);
            for (int r = 0; r < T; ++r)
              tma_load_4d_b(smem_u32(slab + (slot * T + r) * row_bytes + 128), &tmapA, bar, 0, 0, h0 + r, it.n);
            for (int r = 0; r < mirror; ++r)
              tma_load_4d_b(smem_u32(slab + (p.NG * T + r) * row_bytes + 128), &tmapA, bar, 0, 0, h0 + r, it.n);
          }
          trace(p.trace, 0, tr);  // [3k+2] loads issued
        }
        __syncwarp();
        if (++slot == NG) slot = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: converged warp, one elected lane issues =====================
    constexpr uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N);
    mbar_wait(smem_u32(b_bar), 0);
    tcgen05_fence_after();
    const uint32_t slab_lo = sw128_desc_lo(smem_u32(slab)), b_lo0 = sw128_desc_lo(smem_u32(bsm));
    const uint32_t group_lo = static_cast<uint32_t>(T * row_bytes) >> 4;
    uint32_t tap_lo[9];  // descriptor offset of tap (r, s): (r * Wp + s) rows of 128 B
#pragma unroll
    for (int t = 0; t < 9; ++t) tap_lo[t] = static_cast<uint32_t>(((t / 3) * Wp + (t % 3)) * 8);
    uint32_t fslot = 0;                  // ring slot of the first group of the current step
    uint32_t wslot = 0, wphase = 0;      // next group to wait for
    uint32_t acc = 0, acc_phase = 0;
    int tr = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const BandItem it = decode_band(p, item);
      int waited = 0;
      // Two steps per synchronisation round: every barrier wait costs ~100 cycles of this warp even when
      // the barrier has already completed, and while it waits the tensor pipe drains.
      for (int m = 0; m < it.steps; m += 2) {
        const int nb = min(2, it.steps - m);
        if (lane == 0) trace(p.trace, 1, tr);  // [4k] round start
        {
          uint32_t a2 = acc, ph2 = acc_phase;
          for (int b2 = 0; b2 < nb; ++b2) {
            mbar_wait(smem_u32(&tempty_bar[a2]), ph2 ^ 1);
            if (++a2 == kAccBufs) a2 = 0, ph2 ^= 1;
          }
        }
        if (lane == 0) trace(p.trace, 1, tr);  // [4k+1] accumulators free
        while (waited < m + nb + 1) {  // step k reads group k and the first two rows of group k + 1
          mbar_wait(smem_u32(&full_bar[wslot]), wphase);
          if (++wslot == NG) wslot = 0, wphase ^= 1;
          ++waited;
        }
        tcgen05_fence_after();
        if (lane == 0) trace(p.trace, 1, tr);  // [4k+2] operands landed
        const bool leader = elect_one_b();
        for (int b2 = 0; b2 < nb; ++b2) {
          if (leader) {
            const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
            const uint32_t a0 = slab_lo + fslot * group_lo;
            if (!(p.ablate & 2)) {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                const uint32_t a_lo = a0 + tap_lo[t];
                const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(t) * (kTapBytes >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (t == 0 && k == 0)
                    umma_sw128<false>(tmem_d, a_lo, b_lo, idesc);
                  else
                    umma_sw128<true>(tmem_d, a_lo + 2 * k, b_lo + 2 * k, idesc);
                }
              }
            }
            umma_commit(smem_u32(&tfull_bar[acc]));
            umma_commit(smem_u32(&empty_bar[fslot]));  // the step's group is dead after these MMAs
          }
          if (++acc == kAccBufs) acc = 0, acc_phase ^= 1;
          if (++fslot == NG) fslot = 0;
        }
        __syncwarp();
        if (lane == 0) trace(p.trace, 1, tr);  // [4k+3] MMAs and commits issued
      }
      // the band's last group (only its first two rows were read) goes back as well
      if (elect_one_b()) umma_commit(smem_u32(&empty_bar[fslot]));
      __syncwarp();
      if (++fslot == NG) fslot = 0;
    }
  } else {
    // ===================== epilogue: 8 warps = 4 lane groups x 2 channel halves =====================
    constexpr int kNc = BLOCK_N / 2;
    const int lg = warp & 3, ch = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    const int pos = lg * 32 + lane;          // flat slab position (TMEM lane) this thread owns
    const int trow = pos / Wp, q = pos - trow * Wp;
    const bool in_row = q < W && trow < T;
    const int spx = trow * W + q;            // pixel index inside the staging buffer
    const uint32_t stage_addr = smem_u32(stage);
    const uint32_t stage_stride = static_cast<uint32_t>((stage_bytes + 127) / 128 * 128);
    const uint32_t swz = kChunks == 8 ? (spx & 7) : ((spx >> 1) & 3);
    const float alpha = p.alpha;
    const int act = p.act;
    uint32_t acc = 0, acc_phase = 0, sbuf = 0;
    int tr = 0;
    const bool tracer = warp == 2 && lane == 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const BandItem it = decode_band(p, item);
      for (int m = 0; m < it.steps; ++m) {
        const int prow0 = it.p0 + m * T;
        const int n_rows = min(T, it.p1 - prow0);
        if (tracer) trace(p.trace, 2, tr);  // [4k] step start
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        if (tracer) trace(p.trace, 2, tr);  // [4k+1] accumulator complete
        tcgen05_fence_after();
        uint32_t v[kNc];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + acc * BLOCK_N + ch * kNc;
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
        if (++acc == kAccBufs) acc = 0, acc_phase ^= 1;
        const uint32_t sb = stage_addr + sbuf * stage_stride;
        if (in_row && trow < n_rows && !(p.ablate & 4)) {
          const uint32_t my_px = sb + spx * kRowBytes;
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
            act_regs(f, act, alpha);
            const uint32_t cidx = static_cast<uint32_t>(ch * (kNc / 8) + j);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(my_px + ((cidx ^ swz) << 4)),
                         "r"(pack_bf16x2(f[0], f[1])), "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])),
                         "r"(pack_bf16x2(f[6], f[7]))
                         : "memory");
          }
        }
        if (tracer) trace(p.trace, 2, tr);  // [4k+2] staged
        epi_bar_sync_b();  // the step's output rows are complete in the staging buffer
        if (!(p.ablate & 8)) {
          // T consecutive NHWC rows are one contiguous run in HBM: consecutive threads -> consecutive 16 B
          uint8_t* gdst = reinterpret_cast<uint8_t*>(p.out) + (static_cast<size_t>(it.n) * p.H + prow0) * W * kRowBytes;
          const int total = n_rows * W * kChunks;
          for (int i = et; i < total; i += kEpiWarpsB * 32) {
            const uint32_t px = static_cast<uint32_t>(i) / kChunks, cidx = static_cast<uint32_t>(i) % kChunks;
            const uint32_t pswz = kChunks == 8 ? (px & 7) : ((px >> 1) & 3);
            uint4 val;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                         : "r"(sb + px * kRowBytes + ((cidx ^ pswz) << 4)));
            *reinterpret_cast<uint4*>(gdst + static_cast<size_t>(i) * 16) = val;
          }
        }
        sbuf ^= 1;
        if (tracer) trace(p.trace, 2, tr);  // [4k+3] copied out
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

using EncodeTiledFnB = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnB g_encode_b = nullptr;

}  // namespace

bool conv3x3_slab_supported(int Cin, int Cout, int H, int W, int R, int S, int stride, int pad, int dil, int groups,
                            bool residual) {
  if (getenv("TLXCV_NO_SLAB")) return false;
  return Cin == 64 && Cout == 64 && R == 3 && S == 3 && stride == 1 && pad == 1 && dil == 1 && groups == 1 && !residual &&
         W + 2 <= 64 && W >= 8 && H >= 1;
}

cudaError_t conv3x3_slab_set_attributes() {
  return cudaFuncSetAttribute(conv3x3_slab_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimitB);
}

std::string conv3x3_slab_prepare(SlabLaunch& L, int sm_count, const __nv_bfloat16* in, int N, int H, int W, int Cout,
                                 const __nv_bfloat16* packed_w, int Ktot, void* out) {
  if (!g_encode_b) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return "cuTensorMapEncodeTiled is not available from the driver";
    g_encode_b = reinterpret_cast<EncodeTiledFnB>(fn);
  }
  memset(&L, 0, sizeof L);
  SlabParams& p = L.p;
  const int Wp = W + 2;
  p.N = N, p.H = H, p.W = W;
  p.T = kTileM / Wp;
  p.in = in;
  p.out = static_cast<__nv_bfloat16*>(out);
  if (const char* e = getenv("TLXCV_DEBUG_ABLATE_SLAB")) p.ablate = atoi(e);  // timing experiments only: results are wrong
  if (Ktot != 9 * 64) return "slab conv: packed weight K must be 9 x 64";
  const int block_n = 64;
  const int row_bytes = Wp * 128;
  const int stage = (p.T * W * block_n * 2 + 127) / 128 * 128;
  const int fixed = 9 * block_n * 128 + 2 * stage + kMiscB;
  int ng = kMaxGroupsB;
  auto smem_for = [&](int g) { return ((g * p.T + 2) * row_bytes + 1023) / 1024 * 1024 + fixed; };
  while (ng > 2 && smem_for(ng) > kSmemLimitB) --ng;
  if (ng < 3 || smem_for(ng) > kSmemLimitB) return "slab conv: shared memory cannot hold the input-row ring";
  p.NG = ng;
  L.smem = smem_for(ng);
  long long best = -1;
  for (int bands = 1; bands <= H; ++bands) {
    const int rows = (H + bands - 1) / bands;
    if ((H + rows - 1) / rows != bands) continue;
    const long long items = static_cast<long long>(N) * bands;
    const long long per_cta = (items + sm_count - 1) / sm_count;
    const long long steps = (rows + p.T - 1) / p.T;
    const long long cost = per_cta * (steps * 10 + 14);  // steps + band warm-up
    if (best < 0 || cost < best) best = cost, p.bands = bands, p.band_rows = rows;
  }
  L.grid = static_cast<int>(std::min<long long>(static_cast<long long>(N) * p.bands, sm_count));
  L.block_n = block_n;
  L.threads = kThreadsB;
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)W, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_b(&L.tmapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(in), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "slab conv: cuTensorMapEncodeTiled (input rows) failed";
  }
  {
    const int cout_pad = ((Cout + 255) / 256) * 256;
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_b(&L.tmapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(packed_w), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "slab conv: cuTensorMapEncodeTiled (weights) failed";
  }
  return "";
}

cudaError_t conv3x3_slab_launch(const SlabLaunch& L, cudaStream_t st) {
  static const char* trace_path = getenv("TLXCV_DEBUG_TRACE_SLAB");  // debugging: dump CTA 0's pipeline timeline
  if (trace_path != nullptr) {
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 3 * kTraceLen * sizeof(unsigned long long));
    cudaMemsetAsync(dbuf, 0, 3 * kTraceLen * sizeof(unsigned long long), st);
    SlabParams p = L.p;
    p.trace = dbuf;
    conv3x3_slab_kernel<64><<<L.grid, L.threads, L.smem, st>>>(L.tmapA, L.tmapB, p);
    cudaStreamSynchronize(st);
    std::vector<unsigned long long> h(3 * kTraceLen);
    cudaMemcpy(h.data(), dbuf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "wb")) {
      fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
      fclose(f);
    }
    return cudaGetLastError();
  }
  return launch_pdl(conv3x3_slab_kernel<64>, L.grid, L.threads, L.smem, st, L.tmapA, L.tmapB, L.p);
}

}  // namespace tlxcv
