// 3x3 / stride 1 / pad 1 convolution for sm_100a as a SLAB implicit GEMM: 64 output channels per work item from
// one block of CB = 64 or 32 input channels.
//
// Replaces  nn.GroupConv2d(3,1,1[, n_group]) + BatchNorm2d + ReLU/LeakyReLU [+ residual add]  of
//   * the wide early dense layers: `conv2` of ResNet layer1 (classification/resnet.py:111-121, 64 -> 64) and the first
//     DarkNet residual block (detection/backbones/darknet.py:134-146,155-159, 32 -> 64 at 304^2, with the skip add);
//   * the grouped 3x3 convs of ResNeXt (classification/resnext.py:80-89, 32 groups): group boundaries never cross a
//     64-channel block, so block j of the input produces block j of the output through block-diagonal weights;
// where the im2col-mode TMA path of conv_tcgen05.cu is bound by L2 -> SM operand traffic: it fetches every input
// pixel nine times (once per filter tap) plus the weights of every tap for every tile.
//
// Here a persistent CTA walks down a band of output rows of one column segment and keeps
//   * the input rows it needs in a shared-memory SLAB: each row is the segment's Ws pixels plus one halo pixel on
//     either side (TMA box CB channels x (Ws + 2) pixels; pixels outside the image are zero-filled by the hardware,
//     which is exactly the conv's zero padding), 2*CB bytes per pixel in the K-major swizzled layout, consecutive rows
//     contiguous (pitch Wp = Ws + 2 pixels).  Every input row is fetched ONCE;
//   * all nine weight taps of its channel block (9 x [64][CB] K-major tiles) stationary until the block changes.
// The A operand of tap (r, s) for the T = floor(128 / Wp) output rows of a step is then the slab read at a
// ROW-SHIFTED start address  slab + ((row0 + r) * Wp + s) * 2*CB : the hardware swizzle of both TMA and tcgen05.mma is a
// function of the absolute shared-memory address, so a K-major swizzled operand may start at any pixel row
// (tools/micro/umma_shift.cu verifies this on B200).  The (128 - T*Ws) accumulator rows that fall on halo positions
// are simply not written out.
//
// Per step: 9 taps x CB/16 K-steps tcgen05.mma (128 x 64 x 16) into one TMEM accumulator (4 in flight); epilogue warps
// apply scale/shift/activation (+ residual, read straight from HBM: every lane owns one pixel = one 64-byte run), stage
// the output pixels in shared memory and copy them out with 16 B per thread.
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
constexpr int kBlockN = 64;                     // output channels per work item
constexpr int kEpiWarpsB = 8;
constexpr int kThreadsB = (2 + kEpiWarpsB) * 32;  // 320
constexpr int kMaxGroupsB = 16;
constexpr int kSmemLimitB = 232448;
constexpr int kMiscB = 1024;
constexpr int kAccBufsB = 4;

// descriptor low word: start address >> 4 | LBO(ignored) ; high word: SBO (8 rows), version 1, swizzle mode
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
template <int CB>
struct Sw {
  static constexpr int kRowBytes = 2 * CB;  // bytes per pixel of the slab / per weight row of a tap
  static constexpr uint32_t kDescHi = ((8u * kRowBytes) >> 4) | (1u << 14) | ((CB == 64 ? 2u : 4u) << 29);
};

template <bool kAccumulate>
__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "n"(kAccumulate ? 1 : 0)
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
__device__ __forceinline__ void tma_prefetch_4d_b(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// debug timeline (TLXCV_DEBUG_TRACE_SLAB=<file>): CTA 0 records %clock64 at pipeline events, role-major
constexpr int kTraceLen = 4096;
__device__ __forceinline__ void trace(unsigned long long* buf, int role, int& idx) {
  if (buf != nullptr && blockIdx.x == 0 && idx < kTraceLen) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    buf[role * kTraceLen + idx++] = t;
  }
}

__device__ __forceinline__ void epi_bar_sync_b() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// One work item = (channel block, image, band of output rows, column segment); channel block slowest so that a CTA
// changes its stationary weights rarely.
struct BandItem {
  int cb, n, x0;
  int p0, p1;  // output rows [p0, p1)
  int steps;   // ceil((p1 - p0) / T)
  int groups;  // row groups the band loads: steps - 1 + ng (group g = input rows p0 - 1 + g*T ... + T)
};

__device__ __forceinline__ BandItem decode_band(const SlabParams& p, int item) {
  BandItem it;
  const int seg = item % p.segs;
  int t = item / p.segs;
  const int b = t % p.bands;
  t /= p.bands;
  it.n = t % p.N;
  it.cb = t / p.N;
  it.x0 = seg * p.Ws;
  it.p0 = b * p.band_rows;
  it.p1 = min(p.H, it.p0 + p.band_rows);
  it.steps = (it.p1 - it.p0 + p.T - 1) / p.T;
  it.groups = it.steps - 1 + p.ng;
  return it;
}

template <int CB>
__global__ void __launch_bounds__(kThreadsB, 1)
conv3x3_slab_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const SlabParams p) {
  constexpr int kPixBytes = Sw<CB>::kRowBytes;     // bytes per slab pixel
  constexpr int kTapBytes = kBlockN * kPixBytes;   // one weight tap: 64 rows x CB K
  constexpr int kKSteps = CB / 16;
  constexpr int kOutBytes = kBlockN * 2;           // bytes per output pixel of the item's channel block
  constexpr int kChunks = kOutBytes / 16;          // 8
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int T = p.T, Ws = p.Ws, Wp = p.Ws + 2;
  const int row_bytes = Wp * kPixBytes;            // one slab row
  const int slots = p.NG * T + 2;                  // ring rows + two mirror rows
  const int slab_bytes = (slots * row_bytes + 1023) / 1024 * 1024;
  const int stage_bytes = (T * Ws * kOutBytes + 127) / 128 * 128;  // one staging buffer: T output rows of the segment
  uint8_t* slab = smem;
  uint8_t* bsm = slab + slab_bytes;
  uint8_t* stage = bsm + 9 * kTapBytes;
  float* sc_s = reinterpret_cast<float*>(stage + 2 * stage_bytes);
  float* sh_s = sc_s + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sh_s + 64);
  uint64_t* full_bar = bars;                      // [kMaxGroupsB] row group landed
  uint64_t* empty_bar = bars + kMaxGroupsB;       // [kMaxGroupsB] MMAs reading the group retired
  uint64_t* tfull_bar = bars + 2 * kMaxGroupsB;   // [4]
  uint64_t* tempty_bar = tfull_bar + 4;           // [4]
  uint64_t* b_full = tempty_bar + 4;              // weights of the current channel block landed
  uint64_t* b_empty = b_full + 1;                 // every MMA that read the previous block's weights retired
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.cblocks * p.N * p.bands * p.segs;

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
    mbar_init(smem_u32(b_full), 1);
    mbar_init(smem_u32(b_empty), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<kAccBufsB * kBlockN>(smem_u32(tmem_ptr_smem));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t NG = static_cast<uint32_t>(p.NG);
  pdl_wait();               // the previous kernel's activations are complete and visible from here on
  pdl_launch_dependents();  // the next kernel may take this SM as soon as this CTA exits

  if (warp == 0) {
    // ===================== TMA producer: the block's weights, then T input rows per step =====================
    uint32_t slot = 0, phase = 0, b_loads = 0;
    int tr = 0, cur_cb = -1;
    const int ahead = p.NG + 1;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const BandItem it = decode_band(p, item);
      if (it.cb != cur_cb && lane == 0) {
        // new channel block: its nine weight taps replace the stationary ones once the MMAs that read them retired
        if (b_loads > 0) mbar_wait(smem_u32(b_empty), (b_loads - 1) & 1);
        mbar_arrive_expect_tx(smem_u32(b_full), 9 * kTapBytes);
        for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(bsm + t * kTapBytes), &tmapB, smem_u32(b_full), t * CB, it.cb * kBlockN);
        ++b_loads;
      }
      cur_cb = it.cb;
      for (int g = -ahead; g < it.groups; ++g) {
        const int gp = g + ahead;
        if (gp < it.groups && lane < T && !(p.ablate & 16)) {
          // L2 prefetch of a group further down the band (one slab row per lane)
          const int h = it.p0 - 1 + gp * T + lane;
          if (h >= 0 && h < p.H) tma_prefetch_4d_b(&tmapA, it.cb * CB, it.x0 - 1, h, it.n);
        }
        if (g < 0) continue;
        if (lane == 0) {
          trace(p.trace, 0, tr);  // [3k] before the empty wait
          mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1);
          trace(p.trace, 0, tr);  // [3k+1] slot free
          const uint32_t bar = smem_u32(&full_bar[slot]);
          const int h0 = it.p0 - 1 + g * T;
          const int r0 = static_cast<int>(slot) * T;              // ring row of the group's first row
          const int mirror = max(0, min(T, 2 - r0));              // ring rows 0 and 1 are mirrored behind the ring's end
          if (p.ablate & 1) {
            mbar_arrive(bar);
          } else {
            mbar_arrive_expect_tx(bar, (T + mirror) * row_bytes);
            for (int r = 0; r < T; ++r)
              tma_load_4d_b(smem_u32(slab + (r0 + r) * row_bytes), &tmapA, bar, it.cb * CB, it.x0 - 1, h0 + r, it.n);
            for (int r = 0; r < mirror; ++r)
              tma_load_4d_b(smem_u32(slab + (p.NG * T + r0 + r) * row_bytes), &tmapA, bar, it.cb * CB, it.x0 - 1, h0 + r, it.n);
          }
          trace(p.trace, 0, tr);  // [3k+2] loads issued
        }
        __syncwarp();
        if (++slot == NG) slot = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: converged warp, one elected lane issues =====================
    constexpr uint32_t idesc = make_idesc_bf16(kTileM, kBlockN);
    constexpr uint32_t hi = Sw<CB>::kDescHi;
    const uint32_t slab_lo = desc_lo(smem_u32(slab)), b_lo0 = desc_lo(smem_u32(bsm));
    const uint32_t group_lo = static_cast<uint32_t>(T * row_bytes) >> 4;
    const int ng = p.ng;
    uint32_t tap_lo[9];  // descriptor offset of tap (r, s): (r * Wp + s) slab pixels
#pragma unroll
    for (int t = 0; t < 9; ++t) tap_lo[t] = static_cast<uint32_t>(((t / 3) * Wp + (t % 3)) * (kPixBytes >> 4));
    uint32_t fslot = 0;                  // ring slot of the first group of the current step
    uint32_t wslot = 0, wphase = 0;      // next group to wait for
    uint32_t acc = 0, acc_phase = 0, b_loads = 0;
    int tr = 0, cur_cb = -1;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const BandItem it = decode_band(p, item);
      if (it.cb != cur_cb) {
        if (b_loads > 0) {
          // everything issued so far read the previous block's weights: release them, then wait for the new ones
          if (elect_one_b()) umma_commit(smem_u32(b_empty));
          __syncwarp();
        }
        mbar_wait(smem_u32(b_full), b_loads & 1);
        tcgen05_fence_after();
        ++b_loads;
        cur_cb = it.cb;
      }
      int waited = 0;
      // Two steps per synchronisation round: every barrier wait costs ~100 cycles of this warp even when the barrier
      // has already completed, and while it waits the tensor pipe drains.
      for (int m = 0; m < it.steps; m += 2) {
        const int nb = min(2, it.steps - m);
        if (lane == 0) trace(p.trace, 1, tr);  // [4k] round start
        {
          uint32_t a2 = acc, ph2 = acc_phase;
          for (int b2 = 0; b2 < nb; ++b2) {
            mbar_wait(smem_u32(&tempty_bar[a2]), ph2 ^ 1);
            if (++a2 == kAccBufsB) a2 = 0, ph2 ^= 1;
          }
        }
        if (lane == 0) trace(p.trace, 1, tr);  // [4k+1] accumulators free
        while (waited < m + nb - 1 + ng) {  // step k reads groups k .. k + ng - 1
          mbar_wait(smem_u32(&full_bar[wslot]), wphase);
          if (++wslot == NG) wslot = 0, wphase ^= 1;
          ++waited;
        }
        tcgen05_fence_after();
        if (lane == 0) trace(p.trace, 1, tr);  // [4k+2] operands landed
        const bool leader = elect_one_b();
        for (int b2 = 0; b2 < nb; ++b2) {
          if (leader) {
            const uint32_t tmem_d = tmem_base + acc * kBlockN;
            const uint32_t a0 = slab_lo + fslot * group_lo;
            if (!(p.ablate & 2)) {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                const uint32_t a_lo = a0 + tap_lo[t];
                const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(t) * (kTapBytes >> 4);
#pragma unroll
                for (int k = 0; k < kKSteps; ++k) {
                  if (t == 0 && k == 0)
                    umma_lo<false>(tmem_d, a_lo, b_lo, hi, idesc);
                  else
                    umma_lo<true>(tmem_d, a_lo + 2 * k, b_lo + 2 * k, hi, idesc);
                }
              }
            }
            umma_commit(smem_u32(&tfull_bar[acc]));
            umma_commit(smem_u32(&empty_bar[fslot]));  // the step's first group is dead after these MMAs
          }
          if (++acc == kAccBufsB) acc = 0, acc_phase ^= 1;
          if (++fslot == NG) fslot = 0;
        }
        __syncwarp();
        if (lane == 0) trace(p.trace, 1, tr);  // [4k+3] MMAs and commits issued
      }
      // the band's trailing groups (loaded for the last step's lower rows) go back as well
      for (int k = 1; k < ng; ++k) {
        if (elect_one_b()) umma_commit(smem_u32(&empty_bar[fslot]));
        __syncwarp();
        if (++fslot == NG) fslot = 0;
      }
    }
  } else {
    // ===================== epilogue: 8 warps = 4 lane groups x 2 channel halves =====================
    constexpr int kNc = kBlockN / 2;
    const int lg = warp & 3, ch = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    const int pos = lg * 32 + lane;          // flat slab position (TMEM lane) this thread owns
    const int trow = pos / Wp, q = pos - trow * Wp;
    const int spx = trow * Ws + q;           // pixel index inside the staging buffer
    const uint32_t stage_addr = smem_u32(stage);
    const uint32_t swz = spx & 7;
    const float alpha = p.alpha, alpha2 = p.alpha2;
    const int act = p.act, act2 = p.act2;
    const size_t pix_stride = static_cast<size_t>(p.Cout) * 2;  // bytes between output pixels
    uint32_t acc = 0, acc_phase = 0, sbuf = 0;
    int tr = 0, cur_cb = -1;
    const bool tracer = warp == 2 && lane == 0;
    const bool has_res = p.residual != nullptr;
    // Residual: this lane's pixel, its 32 channels = one 64-byte run in HBM, read straight into registers ONE STEP
    // AHEAD (the load for step k + 1 is issued before step k's accumulator is touched), so its latency hides behind a
    // whole step of work instead of stalling every step.
    uint4 res[4], res_next[4];
    auto load_res = [&](const BandItem& bi, int m, uint4 (&dst)[4]) {
      const int prow = bi.p0 + m * T + trow;
      const bool ok = q < min(Ws, p.W - bi.x0) && trow < T && prow < bi.p1;
      if (ok) {
        const uint4* rp = reinterpret_cast<const uint4*>(
            reinterpret_cast<const uint8_t*>(p.residual) +
            ((static_cast<size_t>(bi.n) * p.H + prow) * p.W + bi.x0 + q) * pix_stride + (bi.cb * kBlockN + ch * kNc) * 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = __ldg(rp + j);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(0, 0, 0, 0);
      }
    };
    if (has_res && blockIdx.x < num_items) load_res(decode_band(p, blockIdx.x), 0, res_next);
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const BandItem it = decode_band(p, item);
      const int valid_w = min(Ws, p.W - it.x0);
      if (it.cb != cur_cb) {
        // scale / shift of the new channel block (the barrier also orders this against the previous item's readers)
        epi_bar_sync_b();
        if (et < kBlockN) {
          sc_s[et] = p.scale[it.cb * kBlockN + et];
          sh_s[et] = p.shift[it.cb * kBlockN + et];
        }
        epi_bar_sync_b();
        cur_cb = it.cb;
      }
      const bool in_seg = q < valid_w && trow < T;
      // copy-out plan of this thread for the item (at most 4 chunks of 16 B per step): staging offset, offset from the
      // step's first output pixel, and the output row within the step; -1 = nothing
      constexpr int kCopyIters = (kTileM * kChunks) / (kEpiWarpsB * 32);  // 4
      uint32_t cp_smem[kCopyIters];
      int cp_row[kCopyIters];
      size_t cp_gl[kCopyIters];
      {
        const int per_row = valid_w * kChunks;
#pragma unroll
        for (int j = 0; j < kCopyIters; ++j) {
          const int i = et + j * kEpiWarpsB * 32;
          const int r = i / per_row, k = i - r * per_row;
          const uint32_t w = static_cast<uint32_t>(k) / kChunks, cidx = static_cast<uint32_t>(k) % kChunks;
          const uint32_t px = static_cast<uint32_t>(r * Ws) + w;
          cp_row[j] = r < T ? r : 1 << 20;
          cp_smem[j] = px * kOutBytes + ((cidx ^ (px & 7)) << 4);
          cp_gl[j] = (static_cast<size_t>(r) * p.W + w) * pix_stride + cidx * 16;
        }
      }
      uint8_t* const item_out = reinterpret_cast<uint8_t*>(p.out) +
                                (static_cast<size_t>(it.n) * p.H * p.W + it.x0) * pix_stride + (it.cb * kBlockN) * 2;
      for (int m = 0; m < it.steps; ++m) {
        const int prow0 = it.p0 + m * T;
        const int n_rows = min(T, it.p1 - prow0);
        const bool live = in_seg && trow < n_rows && !(p.ablate & 4);
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) res[j] = res_next[j];
          if (m + 1 < it.steps)
            load_res(it, m + 1, res_next);
          else if (item + static_cast<int>(gridDim.x) < num_items)
            load_res(decode_band(p, item + gridDim.x), 0, res_next);
        }
        if (tracer) trace(p.trace, 2, tr);  // [4k] step start
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        if (tracer) trace(p.trace, 2, tr);  // [4k+1] accumulator complete
        tcgen05_fence_after();
        uint32_t v[kNc];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + acc * kBlockN + ch * kNc;
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
        if (++acc == kAccBufsB) acc = 0, acc_phase ^= 1;
        const uint32_t sb = stage_addr + sbuf * stage_bytes;
        if (live) {
          const uint32_t my_px = sb + spx * kOutBytes;
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
            if (has_res) {
              const __nv_bfloat162* hr = reinterpret_cast<const __nv_bfloat162*>(&res[j]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 rr = __bfloat1622float2(hr[e]);
                f[2 * e] += rr.x;
                f[2 * e + 1] += rr.y;
              }
              act_regs(f, act2, alpha2);
            }
            const uint32_t cidx = static_cast<uint32_t>(ch * (kNc / 8) + j);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(my_px + ((cidx ^ swz) << 4)),
                         "r"(pack_bf16x2(f[0], f[1])), "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])),
                         "r"(pack_bf16x2(f[6], f[7]))
                         : "memory");
          }
        }
        if (tracer) trace(p.trace, 2, tr);  // [4k+2] staged
        epi_bar_sync_b();  // the step's output pixels are complete in the staging buffer
        if (!(p.ablate & 8)) {
          // consecutive threads -> consecutive 16 B: the 8 chunks of a pixel are one 128-byte run, pixels of a row are
          // Cout*2 bytes apart (one contiguous run per row when the layer has a single channel block)
          uint8_t* const step_out = item_out + static_cast<size_t>(prow0) * p.W * pix_stride;
#pragma unroll
          for (int j = 0; j < kCopyIters; ++j) {
            if (cp_row[j] < n_rows) {
              uint4 val;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                           : "r"(sb + cp_smem[j]));
              *reinterpret_cast<uint4*>(step_out + cp_gl[j]) = val;
            }
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
    tmem_dealloc<kAccBufsB * kBlockN>(tmem_base);
  }
}

using EncodeTiledFnB = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnB g_encode_b = nullptr;

}  // namespace

bool conv3x3_slab_supported(int Cin, int Cout, int H, int W, int R, int S, int stride, int pad, int dil, int groups) {
  if (tuning_env("TLXCV_NO_SLAB")) return false;
  if (R != 3 || S != 3 || stride != 1 || pad != 1 || dil != 1 || W < 8 || H < 1) return false;
  if (groups == 1) return (Cin == 64 || Cin == 32) && Cout == 64;
  // grouped: channels-per-group divides 64 and nothing crosses a 64-channel block.  Small maps stay on the im2col
  // path of conv_tcgen05.cu (measured: 14x14 x 512 channels 0.107 ms here against 0.077 ms there).
  const int cpg = Cin / groups;
  return Cin == Cout && Cin % 64 == 0 && cpg * groups == Cin && 64 % cpg == 0 && (W >= 20 || tuning_env("TLXCV_FORCE_SLAB"));
}

cudaError_t conv3x3_slab_set_attributes() {
  cudaError_t e = cudaFuncSetAttribute(conv3x3_slab_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimitB);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv3x3_slab_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimitB);
}

std::string conv3x3_slab_prepare(SlabLaunch& L, int sm_count, const __nv_bfloat16* in, int N, int H, int W, int Cin, int Cout,
                                 int groups, const __nv_bfloat16* packed_w, int Ktot, void* out, const void* residual) {
  if (!g_encode_b) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return "cuTensorMapEncodeTiled is not available from the driver";
    g_encode_b = reinterpret_cast<EncodeTiledFnB>(fn);
  }
  memset(&L, 0, sizeof L);
  SlabParams& p = L.p;
  const int CB = groups > 1 ? 64 : Cin;  // input channels per work item
  p.N = N, p.H = H, p.W = W, p.Cout = Cout;
  p.cblocks = groups > 1 ? Cin / 64 : 1;
  // column segments: as few as possible, equal width, at most 126 output pixels (128 slab pixels with the halo)
  p.segs = (W + 125) / 126;
  p.Ws = (W + p.segs - 1) / p.segs;
  const int Wp = p.Ws + 2;
  p.T = kTileM / Wp;
  p.ng = 1 + (2 + p.T - 1) / p.T;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  if (const char* e = debug_env("TLXCV_DEBUG_ABLATE_SLAB")) p.ablate = atoi(e);  // timing experiments only: results are wrong
  if (Ktot != 9 * CB) return "slab conv: packed weight K does not match 9 taps x channel block";
  const int pix = 2 * CB, row_bytes = Wp * pix;
  const int stage = (p.T * p.Ws * kBlockN * 2 + 127) / 128 * 128;
  const int fixed = 9 * kBlockN * pix + 2 * stage + kMiscB;
  int ng = kMaxGroupsB;
  auto smem_for = [&](int g) { return ((g * p.T + 2) * row_bytes + 1023) / 1024 * 1024 + fixed; };
  while (ng > p.ng + 1 && smem_for(ng) > kSmemLimitB) --ng;
  if (smem_for(ng) > kSmemLimitB) return "slab conv: shared memory cannot hold the input-row ring";
  p.NG = ng;
  L.smem = smem_for(ng);
  long long best = -1;
  for (int bands = 1; bands <= H; ++bands) {
    const int rows = (H + bands - 1) / bands;
    if ((H + rows - 1) / rows != bands) continue;
    const long long items = static_cast<long long>(p.cblocks) * N * bands * p.segs;
    const long long per_cta = (items + sm_count - 1) / sm_count;
    const long long steps = (rows + p.T - 1) / p.T;
    // steps + 0.6 step per band: fitted to bs256 timings (dense 56x56: 1/2/4/7/14 bands = 93/93/85/91/97 us; grouped
    // 56x56: 167/150/156/173/214 us; grouped 28x28: 97/120/153/203 us) - the producer runs ahead across items, so a
    // band costs little beyond its halo rows, and finer bands balance the 148 CTAs better
    const long long cost = per_cta * (steps * 10 + 6);
    if (best < 0 || cost < best) best = cost, p.bands = bands, p.band_rows = rows;
  }
  if (const char* e = tuning_env("TLXCV_DEBUG_SLAB_BANDS")) {  // A/B timing only
    const int bands = std::max(1, std::min(H, atoi(e)));
    p.band_rows = (H + bands - 1) / bands;
    p.bands = (H + p.band_rows - 1) / p.band_rows;
  }
  L.grid = static_cast<int>(std::min<long long>(static_cast<long long>(p.cblocks) * N * p.bands * p.segs, sm_count));
  L.block_n = kBlockN;
  L.cb = CB;
  L.threads = kThreadsB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)Wp, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_b(&L.tmapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(in), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "slab conv: cuTensorMapEncodeTiled (input rows) failed";
  }
  {
    const int cout_pad = ((Cout + 255) / 256) * 256;
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)CB, (cuuint32_t)kBlockN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_b(&L.tmapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(packed_w), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "slab conv: cuTensorMapEncodeTiled (weights) failed";
  }
  return "";
}

cudaError_t conv3x3_slab_launch(const SlabLaunch& L, cudaStream_t st) {
  static const char* trace_path = debug_env("TLXCV_DEBUG_TRACE_SLAB");  // debugging: dump CTA 0's pipeline timeline
  if (trace_path != nullptr) {
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 3 * kTraceLen * sizeof(unsigned long long));
    cudaMemsetAsync(dbuf, 0, 3 * kTraceLen * sizeof(unsigned long long), st);
    SlabParams p = L.p;
    p.trace = dbuf;
    if (L.cb == 64)
      conv3x3_slab_kernel<64><<<L.grid, L.threads, L.smem, st>>>(L.tmapA, L.tmapB, p);
    else
      conv3x3_slab_kernel<32><<<L.grid, L.threads, L.smem, st>>>(L.tmapA, L.tmapB, p);
    cudaStreamSynchronize(st);
    std::vector<unsigned long long> h(3 * kTraceLen);
    cudaMemcpy(h.data(), dbuf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "wb")) {
      fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
      fclose(f);
    }
    return cudaGetLastError();
  }
  if (L.cb == 64) return launch_pdl(conv3x3_slab_kernel<64>, L.grid, L.threads, L.smem, st, L.tmapA, L.tmapB, L.p);
  return launch_pdl(conv3x3_slab_kernel<32>, L.grid, L.threads, L.smem, st, L.tmapA, L.tmapB, L.p);
}

}  // namespace tlxcv
