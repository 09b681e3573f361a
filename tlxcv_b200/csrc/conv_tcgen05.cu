// Implicit-GEMM convolution for sm_100a: NHWC bf16 activations, TMA-staged operand tiles,
// tcgen05.mma accumulating fp32 in TMEM, fused epilogue
//     y = act2( act1( acc * scale + shift ) + residual )
// which is GroupConv2d + BatchNorm2d(eval) + ReLU/ReLU6/LeakyReLU + residual add + ReLU of the
// reference (classification/resnet.py:142-156, resnext.py:109-119, mobilenetv2.py:36-40,
// detection/backbones/darknet.py:54-58,155-159) in one kernel.
//
// GEMM view:  D[M = N*P*Q pixels][Cout] = A[M][K = R*S*Cin] * W[Cout][K]^T
//   A tile  128 pixels x 64 channels of ONE filter tap  (16 KB, SWIZZLE_128B, K-major)
//             kModeTiled    plain 2-D TMA box of the [M][C] matrix (1x1 stride 1)
//             kModeIm2col   im2col-mode TMA: the hardware walks 128 consecutive output pixels of
//                           the NHWC tensor (wrapping rows / images, zero-filling the halo)
//             kModeGatherC4 stems with C_in = 3 (stored as 4): four producer warps gather the
//                           filter rows from NHWC4 global memory straight into the swizzled tile
//   B tile  BLOCK_N filters x 64 K-elements (K-major, SWIZZLE_128B) from the packed weights
//   D       BLOCK_N fp32 TMEM columns x 128 lanes, double buffered (epilogue of tile i overlaps
//           the main loop of tile i+1)
// Warp roles (persistent CTA, one per SM): w0 TMA producer, w1 MMA issuer (converged warp, one elected lane),
// w2 TMEM allocator, w4-7 epilogue (TMEM -> registers -> smem transpose -> coalesced 16 B stores),
// w8-11 gather producers (kModeGatherC4 only).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace tlxcv {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16384
constexpr int kEpiWarps = 8;                    // warps 2..9 (raise to 16 to experiment with p.epi_warps = 16: costs registers)
constexpr int kGatherWarps = 8;                 // warps 10..17 (kModeGatherC4 only)
constexpr int kMaxRing = 4;                     // per epilogue warp: ring of 2 or 4 (32 rows x 64 B) SWIZZLE_64B buffers
#ifdef TLXCV_NARROW_ITEMS                       // A/B builds only
constexpr bool kWideItems = false;
#else
constexpr bool kWideItems = true;               // 4-deep rings of 128/256-wide bf16 tiles run as two 4 KB slots of 32 x 64 items
#endif
constexpr int kWidePrefetchHalf = 0;            // wide items: the next item's residual is requested before this half of the item
// host-side mirror of epilogue_loop's WIDE condition: the output / residual tensor maps then carry 64-column SWIZZLE_128B boxes
inline bool wide_items(int ring, int block_n, bool out_bf16) { return kWideItems && ring == 4 && block_n >= 128 && out_bf16 && kEpiWarps == 8; }
constexpr int kSmemLimit = 232448;              // 227 KB of dynamic shared memory per CTA
constexpr int kScaleBufBytes = 2 * 256 * 4;      // [scale | shift] of one N tile; the kernel has sc_bufs (1 or 2) of them
constexpr int kBarrierBytes = 1024;             // pipeline barriers + kEpiWarps * kRing residual barriers
constexpr int kMaxStages = 8;
constexpr int kThreadsBase = (2 + kEpiWarps) * 32;                // 320
constexpr int kThreadsGather = kThreadsBase + kGatherWarps * 32;  // 576

// KB = channels per K block: 64 (128-byte rows, SWIZZLE_128B), or 32 (64-byte rows, SWIZZLE_64B) for layers with 32 input
// channels, whose 64-wide K blocks would be half zero padding (DarkNet's 32 -> 64 3x3 / stride-2 conv at 608x608)
template <int BLOCK_N, int KB = kBlockK>
struct Cfg {
  static constexpr int kAB = kBlockM * KB * 2;
  static constexpr int kBBytes = BLOCK_N * KB * 2;
  static constexpr int kStageBytes = kAB + kBBytes;
  static constexpr int kStageBytes2 = kAB + kBBytes / 2;  // cta_group::2: each CTA of the pair stages half of B
  // smem layout: [stages x (A | B)] [8 warps x ring x 2 KB] [scale cache] [barriers]; the ring depth
  // and therefore the stage count are chosen per layer (deep ring for residual / HBM-bound layers,
  // more operand stages for MMA-bound ones)
  static constexpr int staging_bytes(int ring, int epi_warps) { return epi_warps * ring * 2048; }
  static constexpr int stages_for(int ring, int sc_bufs, int epi_warps, bool two = false) {
    const int n = (kSmemLimit - staging_bytes(ring, epi_warps) - sc_bufs * kScaleBufBytes - kBarrierBytes) /
                  (two ? kStageBytes2 : kStageBytes);
    return n > kMaxStages ? kMaxStages : n;
  }
  static constexpr int smem_bytes(int ring, int sc_bufs, int epi_warps, bool two = false) {
    return stages_for(ring, sc_bufs, epi_warps, two) * (two ? kStageBytes2 : kStageBytes) + staging_bytes(ring, epi_warps) +
           sc_bufs * kScaleBufBytes + kBarrierBytes;
  }
};

// debug timeline (TLXCV_DEBUG_TRACE_CONV=<file>): CTA 0 records %clock64 at pipeline events, role-major
constexpr int kTraceLenC = 4096;
__device__ __forceinline__ void trace_c(unsigned long long* buf, int role, int& idx) {
  if (buf != nullptr && blockIdx.x == 0 && idx < kTraceLenC) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    buf[role * kTraceLenC + idx++] = t;
  }
}

// ---- cta_group::2 (CTA pair) helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
  return out;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the bytes land in THIS CTA's shared memory, the transaction count goes to `bar`, a
// shared::cluster address that may belong to the peer (the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h,
                                                       int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair when all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
// tcgen05.mma with the descriptors given as (low word, shared high word): no 64-bit arithmetic in the issue loop
template <bool kTwo>
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                               uint32_t accumulate) {
  if (kTwo)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// non-blocking mbarrier phase test, warp-uniform result (all lanes must see the phase complete)
__device__ __forceinline__ bool mbar_test_all(uint32_t bar, uint32_t phase) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
               : "=r"(ok)
               : "r"(bar), "r"(phase)
               : "memory");
  return __all_sync(0xffffffffu, ok != 0);
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

struct PipeState {
  uint32_t stage = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t n_stages) {
    if (++stage == n_stages) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// activation applied to a register tile; the switch is outside the element loop
template <int N>
__device__ __forceinline__ void act_inplace(float (&f)[N], int act, float alpha) {
  if (act == TLXCV_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = fmaxf(f[j], 0.0f);
  } else if (act == TLXCV_ACT_RELU6) {
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = fminf(fmaxf(f[j], 0.0f), 6.0f);
  } else if (act == TLXCV_ACT_LEAKY) {
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = f[j] > 0.0f ? f[j] : f[j] * alpha;
  }
}

// Everything one epilogue warp needs, hoisted out of the loops (kernel parameters live in constant
// memory; re-reading them through the uniform datapath inside the item loop costs latency).
struct EpiArgs {
  uint32_t tmem_base, tmem_full_bar, tmem_empty_bar, ring, res_bar;  // smem addresses are shared-space offsets
  const float* sc_cache;  // smem 2 x [scale(256) | shift(256)]: filled once when the layer has a single N tile (sc_static),
                          // else buffer [acc] is refreshed by the epilogue warps for every tile
  int sc_mode;  // 0: smem buffer filled once; 1: smem buffer [acc] refreshed per tile; 2: global loads per chunk
  const float* scale2;  // DUAL: folded BN of the second accumulator
  const float* shift2;
  int cpw;
  int two, rank;              // cta_group::2: tiles are PAIRS of M tiles, this CTA owns M tile 2 * pair + rank
  uint32_t tmem_empty_remote; // cta_group::2: shared::cluster address of the LEADER's tmem_empty barriers
  const float* scale;
  const float* shift;
  float* out_f32;
  unsigned long long* amax_keys;
  const CUtensorMap* tmap_out;
  const CUtensorMap* tmap_res;
  int M, Cout, n_tiles, num_tiles, first_tile, tile_stride;
  float alpha1, alpha2;
  int ablate;
  unsigned long long* trace;
  // CHAIN kernels (conv -> 1x1 conv in one launch, see conv_chain_kernel): the warp walks whole M tiles (all N chunks of
  // one M tile, then the M tile n_workers further on) and, at the first chunk of every M tile, first turns the finished
  // first-GEMM accumulator into the bf16 A operand of the second GEMM
  int chain_m_tiles;                                      // M tiles of the layer
  uint32_t acc1_full_bar, acc1_empty_bar, a2_full_bar, a2_empty_bar;  // shared-space addresses ([2], [2], [1], [1])
  uint32_t a2_smem;                                       // A2 operand: N1 / 64 SWIZZLE_128B K blocks of 128 rows x 128 B
  const float* sc1_cache;                                 // smem [scale1(N1) | shift1(N1)]
};

// epilogue-warp barrier wait: with TLXCV_LANE0_POLL one lane polls and the warp re-converges (8 polling threads per CTA
// instead of 256); otherwise every lane polls
__device__ __forceinline__ void epi_wait(uint32_t bar, uint32_t parity, int lane) {
#ifdef TLXCV_LANE0_POLL
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
#else
  (void)lane;
  mbar_wait(bar, parity);
#endif
}

template <int ACT>
__device__ __forceinline__ float act1f(float v, float alpha) {
  if (ACT == TLXCV_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == TLXCV_ACT_RELU6) return fminf(fmaxf(v, 0.0f), 6.0f);
  if (ACT == TLXCV_ACT_LEAKY) return v > 0.0f ? v : v * alpha;
  return v;
}

// One epilogue warp: TMEM lane group `lg` (rows), column group `cgroup`.  The warp's work is the
// sequence of its valid 32x32 chunks ("items") over the CTA's tiles.  Per item:
//   tcgen05.ld -> fp32 scale/shift -> act1 -> (+ residual) -> act2 -> bf16 -> TMA store.
// The residual chunk is fetched by TMA into a kRing-deep ring of 2 KB SWIZZLE_64B buffers two
// items ahead; each lane reads ITS row of the buffer, computes, and writes its output row back into
// the same buffer, which one TMA store then drains.  No per-element global addressing, no
// predicates: TMA clips the M and C_out tails.
// DUAL: the tile has TWO accumulators (conv3 of a bottleneck and the block's downsample conv, see the kernel):
//     y = act1( acc1 * scale + shift  +  acc2 * scale2 + shift2 )
// the scale/shift buffer then holds [scale | scale2 | shift + shift2] for the tile's BLOCK_N = 128 channels.
// CHAIN_N1 > 0: second half of a chain kernel (BLOCK_N = 128 chunks of the 1x1 conv), with the first-GEMM hand-over (chain_e1)
// at the start of every M tile; the first GEMM is N1 channels wide.
template <int N1>
__device__ __forceinline__ void chain_e1(const EpiArgs& a, int lg, int cgroup, int lane, uint32_t tile_i);

template <int BLOCK_N, int ACT1, bool RES, int ACT2, bool F32, int kRing, bool DUAL = false, int CHAIN_N1 = 0>
__device__ __forceinline__ void epilogue_loop(const EpiArgs& a, int lg, int cgroup, int lane) {
  static_assert(!DUAL || (BLOCK_N == 128 && !RES && !F32), "dual accumulators: 128-wide bf16 tiles without a residual");
  static_assert(CHAIN_N1 == 0 || (BLOCK_N == 128 && !DUAL && !F32), "chain kernels: 128-wide bf16 chunks");
  constexpr bool CHAIN = CHAIN_N1 > 0;
  constexpr int kAccCols = DUAL ? 2 * BLOCK_N : BLOCK_N;  // TMEM columns per accumulator buffer
  // chunks per warp per tile: the tile's BLOCK_N / 32 chunks over epi_warps / 4 column groups (compile-time for the
  // 8-warp build; a.cpw when kEpiWarps is raised for experiments)
  const int kCpw = kEpiWarps == 8 ? (BLOCK_N / 32) / 2 : a.cpw;
  const int c_first = cgroup * kCpw;
  // WIDE items (the 4-deep ring of 128/256-wide bf16 tiles, i.e. residual and short-K layers, whose pace this loop sets):
  // one item = 32 rows x 64 channels = TWO accumulator chunks behind ONE residual TMA load and ONE TMA store (128-byte
  // rows, SWIZZLE_128B), in a ring of two 4 KB slots.  The per-item fixed costs (bulk-group wait, expect_tx + TMA issue,
  // barrier wait, proxy fence, store issue + commit: ~900 of the ~1850 cycles of a 32-column item, tools/trace_chunks.py)
  // are paid once per 64 columns.
  constexpr bool WIDE = kWideItems && kRing == 4 && BLOCK_N >= 128 && !F32 && kEpiWarps == 8;
  constexpr int kHalves = WIDE ? 2 : 1;            // accumulator chunks per item
  constexpr int kSlotBytes = 2048 * kHalves;
  constexpr uint32_t kSlots = WIDE ? 2 : kRing;    // ring slots
  const int swz_own = WIDE ? (lane & 7) : ((lane >> 1) & 3);
  const uint32_t own_row = a.ring + lane * (64 * kHalves);

  // prefetch cursor (only lane 0 advances it): walks the same item sequence, two items ahead (one for a 2-slot ring)
  // (m_pair, n_tile) of a tile index advance incrementally by (dm, dn) per tile_stride: an integer division per tile
  // costs this warp ~200 dependent cycles, three of them were 8 % of a short-K tile
  // a plain kernel's units are the tiles first_tile, first_tile + tile_stride, ... < num_tiles; a chain kernel's units
  // are the n_tiles chunks of M tile first_tile, then those of M tile first_tile + tile_stride, ...
  const int dm = CHAIN ? 0 : a.tile_stride / a.n_tiles, dn = CHAIN ? 1 : a.tile_stride - dm * a.n_tiles;
  const int carry = CHAIN ? a.tile_stride : 1;
  const int first_m = CHAIN ? a.first_tile : a.first_tile / a.n_tiles, first_n = CHAIN ? 0 : a.first_tile - first_m * a.n_tiles;
  const int n_units = CHAIN ? (a.first_tile < a.chain_m_tiles ? ((a.chain_m_tiles - 1 - a.first_tile) / a.tile_stride + 1) * a.n_tiles : 0)
                            : (a.first_tile < a.num_tiles ? (a.num_tiles - 1 - a.first_tile) / a.tile_stride + 1 : 0);
  int pf_u = 0, pf_ci = 0, pf_m0 = 0, pf_n0 = 0, pf_nmy = 0, pf_mp = first_m, pf_nt = first_n;
  uint32_t pf = 0;
  auto pf_place = [&]() {
    const int m_pair = pf_mp, n_tile = pf_nt;
    const int m_tile = a.two ? 2 * m_pair + a.rank : m_pair;
    pf_m0 = m_tile * kBlockM + lg * 32;
    pf_n0 = n_tile * BLOCK_N;
    pf_nmy = min(min(kCpw, BLOCK_N / 32 - c_first), max(0, (a.Cout - (pf_n0 + c_first * 32) + 31) / 32));
    if (WIDE) pf_nmy = (pf_nmy + 1) >> 1;  // items
  };
  auto pf_issue = [&]() {  // issue the residual load of the next valid item, if any
    while (pf_u < n_units && pf_ci >= pf_nmy) {
      pf_ci = 0;
      ++pf_u;
      pf_mp += dm, pf_nt += dn;
      if (pf_nt >= a.n_tiles) pf_nt -= a.n_tiles, pf_mp += carry;
      if (pf_u < n_units) pf_place();
    }
    if (pf_u >= n_units || (a.ablate & 16)) return;
    const uint32_t slot = pf & (kSlots - 1);
    const uint32_t bar = a.res_bar + slot * 8;
    mbar_arrive_expect_tx(bar, kSlotBytes);
    tma_load_2d(a.ring + slot * kSlotBytes, a.tmap_res, bar, pf_n0 + (c_first + pf_ci * kHalves) * 32, pf_m0);
    ++pf;
    ++pf_ci;
  };
  if (RES && lane == 0) {
    if (pf_u < n_units) pf_place();
#pragma unroll
    for (int k = 0; k < ((kRing == 4 && !WIDE) ? 2 : 1); ++k) pf_issue();
  }

  uint32_t it = 0;  // items processed
  uint32_t acc = 0, acc_phase = 0;
  int tr = 0;
#ifdef TLXCV_FINE_TRACE  // per-chunk events instead of per-tile events (tools/trace_chunks.py); costs a few percent
  const bool fine = (a.ablate & 64) != 0;
#else
  constexpr bool fine = false;
#endif
  const bool tracer = a.trace != nullptr && lg == 2 && cgroup == 0 && lane == 0 && !fine;  // warp 2
  const bool btrace = fine && (a.ablate & 256) != 0;  // tile-boundary events instead of per-chunk events
  const bool ftracer = fine && !btrace && a.trace != nullptr && lg == 2 && cgroup == 0 && lane == 0;
  const bool btracer = btrace && a.trace != nullptr && lg == 2 && cgroup == 0 && lane == 0;
  float4 sc_nx = make_float4(0.f, 0.f, 0.f, 0.f), sh_nx = sc_nx, sc2_nx = sc_nx, sh2_nx = sc_nx;
  const bool sc_lane = a.sc_mode == 1 && lane < kCpw * 8 && c_first * 32 + lane * 4 < BLOCK_N;
  auto fetch_sc = [&](int nt) {  // nt: N tile index
    const int nn0 = nt * BLOCK_N + c_first * 32;
    sc_nx = __ldg(reinterpret_cast<const float4*>(a.scale + nn0) + lane);
    sh_nx = __ldg(reinterpret_cast<const float4*>(a.shift + nn0) + lane);
    if (DUAL) {
      sc2_nx = __ldg(reinterpret_cast<const float4*>(a.scale2 + nn0) + lane);
      sh2_nx = __ldg(reinterpret_cast<const float4*>(a.shift2 + nn0) + lane);
    }
  };
  if (sc_lane && n_units > 0) fetch_sc(first_n);
  int m_pair = first_m, n_tile = first_n;
  uint32_t chain_i = 0;  // CHAIN: M tiles started by this CTA
  for (int unit = 0; unit < n_units; ++unit) {
    if (unit > 0) {
      m_pair += dm, n_tile += dn;
      if (n_tile >= a.n_tiles) n_tile -= a.n_tiles, m_pair += carry;
    }
    if (CHAIN && n_tile == 0) chain_e1<CHAIN_N1 == 0 ? 64 : CHAIN_N1>(a, lg, cgroup, lane, chain_i++);
    const int m_tile = a.two ? 2 * m_pair + a.rank : m_pair;
    const int m0 = m_tile * kBlockM + lg * 32, n0 = n_tile * BLOCK_N;
    const int n_my = max(0, min(min(kCpw, BLOCK_N / 32 - c_first), (a.Cout - (n0 + c_first * 32) + 31) / 32));  // chunks with real channels
    if (tracer) trace_c(a.trace, 2, tr);  // [3k] tile start
    if (btracer) trace_c(a.trace, 2, tr);  // [5k] tile start
    // this warp's slice of the tile's scale / shift was requested one tile ago (an epilogue-bound layer finds its
    // accumulator already complete, so a load issued here would be fully exposed); request the next tile's now
    const float4 sc_pf = sc_nx, sh_pf = sh_nx, sc2_pf = sc2_nx, sh2_pf = sh2_nx;
    if (sc_lane && unit + 1 < n_units) fetch_sc(n_tile + dn >= a.n_tiles ? n_tile + dn - a.n_tiles : n_tile + dn);
    if (btracer) trace_c(a.trace, 2, tr);  // [5k+1] tile set-up done
    epi_wait(a.tmem_full_bar + acc * 8, acc_phase, lane);
    tcgen05_fence_after();
    if (btracer) trace_c(a.trace, 2, tr);  // [5k+2] accumulator complete
    if (tracer) trace_c(a.trace, 2, tr);  // [3k+1] accumulator complete
    float* sc_buf = const_cast<float*>(a.sc_cache) + (a.sc_mode == 1 ? acc * 512 : 0);
    if (a.sc_mode == 1) {
      // Safe to overwrite buffer [acc] now: its previous user (the tile two back) was fully read before every
      // warp released that accumulator (the arrive below comes after the last scale/shift read), and this
      // tile's MMAs could only start after that.  The four warps of a column group write identical values.
      if (sc_lane) {
        reinterpret_cast<float4*>(sc_buf + c_first * 32)[lane] = sc_pf;
        reinterpret_cast<float4*>(sc_buf + 256 + c_first * 32)[lane] = sh_pf;
        if (DUAL) {
          // the two branches' shifts are only ever used as a sum: add them once per tile here, not once per element
          reinterpret_cast<float4*>(sc_buf + 128 + c_first * 32)[lane] = sc2_pf;
          reinterpret_cast<float4*>(sc_buf + 256 + c_first * 32)[lane] =
              make_float4(sh_pf.x + sh2_pf.x, sh_pf.y + sh2_pf.y, sh_pf.z + sh2_pf.z, sh_pf.w + sh2_pf.w);
        }
      }
      __syncwarp();
    }
    auto release_acc = [&]() {  // one arrival per epilogue warp; a CTA pair counts on the leader's barrier
      if (a.two)
        mbar_arrive_cluster(a.tmem_empty_remote + acc * 8);
      else
        mbar_arrive(a.tmem_empty_bar + acc * 8);
    };
    if (n_my == 0 && lane == 0) release_acc();  // nothing to read: release at once
    if (btracer) trace_c(a.trace, 2, tr);  // [5k+3] scale/shift staged, chunk loop starts
    const int n_items = WIDE ? (n_my + 1) >> 1 : n_my;
#pragma unroll 1
    for (int ii = 0; ii < n_items; ++ii, ++it) {
      const uint32_t slot = it & (kSlots - 1);
      const uint32_t row = own_row + slot * kSlotBytes;
      const int cbase0 = n0 + (c_first + ii * kHalves) * 32;  // first channel of the item
      if (ftracer) trace_c(a.trace, 2, tr);  // [6k] item start
#pragma unroll
     for (int hh = 0; hh < kHalves; ++hh) {
      const int ci = ii * kHalves + hh;
      const int chunk = c_first + ci;
      const int cbase = n0 + chunk * 32;
      const bool last_read = ii == n_items - 1 && hh == kHalves - 1;  // this warp's last read of the tile's accumulator
      uint32_t v[32];
      tmem_ld_32x32b_x32(a.tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + acc * kAccCols + chunk * 32, v);
      uint32_t v2[DUAL ? 32 : 1];
      if (DUAL) tmem_ld_32x32b_x32(a.tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + acc * kAccCols + BLOCK_N + chunk * 32,
                                   reinterpret_cast<uint32_t(&)[32]>(v2));
      if (RES) {
        if (kRing == 4 && !WIDE && lane == 0) {
          // four slots: [it-1] draining, [it] in use, [it+1] in flight; request item it+2 into the slot item it-2 used.
          // That store was committed a whole item ago, so this wait does not stall (waiting for the store just
          // committed cost ~300 cycles per item), and the request still leads its use by two items.
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          pf_issue();
        }
        if (WIDE && hh == kWidePrefetchHalf && lane == 0) {
          // two slots: request item it+1 into the slot item it-1 used, once that item's store has read it
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          pf_issue();
        }
        if (hh == 0 && !(a.ablate & 16)) epi_wait(a.res_bar + slot * 8, (it / kSlots) & 1, lane);  // residual item has landed in the ring slot
      } else if (!F32 && hh == 0) {
        // the TMA store that used this slot kSlots items ago must have finished reading it
        if (lane == 0) {
          if (kSlots == 4)
            asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
          else
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
      }
      if (ftracer && hh == 0) trace_c(a.trace, 2, tr);  // [6k+1] residual landed / staging slot free
      tmem_ld_wait();
      if (ftracer && hh == 0) trace_c(a.trace, 2, tr);  // [6k+2] accumulator chunk in registers
      if (a.ablate & 32) {  // timing experiment: accumulator read and dropped
        if (last_read) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) release_acc();
        }
        if (RES && kRing == 2 && lane == 0) pf_issue();
        continue;
      }
      // fp32 pairs (two adjacent channels per 64-bit register): packed FMA / ADD / MUL halve the fp32 instruction
      // count of this loop, which is issue-bound (two epilogue warps per scheduler)
      unsigned long long pr[16];
      if (DUAL) {
        // sum of the two folded-BN branches in fp32, then the block's activation:
        //   acc1 * s1 + (acc2 * s2 + (h1 + h2))
        if (a.sc_mode != 2) {
          const ulonglong2* s1p = reinterpret_cast<const ulonglong2*>(sc_buf + chunk * 32);
          const ulonglong2* s2p = reinterpret_cast<const ulonglong2*>(sc_buf + 128 + chunk * 32);
          const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(sc_buf + 256 + chunk * 32);  // h1 + h2
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const ulonglong2 s1 = s1p[j], s2 = s2p[j], h = hp[j];
            pr[2 * j] = ffma2(pack_u64(v[4 * j], v[4 * j + 1]), s1.x,
                              ffma2(pack_u64(v2[(4 * j) % (DUAL ? 32 : 1)], v2[(4 * j + 1) % (DUAL ? 32 : 1)]), s2.x, h.x));
            pr[2 * j + 1] = ffma2(pack_u64(v[4 * j + 2], v[4 * j + 3]), s1.y,
                                  ffma2(pack_u64(v2[(4 * j + 2) % (DUAL ? 32 : 1)], v2[(4 * j + 3) % (DUAL ? 32 : 1)]), s2.y, h.y));
          }
        } else {
          const ulonglong2* s1p = reinterpret_cast<const ulonglong2*>(a.scale + cbase);
          const ulonglong2* h1p = reinterpret_cast<const ulonglong2*>(a.shift + cbase);
          const ulonglong2* s2p = reinterpret_cast<const ulonglong2*>(a.scale2 + cbase);
          const ulonglong2* h2p = reinterpret_cast<const ulonglong2*>(a.shift2 + cbase);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const ulonglong2 s1 = ldg_u64x2(s1p + j), h1 = ldg_u64x2(h1p + j), s2 = ldg_u64x2(s2p + j), h2 = ldg_u64x2(h2p + j);
            pr[2 * j] = ffma2(pack_u64(v[4 * j], v[4 * j + 1]), s1.x,
                              ffma2(pack_u64(v2[(4 * j) % (DUAL ? 32 : 1)], v2[(4 * j + 1) % (DUAL ? 32 : 1)]), s2.x, fadd2(h1.x, h2.x)));
            pr[2 * j + 1] = ffma2(pack_u64(v[4 * j + 2], v[4 * j + 3]), s1.y,
                                  ffma2(pack_u64(v2[(4 * j + 2) % (DUAL ? 32 : 1)], v2[(4 * j + 3) % (DUAL ? 32 : 1)]), s2.y, fadd2(h1.y, h2.y)));
          }
        }
      } else if (a.sc_mode != 2) {
        const ulonglong2* scp = reinterpret_cast<const ulonglong2*>(sc_buf + chunk * 32);
        const ulonglong2* shp = reinterpret_cast<const ulonglong2*>(sc_buf + 256 + chunk * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const ulonglong2 sc = scp[j], sh = shp[j];
          pr[2 * j] = ffma2(pack_u64(v[4 * j], v[4 * j + 1]), sc.x, sh.x);
          pr[2 * j + 1] = ffma2(pack_u64(v[4 * j + 2], v[4 * j + 3]), sc.y, sh.y);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const ulonglong2 sc = ldg_u64x2(reinterpret_cast<const ulonglong2*>(a.scale + cbase) + j);
          const ulonglong2 sh = ldg_u64x2(reinterpret_cast<const ulonglong2*>(a.shift + cbase) + j);
          pr[2 * j] = ffma2(pack_u64(v[4 * j], v[4 * j + 1]), sc.x, sh.x);
          pr[2 * j + 1] = ffma2(pack_u64(v[4 * j + 2], v[4 * j + 3]), sc.y, sh.y);
        }
      }
      if (last_read) {
        // this warp's last read of the accumulator and of the tile's scale/shift buffer: hand the TMEM
        // buffer back to the MMA warp (one arrival per warp: 256 same-address smem atomics per tile were a
        // measurable cost)
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) release_acc();
      }
      // ReLU / ReLU6 commute with the rounding to bf16 (monotonic, 0 and 6 exact), so they run on packed bf16 pairs
      // after the conversion; anything else, and an activation that precedes the residual add, runs in fp32
      constexpr bool kAct1Packed = !RES && !F32 && (ACT1 == TLXCV_ACT_RELU || ACT1 == TLXCV_ACT_RELU6);
      if (!kAct1Packed) {
#pragma unroll
        for (int j = 0; j < 16; ++j) pr[j] = act_pair_f32<ACT1>(pr[j], a.alpha1);
      }
      if (F32) {
        const int gr = m0 + lane;
        if (gr < a.M && a.amax_keys != nullptr) {
          // fused argmax (tlx.argmax of ImageClassification.predict): best of this lane's 32 columns, first index on ties
          float best = -3.402823466e38f;
          int bi = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(static_cast<uint32_t>(pr[j >> 1] >> ((j & 1) * 32)));
            if (cbase + j < a.Cout && v > best) best = v, bi = j;
          }
          if (cbase < a.Cout) {
            const uint32_t fb = __float_as_uint(best);
            const uint32_t ord = (fb & 0x80000000u) ? ~fb : (fb | 0x80000000u);  // unsigned order == float order
            atomicMax(a.amax_keys + gr, (static_cast<unsigned long long>(ord) << 32) | (0xffffffffu - static_cast<uint32_t>(cbase + bi)));
          }
        }
        if (gr < a.M && a.out_f32 != nullptr) {
          float* dst = a.out_f32 + static_cast<size_t>(gr) * a.Cout + cbase;
          if ((a.Cout & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (cbase + 4 * j < a.Cout) reinterpret_cast<ulonglong2*>(dst)[j] = make_ulonglong2(pr[2 * j], pr[2 * j + 1]);
          } else {  // class counts that are not a multiple of 4 (num_classes=10 heads): rows are not 16-byte aligned
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (cbase + 2 * j < a.Cout) dst[2 * j] = __uint_as_float(static_cast<uint32_t>(pr[j]));
              if (cbase + 2 * j + 1 < a.Cout) dst[2 * j + 1] = __uint_as_float(static_cast<uint32_t>(pr[j] >> 32));
            }
          }
        }
        continue;
      }
      if (ftracer && hh == kHalves - 1) trace_c(a.trace, 2, tr);  // [6k+3] scale/shift/act done (approximately: the compiler may move math)
      uint4 rv[RES ? 4 : 1];
      if (RES) {
        // the four 16-byte pieces of this lane's residual row, requested back to back (volatile asm keeps program order:
        // a load per loop iteration below would serialise load -> math -> store four times)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(rv[j % (RES ? 4 : 1)].x), "=r"(rv[j % (RES ? 4 : 1)].y), "=r"(rv[j % (RES ? 4 : 1)].z), "=r"(rv[j % (RES ? 4 : 1)].w)
                       : "r"(row + (((4 * hh + j) ^ swz_own) << 4)));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t addr = row + (((4 * hh + j) ^ swz_own) << 4);
        uint32_t o[4];
        if (RES) {
          const uint4 val = rv[j % (RES ? 4 : 1)];
          const uint32_t h[4] = {val.x, val.y, val.z, val.w};
          constexpr bool kAct2Packed = ACT2 == TLXCV_ACT_RELU || ACT2 == TLXCV_ACT_RELU6;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            unsigned long long t = fadd2(pr[4 * j + e], bf16x2_to_f32x2(h[e]));  // the add stays in fp32
            if (!kAct2Packed) t = act_pair_f32<ACT2>(t, a.alpha2);
            o[e] = kAct2Packed ? pack_pair_bf16_act<ACT2>(t) : pack_pair_bf16(t);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            o[e] = kAct1Packed ? pack_pair_bf16_act<ACT1>(pr[4 * j + e]) : pack_pair_bf16(pr[4 * j + e]);
          }
        }
        // each lane only ever touches ITS row of the slot, so no warp sync is needed between the
        // residual read and this write
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
      }
     }  // halves
      if (F32 || (a.ablate & 32)) continue;
      if (!(a.ablate & 128)) fence_proxy_async_smem();  // generic-proxy stores -> visible to the TMA (async proxy) read
      __syncwarp();
      if (ftracer) trace_c(a.trace, 2, tr);  // [6k+4] staged
      if (lane == 0) {
        if (!(a.ablate & 2)) tma_store_2d(a.tmap_out, a.ring + slot * kSlotBytes, cbase0, m0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (RES && kRing == 2) {
          // two slots: refill the slot item it-1 used with item it+1 as soon as its store has read the buffer
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          pf_issue();
        }
      }
      if (ftracer) trace_c(a.trace, 2, tr);  // [6k+5] store issued, next residual requested
    }
    if (btracer) trace_c(a.trace, 2, tr);  // [5k+4] chunk loop done
    if (++acc == 2) {
      acc = 0;
      acc_phase ^= 1;
    }
    if (tracer) trace_c(a.trace, 2, tr);  // [3k+2] all chunks of the tile processed
  }
  if (!F32 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// DUAL (BLOCK_N = 128, MODE = kModeTiled): two GEMMs per output tile into two TMEM accumulators,
//   acc1 = A[M][K1] * W1^T   (K blocks 0 .. num_kb1-1: tmapA / tmapB, the bottleneck's last 1x1 conv)
//   acc2 = A2      * W2^T   (remaining K blocks: tmapA2 / tmapB2, the block's 1x1 stride-s downsample conv read
//                            either as a plain [M][C2] matrix (s = 1) or through im2col-mode TMA (s = 2)),
// combined by the epilogue.  It replaces `out = relu(bn3(conv3(x2)) + bn_d(conv_d(x)))` of a ResNet / ResNeXt
// stage's first block (classification/resnet.py:142-156, :246-261) without the downsample map ever reaching HBM.
// TWO (BLOCK_N = 256): the kernel runs as CTA PAIRS (cluster of 2, `tcgen05.mma.cta_group::2`): a pair owns a
// 256-pixel x 256-channel tile, each CTA stages its own 128 pixel rows of A and HALF of the weight tile, the leader
// issues 256x256x16 MMAs that read both CTAs' shared memory and write both CTAs' TMEM, every CTA drains its own 128
// accumulator rows.  Per CTA and K block that is 16 KB + 16 KB from L2 instead of 16 KB + 32 KB: the MMA-bound
// layers are limited by exactly that L2 -> SM operand traffic.
template <int BLOCK_N, int MODE, bool DUAL = false, bool TWO = false, int KB = kBlockK>
__global__ void __launch_bounds__(MODE == kModeGatherC4 ? kThreadsGather : kThreadsBase, 1)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                    const __grid_constant__ CUtensorMap tmapOut, const __grid_constant__ CUtensorMap tmapRes,
                    const __grid_constant__ CUtensorMap tmapA2, const __grid_constant__ CUtensorMap tmapB2,
                    const ConvKernelParams p) {
  using C = Cfg<BLOCK_N, KB>;
  static_assert(KB == kBlockK || (KB == 32 && MODE == kModeIm2col && !DUAL && !TWO), "32-channel K blocks: plain im2col tiles");
  constexpr int kAB = C::kAB;  // bytes of one A tile
  static_assert(!DUAL || (BLOCK_N == 128 && MODE == kModeTiled), "dual accumulators: 128-wide tiles over a tiled first operand");
  static_assert(!TWO || ((BLOCK_N == 256 || BLOCK_N == 128) && !DUAL && MODE != kModeGatherC4), "CTA pairs: 128/256-wide plain tiles");
  constexpr int kAccCols = DUAL ? 2 * BLOCK_N : BLOCK_N;  // TMEM columns per accumulator buffer
  constexpr int kTmemColsK = 2 * kAccCols;
  constexpr int kStageB = TWO ? C::kStageBytes2 : C::kStageBytes;
  const uint32_t rank = TWO ? cluster_ctarank() : 0;   // 0 = leader of the pair
  // SWIZZLE_128B operand tiles need 1024-byte alignment; no pointer casts through integers here, so
  // that the compiler keeps the shared address space (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int n_stages = p.stages, ring = p.ring;
  // K blocks per pipeline slot.  A barrier round (wait, expect_tx / poll, fence, election, commit) costs the producer and
  // the MMA warp ~300-400 cycles whatever it carries, and a 64-wide K block is 128 tensor-pipe cycles: 64-wide tiles with
  // a K loop put TWO K blocks behind one full / empty barrier pair (ResNeXt 512-channel grouped 3x3, bs256: 42 of its 81 us
  // were this skeleton with loads, MMAs and epilogue math all ablated)
  const int kgroup = (!TWO && !DUAL && MODE != kModeGatherC4 && p.kgroup == 2) ? 2 : 1;
  const int slot_b = kStageB * kgroup;
  const int epi_warps = p.epi_warps;
  const int staging_bytes = epi_warps * ring * 2048;
  uint8_t* staging = smem + n_stages * slot_b;
  float* sc_cache = reinterpret_cast<float*>(staging + staging_bytes);  // [BLOCK_N scale | 256: BLOCK_N shift]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + staging_bytes + p.sc_bufs * kScaleBufBytes);
  uint64_t* full_bar = bars;                       // [kStages]  operands landed
  uint64_t* empty_bar = bars + kMaxStages;         // [kStages]  MMAs that read the stage retired
  uint64_t* tmem_full_bar = bars + 2 * kMaxStages; // [2]        accumulator complete
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]        accumulator drained by the epilogue
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint64_t* res_bar = bars + 2 * kMaxStages + 8;   // [kEpiWarps][kMaxRing] residual chunk landed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work units: tiles, or (pair of M tiles) x N tile for CTA pairs; unit u belongs to worker u mod n_workers
  const int num_tiles = TWO ? ((p.m_tiles + 1) / 2) * p.n_tiles : p.m_tiles * p.n_tiles;
  const int worker = TWO ? blockIdx.x >> 1 : blockIdx.x, n_workers = TWO ? gridDim.x >> 1 : gridDim.x;

  if (warp == 1 && lane == 0) {
    if (MODE != kModeGatherC4) tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    if (DUAL) {
      tma_prefetch_desc(&tmapA2);
      tma_prefetch_desc(&tmapB2);
    }
    if (!p.out_f32) tma_prefetch_desc(&tmapOut);
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), MODE == kModeGatherC4 ? 1 + kGatherWarps * 32 : 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), TWO ? 2 * epi_warps : epi_warps);  // one arrival per epilogue warp (of the pair)
    }
    for (int i = 0; i < kEpiWarps * kMaxRing; ++i) mbar_init(smem_u32(&res_bar[i]), 1);
    if (!p.out_f32 && p.residual != nullptr) tma_prefetch_desc(&tmapRes);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (TWO)
      tmem_alloc_2sm<kTmemColsK>(smem_u32(tmem_ptr_smem));
    else
      tmem_alloc<kTmemColsK>(smem_u32(tmem_ptr_smem));
  }
  const bool sc_cached = p.n_tiles == 1 && !DUAL;  // one N tile: scale/shift never change, keep them in smem
  if (sc_cached && warp >= 2 && warp < 2 + epi_warps) {
    for (int i = threadIdx.x - 64; i < BLOCK_N; i += epi_warps * 32) {
      sc_cache[i] = p.scale[i];
      sc_cache[256 + i] = p.shift[i];
    }
  }
  tcgen05_fence_before();
  if (TWO)
    cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrival / multicast commit
  else
    __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();               // the previous kernel's activations are complete and visible from here on
  pdl_launch_dependents();  // the next kernel may take this SM as soon as this CTA exits

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      PipeState ps;
      const int PQ = p.P * p.Q;
      int tr = 0;
      for (int tile = worker; tile < num_tiles; tile += n_workers) {
        if (!(p.ablate & 512)) trace_c(p.trace, 0, tr);  // [2k] tile start
        const int m_pair = tile / p.n_tiles, n_tile = tile - m_pair * p.n_tiles;
        const int m_tile = TWO ? 2 * m_pair + static_cast<int>(rank) : m_pair;  // a pair's second tile may lie past M: zero-filled
        const int m0 = m_tile * kBlockM, n0 = n_tile * BLOCK_N;
        int img = 0, base_h = 0, base_w = 0;
        if (MODE == kModeIm2col || (DUAL && p.a2_im2col)) {
          img = m0 / PQ;
          const int rem = m0 - img * PQ;
          const int op = rem / p.Q, oq = rem - op * p.Q;
          base_h = op * p.stride - p.pad;
          base_w = oq * p.stride - p.pad;
        }
        const int c_base = p.a_chan_from_n ? n_tile * BLOCK_N : 0;
        // no divisions in the K loop: this single thread's per-iteration latency bounds small-N tiles
        int r = 0, sx = 0, cb = 0;
        if (kgroup == 2) {  // two K blocks per slot (plain 64-wide tiles): one wait and one expect_tx per pair
          for (int kb0 = 0; kb0 < p.num_kb; kb0 += 2) {
            const int nk = min(2, p.num_kb - kb0);
            mbar_wait(smem_u32(&empty_bar[ps.stage]), ps.phase ^ 1);
            const uint32_t bar = smem_u32(&full_bar[ps.stage]);
            if (p.ablate & 8) {  // timing experiment: operands never loaded
              mbar_arrive(bar);
              ps.advance(n_stages);
              continue;
            }
            mbar_arrive_expect_tx(bar, nk * C::kStageBytes);
            for (int j = 0; j < nk; ++j) {
              const int kb = kb0 + j;
              const uint32_t a_dst = smem_u32(smem + ps.stage * slot_b + j * kStageB);
              if (MODE == kModeTiled) {
                tma_load_2d(a_dst, &tmapA, bar, kb * KB, m0);
              } else if (KB == kBlockK && p.pair_taps != 0) {
                const uint32_t e = p.pair_taps >> (4 * kb);
                tma_load_im2col_4d(a_dst, (e & 1) ? &tmapA2 : &tmapA, bar, 0, base_w, base_h, img, static_cast<uint16_t>((e >> 1) & 1),
                                   static_cast<uint16_t>((e >> 2) & 1));
              } else {
                tma_load_im2col_4d(a_dst, &tmapA, bar, c_base + cb * KB, base_w, base_h, img,
                                   static_cast<uint16_t>(sx * p.dil), static_cast<uint16_t>(r * p.dil));
                if (++cb == p.kb_per_tap) {
                  cb = 0;
                  if (++sx == p.S) sx = 0, ++r;
                }
              }
              tma_load_2d(a_dst + kAB, &tmapB, bar, kb * KB, n0);
            }
            ps.advance(n_stages);
          }
          if (!(p.ablate & 512)) trace_c(p.trace, 0, tr);  // [2k+1] all loads of the tile issued
          continue;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
#ifdef TLXCV_FINE_TRACE
          if (p.ablate & 512) trace_c(p.trace, 0, tr);  // round trace: before the empty wait
#endif
          mbar_wait(smem_u32(&empty_bar[ps.stage]), ps.phase ^ 1);
#ifdef TLXCV_FINE_TRACE
          if (p.ablate & 512) trace_c(p.trace, 0, tr);  // round trace: slot free
#endif
          const uint32_t bar = smem_u32(&full_bar[ps.stage]);
          const uint32_t a_dst = smem_u32(smem + ps.stage * slot_b);
          if constexpr (TWO) {
            // both CTAs' bytes are counted on the LEADER's full barrier, which alone is armed (for both halves)
            const uint32_t lbar = mapa_u32(bar, 0);
            if (p.ablate & 8) {  // timing experiment: operands never loaded
              if (rank == 0) mbar_arrive(bar);
              ps.advance(n_stages);
              continue;
            }
            if (rank == 0) mbar_arrive_expect_tx(bar, 2 * kStageB);
            if (MODE == kModeTiled) {
              tma_load_2d_2sm(a_dst, &tmapA, lbar, kb * KB, m0);
            } else {
              tma_load_im2col_4d_2sm(a_dst, &tmapA, lbar, c_base + cb * KB, base_w, base_h, img,
                                     static_cast<uint16_t>(sx * p.dil), static_cast<uint16_t>(r * p.dil));
              if (++cb == p.kb_per_tap) {
                cb = 0;
                if (++sx == p.S) sx = 0, ++r;
              }
            }
            tma_load_2d_2sm(a_dst + kAB, &tmapB, lbar, kb * KB, n0 + static_cast<int>(rank) * (BLOCK_N / 2));
            ps.advance(n_stages);
            continue;
          }
          if (p.ablate & 8) {  // timing experiment: operands never loaded
            mbar_arrive(bar);
            ps.advance(n_stages);
            continue;
          }
          mbar_arrive_expect_tx(bar, MODE == kModeGatherC4 ? C::kBBytes : C::kStageBytes);
          if (DUAL && kb >= p.num_kb1) {
            // second GEMM: 1x1 filter, so the K block is just a 64-channel slice of the (strided) input pixels
            const int kb2 = kb - p.num_kb1;
            if (p.a2_im2col)
              tma_load_im2col_4d(a_dst, &tmapA2, bar, kb2 * KB, base_w, base_h, img, 0, 0);
            else
              tma_load_2d(a_dst, &tmapA2, bar, kb2 * KB, m0);
            tma_load_2d(a_dst + kAB, &tmapB2, bar, kb2 * KB, n0);
            ps.advance(n_stages);
            continue;
          }
          if (MODE == kModeTiled) {
            tma_load_2d(a_dst, &tmapA, bar, kb * KB, m0);
          } else if (MODE == kModeIm2col && KB == kBlockK && p.pair_taps != 0) {
            // pixel-pair layout: K block -> (even / odd input rows map, pair offset, row offset); base_w / base_h are in pair space
            const uint32_t e = p.pair_taps >> (4 * kb);
            tma_load_im2col_4d(a_dst, (e & 1) ? &tmapA2 : &tmapA, bar, 0, base_w, base_h, img, static_cast<uint16_t>((e >> 1) & 1),
                               static_cast<uint16_t>((e >> 2) & 1));
          } else if (MODE == kModeIm2col) {
            tma_load_im2col_4d(a_dst, &tmapA, bar, c_base + cb * KB, base_w, base_h, img,
                               static_cast<uint16_t>(sx * p.dil), static_cast<uint16_t>(r * p.dil));
            if (++cb == p.kb_per_tap) {
              cb = 0;
              if (++sx == p.S) sx = 0, ++r;
            }
          }
          tma_load_2d(a_dst + kAB, &tmapB, bar, kb * KB, n0);
          ps.advance(n_stages);
        }
        if (!(p.ablate & 512)) trace_c(p.trace, 0, tr);  // [2k+1] all loads of the tile issued
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the warp stays converged, one elected lane issues =====================
    // (issued from a divergent `lane == 0` branch, every tcgen05.mma operand went through an ELECT + R2UR "waterfall"
    // loop: ~100 cycles per instruction, which bounded every 64/128-wide tile; in a converged warp the descriptors
    // live in uniform registers)
    if (rank == 0) {  // in a CTA pair only the leader issues
      constexpr uint32_t idesc = make_idesc_bf16(TWO ? 2 * kBlockM : kBlockM, BLOCK_N);
      // make_kmajor_sw128_desc split in words: high = SBO 1024 B | version 1 | SWIZZLE_128B, low = address >> 4 | LBO 1
      // (64-byte rows: 8-row atoms of 512 B, SWIZZLE_64B)
      constexpr uint32_t desc_hi = KB == 64 ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : ((512u >> 4) | (1u << 14) | (4u << 29));
      const uint32_t smem_lo = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (1u << 16);  // descriptor low word of stage 0
      const bool rtrace = (p.ablate & 512) != 0;  // per-round events instead of per-tile events (fine trace builds)
      const bool tracer1 = lane == 0 && p.trace != nullptr && !rtrace;
#ifdef TLXCV_FINE_TRACE
      const bool rtracer = lane == 0 && p.trace != nullptr && rtrace;
#endif
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      const int num_kb = p.num_kb, num_kb1 = p.num_kb1;
#ifdef TLXCV_DEBUG_TOOLS
      const bool no_mma = (p.ablate & 4) != 0;  // timing experiments
#else
      constexpr bool no_mma = false;
#endif
      PipeState ps;
      uint32_t acc = 0, acc_phase = 0;
      int tr = 0;
      for (int tile = worker; tile < num_tiles; tile += n_workers) {
        if (tracer1) trace_c(p.trace, 1, tr);  // [4k] tile start
        mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
        tcgen05_fence_after();
        if (tracer1) trace_c(p.trace, 1, tr);  // [4k+1] accumulator free
        const uint32_t tmem_d1 = tmem_base + acc * kAccCols;
        if (kgroup == 2) {  // two K blocks per slot: one wait, one election, one commit per pair
          auto issue_pair_kb = [&](int kk, uint32_t a_lo) {
            const uint32_t b_lo = a_lo + (kAB >> 4);
            if (!no_mma) {
#pragma unroll
              for (int k = 0; k < KB / 16; ++k) {
                if (k == 0)
                  umma_bf16_lohi<TWO>(tmem_d1, a_lo, b_lo, desc_hi, idesc, kk != 0);
                else
                  umma_bf16_lohi<TWO>(tmem_d1, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, 1);
              }
            }
          };
          for (int kb = 0; kb < num_kb; kb += 2) {
            const uint32_t st = ps.stage;
            mbar_wait(full0 + st * 8, ps.phase);
            ps.advance(n_stages);
            tcgen05_fence_after();
            if (kb == 0 && tracer1) trace_c(p.trace, 1, tr);  // [4k+2] first operands landed
            const uint32_t a_lo = smem_lo + st * (static_cast<uint32_t>(slot_b) >> 4);
            if (kb + 1 < num_kb) {  // warp-uniform branch outside the election
              if (elect_one_sync()) {
                issue_pair_kb(kb, a_lo);
                issue_pair_kb(kb + 1, a_lo + (kStageB >> 4));
                umma_commit(empty0 + st * 8);
              }
            } else {
              if (elect_one_sync()) {
                issue_pair_kb(kb, a_lo);
                umma_commit(empty0 + st * 8);
              }
            }
            __syncwarp();
          }
        }
        // Up to two K blocks per round: this warp shares its scheduler with two epilogue warps, and the fixed cost of a round
        // (barrier poll, fence, election, uniform-register setup, commit) otherwise exceeds the 256 tensor-core cycles
        // of a 128-wide K block.
        for (int kb = kgroup == 2 ? num_kb : 0; kb < num_kb;) {
          const uint32_t st0 = ps.stage;
#ifdef TLXCV_FINE_TRACE
          if (rtracer) trace_c(p.trace, 1, tr);  // round trace: before the full wait
#endif
          mbar_wait(full0 + st0 * 8, ps.phase);
#ifdef TLXCV_FINE_TRACE
          if (rtracer) trace_c(p.trace, 1, tr);  // round trace: operands landed
#endif
          ps.advance(n_stages);
          const uint32_t st1 = ps.stage;
          // the next K block joins this round only if its operands have landed already (never wait for it: with three
          // 48 KB stages the loads are the critical path)
          const bool two_kb = kb + 1 < num_kb && mbar_test_all(full0 + st1 * 8, ps.phase);
          if (two_kb) ps.advance(n_stages);
          tcgen05_fence_after();
          if (kb == 0 && tracer1) trace_c(p.trace, 1, tr);  // [4k+2] first operands landed
          // The elected block is straight-line code: a conditional commit (or a loop with a break) inside it costs ~100
          // cycles per round in divergence bookkeeping (tools/micro/ctrl_cost4.cu), so the one / two K block cases branch
          // OUTSIDE the election (warp-uniform) and the accumulator hand-over has its own election after the K loop.
          auto issue_kb = [&](int kk, uint32_t st) {
            const bool second = DUAL && kk >= num_kb1;
            const uint32_t tmem_d = second ? tmem_d1 + BLOCK_N : tmem_d1;
            const int kbl = second ? kk - num_kb1 : kk;  // first K block of an accumulator overwrites it
            const uint32_t a_lo = smem_lo + st * (static_cast<uint32_t>(slot_b) >> 4);
            const uint32_t b_lo = a_lo + (kAB >> 4);
            if (!no_mma) {
#pragma unroll
              for (int k = 0; k < KB / 16; ++k) {
                // +32 B per 16-element K step inside the 128 B swizzle atom: +2 in the (addr >> 4) field
                if (k == 0)
                  umma_bf16_lohi<TWO>(tmem_d, a_lo, b_lo, desc_hi, idesc, kbl != 0);
                else
                  umma_bf16_lohi<TWO>(tmem_d, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, 1);
              }
            }
            // frees the smem stage (of both CTAs of a pair) when these MMAs retire
            if (TWO)
              umma_commit_2sm(empty0 + st * 8);
            else
              umma_commit(empty0 + st * 8);
          };
          if (two_kb) {
            if (elect_one_sync()) {
              issue_kb(kb, st0);
              issue_kb(kb + 1, st1);
            }
          } else {
            if (elect_one_sync()) issue_kb(kb, st0);
          }
          __syncwarp();
#ifdef TLXCV_FINE_TRACE
          if (rtracer) {
            trace_c(p.trace, 1, tr);  // round trace: round issued + committed
            if (blockIdx.x == 0 && tr < kTraceLenC) p.trace[kTraceLenC + tr++] = two_kb ? 2 : 1;  // K blocks in this round
          }
#endif
          kb += two_kb ? 2 : 1;
        }
        // accumulator ready for the epilogue (of both CTAs of a pair)
        if (elect_one_sync()) {
          if (TWO)
            umma_commit_2sm(smem_u32(&tmem_full_bar[acc]));
          else
            umma_commit(smem_u32(&tmem_full_bar[acc]));
        }
        __syncwarp();
        if (tracer1) trace_c(p.trace, 1, tr);  // [4k+3] all MMAs of the tile issued
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp < 2 + epi_warps) {
    // ===================== epilogue: 8 warps (see epilogue_loop) =====================
    constexpr bool kRes = MODE != kModeGatherC4;  // stems never carry a residual
    EpiArgs a;
    a.tmem_base = tmem_base;
    a.tmem_full_bar = smem_u32(tmem_full_bar);
    a.tmem_empty_bar = smem_u32(tmem_empty_bar);
    a.ring = smem_u32(staging + (warp - 2) * (ring * 2048));
    a.res_bar = smem_u32(res_bar + (warp - 2) * kMaxRing);
    a.sc_cache = sc_cache;
    a.sc_mode = sc_cached ? 0 : (p.sc_bufs == 2 ? 1 : 2);
    a.scale2 = p.scale2, a.shift2 = p.shift2;
    a.cpw = max(1, (BLOCK_N / 32) / (epi_warps / 4));
    a.scale = p.scale, a.shift = p.shift;
    a.out_f32 = reinterpret_cast<float*>(p.out);
    a.amax_keys = p.amax_keys;
    a.tmap_out = &tmapOut, a.tmap_res = &tmapRes;
    a.M = p.M, a.Cout = p.Cout, a.n_tiles = p.n_tiles, a.num_tiles = num_tiles;
    a.first_tile = worker, a.tile_stride = n_workers;
    a.two = TWO ? 1 : 0, a.rank = static_cast<int>(rank);
    a.tmem_empty_remote = TWO ? mapa_u32(smem_u32(tmem_empty_bar), 0) : 0;
    a.alpha1 = p.alpha1, a.alpha2 = p.alpha2;
    a.ablate = p.ablate;
    a.trace = p.trace;
    const int lg = warp & 3, cgroup = (warp - 2) >> 2;
    const bool res = kRes && p.residual != nullptr;
    if constexpr (DUAL) {
      // the dual tile is a plain (no residual) bf16 tile whose pre-activation value is the sum of both branches
      if (p.act1 == TLXCV_ACT_RELU) {
        if (ring == 2) epilogue_loop<BLOCK_N, TLXCV_ACT_RELU, false, TLXCV_ACT_NONE, false, 2, true>(a, lg, cgroup, lane);
        else epilogue_loop<BLOCK_N, TLXCV_ACT_RELU, false, TLXCV_ACT_NONE, false, 4, true>(a, lg, cgroup, lane);
      } else {
        if (ring == 2) epilogue_loop<BLOCK_N, TLXCV_ACT_NONE, false, TLXCV_ACT_NONE, false, 2, true>(a, lg, cgroup, lane);
        else epilogue_loop<BLOCK_N, TLXCV_ACT_NONE, false, TLXCV_ACT_NONE, false, 4, true>(a, lg, cgroup, lane);
      }
    } else {
#define TLXCV_EPI(A1)                                                                                          \
  if (p.out_f32) epilogue_loop<BLOCK_N, A1, false, TLXCV_ACT_NONE, true, 2>(a, lg, cgroup, lane);              \
  else if (!res && ring == 2) epilogue_loop<BLOCK_N, A1, false, TLXCV_ACT_NONE, false, 2>(a, lg, cgroup, lane); \
  else if (!res) epilogue_loop<BLOCK_N, A1, false, TLXCV_ACT_NONE, false, 4>(a, lg, cgroup, lane);             \
  else if (p.act2 == TLXCV_ACT_RELU && ring == 2) epilogue_loop<BLOCK_N, A1, kRes, TLXCV_ACT_RELU, false, 2>(a, lg, cgroup, lane); \
  else if (p.act2 == TLXCV_ACT_RELU) epilogue_loop<BLOCK_N, A1, kRes, TLXCV_ACT_RELU, false, 4>(a, lg, cgroup, lane); \
  else if (ring == 2) epilogue_loop<BLOCK_N, A1, kRes, TLXCV_ACT_NONE, false, 2>(a, lg, cgroup, lane); \
  else epilogue_loop<BLOCK_N, A1, kRes, TLXCV_ACT_NONE, false, 4>(a, lg, cgroup, lane);
    switch (p.act1) {
      case TLXCV_ACT_RELU: TLXCV_EPI(TLXCV_ACT_RELU) break;
      case TLXCV_ACT_RELU6: TLXCV_EPI(TLXCV_ACT_RELU6) break;
      case TLXCV_ACT_LEAKY: TLXCV_EPI(TLXCV_ACT_LEAKY) break;
      default: TLXCV_EPI(TLXCV_ACT_NONE) break;
    }
#undef TLXCV_EPI
    }
  } else if (MODE == kModeGatherC4 && warp >= 2 + kEpiWarps) {
    // ===================== gather producers (C_in <= 4 stems): 8 warps =====================
    // K layout of one 64-wide block: r_per_kb filter rows x KR elements, element = s*4 + c.
    // Thread -> (A-tile row, 4 of the 8 16-byte chunks of that row): all 8 loads of a K block are
    // issued before the 4 swizzled 16 B stores.
    const int g = threadIdx.x - kThreadsBase;
    const int t = g & 127;         // A-tile row
    const int chalf = g >> 7;      // chunks [4*chalf, 4*chalf+4)
    const int r_per_kb = kBlockK / p.KR;
    const int cpr_shift = p.KR == 16 ? 1 : 2;  // log2(16-byte chunks per filter row)
    const int PQ = p.P * p.Q;
    // Fully asynchronous gather: every tap is an 8-byte cp.async (zero-filled outside the image)
    // straight into the swizzled A tile, and the stage's full barrier is armed with
    // cp.async.mbarrier.arrive.noinc, so a thread never waits for its own loads: up to n_stages
    // K blocks of global latency are in flight per thread, with no registers and no proxy fence
    // (a fence.proxy.async here is a MEMBAR that would drain the outstanding loads).
    // Address generation is the cost here (8 taps per thread per K block), so everything that does
    // not depend on the K block is hoisted: a thread's 4 chunks lie in at most two filter-row slots
    // (A for chunks 0-1, B for chunks 2-3; A == B when a filter row spans 4 chunks), and the column
    // offsets / column validity of its 8 taps are per-tile constants.
    const int slotA = (chalf * 4) >> cpr_shift, slotB = (chalf * 4 + 3) >> cpr_shift;
    const int jmask = (1 << cpr_shift) - 1;
    const int Wd = p.W, Hd = p.H, dil = p.dil;
    PipeState ps;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m = (tile / p.n_tiles) * kBlockM + t;
      const bool row_ok = m < p.M;
      const int img = row_ok ? m / PQ : 0;
      const int rem = m - img * PQ;
      const int op = rem / p.Q, oq = rem - op * p.Q;
      const int ih0 = op * p.stride - p.pad, iw0 = oq * p.stride - p.pad;
      const uint2* img_base = reinterpret_cast<const uint2*>(p.in_c4) + static_cast<size_t>(img) * Hd * Wd;
      int woff[8];
      uint32_t okw = 0;  // bit 2u: first tap of chunk u is inside the filter and the image row; bit 2u+1: second tap
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = (chalf * 4 + u) & jmask;
        const int s0 = 2 * j, s1 = s0 + 1;
        const int w0 = iw0 + s0 * dil, w1 = w0 + dil;
        woff[2 * u] = w0, woff[2 * u + 1] = w1;
        if (row_ok && s0 < p.S && w0 >= 0 && w0 < Wd) okw |= 1u << (2 * u);
        if (row_ok && s1 < p.S && w1 >= 0 && w1 < Wd) okw |= 1u << (2 * u + 1);
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int rA = kb * r_per_kb + slotA, rB = kb * r_per_kb + slotB;
        const int ihA = ih0 + rA * dil, ihB = ih0 + rB * dil;
        const bool okA = rA < p.R && ihA >= 0 && ihA < Hd, okB = rB < p.R && ihB >= 0 && ihB < Hd;
        const uint2* rowA = img_base + static_cast<size_t>(okA ? ihA : 0) * Wd;
        const uint2* rowB = img_base + static_cast<size_t>(okB ? ihB : 0) * Wd;
        mbar_wait(smem_u32(&empty_bar[ps.stage]), ps.phase ^ 1);
        const uint32_t a_row = smem_u32(smem + ps.stage * slot_b + t * 128);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint2* rowp = u < 2 ? rowA : rowB;
          const bool rok = u < 2 ? okA : okB;
          const bool ok0 = rok && ((okw >> (2 * u)) & 1u), ok1 = rok && ((okw >> (2 * u + 1)) & 1u);
          const uint32_t dst = a_row + (((chalf * 4 + u) ^ (t & 7)) << 4);
          if (p.ablate & 1) continue;
          cp_async_8_zfill(dst, ok0 ? rowp + woff[2 * u] : img_base, ok0);
          cp_async_8_zfill(dst + 8, ok1 ? rowp + woff[2 * u + 1] : img_base, ok1);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[ps.stage])) : "memory");
        ps.advance(n_stages);
      }
    }
  }

  tcgen05_fence_before();
  if (TWO)
    cluster_sync_all();  // the peer's shared memory and TMEM stay valid until both CTAs are done
  else
    __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    if (TWO)
      tmem_dealloc_2sm<kTmemColsK>(tmem_base);
    else
      tmem_dealloc<kTmemColsK>(tmem_base);
  }
  if (!TWO && !DUAL && p.amax_keys != nullptr) {
    // fused argmax: the last CTA to get here decodes every row's key and leaves keys / ticket zero for the next launch
    // (the flag lives in the spare tail of the barrier block: a static __shared__ variable would push the kernel past
    // the 227 KB it already requests dynamically)
    volatile uint32_t* last_cta = reinterpret_cast<volatile uint32_t*>(bars + 120);
    if (threadIdx.x == 0) {
      __threadfence();
      *last_cta = atomicAdd(p.amax_ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (*last_cta) {
      __threadfence();
      for (int r = threadIdx.x; r < p.M; r += blockDim.x) {
        const unsigned long long key = __ldcg(p.amax_keys + r);
        p.amax_out[r] = static_cast<long long>(0xffffffffu - static_cast<uint32_t>(key));
        p.amax_keys[r] = 0ull;
      }
      if (threadIdx.x == 0) *p.amax_ticket = 0u;
    }
  }
}


// ------------------------------------------------------------------------------------------------
// CHAIN: conv (3x3 / strided / any im2col geometry, N1 = 64 or 128 output channels) + BN + ReLU  ->  1x1 conv + BN
// (+ residual) (+ ReLU) in ONE launch: the tail of a ResNet bottleneck, `relu(bn3(conv3(relu(bn2(conv2(h))))) + x)`
// (classification/resnet.py:146-155), without the N1-channel intermediate ever reaching HBM, and with the MMA-bound
// 3x3 running under the HBM-bound 1x1's output traffic.
//   GEMM1  acc1[128 px][N1]   = im2col(h)[128][K1] * W1[N1][K1]^T     (TMEM columns 256.., double buffered)
//   E1     A2 = bf16(relu(acc1 * scale1 + shift1))  written by the epilogue warps straight into shared memory in the
//          SWIZZLE_128B K-major layout tcgen05.mma reads (N1 / 64 K blocks of 128 rows x 128 B)
//   GEMM2  acc2[128 px][128]  = A2[128][N1] * W2[chunk][N1]^T         per 128-channel chunk of the 1x1 conv (TMEM columns
//          0..255, double buffered), W2 chunks streamed through the operand ring
//   E2     the ordinary epilogue (scale / shift, residual by TMA, ReLU, TMA store) per chunk
// One operand ring serves both GEMMs; producer and MMA warp walk the same static item order per CTA: the K blocks of G1(i)
// with the chunks of G2(i-1) slotted in at even spacing, then G2(last) - the tensor pipe works on G1(i) while the epilogue
// warps drain the chunks of tile i-1, and turns acc1(i) into A2 while G1(i+1) has already begun.
// (Measured on the first version, which issued all of G2(i-1) after G1(i): the epilogue warps idled for the whole first GEMM.)
// ------------------------------------------------------------------------------------------------
constexpr int kChainBarBase = 2 * kMaxStages + 8 + kEpiWarps * kMaxRing;  // after the residual barriers (index 56)

template <int N1>
__device__ __forceinline__ void chain_e1(const EpiArgs& a, int lg, int cgroup, int lane, uint32_t tile_i) {
  constexpr int kChunks = N1 / 32;            // 32-channel chunks of acc1
  constexpr int kCpw1 = kChunks / 2;          // per warp (two column groups)
  const uint32_t b = tile_i & 1u;
  epi_wait(a.acc1_full_bar + b * 8, (tile_i >> 1) & 1u, lane);
  tcgen05_fence_after();
  // the second GEMM of the previous tile has finished reading A2 (first tile: passes at once)
  epi_wait(a.a2_empty_bar, (tile_i & 1u) ^ 1u, lane);
  const int row = lg * 32 + lane;
  const uint32_t row_base = a.a2_smem + static_cast<uint32_t>(row) * 128u;
  const uint32_t swz = static_cast<uint32_t>(row & 7);
#pragma unroll
  for (int ci = 0; ci < kCpw1; ++ci) {
    const int chunk = cgroup * kCpw1 + ci;
    uint32_t v[32];
    tmem_ld_32x32b_x32(a.tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + 256u + b * N1 + chunk * 32, v);
    tmem_ld_wait();
    const ulonglong2* scp = reinterpret_cast<const ulonglong2*>(a.sc1_cache + chunk * 32);
    const ulonglong2* shp = reinterpret_cast<const ulonglong2*>(a.sc1_cache + N1 + chunk * 32);
    const uint32_t kb_base = row_base + static_cast<uint32_t>((chunk * 32) / 64) * kABytes;
    const uint32_t j0 = static_cast<uint32_t>(((chunk * 32) % 64) / 8);
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // 8 channels = one 16-byte piece of the row
      const ulonglong2 sc0 = scp[2 * q], sc1 = scp[2 * q + 1], sh0 = shp[2 * q], sh1 = shp[2 * q + 1];
      const uint32_t o0 = pack_pair_bf16_act<TLXCV_ACT_RELU>(ffma2(pack_u64(v[8 * q], v[8 * q + 1]), sc0.x, sh0.x));
      const uint32_t o1 = pack_pair_bf16_act<TLXCV_ACT_RELU>(ffma2(pack_u64(v[8 * q + 2], v[8 * q + 3]), sc0.y, sh0.y));
      const uint32_t o2 = pack_pair_bf16_act<TLXCV_ACT_RELU>(ffma2(pack_u64(v[8 * q + 4], v[8 * q + 5]), sc1.x, sh1.x));
      const uint32_t o3 = pack_pair_bf16_act<TLXCV_ACT_RELU>(ffma2(pack_u64(v[8 * q + 6], v[8 * q + 7]), sc1.y, sh1.y));
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(kb_base + (((j0 + q) ^ swz) << 4)), "r"(o0), "r"(o1), "r"(o2), "r"(o3)
                   : "memory");
    }
  }
  fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
  tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) {
    mbar_arrive(a.acc1_empty_bar + b * 8);  // acc1[b] may be overwritten by the first GEMM of tile i + 2
    mbar_arrive(a.a2_full_bar);             // one arrival per epilogue warp: A2 complete when all eight are in
  }
}

// W1RES: the first conv's weights (num_kb1 K blocks of N1 x 64) stay RESIDENT in shared memory for the whole launch instead
// of being streamed again for every tile: the chain kernels are bound by the bytes an SM can take in (~45 B/clk measured,
// tools/micro/l2_feed.cu - the same for data every SM reads and for data only one reads, multicast does not help), and for
// the 64 -> 64 3x3 the weights are 72 KB of the 376 KB a tile pulls in.
template <int N1, bool W1RES>
struct ChainCfg {
  static constexpr int kB1Bytes = N1 * 128;                  // one K block of W1
  static constexpr int kSlot = W1RES ? kABytes : kABytes + kB1Bytes;  // A tile [| W1 K block]; a W2 item uses the A part
  static constexpr int kA2Bytes = (N1 / 64) * kABytes;
  static constexpr int kSc1Bytes = 1024;                     // [scale1 | shift1], N1 <= 128 floats each
  static constexpr int fixed_bytes(int ring, int num_kb1) {
    return (W1RES ? num_kb1 * kB1Bytes : 0) + kA2Bytes + kEpiWarps * ring * 2048 + 2 * kScaleBufBytes + kSc1Bytes + kBarrierBytes;
  }
  static constexpr int stages_for(int ring, int num_kb1) {
    const int n = (kSmemLimit - fixed_bytes(ring, num_kb1)) / kSlot;
    return n > kMaxStages ? kMaxStages : n;
  }
  static constexpr int smem_bytes(int ring, int num_kb1) { return stages_for(ring, num_kb1) * kSlot + fixed_bytes(ring, num_kb1); }
};

template <int N1, bool W1RES>
__global__ void __launch_bounds__(kThreadsBase, 1)
conv_chain_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                  const __grid_constant__ CUtensorMap tmapB2, const __grid_constant__ CUtensorMap tmapOut,
                  const __grid_constant__ CUtensorMap tmapRes, const ConvKernelParams p) {
  using CC = ChainCfg<N1, W1RES>;
  constexpr int BLOCK_N = 128;      // chunk width of the second GEMM
  constexpr int kKb2 = N1 / 64;     // K blocks of the second GEMM
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int n_stages = p.stages, ring = p.ring;
  uint8_t* w1res = smem_raw;                                              // W1RES: num_kb1 K blocks of W1
  uint8_t* smem = smem_raw + (W1RES ? p.num_kb1 * CC::kB1Bytes : 0);      // operand ring
  uint8_t* a2 = smem + n_stages * CC::kSlot;
  uint8_t* staging = a2 + CC::kA2Bytes;
  const int staging_bytes = kEpiWarps * ring * 2048;
  float* sc_cache = reinterpret_cast<float*>(staging + staging_bytes);
  float* sc1_cache = reinterpret_cast<float*>(staging + staging_bytes + 2 * kScaleBufBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + staging_bytes + 2 * kScaleBufBytes + CC::kSc1Bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full_bar = bars + 2 * kMaxStages;   // acc2
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint64_t* res_bar = bars + 2 * kMaxStages + 8;
  uint64_t* acc1_full_bar = bars + kChainBarBase;    // [2]
  uint64_t* acc1_empty_bar = acc1_full_bar + 2;      // [2]
  uint64_t* a2_full_bar = acc1_empty_bar + 2;        // [1]
  uint64_t* a2_empty_bar = a2_full_bar + 1;          // [1]
  uint64_t* w1_full_bar = a2_empty_bar + 1;          // [1]  W1RES: resident weights landed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int worker = blockIdx.x, n_workers = gridDim.x;
  const int m_tiles = p.m_tiles, n_chunks = p.n_tiles;
  const int my_tiles = worker < m_tiles ? (m_tiles - 1 - worker) / n_workers + 1 : 0;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    tma_prefetch_desc(&tmapB2);
    tma_prefetch_desc(&tmapOut);
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), kEpiWarps);
      mbar_init(smem_u32(&acc1_full_bar[i]), 1);
      mbar_init(smem_u32(&acc1_empty_bar[i]), kEpiWarps);
    }
    mbar_init(smem_u32(a2_full_bar), kEpiWarps);
    mbar_init(smem_u32(a2_empty_bar), 1);
    mbar_init(smem_u32(w1_full_bar), 1);
    for (int i = 0; i < kEpiWarps * kMaxRing; ++i) mbar_init(smem_u32(&res_bar[i]), 1);
    if (p.residual != nullptr) tma_prefetch_desc(&tmapRes);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(tmem_ptr_smem));
  const bool sc_cached = n_chunks == 1;
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < N1; i += kEpiWarps * 32) {  // first BN: N1 scale / shift pairs, constant for the launch
      sc1_cache[i] = p.scale2[i];
      sc1_cache[N1 + i] = p.shift2[i];
    }
    if (sc_cached)
      for (int i = threadIdx.x - 64; i < BLOCK_N; i += kEpiWarps * 32) {
        sc_cache[i] = p.scale[i];
        sc_cache[256 + i] = p.shift[i];
      }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      PipeState ps;
      const int PQ = p.P * p.Q;
      // Static item order shared with the MMA warp: the K blocks of G1(i), with the chunks of G2(i - 1) slotted in after K block
      // (c + 1) * K1 / (n_chunks + 1) - 1, so that the epilogue warps drain tile i - 1 while the tensor pipe works on tile i.
      auto load_g2_chunk = [&](int c) {
        for (int kb2 = 0; kb2 < kKb2; ++kb2) {
          mbar_wait(smem_u32(&empty_bar[ps.stage]), ps.phase ^ 1);
          const uint32_t bar = smem_u32(&full_bar[ps.stage]);
          if (p.ablate & 8) {  // timing experiment (debug builds): operands never loaded
            mbar_arrive(bar);
            ps.advance(n_stages);
            continue;
          }
          mbar_arrive_expect_tx(bar, kABytes);
          tma_load_2d(smem_u32(smem + ps.stage * CC::kSlot), &tmapB2, bar, kb2 * kBlockK, c * BLOCK_N);
          ps.advance(n_stages);
        }
      };
      if (W1RES && my_tiles > 0) {  // weights are parameters, not activations: they could even precede pdl_wait
        mbar_arrive_expect_tx(smem_u32(w1_full_bar), p.num_kb1 * CC::kB1Bytes);
        for (int kb = 0; kb < p.num_kb1; ++kb) tma_load_2d(smem_u32(w1res + kb * CC::kB1Bytes), &tmapB, smem_u32(w1_full_bar), kb * kBlockK, 0);
      }
      int tr = 0;
      for (int i = 0; i < my_tiles; ++i) {
        trace_c(p.trace, 0, tr);  // [2i] tile start
        const int m0 = (worker + i * n_workers) * kBlockM;
        const int img = m0 / PQ;
        const int rem = m0 - img * PQ;
        const int op = rem / p.Q, oq = rem - op * p.Q;
        const int base_h = op * p.stride - p.pad, base_w = oq * p.stride - p.pad;
        int r = 0, sx = 0, cb = 0, c_next = 0;
        int slot_at = p.num_kb1 / (n_chunks + 1);  // K block in front of which the next G2 chunk goes (no division per K block)
        for (int kb = 0; kb < p.num_kb1; ++kb) {
          if (i > 0 && c_next < n_chunks && kb == slot_at) {
            load_g2_chunk(c_next++);
            slot_at = ((c_next + 1) * p.num_kb1) / (n_chunks + 1);
          }
          mbar_wait(smem_u32(&empty_bar[ps.stage]), ps.phase ^ 1);
          const uint32_t bar = smem_u32(&full_bar[ps.stage]);
          const uint32_t dst = smem_u32(smem + ps.stage * CC::kSlot);
          if (p.ablate & 8) {  // timing experiment (debug builds): operands never loaded
            mbar_arrive(bar);
            ps.advance(n_stages);
            continue;
          }
          mbar_arrive_expect_tx(bar, CC::kSlot);  // == kABytes when the weights are resident
          tma_load_im2col_4d(dst, &tmapA, bar, cb * kBlockK, base_w, base_h, img, static_cast<uint16_t>(sx * p.dil),
                             static_cast<uint16_t>(r * p.dil));
          if (++cb == p.kb_per_tap) {
            cb = 0;
            if (++sx == p.S) sx = 0, ++r;
          }
          if (!W1RES) tma_load_2d(dst + kABytes, &tmapB, bar, kb * kBlockK, 0);
          ps.advance(n_stages);
        }
        if (i > 0)
          while (c_next < n_chunks) load_g2_chunk(c_next++);
        trace_c(p.trace, 0, tr);  // [2i+1] all loads of the tile issued
      }
      if (my_tiles > 0)
        for (int c = 0; c < n_chunks; ++c) load_g2_chunk(c);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane) =====================
    constexpr uint32_t idesc1 = make_idesc_bf16(kBlockM, N1);
    constexpr uint32_t idesc2 = make_idesc_bf16(kBlockM, BLOCK_N);
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t smem_lo = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a2_lo = ((smem_u32(a2) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w1_lo = ((smem_u32(w1res) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    PipeState ps;
    uint32_t acc2 = 0, acc2_phase = 0;
    if (W1RES && my_tiles > 0) {
      mbar_wait(smem_u32(w1_full_bar), 0);
      tcgen05_fence_after();
    }
    int tr = 0;
    const bool tracer1 = lane == 0 && p.trace != nullptr;
#ifdef TLXCV_DEBUG_TOOLS
    const bool no_mma = (p.ablate & 4) != 0;  // timing experiments
#else
    constexpr bool no_mma = false;
#endif
    auto issue_g2_chunk = [&](uint32_t j, int c) {
      if (c == 0) {
        if (tracer1) trace_c(p.trace, 1, tr);  // [6i+2] before the A2 wait
        mbar_wait(smem_u32(a2_full_bar), j & 1u);  // the epilogue warps have written A2 of tile j
        tcgen05_fence_after();
        if (tracer1) trace_c(p.trace, 1, tr);  // [6i+3] A2 of the previous tile is ready
      }
      mbar_wait(smem_u32(&tmem_empty_bar[acc2]), acc2_phase ^ 1);
      tcgen05_fence_after();
      if (c == 0 && tracer1) trace_c(p.trace, 1, tr);  // [6i+4] acc2 buffer free
      const uint32_t tmem_d = tmem_base + acc2 * BLOCK_N;
      for (int kb2 = 0; kb2 < kKb2; ++kb2) {
        const uint32_t st = ps.stage;
        mbar_wait(full0 + st * 8, ps.phase);
        ps.advance(n_stages);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = a2_lo + kb2 * (kABytes >> 4);
          const uint32_t b_lo = smem_lo + st * (CC::kSlot >> 4);
          if (!no_mma)
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_lohi<false>(tmem_d, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc2, (kb2 != 0 || k != 0) ? 1u : 0u);
          umma_commit(empty0 + st * 8);
        }
        __syncwarp();
      }
      // hand-overs outside the per-K-block election (a conditional commit inside it costs ~100 cycles per round)
      if (c == n_chunks - 1) {
        if (elect_one_sync()) {
          umma_commit(smem_u32(&tmem_full_bar[acc2]));
          umma_commit(smem_u32(a2_empty_bar));  // every MMA that reads A2 of this tile has retired
        }
      } else {
        if (elect_one_sync()) umma_commit(smem_u32(&tmem_full_bar[acc2]));
      }
      __syncwarp();
      if (++acc2 == 2) acc2 = 0, acc2_phase ^= 1;
    };
    for (int i = 0; i < my_tiles; ++i) {
      const uint32_t b = static_cast<uint32_t>(i) & 1u;
      if (tracer1) trace_c(p.trace, 1, tr);  // [6i] tile start
      mbar_wait(smem_u32(&acc1_empty_bar[b]), ((static_cast<uint32_t>(i) >> 1) & 1u) ^ 1u);
      tcgen05_fence_after();
      if (tracer1) trace_c(p.trace, 1, tr);  // [6i+1] acc1 buffer free
      if (i == 0 && tracer1) { trace_c(p.trace, 1, tr); trace_c(p.trace, 1, tr); trace_c(p.trace, 1, tr); }  // no G2 inside the first tile
      const uint32_t tmem_d = tmem_base + 256u + b * N1;
      int c_next = 0;
      int slot_at = p.num_kb1 / (n_chunks + 1);
      for (int kb = 0; kb < p.num_kb1; ++kb) {
        if (i > 0 && c_next < n_chunks && kb == slot_at) {
          issue_g2_chunk(static_cast<uint32_t>(i - 1), c_next);
          ++c_next;
          slot_at = ((c_next + 1) * p.num_kb1) / (n_chunks + 1);
        }
        const uint32_t st = ps.stage;
        mbar_wait(full0 + st * 8, ps.phase);
        ps.advance(n_stages);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = smem_lo + st * (CC::kSlot >> 4);
          const uint32_t b_lo = W1RES ? w1_lo + kb * (CC::kB1Bytes >> 4) : a_lo + (kABytes >> 4);
          if (!no_mma)
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_lohi<false>(tmem_d, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc1, (kb != 0 || k != 0) ? 1u : 0u);
          umma_commit(empty0 + st * 8);
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(smem_u32(&acc1_full_bar[b]));  // first GEMM of the tile complete -> epilogue warps (chain_e1)
      __syncwarp();
      if (i > 0)
        for (; c_next < n_chunks; ++c_next) issue_g2_chunk(static_cast<uint32_t>(i - 1), c_next);
      if (tracer1) trace_c(p.trace, 1, tr);  // [6i+5] all MMAs of the iteration issued
    }
    if (my_tiles > 0)
      for (int c = 0; c < n_chunks; ++c) issue_g2_chunk(static_cast<uint32_t>(my_tiles - 1), c);
  } else {
    // ===================== epilogue: 8 warps =====================
    EpiArgs a;
    a.tmem_base = tmem_base;
    a.tmem_full_bar = smem_u32(tmem_full_bar);
    a.tmem_empty_bar = smem_u32(tmem_empty_bar);
    a.ring = smem_u32(staging + (warp - 2) * (ring * 2048));
    a.res_bar = smem_u32(res_bar + (warp - 2) * kMaxRing);
    a.sc_cache = sc_cache;
    a.sc_mode = sc_cached ? 0 : 1;
    a.scale2 = nullptr, a.shift2 = nullptr;
    a.cpw = 2;
    a.scale = p.scale, a.shift = p.shift;
    a.out_f32 = nullptr;
    a.amax_keys = nullptr;
    a.tmap_out = &tmapOut, a.tmap_res = &tmapRes;
    a.M = p.M, a.Cout = p.Cout, a.n_tiles = n_chunks, a.num_tiles = m_tiles * n_chunks;
    a.first_tile = worker, a.tile_stride = n_workers;
    a.two = 0, a.rank = 0, a.tmem_empty_remote = 0;
    a.alpha1 = p.alpha1, a.alpha2 = p.alpha2;
    a.ablate = p.ablate;
    a.trace = p.trace;
    a.chain_m_tiles = m_tiles;
    a.acc1_full_bar = smem_u32(acc1_full_bar), a.acc1_empty_bar = smem_u32(acc1_empty_bar);
    a.a2_full_bar = smem_u32(a2_full_bar), a.a2_empty_bar = smem_u32(a2_empty_bar);
    a.a2_smem = smem_u32(a2);
    a.sc1_cache = sc1_cache;
    const int lg = warp & 3, cgroup = (warp - 2) >> 2;
    const bool res = p.residual != nullptr;
    const bool relu2 = p.act2 == TLXCV_ACT_RELU;
    // y = act2(acc2 * scale + shift + residual) or act1(acc2 * scale + shift); act in {none, relu} (checked by the planner)
    if (res) {
      if (relu2) {
        if (ring == 2) epilogue_loop<BLOCK_N, TLXCV_ACT_NONE, true, TLXCV_ACT_RELU, false, 2, false, N1>(a, lg, cgroup, lane);
        else epilogue_loop<BLOCK_N, TLXCV_ACT_NONE, true, TLXCV_ACT_RELU, false, 4, false, N1>(a, lg, cgroup, lane);
      } else {
        if (ring == 2) epilogue_loop<BLOCK_N, TLXCV_ACT_NONE, true, TLXCV_ACT_NONE, false, 2, false, N1>(a, lg, cgroup, lane);
        else epilogue_loop<BLOCK_N, TLXCV_ACT_NONE, true, TLXCV_ACT_NONE, false, 4, false, N1>(a, lg, cgroup, lane);
      }
    } else if (p.act1 == TLXCV_ACT_RELU) {
      if (ring == 2) epilogue_loop<BLOCK_N, TLXCV_ACT_RELU, false, TLXCV_ACT_NONE, false, 2, false, N1>(a, lg, cgroup, lane);
      else epilogue_loop<BLOCK_N, TLXCV_ACT_RELU, false, TLXCV_ACT_NONE, false, 4, false, N1>(a, lg, cgroup, lane);
    } else {
      if (ring == 2) epilogue_loop<BLOCK_N, TLXCV_ACT_NONE, false, TLXCV_ACT_NONE, false, 2, false, N1>(a, lg, cgroup, lane);
      else epilogue_loop<BLOCK_N, TLXCV_ACT_NONE, false, TLXCV_ACT_NONE, false, 4, false, N1>(a, lg, cgroup, lane);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;

std::string load_driver_entry_points() {
  if (g_encode_tiled && g_encode_im2col) return "";
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
    return "cuTensorMapEncodeTiled is not available from the driver";
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
    return "cuTensorMapEncodeIm2col is not available from the driver";
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  return "";
}

std::string encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_bytes,
                      uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu) stride=%llu box=(%u,%u)", int(r),
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_bytes, box_inner, box_outer);
    return buf;
  }
  return "";
}

// row_bytes / image_bytes: byte strides of the H and N dimensions (0 = dense NHWC); pad_hi: padding on the high side (-1 = pad)
std::string encode_im2col(CUtensorMap* map, const void* base, int N, int H, int W, int C, int R, int S, int stride,
                          int pad, int dil, int kb = kBlockK, uint64_t row_bytes = 0, uint64_t image_bytes = 0, int pad_hi = -1) {
  if (pad_hi < 0) pad_hi = pad;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, row_bytes ? row_bytes : (cuuint64_t)W * C * 2,
                           image_bytes ? image_bytes : (cuuint64_t)H * W * C * 2};
  // bounding box of base pixels: lower = -pad, upper = pad_hi - (filter-1)*dilation  (W, H order)
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad_hi - (S - 1) * dil, pad_hi - (R - 1) * dil};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                               upper, kb, kBlockM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               kb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeIm2col failed (%d) NHWC=(%d,%d,%d,%d) RS=(%d,%d) stride=%d pad=%d", int(r),
             N, H, W, C, R, S, stride, pad);
    return buf;
  }
  // Known driver quirk for small tensors (the CUTLASS im2col descriptor builder applies the same fix).
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010) {
    const size_t bytes = static_cast<size_t>(N) * H * W * C * 2;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  }
  return "";
}

inline int& conv_launch_counter() {
  static int n = 0;
  return n;
}

// cluster of 2 + programmatic dependent launch
template <typename... KArgs, typename... Args>
cudaError_t launch_pair(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr, cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <int BLOCK_N, int MODE, bool DUAL = false, bool TWO = false, int KB = kBlockK>
cudaError_t launch_t(const TcConvLaunch& L, cudaStream_t st) {
  // debugging: dump CTA 0's timeline of conv launch number TLXCV_DEBUG_TRACE_CONV_INDEX (default: every launch, so the
  // file holds the last one)
  static const char* trace_path = debug_env("TLXCV_DEBUG_TRACE_CONV");
  static const int trace_index = debug_env("TLXCV_DEBUG_TRACE_CONV_INDEX") ? atoi(debug_env("TLXCV_DEBUG_TRACE_CONV_INDEX")) : -1;
  static int& launch_counter = conv_launch_counter();
  const int my_index = launch_counter++;
  if (trace_path != nullptr && (trace_index < 0 || trace_index == my_index)) {
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 3 * kTraceLenC * sizeof(unsigned long long));
    cudaMemsetAsync(dbuf, 0, 3 * kTraceLenC * sizeof(unsigned long long), st);
    ConvKernelParams p = L.p;
    p.trace = dbuf;
    if (TWO)
      launch_pair(conv_tcgen05_kernel<BLOCK_N, MODE, DUAL, TWO, KB>, L.grid, L.threads, L.smem, st, L.tmapA, L.tmapB, L.tmapOut, L.tmapRes,
                  L.tmapA2, L.tmapB2, p);
    else
      conv_tcgen05_kernel<BLOCK_N, MODE, DUAL, TWO, KB><<<L.grid, L.threads, L.smem, st>>>(L.tmapA, L.tmapB, L.tmapOut, L.tmapRes, L.tmapA2, L.tmapB2, p);
    cudaStreamSynchronize(st);
    std::vector<unsigned long long> h(3 * kTraceLenC);
    cudaMemcpy(h.data(), dbuf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "wb")) {
      fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
      fclose(f);
    }
    return cudaGetLastError();
  }
  if (TWO)
    return launch_pair(conv_tcgen05_kernel<BLOCK_N, MODE, DUAL, TWO, KB>, L.grid, L.threads, L.smem, st, L.tmapA, L.tmapB, L.tmapOut,
                       L.tmapRes, L.tmapA2, L.tmapB2, L.p);
  return launch_pdl(conv_tcgen05_kernel<BLOCK_N, MODE, DUAL, TWO, KB>, L.grid, L.threads, L.smem, st, L.tmapA, L.tmapB, L.tmapOut,
                    L.tmapRes, L.tmapA2, L.tmapB2, L.p);
}

template <int BLOCK_N, int MODE, bool DUAL = false, bool TWO = false, int KB = kBlockK>
cudaError_t set_attr_t() {
  return cudaFuncSetAttribute(conv_tcgen05_kernel<BLOCK_N, MODE, DUAL, TWO, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              kSmemLimit);
}

int smem_for(int block_n, int ring, int sc_bufs, int ew, bool two = false, int kb = kBlockK) {
  if (kb == 32) return Cfg<64, 32>::smem_bytes(ring, sc_bufs, ew);
  if (two) return block_n == 256 ? Cfg<256>::smem_bytes(ring, sc_bufs, ew, true) : Cfg<128>::smem_bytes(ring, sc_bufs, ew, true);
  return block_n == 256 ? Cfg<256>::smem_bytes(ring, sc_bufs, ew)
                        : (block_n == 128 ? Cfg<128>::smem_bytes(ring, sc_bufs, ew) : Cfg<64>::smem_bytes(ring, sc_bufs, ew));
}
int stages_for(int block_n, int ring, int sc_bufs, int ew, bool two = false, int kb = kBlockK) {
  if (kb == 32) return Cfg<64, 32>::stages_for(ring, sc_bufs, ew);
  if (two) return block_n == 256 ? Cfg<256>::stages_for(ring, sc_bufs, ew, true) : Cfg<128>::stages_for(ring, sc_bufs, ew, true);
  return block_n == 256 ? Cfg<256>::stages_for(ring, sc_bufs, ew)
                        : (block_n == 128 ? Cfg<128>::stages_for(ring, sc_bufs, ew) : Cfg<64>::stages_for(ring, sc_bufs, ew));
}

// Epilogue configuration of a launch: 8 epilogue warps; a 4-deep store / residual ring for residual layers and for
// layers with few K blocks per tile (epilogue / HBM bound), a 2-deep ring for MMA-bound layers, which need the shared
// memory for operand stages instead.  16 epilogue warps (TLXCV_DEBUG_EPI_WARPS=16) were measured on B200 and are NOT
// faster: ResNet-50 bs256 3.67 ms against 3.51 ms - the per-SM epilogue rate is set by shared-memory traffic (staging
// stores, TMA store reads, scale/shift broadcasts next to the MMA operand reads) and the per-chunk proxy fence, not by
// the number of warps issuing.
void choose_epilogue(ConvKernelParams& p, int block_n, bool residual, bool out_bf16, bool two = false, int kb = kBlockK) {
  static const int force = tuning_env("TLXCV_DEBUG_EPI_WARPS") ? atoi(tuning_env("TLXCV_DEBUG_EPI_WARPS")) : 0;
  const bool light = p.num_kb <= 8;
  p.epi_warps = 8;
  if (force == 8 || (force == 16 && kEpiWarps >= 16)) p.epi_warps = force;
  if (block_n == 64 && p.epi_warps == 16) p.epi_warps = 8;  // two 32-column chunks per tile: nothing for 16 warps to share
  p.ring = 2;
  if (p.epi_warps == 8 && (residual || light)) p.ring = 4;
  if (!out_bf16) p.ring = 2;
  // scale/shift: one smem buffer filled once (single N tile); two buffers refreshed per tile by the epilogue
  // warps (several N tiles) unless that second buffer would cost an operand stage: then read through __ldg
  p.sc_bufs = (p.n_tiles > 1 && stages_for(block_n, p.ring, 2, p.epi_warps, two, kb) == stages_for(block_n, p.ring, 1, p.epi_warps, two, kb)) ? 2 : 1;
  if (const char* e = tuning_env("TLXCV_DEBUG_SC_BUFS")) p.sc_bufs = atoi(e) == 2 && p.n_tiles > 1 ? 2 : 1;  // A/B timing only
  if (const char* e = tuning_env("TLXCV_DEBUG_RING")) p.ring = (atoi(e) == 4 && out_bf16) ? 4 : 2;  // A/B timing only
  p.stages = stages_for(block_n, p.ring, p.sc_bufs, p.epi_warps, two, kb);
  if (const char* e = tuning_env("TLXCV_DEBUG_STAGES")) p.stages = std::max(2, std::min(p.stages, atoi(e)));  // A/B timing only
}

}  // namespace

int tc_conv_mode(int Cin, int R, int S, int stride, int pad, int groups) {
  if (Cin <= 4) return kModeGatherC4;
  if (R == 1 && S == 1 && stride == 1 && pad == 0 && groups == 1) return kModeTiled;
  return kModeIm2col;
}

int tc_conv_layout(int Cin, int Cout, int R, int S, int stride, int pad, int dil, int groups, int H, int W) {
  if (tc_conv_mode(Cin, R, S, stride, pad, groups) != kModeIm2col || groups != 1 || Cin != 32) return kLayoutPlain;
  if (R == 3 && S == 3 && stride == 2 && pad == 1 && dil == 1 && H % 2 == 0 && W % 2 == 0 && !tuning_env("TLXCV_NO_PIXEL_PAIRS"))
    return kLayoutPixelPairs;
  if (Cout <= 64 && !tuning_env("TLXCV_NO_KB32")) return kLayoutKb32;
  return kLayoutPlain;
}

int tc_conv_packed_k(int Cin, int R, int S, int groups, int mode, int layout) {
  if (layout == kLayoutKb32) return R * S * 32;
  if (layout == kLayoutPixelPairs) return 6 * kBlockK;
  if (mode == kModeGatherC4) {
    const int KR = (S * 4 <= 16) ? 16 : 32;
    const int r_per_kb = kBlockK / KR;
    return ((R + r_per_kb - 1) / r_per_kb) * kBlockK;
  }
  if (groups > 1) return R * S * kBlockK;
  return R * S * ((Cin + kBlockK - 1) / kBlockK) * kBlockK;
}

cudaError_t tc_conv_set_attributes() {
  cudaError_t e;
#define TLXCV_SET(BN, MD) \
  if ((e = set_attr_t<BN, MD>()) != cudaSuccess) return e;
  TLXCV_SET(64, kModeTiled) TLXCV_SET(128, kModeTiled) TLXCV_SET(256, kModeTiled)
  TLXCV_SET(64, kModeIm2col) TLXCV_SET(128, kModeIm2col) TLXCV_SET(256, kModeIm2col)
  TLXCV_SET(64, kModeGatherC4) TLXCV_SET(128, kModeGatherC4)
#undef TLXCV_SET
  if ((e = set_attr_t<128, kModeTiled, true>()) != cudaSuccess) return e;
  if ((e = set_attr_t<256, kModeTiled, false, true>()) != cudaSuccess) return e;
  if ((e = set_attr_t<256, kModeIm2col, false, true>()) != cudaSuccess) return e;
  if ((e = set_attr_t<128, kModeTiled, false, true>()) != cudaSuccess) return e;
  if ((e = set_attr_t<128, kModeIm2col, false, true>()) != cudaSuccess) return e;
  if ((e = set_attr_t<64, kModeIm2col, false, false, 32>()) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(conv_chain_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(conv_chain_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(conv_chain_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)) != cudaSuccess) return e;
  return cudaSuccess;
}

std::string tc_conv_prepare(TcConvLaunch& L, int sm_count, const __nv_bfloat16* act_in, int N, int H, int W, int Cin,
                            int Cin_storage, const __nv_bfloat16* packed_w, int Ktot, int Cout, int R, int S, int stride,
                            int pad, int dil, int groups, int force_block_n, void* out_bf16, int Cout_storage,
                            const void* residual_bf16) {
  std::string err = load_driver_entry_points();
  if (!err.empty()) return err;
  memset(&L, 0, sizeof L);
  const int mode = tc_conv_mode(Cin, R, S, stride, pad, groups);
  const int P = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const int Q = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  const long long M = static_cast<long long>(N) * P * Q;
  if (M <= 0 || M > 0x7fffffffLL) return "conv: output pixel count out of range";
  ConvKernelParams& p = L.p;
  p.M = static_cast<int>(M);
  p.Cout = Cout;
  p.S = S, p.R = R, p.P = P, p.Q = Q, p.H = H, p.W = W, p.stride = stride, p.pad = pad, p.dil = dil;
  // bf16 maps: rows padded to a multiple of 8 channels (16 bytes: the TMA store's global stride unit), the store clips at
  // C_out; fp32 outputs (logits of any class count) are written with direct stores
  if (out_bf16 != nullptr && (Cout_storage % 8 || Cout_storage < Cout)) return "conv: output rows must be padded to a multiple of 8 channels";
  if (residual_bf16 != nullptr && Cout_storage != Cout) return "conv: a residual needs an unpadded output";

  int block_n;
  int kb = kBlockK;  // channels per K block
  int layout = kLayoutPlain;
  if (mode == kModeGatherC4) {
    if (S * 4 > 32) return "stem conv: filter width > 8 is not supported";
    if (Cin_storage != 4) return "stem conv: input must be stored as NHWC4";
    if (groups != 1) return "stem conv: groups must be 1";
    p.KR = (S * 4 <= 16) ? 16 : 32;
    p.num_kb = Ktot / kBlockK;
    p.kb_per_tap = 1;
    p.in_c4 = act_in;
    block_n = Cout <= 64 ? 64 : 128;
  } else if (groups > 1) {
    const int cpg = Cin / groups;
    if (Cin != Cout || cpg * groups != Cin || (kBlockK % cpg) != 0 || (Cin % kBlockK) != 0)
      return "grouped conv: only C_in == C_out with channels-per-group dividing 64 is on the tensor-core path";
    if (dil != 1) return "grouped conv: dilation must be 1";
    p.a_chan_from_n = 1;
    p.kb_per_tap = 1;
    p.num_kb = R * S;
    block_n = 64;
  } else {
    if (Cin % 8) return "conv: C_in must be a multiple of 8 on the tensor-core path";
    layout = tc_conv_layout(Cin, Cout, R, S, stride, pad, dil, groups, H, W);
    kb = layout == kLayoutKb32 ? 32 : kBlockK;
    p.kb_per_tap = (Cin + kb - 1) / kb;
    p.num_kb = R * S * p.kb_per_tap;
    if (layout == kLayoutPixelPairs) {
      // K block -> {valid | row offset | pair offset | odd-row map}: filter rows 1, 2, 0 read input rows 2*oy (even map, row oy),
      // 2*oy + 1 (odd map, row oy) and 2*oy - 1 (odd map, row oy - 1); bases are (oy - 1, ox - 1) in pair space
      p.num_kb = 6, p.kb_per_tap = 1;
      static const unsigned taps[6] = {8 | 4 | 2 | 0, 8 | 4 | 2 | 1, 8 | 0 | 2 | 1, 8 | 4 | 0 | 0, 8 | 4 | 0 | 1, 8 | 0 | 0 | 1};
      for (int j = 0; j < 6; ++j) p.pair_taps |= taps[j] << (4 * j);
      p.stride = 1, p.pad = 1;  // the producer's base coordinates: pair space, stride 1, one pair / row of padding on the low side
    }
    // tile width: minimise (waves x per-tile MMA time); N=64 tiles are shared-memory-bandwidth limited
    const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
    double best = 1e30;
    block_n = 64;
    for (int bn : {256, 128, 64}) {
      if (bn > 64 && bn / 2 >= Cout) continue;  // do not pad C_out by 2x or more
      const long long tiles = static_cast<long long>(m_tiles) * ((Cout + bn - 1) / bn);
      const double waves = static_cast<double>((tiles + sm_count - 1) / sm_count);
      const double mma = p.num_kb * 4.0 * std::max(bn / 2, 48);                       // cycles per tile (tensor pipe)
      const double mem = (kABytes * p.num_kb + 2.0 * kBlockM * bn * 2) / 40.0;          // cycles per tile at ~40 B/clk/SM
      const double cost = waves * (std::max(mma, mem) + 600.0);
      if (cost < best) best = cost, block_n = bn;
    }
  }
  if (force_block_n == 64 || force_block_n == 128 || force_block_n == 256) {
    if (!(groups > 1 && force_block_n != 64) && !(mode == kModeGatherC4 && force_block_n == 256))
      block_n = force_block_n;
  }
  if (kb == 32 && block_n != 64) return "conv: 32-channel K blocks need a 64-wide tile";
  if (Ktot != p.num_kb * kb) return "conv: packed weight K does not match the kernel's K blocking";
  L.kblock = kb;
  p.m_tiles = (p.M + kBlockM - 1) / kBlockM;
  p.n_tiles = (Cout + block_n - 1) / block_n;
  L.mode = mode;
  L.block_n = block_n;
  L.threads = mode == kModeGatherC4 ? kThreadsGather : kThreadsBase;
  // residual layers and short-K (HBM / epilogue bound) layers get the deep store ring; long-K
  // (MMA bound) layers trade it for one more operand stage
  if (const char* e = debug_env("TLXCV_DEBUG_ABLATE")) p.ablate = atoi(e);  // timing experiments only: results are wrong
  // CTA pairs (cta_group::2) for the 256-wide layers with enough K per tile to be bound by MMA / L2 operand traffic
  // rather than by the epilogue; TLXCV_DEBUG_2SM=0/1 forces it off / on where legal
  // (measured on B200, bs256: 14x14 maps 1024->256 35.8 -> 33.8 us, 512->1024 58.4 -> 54.3 us; 7x7 maps with their 98
  //  M tiles lose 3-5 %: pairs halve the number of schedulable units)
  const bool pairable = (block_n == 256 || block_n == 128) && groups == 1 && mode != kModeGatherC4 && out_bf16 != nullptr &&
                        layout == kLayoutPlain;
  // 128-wide pairs measured: no gain.  Few M tiles (7x7 maps: 98) pair only with a long K loop: bs256 512->512 3x3
  // 63.5 -> 57.3 us, 2048->512 33.8 -> 31.7 us, but 512->2048 + residual (8 K blocks) 38.9 -> 43.9 us.
  bool two = pairable && kb == kBlockK && block_n == 256 && ((p.num_kb >= 8 && p.m_tiles >= 256) || (p.num_kb >= 16 && p.m_tiles >= 64));
  // 128-wide 3x3 layers on very large maps (DarkNet / YOLOv3 at 608x608: 64 -> 128 at 152x152, 304 -> 152): bound by the
  // operand feed, the pair halves the weight bytes per SM: 0.36 -> 0.32 / 0.37 -> 0.35 ms (ResNet-50's 28x28 layers: no gain)
  if (pairable && kb == kBlockK && block_n == 128 && mode == kModeIm2col && p.num_kb >= 8 && p.m_tiles >= 4096) two = true;
  // measured again after the MMA-issue and epilogue work (ResNet-50 bs256): 128-wide layers with >= 8 K blocks gain 3-7 %
  // as pairs (3x3 stride-2 128->128 86.6 -> 80.8 us, 1x1 512->128 54.3 -> 52.3), a 256-wide 1x1 WITHOUT a residual gains
  // 10 % on 98 M tiles (512->2048: 37.9 -> 33.9), with a residual it loses (54 -> 60 on 14x14, 41.9 -> 43.8 on 7x7)
  if (pairable && kb == kBlockK && block_n == 128 && p.num_kb >= 8 && p.m_tiles >= 256) two = true;
  if (pairable && kb == kBlockK && block_n == 256 && p.num_kb >= 8 && p.m_tiles >= 64 && residual_bf16 == nullptr) two = true;
  if (const char* e = tuning_env("TLXCV_DEBUG_2SM")) two = atoi(e) != 0 && pairable && kb == kBlockK && p.m_tiles >= 2;
  L.two = two ? 1 : 0;
  choose_epilogue(p, block_n, residual_bf16 != nullptr, out_bf16 != nullptr, two, kb);
  L.smem = smem_for(block_n, p.ring, p.sc_bufs, p.epi_warps, two, kb);
  // 64-wide tiles with a K loop: two K blocks per pipeline slot (half as many barrier rounds); at least two slots
  if (block_n <= (tuning_env("TLXCV_KGROUP_128") ? 128 : 64) && mode != kModeGatherC4 && !two && p.num_kb >= (tuning_env("TLXCV_KGROUP_MIN_KB") ? atoi(tuning_env("TLXCV_KGROUP_MIN_KB")) : 4) && p.stages >= 4 &&
      !tuning_env("TLXCV_NO_KGROUP")) {
    p.kgroup = 2;
    p.stages /= 2;
  }
  const long long tiles = two ? static_cast<long long>((p.m_tiles + 1) / 2) * p.n_tiles : static_cast<long long>(p.m_tiles) * p.n_tiles;
  L.grid = two ? 2 * static_cast<int>(std::min<long long>(tiles, sm_count / 2)) : static_cast<int>(std::min<long long>(tiles, sm_count));

  // B: packed weights [Cout_pad][Ktot], K-major; Cout_pad is a multiple of 256 rows so any tile box is in bounds
  const int cout_pad = ((Cout + 255) / 256) * 256;
  err = encode_2d(&L.tmapB, packed_w, Ktot, cout_pad, static_cast<uint64_t>(Ktot) * 2, kb, L.two ? block_n / 2 : block_n,
                  kb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  if (!err.empty()) return err;
  if (mode == kModeTiled) {
    err = encode_2d(&L.tmapA, act_in, Cin, p.M, static_cast<uint64_t>(Cin_storage) * 2, kBlockK, kBlockM);
  } else if (mode == kModeIm2col) {
    if (Cin_storage != Cin) return "conv: padded channel storage is only supported for stems";
    if (layout == kLayoutPixelPairs) {
      // two maps over (64, W/2, H/2, N): even and odd input rows; a 2 x 2 "filter" with one pair / row of padding on the low side
      const uint64_t row = static_cast<uint64_t>(W) * Cin * 2;
      err = encode_im2col(&L.tmapA, act_in, N, H / 2, W / 2, 64, 2, 2, 1, 1, 1, kBlockK, 2 * row, static_cast<uint64_t>(H) * row, 0);
      if (err.empty())
        err = encode_im2col(&L.tmapA2, reinterpret_cast<const uint8_t*>(act_in) + row, N, H / 2, W / 2, 64, 2, 2, 1, 1, 1, kBlockK,
                            2 * row, static_cast<uint64_t>(H) * row, 0);
    } else {
      err = encode_im2col(&L.tmapA, act_in, N, H, W, Cin, R, S, stride, pad, dil, kb);
    }
  } else {
    L.tmapA = L.tmapB;  // unused
  }
  if (!err.empty()) return err;
  // output [M][Cout] bf16 written by per-warp TMA stores of 32 rows x 32 channels (64 B rows, SWIZZLE_64B) or, for the
  // wide items of a 4-deep ring, 32 rows x 64 channels (128 B rows, SWIZZLE_128B);
  // fp32 outputs (logits) are written with direct stores and leave the map unused
  const bool wide = wide_items(p.ring, block_n, out_bf16 != nullptr);
  if (out_bf16)
    err = encode_2d(&L.tmapOut, out_bf16, Cout, p.M, static_cast<uint64_t>(Cout_storage) * 2, wide ? 64 : 32, 32,
                    wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  else
    L.tmapOut = L.tmapB;
  if (!err.empty()) return err;
  // residual [M][Cout] bf16, fetched by the epilogue warps with the same 32 x 32 boxes
  if (residual_bf16) {
    if (!out_bf16) return "conv: a residual needs a bf16 output";
    if (mode == kModeGatherC4) return "stem conv: residual inputs are not supported";
    err = encode_2d(&L.tmapRes, residual_bf16, Cout, p.M, static_cast<uint64_t>(Cout) * 2, wide ? 64 : 32, 32,
                    wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    p.residual = static_cast<const __nv_bfloat16*>(residual_bf16);
  } else {
    L.tmapRes = L.tmapB;
  }
  if (layout != kLayoutPixelPairs) L.tmapA2 = L.tmapB;  // otherwise only read by dual launches
  L.tmapB2 = L.tmapB;
  return err;
}

std::string tc_conv_prepare_dual(TcConvLaunch& L, int sm_count, const __nv_bfloat16* a1, int M, int K1,
                                 const __nv_bfloat16* w1, const __nv_bfloat16* a2, int N, int H2, int W2, int C2, int stride2,
                                 const __nv_bfloat16* w2, int Cout, void* out_bf16) {
  std::string err = load_driver_entry_points();
  if (!err.empty()) return err;
  memset(&L, 0, sizeof L);
  ConvKernelParams& p = L.p;
  constexpr int block_n = 128;
  const int P = (H2 - 1) / stride2 + 1, Q = (W2 - 1) / stride2 + 1;
  if (static_cast<long long>(N) * P * Q != M) return "dual conv: the two branches do not produce the same pixels";
  if (K1 % 8 || C2 % 8 || Cout % 8) return "dual conv: channel counts must be multiples of 8";
  p.M = M, p.Cout = Cout;
  p.S = 1, p.R = 1, p.P = P, p.Q = Q, p.H = H2, p.W = W2, p.stride = stride2, p.pad = 0, p.dil = 1;
  p.kb_per_tap = 1;
  p.num_kb1 = (K1 + kBlockK - 1) / kBlockK;
  p.num_kb = p.num_kb1 + (C2 + kBlockK - 1) / kBlockK;
  p.a2_im2col = stride2 != 1;
  p.m_tiles = (M + kBlockM - 1) / kBlockM;
  p.n_tiles = (Cout + block_n - 1) / block_n;
  L.mode = kModeTiled, L.block_n = block_n, L.dual = 1, L.threads = kThreadsBase;
  if (const char* e = debug_env("TLXCV_DEBUG_ABLATE")) p.ablate = atoi(e);
  choose_epilogue(p, block_n, false, true);
  if (p.sc_bufs != 2 && stages_for(block_n, p.ring, 2, p.epi_warps) >= 2) {  // the dual epilogue reads four vectors: keep them in smem
    p.sc_bufs = 2;
    p.stages = stages_for(block_n, p.ring, 2, p.epi_warps);
  }
  L.smem = smem_for(block_n, p.ring, p.sc_bufs, p.epi_warps);
  L.grid = static_cast<int>(std::min<long long>(static_cast<long long>(p.m_tiles) * p.n_tiles, sm_count));
  const int cout_pad = ((Cout + 255) / 256) * 256;
  const int k1p = p.num_kb1 * kBlockK, k2p = (p.num_kb - p.num_kb1) * kBlockK;
  if (!(err = encode_2d(&L.tmapB, w1, k1p, cout_pad, static_cast<uint64_t>(k1p) * 2, kBlockK, block_n)).empty()) return err;
  if (!(err = encode_2d(&L.tmapB2, w2, k2p, cout_pad, static_cast<uint64_t>(k2p) * 2, kBlockK, block_n)).empty()) return err;
  if (!(err = encode_2d(&L.tmapA, a1, K1, M, static_cast<uint64_t>(K1) * 2, kBlockK, kBlockM)).empty()) return err;
  if (p.a2_im2col)
    err = encode_im2col(&L.tmapA2, a2, N, H2, W2, C2, 1, 1, stride2, 0, 1);
  else
    err = encode_2d(&L.tmapA2, a2, C2, M, static_cast<uint64_t>(C2) * 2, kBlockK, kBlockM);
  if (!err.empty()) return err;
  const bool wide = wide_items(p.ring, block_n, true);
  if (!(err = encode_2d(&L.tmapOut, out_bf16, Cout, M, static_cast<uint64_t>(Cout) * 2, wide ? 64 : 32, 32,
                        wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)).empty())
    return err;
  L.tmapRes = L.tmapB;
  return "";
}

bool tc_chain_supported(int Cin, int N1, int N2) {
  return (N1 == 64 || N1 == 128) && Cin % 8 == 0 && Cin > 4 && N2 % 8 == 0 && N2 >= 64;
}

std::string tc_chain_prepare(TcConvLaunch& L, int sm_count, const __nv_bfloat16* act_in, int N, int H, int W, int Cin,
                             const __nv_bfloat16* w1, int K1tot, int N1, int R, int S, int stride, int pad, int dil,
                             const __nv_bfloat16* w2, int N2, void* out_bf16, const void* residual_bf16) {
  std::string err = load_driver_entry_points();
  if (!err.empty()) return err;
  memset(&L, 0, sizeof L);
  if (!tc_chain_supported(Cin, N1, N2)) return "chain conv: unsupported channel counts";
  ConvKernelParams& p = L.p;
  const int P = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const int Q = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  const long long M = static_cast<long long>(N) * P * Q;
  if (M <= 0 || M > 0x7fffffffLL) return "conv: output pixel count out of range";
  constexpr int block_n = 128;
  p.M = static_cast<int>(M), p.Cout = N2;
  p.S = S, p.R = R, p.P = P, p.Q = Q, p.H = H, p.W = W, p.stride = stride, p.pad = pad, p.dil = dil;
  p.kb_per_tap = (Cin + kBlockK - 1) / kBlockK;
  p.num_kb1 = R * S * p.kb_per_tap;
  p.num_kb = p.num_kb1 + N1 / kBlockK;
  if (K1tot != p.num_kb1 * kBlockK) return "chain conv: packed weight K does not match the kernel's K blocking";
  p.m_tiles = (p.M + kBlockM - 1) / kBlockM;
  p.n_tiles = (N2 + block_n - 1) / block_n;
  p.epi_warps = kEpiWarps;
  p.sc_bufs = 2;
  if (const char* e = debug_env("TLXCV_DEBUG_ABLATE")) p.ablate = atoi(e);  // timing experiments only (debug build): results are wrong
  // resident first-conv weights (opt-in, TLXCV_CHAIN_W1RES=1): measured on B200 for the 64 -> 64 3x3 + 64 -> 256 chain at
  // 56x56, bs256: 227 us resident against 219 us streamed - the chain is paced by its epilogue warps, not by what the SM
  // takes in, so the 72 KB of shared memory are better spent on operand stages
  const int w1_bytes = p.num_kb1 * N1 * 128;
  const bool w1res = N1 == 64 && w1_bytes <= 80 * 1024 && tuning_env("TLXCV_CHAIN_W1RES") != nullptr;
  auto stages_of = [&](int ring) {
    return N1 == 64 ? (w1res ? ChainCfg<64, true>::stages_for(ring, p.num_kb1) : ChainCfg<64, false>::stages_for(ring, p.num_kb1))
                    : ChainCfg<128, false>::stages_for(ring, p.num_kb1);
  };
  p.ring = (residual_bf16 != nullptr && stages_of(4) >= 3) ? 4 : 2;
  if (const char* e = tuning_env("TLXCV_DEBUG_RING")) p.ring = atoi(e) == 4 ? 4 : 2;
  p.stages = stages_of(p.ring);
  if (p.stages < 2) return "chain conv: not enough shared memory for the operand ring";
  L.smem = N1 == 64 ? (w1res ? ChainCfg<64, true>::smem_bytes(p.ring, p.num_kb1) : ChainCfg<64, false>::smem_bytes(p.ring, p.num_kb1))
                    : ChainCfg<128, false>::smem_bytes(p.ring, p.num_kb1);
  L.mode = kModeIm2col, L.block_n = block_n, L.threads = kThreadsBase, L.chain_n1 = N1, L.chain_w1res = w1res ? 1 : 0;
  L.grid = static_cast<int>(std::min<long long>(p.m_tiles, sm_count));
  const int n2_pad = ((N2 + 255) / 256) * 256;
  if (!(err = encode_2d(&L.tmapB, w1, K1tot, 256, static_cast<uint64_t>(K1tot) * 2, kBlockK, N1)).empty()) return err;
  if (!(err = encode_2d(&L.tmapB2, w2, N1, n2_pad, static_cast<uint64_t>(N1) * 2, kBlockK, block_n)).empty()) return err;
  if (!(err = encode_im2col(&L.tmapA, act_in, N, H, W, Cin, R, S, stride, pad, dil)).empty()) return err;
  const bool wide = wide_items(p.ring, block_n, true);
  if (!(err = encode_2d(&L.tmapOut, out_bf16, N2, p.M, static_cast<uint64_t>(N2) * 2, wide ? 64 : 32, 32,
                        wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)).empty())
    return err;
  if (residual_bf16) {
    if (!(err = encode_2d(&L.tmapRes, residual_bf16, N2, p.M, static_cast<uint64_t>(N2) * 2, wide ? 64 : 32, 32,
                          wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)).empty())
      return err;
    p.residual = static_cast<const __nv_bfloat16*>(residual_bf16);
  } else {
    L.tmapRes = L.tmapB;
  }
  L.tmapA2 = L.tmapB;
  return "";
}

template <int N1, bool W1RES>
cudaError_t launch_chain(const TcConvLaunch& L, cudaStream_t st) {
  // debugging (debug builds): CTA 0's timeline of the last chain launch, TLXCV_DEBUG_TRACE_CHAIN=<file> (tools/trace_chain.py)
  static const char* trace_path = debug_env("TLXCV_DEBUG_TRACE_CHAIN");
  if (trace_path != nullptr) {
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 3 * kTraceLenC * sizeof(unsigned long long));
    cudaMemsetAsync(dbuf, 0, 3 * kTraceLenC * sizeof(unsigned long long), st);
    ConvKernelParams p = L.p;
    p.trace = dbuf;
    conv_chain_kernel<N1, W1RES><<<L.grid, L.threads, L.smem, st>>>(L.tmapA, L.tmapB, L.tmapB2, L.tmapOut, L.tmapRes, p);
    cudaStreamSynchronize(st);
    std::vector<unsigned long long> h(3 * kTraceLenC);
    cudaMemcpy(h.data(), dbuf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "wb")) {
      fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
      fclose(f);
    }
    return cudaGetLastError();
  }
  return launch_pdl(conv_chain_kernel<N1, W1RES>, L.grid, L.threads, L.smem, st, L.tmapA, L.tmapB, L.tmapB2, L.tmapOut, L.tmapRes, L.p);
}

cudaError_t tc_conv_launch(const TcConvLaunch& L, cudaStream_t st) {
  if (L.chain_n1 == 64) return L.chain_w1res ? launch_chain<64, true>(L, st) : launch_chain<64, false>(L, st);
  if (L.chain_n1 == 128) return launch_chain<128, false>(L, st);
  if (L.dual) return launch_t<128, kModeTiled, true>(L, st);
  if (L.two && L.block_n == 256)
    return L.mode == kModeTiled ? launch_t<256, kModeTiled, false, true>(L, st) : launch_t<256, kModeIm2col, false, true>(L, st);
  if (L.two) return L.mode == kModeTiled ? launch_t<128, kModeTiled, false, true>(L, st) : launch_t<128, kModeIm2col, false, true>(L, st);
  if (L.kblock == 32) return launch_t<64, kModeIm2col, false, false, 32>(L, st);
#define TLXCV_CASE(BN, MD) \
  if (L.block_n == BN && L.mode == MD) return launch_t<BN, MD>(L, st);
  TLXCV_CASE(64, kModeTiled) TLXCV_CASE(128, kModeTiled) TLXCV_CASE(256, kModeTiled)
  TLXCV_CASE(64, kModeIm2col) TLXCV_CASE(128, kModeIm2col) TLXCV_CASE(256, kModeIm2col)
  TLXCV_CASE(64, kModeGatherC4) TLXCV_CASE(128, kModeGatherC4)
#undef TLXCV_CASE
  return cudaErrorInvalidValue;
}

}  // namespace tlxcv
