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
// GEMM per STEP of two output rows p, p+1 (one tcgen05.mma issuing thread is the scarce resource, so
// the instruction count per output row is what bounds this kernel):
//   D[128 windows][2*BLOCK_N] = sum over the (sv + R) input rows h of the step
//                               A_h[128][32] * [ B_{h-top} ; B_{h-top-sv} ][2*BLOCK_N][32]^T
// i.e. the slice of input row h is multiplied ONCE by the vertical stack of the two filter rows it
// meets in output rows p and p+1 (zero block where it meets none): 2*(sv+R) MMAs per two rows instead
// of 4*R.  The stacked B operands are contiguous windows of a per-parity "chain"
// [0, B_rmax, B_rmax-sv, ..., 0] kept stationary in shared memory for the whole kernel.
// The slice ring is organised in groups of 2*sv input rows = what one step newly needs: one TMA
// (box 32 x windows x 2sv rows), one full and one empty barrier per step; groups further down the
// band are prefetched into L2 (cp.async.bulk.prefetch.tensor) so the ring can stay shallow.
// Epilogue (8 warps, 4 per output row of the step): tcgen05.ld -> fp32 scale/shift -> activation ->
// bf16 -> swizzled row buffer in shared memory (4 rows) -> either a coalesced 16 B/thread copy of
// the rows to HBM, or (pool) the 3x3/s2/p1 maximum over the last three conv rows, written as the
// pooled row.  The conv activation map never reaches HBM in the pooled variant.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace tlxcv {

namespace {

constexpr int kTileM = 128;        // windows (A rows) per MMA
constexpr int kT = 2;              // output rows per step
constexpr int kMaxR = 7;
constexpr int kRowSlots = 4;       // conv-row buffers (bf16) for the epilogue / pool
constexpr int kEpiWarpsS = 8;
constexpr int kThreadsS = (2 + kEpiWarpsS) * 32;  // 320
constexpr int kMaxGroups = 8;
constexpr int kSmemLimitS = 232448;
constexpr int kMiscBytes = 1024;   // scale/shift (2 x 64 floats) + barriers

template <int BLOCK_N>
struct SCfg {
  static constexpr int kRowBytes = BLOCK_N * 2;         // bytes per window in one output row
  static constexpr int kBlkBytes = BLOCK_N * 64;        // one filter-row block of B: BLOCK_N rows x 32 K
  static constexpr int kRowBuf = kTileM * kRowBytes;    // one conv-row buffer
  static constexpr int kAccBufs = 512 / (kT * BLOCK_N) > 4 ? 4 : 512 / (kT * BLOCK_N);
};

// descriptor low word: start address >> 4 | LBO(ignored)=1 << 16 ; high word: SBO = 512 B, version 1, SWIZZLE_64B
__device__ __forceinline__ uint32_t sw64_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t kSw64DescHi = (512u >> 4) | (1u << 14) | (4u << 29);

template <bool kAccumulate>
__device__ __forceinline__ void umma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(kSw64DescHi), "r"(idesc), "n"(kAccumulate ? 1 : 0)
      : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// One work item = (image, band of output rows, 128-window column tile).
struct Item {
  int n, qt;
  int c0, c1;   // conv rows [c0, c1) wanted from this band
  int j0, j1;   // pooled rows [j0, j1) (pool only)
  int steps;    // ceil((c1 - c0) / 2)
  int hb;       // input row of ring group 0 (may be negative: zero-filled by TMA)
  int groups;   // ring groups the band loads: steps - 1 + ng
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
  it.steps = (it.c1 - it.c0 + kT - 1) / kT;
  it.hb = it.c0 * p.sv - p.pad;
  it.groups = it.steps - 1 + p.ng;
  return it;
}

// first chain block of the stacked operand [B_i ; B_{i-SV}] for input-row position i of a step (see build_chain)
template <int R, int SV>
__host__ __device__ constexpr int chain_block(int i) {
  int off = 0;
  for (int rho = 0; rho < SV; ++rho) {
    const int rmax = (R - 1) - ((R - 1 - rho) % SV);
    if (rho == i % SV) return off + (kT - 1) + (rmax - i) / SV;
    off += 2 * (kT - 1) + (rmax - rho) / SV + 1;
  }
  return 0;
}

template <int BLOCK_N, int R, int SV>
__global__ void __launch_bounds__(kThreadsS, 1)
stem_rowring_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                    const __grid_constant__ CUtensorMap tmapOut, const StemParams p) {
  using C = SCfg<BLOCK_N>;
  constexpr int kRowBytes = C::kRowBytes;
  constexpr int kChunks = kRowBytes / 16;  // 16-byte chunks per window in one output row (4 or 8)
  constexpr int kAccBufs = C::kAccBufs;
  constexpr int kAccCols = kT * BLOCK_N;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int group_bytes = p.G * p.slice_bytes;
  uint8_t* ring = smem;
  uint8_t* bsm = ring + p.NG * group_bytes;
  uint8_t* rows = bsm + p.nblk * C::kBlkBytes;
  float* sc_s = reinterpret_cast<float*>(rows + kRowSlots * C::kRowBuf);
  float* sh_s = sc_s + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sh_s + 64);
  uint64_t* full_bar = bars;                      // [kMaxGroups] group landed
  uint64_t* empty_bar = bars + kMaxGroups;        // [kMaxGroups] MMAs reading the group retired
  uint64_t* tfull_bar = bars + 2 * kMaxGroups;    // [4]
  uint64_t* tempty_bar = tfull_bar + 4;           // [4]
  uint64_t* b_bar = tempty_bar + 4;               // weights landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.N * p.bands * p.q_tiles;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    for (int i = 0; i < kMaxGroups; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), kEpiWarpsS);
    }
    mbar_init(smem_u32(b_bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<kAccBufs * kAccCols>(smem_u32(tmem_ptr_smem));
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
  pdl_wait();               // the previous kernel's output (the padded input) is complete and visible from here on
  pdl_launch_dependents();  // the next kernel may take this SM as soon as this CTA exits

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one group of 2*sv input rows per step =====================
    // The whole warp walks the loop: lane 0 waits / issues the TMA, all lanes prefetch the input rows of
    // the group `ahead` steps further down the band into L2 with plain prefetch instructions (the TMA
    // unit's per-row request rate is the scarce resource, so it only carries the real loads).
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(b_bar), p.nblk * C::kBlkBytes);
      for (int b = 0; b < p.nblk; ++b) tma_load_2d(smem_u32(bsm + b * C::kBlkBytes), &tmapB, smem_u32(b_bar), 0, b * BLOCK_N);
    }
    uint32_t slot = 0, phase = 0;
    const int ahead = p.NG + 2;
    const size_t row_pitch = static_cast<size_t>(p.Wp) * 8;
    const int pf_bytes = p.slice_bytes / 4 + 48;            // bytes of one input row a slice touches
    const int pf_lines = (pf_bytes + 127) / 128 + 1;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item it = decode_item(p, item);
      const int w0 = it.qt * kTileM;
      const uint8_t* img = reinterpret_cast<const uint8_t*>(p.in) + static_cast<size_t>(it.n) * p.H * row_pitch + static_cast<size_t>(w0) * 16;
      for (int g = -ahead; g < it.groups; ++g) {
        const int gp = g + ahead;
        if (gp < it.groups && !(p.ablate & 16)) {
          for (int k = lane; k < p.G * pf_lines; k += 32) {
            const int r = k / pf_lines, l = k - r * pf_lines;
            const int h = it.hb + gp * p.G + r;
            if (h >= 0 && h < p.H && l * 128 < pf_bytes + 127)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(img + static_cast<size_t>(h) * row_pitch + l * 128));
          }
        }
        if (g < 0) continue;
        if (lane == 0) {
          mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1);
          const uint32_t bar = smem_u32(&full_bar[slot]);
          if (p.ablate & 1) {
            mbar_arrive(bar);
          } else {
            mbar_arrive_expect_tx(bar, group_bytes);
            tma_load_4d(smem_u32(ring + slot * group_bytes), &tmapA, bar, 0, w0, it.hb + g * p.G, it.n);
          }
        }
        __syncwarp();
        if (++slot == static_cast<uint32_t>(p.NG)) slot = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the warp stays converged, one elected lane issues =====================
    // Everything per MMA is a compile-time offset from three per-step group bases: the instruction
    // stream of this warp is what bounds the kernel.  Input rows outside the image were zero-filled by
    // TMA, so no MMA is conditional.
    constexpr uint32_t idesc = make_idesc_bf16(kTileM, kAccCols);
    constexpr int G = kT * SV, n_in = (kT - 1) * SV + R, ng = (n_in + G - 1) / G;
    mbar_wait(smem_u32(b_bar), 0);
    tcgen05_fence_after();
    const uint32_t NG = static_cast<uint32_t>(p.NG);
    const uint32_t ring_lo = sw64_desc_lo(smem_u32(ring)), b_lo0 = sw64_desc_lo(smem_u32(bsm));
    const uint32_t slice_lo = static_cast<uint32_t>(p.slice_bytes) >> 4, group_lo = static_cast<uint32_t>(group_bytes) >> 4;
    uint32_t fslot = 0;                  // ring slot of the first group of the current step
    uint32_t wslot = 0, wphase = 0;      // next group to wait for
    uint32_t acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item it = decode_item(p, item);
      int waited = 0;  // groups of this band known to have landed
      for (int m = 0; m < it.steps; ++m) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        while (waited < m + ng) {
          mbar_wait(smem_u32(&full_bar[wslot]), wphase);
          if (++wslot == NG) wslot = 0, wphase ^= 1;
          ++waited;
        }
        tcgen05_fence_after();
        uint32_t gbase[ng];
        {
          uint32_t s = fslot;
#pragma unroll
          for (int k = 0; k < ng; ++k) {
            gbase[k] = ring_lo + s * group_lo;
            if (++s == NG) s = 0;
          }
        }
        if (elect_one()) {
          const uint32_t tmem_d = tmem_base + acc * kAccCols;
          if (!(p.ablate & 2)) {
#pragma unroll
            for (int i = 0; i < n_in; ++i) {
              const uint32_t a_lo = gbase[i / G] + static_cast<uint32_t>(i % G) * slice_lo;
              const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(chain_block<R, SV>(i)) * (C::kBlkBytes >> 4);
              if (i == 0)
                umma_bf16_lo<false>(tmem_d, a_lo, b_lo, idesc);
              else
                umma_bf16_lo<true>(tmem_d, a_lo, b_lo, idesc);
              umma_bf16_lo<true>(tmem_d, a_lo + 2, b_lo + 2, idesc);  // second K=16 step: +32 B inside the swizzle atom
            }
          }
          umma_commit(smem_u32(&tfull_bar[acc]));
          umma_commit(smem_u32(&empty_bar[fslot]));  // the step's first group is dead after these MMAs
        }
        __syncwarp();
        if (++acc == kAccBufs) acc = 0, acc_phase ^= 1;
        if (++fslot == NG) fslot = 0;
      }
      // the last step's other ng-1 groups were loaded for rows below the band: hand them back as well
      for (int k = 1; k < ng; ++k) {
        if (elect_one()) umma_commit(smem_u32(&empty_bar[fslot]));
        __syncwarp();
        if (++fslot == NG) fslot = 0;
      }
    }
  } else {
    // ===================== epilogue: 8 warps, 4 per output row of the step =====================
    const int lg = warp & 3, ch = (warp - 2) >> 2;   // TMEM lane group; output row within the step
    const int et = threadIdx.x - 64;                  // 0..255
    const int px = lg * 32 + lane;                    // window (TMEM lane) this thread owns
    const uint32_t swz = kChunks == 8 ? (px & 7) : ((px >> 1) & 3);
    const uint32_t rows_addr = smem_u32(rows);
    const float alpha = p.alpha;
    const int act = p.act;
    uint32_t acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item it = decode_item(p, item);
      const int valid_w = min(kTileM, p.Qw - it.qt * kTileM);  // windows of this tile that exist
      for (int m = 0; m < it.steps; ++m) {
        const int cs = it.c0 + m * kT;
        const int c = cs + ch;  // this warp's conv row
        // the row buffers written now were last used two steps ago: their TMA stores (unpooled stems) must have
        // finished reading shared memory; only the issuing thread tracks them, the barrier publishes it
        if (!p.pool && et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        epi_bar_sync();         // the previous step's consumers are done with the row buffers
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + acc * kAccCols + ch * BLOCK_N;
        const uint32_t my_row = rows_addr + static_cast<uint32_t>((c + 1) & (kRowSlots - 1)) * C::kRowBuf + px * kRowBytes;
        const bool live = c < it.c1 && !(p.ablate & 4);
#pragma unroll
        for (int half = 0; half < BLOCK_N / 32; ++half) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + half * 32, v);
          tmem_ld_wait();
          if (half == BLOCK_N / 32 - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
          }
          if (live) {
            // packed fp32x2 FMA; ReLU / ReLU6 inside the bf16 conversion, LeakyReLU as max(v, alpha v) (slope in [0, 1],
            // checked by the planner): same values as scale/shift -> activation -> rounding, a third of the instructions
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const ulonglong2 s0 = *reinterpret_cast<const ulonglong2*>(sc_s + half * 32 + 8 * j);
              const ulonglong2 s1 = *reinterpret_cast<const ulonglong2*>(sc_s + half * 32 + 8 * j + 4);
              const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(sh_s + half * 32 + 8 * j);
              const ulonglong2 h1 = *reinterpret_cast<const ulonglong2*>(sh_s + half * 32 + 8 * j + 4);
              unsigned long long q[4];
              q[0] = ffma2(pack_u64(v[8 * j + 0], v[8 * j + 1]), s0.x, h0.x);
              q[1] = ffma2(pack_u64(v[8 * j + 2], v[8 * j + 3]), s0.y, h0.y);
              q[2] = ffma2(pack_u64(v[8 * j + 4], v[8 * j + 5]), s1.x, h1.x);
              q[3] = ffma2(pack_u64(v[8 * j + 6], v[8 * j + 7]), s1.y, h1.y);
              uint32_t o[4];
              if (act == TLXCV_ACT_RELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = pack_pair_bf16_act<TLXCV_ACT_RELU>(q[e]);
              } else if (act == TLXCV_ACT_RELU6) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = pack_pair_bf16_act<TLXCV_ACT_RELU6>(q[e]);
              } else if (act == TLXCV_ACT_LEAKY) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = pack_pair_bf16(act_pair_f32<TLXCV_ACT_LEAKY>(q[e], alpha));
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = pack_pair_bf16(q[e]);
              }
              const uint32_t cidx = static_cast<uint32_t>(half * 4 + j);
              const uint32_t addr = my_row + ((cidx ^ swz) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                           : "memory");
            }
          }
        }
        if (++acc == kAccBufs) acc = 0, acc_phase ^= 1;
        if (!p.pool) fence_proxy_async_smem();  // the rows are read by TMA stores below
        epi_bar_sync();  // both conv rows of the step are complete in their buffers
        if (p.ablate & 8) continue;
        if (!p.pool && !(p.ablate & 32)) {
          // One TMA store per conv row, issued by one thread: the row buffer already is the SWIZZLE_64B / 128B image of
          // [128 windows][BLOCK_N channels], and the copy runs behind the next step's accumulator reads (as a loop of
          // 16-byte loads and stores by all epilogue threads it was serialised with them: 214 of 590 us on the
          // 608x608 DarkNet stem).  Columns past the row end and rows past the band are clipped / skipped.
          if (et == 0) {
            const int n_rows = min(kT, it.c1 - cs);
            for (int t = 0; t < n_rows; ++t)
              tma_store_4d(&tmapOut, rows_addr + static_cast<uint32_t>((cs + t + 1) & (kRowSlots - 1)) * C::kRowBuf, 0,
                           it.qt * kTileM, cs + t, it.n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else if (!p.pool) {
          // copy the rows out: consecutive threads -> consecutive 16 B of the NHWC row
          const int n_rows = min(kT, it.c1 - cs);
          const int per_row = valid_w * kChunks;
          for (int i = et; i < n_rows * per_row; i += kEpiWarpsS * 32) {
            const int t = i >= per_row ? 1 : 0;
            const int k = i - t * per_row;
            const uint32_t w = static_cast<uint32_t>(k) / kChunks, cidx = static_cast<uint32_t>(k) % kChunks;
            const uint32_t wswz = kChunks == 8 ? (w & 7) : ((w >> 1) & 3);
            const uint32_t sa = rows_addr + static_cast<uint32_t>((cs + t + 1) & (kRowSlots - 1)) * C::kRowBuf;
            uint4 val;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                         : "r"(sa + w * kRowBytes + ((cidx ^ wswz) << 4)));
            uint8_t* grow = reinterpret_cast<uint8_t*>(p.out) +
                            ((static_cast<size_t>(it.n) * p.P + cs + t) * p.Qw + static_cast<size_t>(it.qt) * kTileM) * kRowBytes;
            *reinterpret_cast<uint4*>(grow + static_cast<size_t>(k) * 16) = val;
          }
        } else {
          // pooled row j = max over conv rows 2j-1, 2j, 2j+1 and columns 2i-1, 2i, 2i+1.  Padding is
          // -inf, i.e. positions outside the map do not take part: clamping the index re-reads an
          // element that is already in the window, which leaves the maximum unchanged.
#pragma unroll 1
          for (int t = 0; t < kT; ++t) {
            const int cc = cs + t;
            if (cc >= it.c1) break;
            const bool completes = (cc & 1) ? true : (cc == p.P - 1);
            const int j = cc >> 1;
            if (!completes || j < it.j0 || j >= it.j1) continue;
            uint32_t sa[3];
            sa[0] = rows_addr + static_cast<uint32_t>((max(2 * j - 1, 0) + 1) & (kRowSlots - 1)) * C::kRowBuf;
            sa[1] = rows_addr + static_cast<uint32_t>((2 * j + 1) & (kRowSlots - 1)) * C::kRowBuf;
            sa[2] = rows_addr + static_cast<uint32_t>((min(2 * j + 1, p.P - 1) + 1) & (kRowSlots - 1)) * C::kRowBuf;
            uint8_t* grow = reinterpret_cast<uint8_t*>(p.out) + (static_cast<size_t>(it.n) * p.Pp + j) * p.Qp * kRowBytes;
            for (int i = et; i < p.Qp * kChunks; i += kEpiWarpsS * 32) {
              const int qo = i / kChunks;
              const uint32_t cidx = static_cast<uint32_t>(i % kChunks);
              uint32_t woff[3];
#pragma unroll
              for (int dc = 0; dc < 3; ++dc) {
                const uint32_t w = static_cast<uint32_t>(min(max(2 * qo - 1 + dc, 0), p.Qw - 1));
                const uint32_t wswz = kChunks == 8 ? (w & 7) : ((w >> 1) & 3);
                woff[dc] = w * kRowBytes + ((cidx ^ wswz) << 4);
              }
              __nv_bfloat162 mx[4];
#pragma unroll
              for (int dr = 0; dr < 3; ++dr) {
#pragma unroll
                for (int dc = 0; dc < 3; ++dc) {
                  uint4 val;
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                               : "r"(sa[dr] + woff[dc]));
                  const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&val);
                  if (dr == 0 && dc == 0) {
                    mx[0] = hv[0], mx[1] = hv[1], mx[2] = hv[2], mx[3] = hv[3];
                  } else {
                    mx[0] = __hmax2(mx[0], hv[0]), mx[1] = __hmax2(mx[1], hv[1]);
                    mx[2] = __hmax2(mx[2], hv[2]), mx[3] = __hmax2(mx[3], hv[3]);
                  }
                }
              }
              *reinterpret_cast<uint4*>(grow + static_cast<size_t>(i) * 16) = *reinterpret_cast<const uint4*>(mx);
            }
          }
        }
      }
    }
  }

  if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the row stores of the last steps
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<kAccBufs * kAccCols>(tmem_base);
  }
}

// weights OIHW fp32 -> chain blocks [nblk][BLOCK_N][8 window pixels][4] bf16 (blk_r[b] = filter row of block b,
// -1 for a zero block); row o' = (parity e, output channel) in pair mode; window position x holds filter
// column s = x - xoff - e
struct StemBlocks {
  int r[kMaxR + 2 * 2 * (kT - 1)];
};
__global__ void pack_stem_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin,
                                         int R, int S, int block_n, int pairs, int xoff, int nblk, StemBlocks blk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nblk * block_n * 32) return;
  const int b = idx / (block_n * 32), o = (idx / 32) % block_n, k = idx % 32;
  const int r = blk.r[b];
  const int x = k / 4, c = k % 4;
  const int e = pairs ? o / Cout : 0, co = pairs ? o % Cout : o;
  const int s = x - xoff - e;
  float v = 0.0f;
  if (r >= 0 && co < Cout && e < 2 && c < Cin && s >= 0 && s < S) v = w[((static_cast<size_t>(co) * Cin + c) * R + r) * S + s];
  dst[idx] = __float2bfloat16_rn(v);
}

// NCHW fp32 (C <= 4) -> [N][H][Wp][4] bf16 with zero pad columns
__global__ void import_nchw_c4_padded_kernel(const float* __restrict__ src, uint2* __restrict__ dst, int C, int H, int W,
                                             int Wp, int pad_l, size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int wp = static_cast<int>(idx % Wp);
  const size_t nh = idx / Wp;
  const int w = wp - pad_l;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
  if (w >= 0 && w < W) {
    const size_t n = nh / H, h = nh % H;
    const size_t HW = static_cast<size_t>(H) * W;
    const float* s = src + n * C * HW + h * W + w;
    v0 = __ldg(s);
    if (C > 1) v1 = __ldg(s + HW);
    if (C > 2) v2 = __ldg(s + 2 * HW);
    if (C > 3) v3 = __ldg(s + 3 * HW);
  }
  dst[idx] = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3));
}

// vectorised variant: 4 consecutive stored pixels per thread (3 x 16 B loads, 2 x 16 B stores);
// needs W % 4 == 0, pad_l % 4 == 0, Wp % 4 == 0 so that no group of 4 straddles the image edge
__global__ void import_nchw_c4_padded_x4_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int C, int H, int W,
                                                int Wp4, int pad_l, size_t total) {
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int wq = static_cast<int>(idx % Wp4);
  const size_t nh = idx / Wp4;
  const int w = wq * 4 - pad_l;
  float4 v[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (w >= 0 && w < W) {
    const size_t n = nh / H, h = nh % H;
    const size_t HW = static_cast<size_t>(H) * W;
    const float* s = src + n * C * HW + h * W + w;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < C) v[c] = __ldg(reinterpret_cast<const float4*>(s + c * HW));
  }
  uint4 o0, o1;
  o0.x = pack_bf16x2(v[0].x, v[1].x), o0.y = pack_bf16x2(v[2].x, v[3].x);
  o0.z = pack_bf16x2(v[0].y, v[1].y), o0.w = pack_bf16x2(v[2].y, v[3].y);
  o1.x = pack_bf16x2(v[0].z, v[1].z), o1.y = pack_bf16x2(v[2].z, v[3].z);
  o1.z = pack_bf16x2(v[0].w, v[1].w), o1.w = pack_bf16x2(v[2].w, v[3].w);
  dst[idx * 2] = o0;
  dst[idx * 2 + 1] = o1;
}

using EncodeTiledFnS = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnS g_encode = nullptr;

// chain layout of the stationary weights: for each row parity class rho (mod sv):
//   [zero] [B_rmax(rho), B_rmax-sv, ..., B_rho] [zero]
// returns the number of blocks; blk_r[b] = filter row of block b (-1 = zero), bblk[i] = first block of the
// stacked operand [B_i ; B_{i-sv}] of input-row position i in a step
int build_chain(int R, int sv, int* blk_r, int* bblk) {
  int nblk = 0;
  int chain_off[2] = {0, 0}, rmax[2] = {0, 0};
  for (int rho = 0; rho < sv; ++rho) {
    chain_off[rho] = nblk;
    rmax[rho] = rho > R - 1 ? -1 : (R - 1) - ((R - 1 - rho) % sv);
    for (int k = 0; k < kT - 1; ++k) blk_r[nblk++] = -1;
    for (int r = rmax[rho]; r >= 0; r -= sv) blk_r[nblk++] = r;
    for (int k = 0; k < kT - 1; ++k) blk_r[nblk++] = -1;
  }
  const int n_in = (kT - 1) * sv + R;
  for (int i = 0; i < kT + kMaxR; ++i) {
    bblk[i] = 0;
    if (i >= n_in) continue;
    const int rho = i % sv;
    bblk[i] = chain_off[rho] + (kT - 1) + (rmax[rho] - i) / sv;  // (rmax - i) is a multiple of sv, possibly negative
  }
  return nblk;
}

}  // namespace

bool stem_rowring_geometry(StemGeometry& g, int Cin, int Cout, int H, int W, int R, int S, int stride, int pad, int dil,
                           int groups) {
  memset(&g, 0, sizeof g);
  if (Cin > 4 || groups != 1 || dil != 1 || R != S || (stride != 1 && stride != 2) || pad >= R) return false;
  if (R != 3 && R != 5 && R != 7) return false;  // instantiated filter sizes
  g.pairs = stride == 1 ? 1 : 0;
  g.block_n = Cout * (g.pairs ? 2 : 1);
  if (g.block_n != 32 && g.block_n != 64) return false;
  g.P = (H + 2 * pad - R) / stride + 1;
  g.Q = (W + 2 * pad - S) / stride + 1;
  if (g.P < 1 || g.Q < 1) return false;
  if (g.pairs && (g.Q & 1)) return false;
  g.Qw = g.pairs ? g.Q / 2 : g.Q;
  g.pad_l = pad + (pad & 1);
  g.xoff = g.pad_l - pad;
  if (S + g.xoff + g.pairs > 8) return false;
  g.Wp = std::max(2 * (g.Qw - 1) + 8, W + g.pad_l);
  g.Wp = (g.Wp + 3) / 4 * 4;
  return true;
}

#define TLXCV_STEM_INSTANCES(X) \
  X(64, 7, 2) X(32, 7, 2) X(64, 5, 2) X(32, 5, 2) X(64, 3, 2) X(32, 3, 2) \
  X(64, 7, 1) X(32, 7, 1) X(64, 5, 1) X(32, 5, 1) X(64, 3, 1) X(32, 3, 1)

cudaError_t stem_rowring_set_attributes() {
  cudaError_t e;
#define TLXCV_X(BN, RR, SS)                                                                                          \
  e = cudaFuncSetAttribute(stem_rowring_kernel<BN, RR, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimitS); \
  if (e != cudaSuccess) return e;
  TLXCV_STEM_INSTANCES(TLXCV_X)
#undef TLXCV_X
  return cudaSuccess;
}

int stem_rowring_weight_elems(const StemGeometry& g, int R, int stride) {
  int blk_r[32], bblk[kT + kMaxR];
  return build_chain(R, stride, blk_r, bblk) * g.block_n * 32;
}

cudaError_t pack_stem_weights(const float* oihw, __nv_bfloat16* dst, int Cout, int Cin, int R, int S, int stride,
                              const StemGeometry& g, cudaStream_t st) {
  StemBlocks blk;
  int bblk[kT + kMaxR];
  const int nblk = build_chain(R, stride, blk.r, bblk);
  const int total = nblk * g.block_n * 32;
  pack_stem_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(oihw, dst, Cout, Cin, R, S, g.block_n, g.pairs, g.xoff, nblk, blk);
  return cudaGetLastError();
}

cudaError_t import_nchw_c4_padded(const float* src, void* dst, int N, int C, int H, int W, int Wp, int pad_l, cudaStream_t st) {
  if (W % 4 == 0 && pad_l % 4 == 0 && Wp % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const size_t total = static_cast<size_t>(N) * H * (Wp / 4);
    return launch_pdl(import_nchw_c4_padded_x4_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, st, src,
                      static_cast<uint4*>(dst), C, H, W, Wp / 4, pad_l, total);
  }
  const size_t total = static_cast<size_t>(N) * H * Wp;
  return launch_pdl(import_nchw_c4_padded_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, st, src,
                    static_cast<uint2*>(dst), C, H, W, Wp, pad_l, total);
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
  p.in = in_padded, p.Wp = g.Wp;
  if (const char* e = debug_env("TLXCV_DEBUG_ABLATE_STEM")) p.ablate = atoi(e);  // timing experiments only: results are wrong
  if (pool && (p.q_tiles != 1 || g.pairs)) return "stem: the fused max-pool needs a single column tile";
  // step / ring geometry
  p.G = kT * stride;
  p.n_in = (kT - 1) * stride + R;
  p.ng = (p.n_in + p.G - 1) / p.G;
  const int wtile = std::min(kTileM, (g.Qw + 7) / 8 * 8);  // windows per slice (whole 8-row swizzle atoms)
  p.slice_bytes = wtile * 64;
  int blk_r[32];
  p.nblk = build_chain(R, stride, blk_r, p.bblk);
#define TLXCV_X(BN, RR, SS)                                                       \
  if (g.block_n == BN && R == RR && stride == SS)                                 \
    for (int i = 0; i < p.n_in; ++i)                                              \
      if (chain_block<RR, SS>(i) != p.bblk[i]) return "stem: weight chain layout mismatch";
  TLXCV_STEM_INSTANCES(TLXCV_X)
#undef TLXCV_X
  const int blk_bytes = g.block_n * 64, rowbuf = kTileM * g.block_n * 2;
  const int fixed = p.nblk * blk_bytes + kRowSlots * rowbuf + kMiscBytes;
  p.NG = std::min(kMaxGroups, (kSmemLimitS - fixed) / (p.G * p.slice_bytes));
  if (p.NG < p.ng + 1) return "stem: shared memory cannot hold the input-row ring";
  L.smem = p.NG * p.G * p.slice_bytes + fixed;
  // band height: minimise the steps the busiest CTA walks
  const int units = pool ? Pp : g.P;  // rows a band is measured in
  long long best = -1;
  for (int bands = 1; bands <= units; ++bands) {
    const int rows = (units + bands - 1) / bands;
    if ((units + rows - 1) / rows != bands) continue;
    const long long items = static_cast<long long>(N) * bands * p.q_tiles;
    const long long per_cta = (items + sm_count - 1) / sm_count;
    const long long steps = pool ? rows + 1 : (rows + kT - 1) / kT;
    // steps + 0.4 step per band, fitted to bs256/512 timings (ResNet stem 1/2/4/8 bands = 186/190/174/184 us, MobileNet
    // stem 220/197/201/211 us): the producer runs ahead across items, a band costs little beyond its halo rows
    const long long cost = per_cta * (steps * 10 + 4);
    if (best < 0 || cost < best) best = cost, p.bands = bands, p.band_rows = rows;
  }
  if (const char* e = tuning_env("TLXCV_DEBUG_STEM_BANDS")) {  // A/B timing only
    const int bands = std::max(1, std::min(units, atoi(e)));
    p.band_rows = (units + bands - 1) / bands;
    p.bands = (units + p.band_rows - 1) / p.band_rows;
  }
  const long long items = static_cast<long long>(N) * p.bands * p.q_tiles;
  L.grid = static_cast<int>(std::min<long long>(items, sm_count));
  L.block_n = g.block_n;
  L.threads = kThreadsS;

  // A: overlapping windows. dim0 = 32 bf16 (8 pixels x 4 channels), dim1 = window q (stride 2 pixels = 16 B),
  // dim2 = input row, dim3 = image.  Out-of-range rows / windows are zero-filled by the hardware.
  {
    cuuint64_t dims[4] = {32, (cuuint64_t)g.Qw, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {16, (cuuint64_t)g.Wp * 8, (cuuint64_t)H * g.Wp * 8};
    cuuint32_t box[4] = {32, (cuuint32_t)wtile, (cuuint32_t)p.G, 1};
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
    cuuint64_t dims[2] = {32, (cuuint64_t)p.nblk * g.block_n};
    cuuint64_t strides[1] = {64};
    cuuint32_t box[2] = {32, (cuuint32_t)g.block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&L.tmapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(packed_w), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "stem: cuTensorMapEncodeTiled (weights) failed";
  }
  L.tmapOut = L.tmapB;  // unused with the fused max-pool (the pooled row is written with plain stores)
  if (!pool) {
    // conv output [N][P][Qw][block_n] (NHWC; pair mode: Qw pixel pairs of 2 x C_out channels), stored one 128-window row
    // at a time from the swizzled row buffer
    cuuint64_t dims[4] = {(cuuint64_t)g.block_n, (cuuint64_t)p.Qw, (cuuint64_t)p.P, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)g.block_n * 2, (cuuint64_t)p.Qw * g.block_n * 2, (cuuint64_t)p.P * p.Qw * g.block_n * 2};
    cuuint32_t box[4] = {(cuuint32_t)g.block_n, (cuuint32_t)kTileM, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&L.tmapOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, g.block_n == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "stem: cuTensorMapEncodeTiled (output rows) failed";
  }
  return "";
}

cudaError_t stem_rowring_launch(const StemLaunch& L, cudaStream_t st) {
#define TLXCV_X(BN, RR, SS)                                                                              \
  if (L.block_n == BN && L.p.R == RR && L.p.sv == SS) {                                                  \
    return launch_pdl(stem_rowring_kernel<BN, RR, SS>, L.grid, L.threads, L.smem, st, L.tmapA, L.tmapB, L.tmapOut, L.p); \
  }
  TLXCV_STEM_INSTANCES(TLXCV_X)
#undef TLXCV_X
  return cudaErrorInvalidValue;
}

}  // namespace tlxcv
