// Shared device helpers for the sm_100a kernels: PTX wrappers for mbarrier, TMA, tcgen05/TMEM,
// 128-bit vector access and activation functions.  Hand-written; no CUTLASS/CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/tlxcv_b200.h"
#include "env.h"

namespace tlxcv {

// ------------------------------------------------------------------------------------------------
// activations (epilogue math is fp32)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  switch (act) {
    case TLXCV_ACT_RELU: return fmaxf(v, 0.0f);
    case TLXCV_ACT_RELU6: return fminf(fmaxf(v, 0.0f), 6.0f);
    case TLXCV_ACT_LEAKY: return v > 0.0f ? v : v * alpha;
    default: return v;
  }
}

// activation over a register tile; the switch is outside the element loop
template <int N>
__device__ __forceinline__ void act_regs(float (&f)[N], int act, float alpha) {
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

// ------------------------------------------------------------------------------------------------
// 8-element activation vectors: bf16 -> one 16 B access, fp32 -> two
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Vec8;

template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// programmatic dependent launch: every forward kernel is launched with the stream-serialisation
// attribute, so its CTAs may start (barrier init, TMEM alloc, descriptor prefetch) while the previous
// kernel of the stream is still draining; pdl_wait() blocks until that kernel has completed and its
// memory is visible, and must precede the first access to any activation buffer.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = tuning_env("TLXCV_NO_PDL") == nullptr;
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------
// shared-memory addresses, mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline traps (surfacing as a CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 50000000u) __trap();
  }
}

// 8-byte asynchronous global -> shared copy; `valid == false` writes zeros and reads nothing
__device__ __forceinline__ void cp_async_8_zfill(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 8u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// smem -> global tile store (bulk async group); out-of-bounds rows / columns are clipped by the hardware
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// im2col-mode load of a (pixels x channels) tile of an NHWC tensor: coordinates are the base input
// pixel {c, w, h, n} of the filter window, offsets {s*dil, r*dil} select the tap.
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h,
                                                   int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte rows, SWIZZLE_128B (what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B):
// 8-row atoms of 1024 B; SBO = 1024 B; LBO unused for swizzled K-major; descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (ignored), bits [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // version = 1, bits [46,48)
  d |= static_cast<uint64_t>(2) << 61;                      // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N tile
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// packed fp32x2 / bf16x2 helpers of the epilogues (sm_100: fma/add/mul.f32x2 work on 64-bit register pairs)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_u64(uint32_t lo, uint32_t hi) {
  return static_cast<unsigned long long>(lo) | (static_cast<unsigned long long>(hi) << 32);
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long bf16x2_to_f32x2(uint32_t u) {  // low half -> first float
  return pack_u64(u << 16, u & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_pair_bf16(unsigned long long p) {
  return pack_bf16x2(__uint_as_float(static_cast<uint32_t>(p)), __uint_as_float(static_cast<uint32_t>(p >> 32)));
}
__device__ __forceinline__ ulonglong2 ldg_u64x2(const ulonglong2* p) {
  ulonglong2 r;
  asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}
// activation of an fp32 pair.  LeakyReLU is max(v, alpha * v): exact for 0 <= alpha <= 1, which the planner enforces.
template <int ACT>
__device__ __forceinline__ unsigned long long act_pair_f32(unsigned long long p, float alpha) {
  if (ACT == TLXCV_ACT_NONE) return p;
  float lo = __uint_as_float(static_cast<uint32_t>(p)), hi = __uint_as_float(static_cast<uint32_t>(p >> 32));
  if (ACT == TLXCV_ACT_RELU) {
    lo = fmaxf(lo, 0.0f), hi = fmaxf(hi, 0.0f);
  } else if (ACT == TLXCV_ACT_RELU6) {
    lo = fminf(fmaxf(lo, 0.0f), 6.0f), hi = fminf(fmaxf(hi, 0.0f), 6.0f);
  } else if (ACT == TLXCV_ACT_LEAKY) {
    const uint32_t al = __float_as_uint(alpha);
    const unsigned long long m = fmul2(p, pack_u64(al, al));
    lo = fmaxf(lo, __uint_as_float(static_cast<uint32_t>(m))), hi = fmaxf(hi, __uint_as_float(static_cast<uint32_t>(m >> 32)));
  }
  return pack_u64(__float_as_uint(lo), __float_as_uint(hi));
}
// fp32 pair -> packed bf16 pair with ReLU / ReLU6 applied after the rounding: cvt.rn.relu clamps in the conversion
// itself (one instruction per pair), ReLU6 adds a packed min
template <int ACT>
__device__ __forceinline__ uint32_t pack_pair_bf16_act(unsigned long long p) {
  if (ACT != TLXCV_ACT_RELU && ACT != TLXCV_ACT_RELU6) return pack_pair_bf16(p);
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;"
      : "=r"(r)
      : "f"(__uint_as_float(static_cast<uint32_t>(p >> 32))), "f"(__uint_as_float(static_cast<uint32_t>(p))));
  if (ACT == TLXCV_ACT_RELU6) asm("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(0x40C040C0u));  // 6.0 | 6.0
  return r;
}


}  // namespace tlxcv
