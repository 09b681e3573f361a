// Micro-experiment: can a NO-SWIZZLE K-major UMMA A operand have OVERLAPPING rows?
//   smem holds one raw row X[e] of bf16 elements.  A[m][k] := X[8 m + k]  (row pitch 16 B = 8 elements, K = 32 elements):
//   the 7x7/stride-2 stem's window operand (8 pixels x 4 channels per window, windows 2 pixels apart) read in place from
//   the padded input row instead of from a 4x expanded copy made by an overlapping-window TMA map.
//   Canonical no-swizzle K-major layout: core matrix = 8 rows x 16 B, rows 16 B apart; LBO = byte distance between core
//   matrices adjacent in K, SBO = between core matrices adjacent in M.  Here LBO = 16 B (!), SBO = 128 B.
//   D = A * I^T with B = 32x32 identity (ordinary no-swizzle layout) must give D[m][n] = X[8 m + n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_overlap umma_overlap.cu && ./umma_overlap
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#include "../../tlxcv_b200/csrc/common.cuh"

using namespace tlxcv;

__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // version 1; layout type 0 = no swizzle
  return d;
}

constexpr int kX = 8 * 127 + 32;  // elements of the raw row that the 128 windows touch

__global__ void __launch_bounds__(128, 1) overlap_kernel(const __nv_bfloat16* x, float* out, int swap_lbo_sbo, int swap_b) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);                // raw row, 2 KB + 64 B
  __nv_bfloat16* bs = reinterpret_cast<__nv_bfloat16*>(smem + 4096);         // identity, no-swizzle K-major: 2 KB
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 8192);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(mbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) xs[i] = i < kX ? x[i] : __float2bfloat16(0.f);
  // B[n][k] (32 x 32): core matrix (n / 8, k / 8) at ((n / 8) * 4 + k / 8) * 128 B, row n % 8 at +16 B each
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int n = i / 32, k = i % 32;
    bs[((n / 8) * 4 + k / 8) * 64 + (n % 8) * 8 + (k % 8)] = __float2bfloat16(n == k ? 1.0f : 0.0f);
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(mbar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<32>(smem_u32(tptr));
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 32);
    for (int k = 0; k < 2; ++k) {  // two K = 16 steps: +32 B along the row
      const uint64_t adesc = swap_lbo_sbo ? desc_nosw(smem_u32(xs) + 32 * k, 128, 16) : desc_nosw(smem_u32(xs) + 32 * k, 16, 128);
      const uint64_t bdesc = swap_b ? desc_nosw(smem_u32(bs) + 256 * k, 512, 128) : desc_nosw(smem_u32(bs) + 256 * k, 128, 512);  // next K pair of core matrices: +2 x 128 B
      umma_bf16(tmem, adesc, bdesc, idesc, k != 0);
    }
    umma_commit(smem_u32(mbar));
  }
  __syncwarp();
  mbar_wait(smem_u32(mbar), 0);
  tcgen05_fence_after();
  uint32_t v[32];
  tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem);
}

int main() {
  std::vector<__nv_bfloat16> X(kX);
  for (int i = 0; i < kX; ++i) X[i] = __float2bfloat16(float((i * 5) % 241) - 120.0f);
  __nv_bfloat16* dX;
  float* dO;
  cudaMalloc(&dX, X.size() * 2);
  cudaMalloc(&dO, 128 * 32 * 4);
  cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 64);
  std::vector<float> O(128 * 32);
  for (int bmode = 0; bmode < 2; ++bmode)
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dO, 0, O.size() * 4);
    overlap_kernel<<<1, 128, 8192 + 64>>>(dX, dO, mode, bmode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n)
        if (O[m * 32 + n] != __bfloat162float(X[8 * m + n])) ++bad;
    printf("B %s | A rows 16 B apart (overlapping), %s: %d / %d mismatches; D[1][0..3] = %g %g %g %g (expect %g %g %g %g)\n",
           bmode ? "LBO=512 SBO=128" : "LBO=128 SBO=512", mode ? "LBO=128 SBO=16" : "LBO=16 SBO=128", bad, 128 * 32, O[32], O[33], O[34], O[35], __bfloat162float(X[8]),
           __bfloat162float(X[9]), __bfloat162float(X[10]), __bfloat162float(X[11]));
  }
  return 0;
}
