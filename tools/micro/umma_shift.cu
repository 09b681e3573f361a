// Micro-experiment: can a SWIZZLE_128B K-major UMMA operand start at an arbitrary 128-byte row of a
// shared-memory slab that TMA filled at a destination that is NOT 1024-byte aligned?
//   * TMA loads X[rows][64] bf16 (128 B rows, SWIZZLE_128B) to  slab + 128*d
//   * UMMA (M=128, N=64, K=64) computes D = A * I^T with A's descriptor start = slab + 128*(d + sh)
//     and base_offset field either 0 or ((start >> 7) & 7)
//   * D[m][n] must equal X[m + sh][n] for every m with m + sh < rows
// Prints the number of mismatching elements for each (d, sh, base_offset mode).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift umma_shift.cu -lcuda && ./umma_shift
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../tlxcv_b200/csrc/common.cuh"

using namespace tlxcv;

constexpr int kRows = 184;   // rows TMA brings in (box height)
constexpr int kSlab = 40960; // bytes of slab

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapI, float* out, int d, int sh,
             int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* slab = smem;                 // 40 KB
  uint8_t* ident = smem + kSlab;        // 64 x 128 B = 8 KB, 1024-aligned
  uint64_t* bar = reinterpret_cast<uint64_t*>(ident + 8192);
  uint64_t* mbar = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kSlab / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(slab)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_init(smem_u32(mbar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<64>(smem_u32(tptr));
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(bar), kRows * 128 + 8192);
    tma_load_2d(smem_u32(slab) + 128 * d, &tmapX, smem_u32(bar), 0, 0);
    tma_load_2d(smem_u32(ident), &tmapI, smem_u32(bar), 0, 0);
    mbar_wait(smem_u32(bar), 0);
    tcgen05_fence_after();
    const uint32_t a_addr = smem_u32(slab) + 128 * (d + sh);
    const uint32_t bo = mode ? ((a_addr >> 7) & 7) : 0;
    const uint64_t adesc = desc_sw128(a_addr, bo), bdesc = desc_sw128(smem_u32(ident), 0);
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
    umma_commit(smem_u32(mbar));
  }
  __syncwarp();
  mbar_wait(smem_u32(mbar), 0);
  tcgen05_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 2; ++c) {
    tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  Fn enc = reinterpret_cast<Fn>(fn);
  std::vector<__nv_bfloat16> X(kRows * 64), I(64 * 64);
  for (int r = 0; r < kRows; ++r)
    for (int c = 0; c < 64; ++c) X[r * 64 + c] = __float2bfloat16(float((r * 7 + c * 3) % 251) - 125.0f);
  for (int r = 0; r < 64; ++r)
    for (int c = 0; c < 64; ++c) I[r * 64 + c] = __float2bfloat16(r == c ? 1.0f : 0.0f);
  __nv_bfloat16 *dX, *dI;
  float* dO;
  cudaMalloc(&dX, X.size() * 2);
  cudaMalloc(&dI, I.size() * 2);
  cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dI, I.data(), I.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mX, mI;
  {
    cuuint64_t dims[2] = {64, kRows}, strides[1] = {128};
    cuuint32_t box[2] = {64, kRows}, es[2] = {1, 1};
    CUresult r = enc(&mX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode X failed %d\n", int(r)); return 1; }
  }
  {
    cuuint64_t dims[2] = {64, 64}, strides[1] = {128};
    cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
    CUresult r = enc(&mI, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dI, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode I failed %d\n", int(r)); return 1; }
  }
  const int smem = kSlab + 8192 + 64;
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * 64);
  for (int mode = 0; mode < 2; ++mode)
    for (int d : {0, 1, 3, 8})
      for (int sh : {0, 1, 2, 5, 8, 29}) {
        cudaMemset(dO, 0, O.size() * 4);
        shift_kernel<<<1, 128, smem>>>(mX, mI, dO, d, sh, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("d=%d sh=%d mode=%d: CUDA error %s\n", d, sh, mode, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, checked = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            if (m + sh >= kRows) continue;
            ++checked;
            if (O[m * 64 + n] != __bfloat162float(X[(m + sh) * 64 + n])) ++bad;
          }
        printf("tma dst +%d rows, A start +%d rows, base_offset %s: %d / %d mismatches\n", d, sh, mode ? "computed" : "0", bad, checked);
      }
  return 0;
}
