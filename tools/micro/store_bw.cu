// Microbenchmark: how fast can 8 epilogue-style warps per SM write a [M][C] bf16 matrix?
//   mode 0: TMA store, 32 rows x 32 cols (64 B rows, SWIZZLE_64B)   <- conv epilogue v1/v2
//   mode 1: TMA store, 32 rows x 64 cols (128 B rows, SWIZZLE_128B)
//   mode 2: st.global.v4 from smem, 64 B row segments (8 rows per instruction)
//   mode 3: st.global.v4 from smem, 128 B row segments (4 rows per instruction)
//   mode 4: TMA store, 128 rows x 64 cols issued by one thread per CTA (16 KB boxes)
//   mode 5: st.global.v4 straight from registers, every lane its own row's 64 B (4 instructions per 32 x 32 item, rows C*2 bytes apart)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m), "r"(src), "r"(c0), "r"(c1) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) store_kernel(const __grid_constant__ CUtensorMap tmap, uint16_t* out, int M, int C, int m_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lg = warp & 3, half = warp >> 2;
  constexpr int W = (MODE == 0 || MODE == 2 || MODE == 5) ? 32 : 64;       // columns per item
  constexpr int ROWB = W * 2;
  uint8_t* stg = smem + warp * 8192;                          // 2 buffers x 4 KB
  uint32_t nstore = 0;
  const int items_per_tile = (C / 2) / W;                     // this warp's column half
  for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
    const int m0 = tile * 128 + lg * 32;
    if (MODE == 4) {
      // whole CTA: 128 rows x 64 cols boxes by warp 0 lane 0, others just fill smem
      for (int cb = 0; cb < C; cb += 64) {
        uint8_t* buf = smem + (nstore & 1) * 16384;
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        for (int i = threadIdx.x; i < 1024; i += 256) reinterpret_cast<uint4*>(buf)[i] = make_uint4(tile, cb, i, 7);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
          tma_store_2d(&tmap, smem_u32(buf), cb, tile * 128);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++nstore;
      }
      continue;
    }
    for (int it = 0; it < items_per_tile; ++it) {
      const int cbase = half * (C / 2) + it * W;
      uint8_t* buf = stg + (nstore & 1) * 4096;
      if (MODE == 5) {
        if (m0 + lane < M) {
          uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(m0 + lane) * C + cbase);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(tile, it, lane, j);
        }
        ++nstore;
        continue;
      }
      if (MODE <= 1) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
      }
      // each lane writes its row (ROWB bytes), swizzled like the conv epilogue
#pragma unroll
      for (int j = 0; j < ROWB / 16; ++j) {
        const int swz = (ROWB == 64) ? ((lane >> 1) & 3) : (lane & 7);
        *reinterpret_cast<uint4*>(buf + lane * ROWB + ((j ^ swz) << 4)) = make_uint4(tile, it, lane, j);
      }
      if (MODE <= 1) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap, smem_u32(buf), cbase, m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        __syncwarp();
        constexpr int LPR = ROWB / 16;          // lanes per row
        constexpr int RPI = 32 / LPR;           // rows per instruction
#pragma unroll
        for (int i = 0; i < 32 / RPI; ++i) {
          const int r = lane / LPR + RPI * i, q = lane % LPR;
          const int swz = (ROWB == 64) ? ((r >> 1) & 3) : (r & 7);
          const uint4 v = *reinterpret_cast<const uint4*>(buf + r * ROWB + ((q ^ swz) << 4));
          if (m0 + r < M) *reinterpret_cast<uint4*>(out + (size_t)(m0 + r) * C + cbase + q * 8) = v;
        }
        __syncwarp();
      }
      ++nstore;
    }
  }
  if (MODE <= 1 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (MODE == 4 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 802816, C = argc > 2 ? atoi(argv[2]) : 256;
  uint16_t* out;
  cudaMalloc(&out, (size_t)M * C * 2);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  auto make = [&](int bw, int bh, CUtensorMapSwizzle sw) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M}, strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    return m;
  };
  const int m_tiles = (M + 127) / 128;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const char* names[] = {"tma 32x32 (64B rows)", "tma 32x64 (128B rows)", "stg 64B segments", "stg 128B segments", "tma 128x64 per CTA",
                         "stg per-lane rows (regs)"};
  for (int mode = 0; mode < 6; ++mode) {
    CUtensorMap tm = mode == 0 ? make(32, 32, CU_TENSOR_MAP_SWIZZLE_64B) : (mode == 4 ? make(64, 128, CU_TENSOR_MAP_SWIZZLE_128B) : make(64, 32, CU_TENSOR_MAP_SWIZZLE_128B));
    auto launch = [&]() {
      switch (mode) {
        case 0: store_kernel<0><<<148, 256, 65536>>>(tm, out, M, C, m_tiles); break;
        case 1: store_kernel<1><<<148, 256, 65536>>>(tm, out, M, C, m_tiles); break;
        case 2: store_kernel<2><<<148, 256, 65536>>>(tm, out, M, C, m_tiles); break;
        case 3: store_kernel<3><<<148, 256, 65536>>>(tm, out, M, C, m_tiles); break;
        case 4: store_kernel<4><<<148, 256, 65536>>>(tm, out, M, C, m_tiles); break;
        case 5: store_kernel<5><<<148, 256, 65536>>>(tm, out, M, C, m_tiles); break;
      }
    };
    cudaFuncSetAttribute(store_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(store_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(store_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(store_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(store_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(store_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("M=%d C=%d mode %d %-26s %8.1f us  %7.0f GB/s  (%s)\n", M, C, mode, names[mode], ms / 10 * 1e3,
           (double)M * C * 2 / (ms / 10 * 1e-3) / 1e9, cudaGetErrorString(err));
  }
  return 0;
}
