// Microbenchmark: what bounds the L2 -> SM operand feed of the conv kernels?  148 persistent CTAs (or 74 clusters of 2), one
// producer thread each, stream 16 KB TMA boxes ([128 rows][64 bf16], SWIZZLE_128B) through a 6-slot ring and drop them.
//   mode 0: every CTA reads its OWN L2-resident region (activations: unique data per SM)
//   mode 1: every CTA reads the SAME region (weights: identical data for all SMs)
//   mode 2: half own, half shared (a conv K block: A tile + weight tile)
//   mode 3: shared region, cluster of 2, each CTA loads HALF of every box and multicasts it to both (weights via multicast)
//   mode 4: half own (unicast) + half shared by multicast
// Reports GB/s landed in shared memory, summed over the chip.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_feed l2_feed.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (++spins > 100000000u) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(addr), "r"(rank)); return o; }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

constexpr int kSlots = 6, kBox = 16384;

template <int MODE>
__global__ void __launch_bounds__(64, 1) feed_kernel(const __grid_constant__ CUtensorMap own, const __grid_constant__ CUtensorMap shared_full,
                                                      const __grid_constant__ CUtensorMap shared_half, int iters, int own_rows, int shared_rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSlots * kBox);
  uint64_t* empty = full + kSlots;
  constexpr bool MC = MODE >= 3;
  const uint32_t rank = MC ? ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), MC ? 2 : 1);   // multicast: both CTAs of the pair must have released the slot
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (MC) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  const int own_base = blockIdx.x * own_rows;
  if (threadIdx.x == 0) {          // producer
    uint32_t slot = 0, phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(smem_u32(&empty[slot]), phase ^ 1);
      const uint32_t bar = smem_u32(&full[slot]), dst = smem_u32(smem + slot * kBox);
      mbar_expect(bar, kBox);
      const bool use_shared = MODE == 1 || MODE == 3 || ((MODE == 2 || MODE == 4) && (i & 1));
      if (!use_shared) {
        tma_load_2d(dst, &own, bar, 0, own_base + (i * 128) % own_rows);
      } else if (!MC) {
        tma_load_2d(dst, &shared_full, bar, 0, (i * 128) % shared_rows);
      } else {
        // this CTA fetches rows [64 * rank, 64 * rank + 64) of the box and multicasts them to both CTAs of the pair
        tma_load_2d_mc(dst + rank * (kBox / 2), &shared_half, bar, 0, (i * 128) % shared_rows + rank * 64, 3);
      }
      if (++slot == kSlots) slot = 0, phase ^= 1;
    }
  } else if (threadIdx.x == 32) {  // consumer: drop the data, hand the slot back
    uint32_t slot = 0, phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(smem_u32(&full[slot]), phase);
      const uint32_t e = smem_u32(&empty[slot]);
      if (MC) {
        mbar_arrive_cluster(mapa(e, 0));
        mbar_arrive_cluster(mapa(e, 1));
      } else {
        mbar_arrive_cluster(mapa(e, 0) * 0 + e);
      }
      if (++slot == kSlots) slot = 0, phase ^= 1;
    }
  }
  if (MC) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void* base, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int MODE>
static void run(EncodeFn enc, void* own_buf, void* sh_buf, int own_rows, int shared_rows, int iters, int grid) {
  CUtensorMap own = make_map(enc, own_buf, (uint64_t)own_rows * grid, 128), shf = make_map(enc, sh_buf, shared_rows, 128),
              shh = make_map(enc, sh_buf, shared_rows, 64);
  const int smem = kSlots * kBox + 1024;
  cudaFuncSetAttribute(feed_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(64), cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MODE >= 3 ? 2 : 1, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, feed_kernel<MODE>, own, shf, shh, iters, own_rows, shared_rows);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("mode %d failed: %s\n", MODE, cudaGetErrorString(err)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double bytes = (double)grid * iters * kBox;
  printf("mode %d: own region %4d KB/CTA, shared region %5d KB: %8.3f ms  %8.1f GB/s landed in smem (%.1f B/clk/SM at 1.9 GHz)\n", MODE,
         own_rows * 128 / 1024, shared_rows * 128 / 1024, best, bytes / best / 1e6, bytes / best / 1e6 / grid / 1.9);
}

int main(int argc, char** argv) {
  cudaFree(0);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  const int grid = 148, iters = 4000;
  for (int own_kb : {128, 512}) {
    for (int sh_kb : {128, 1024}) {
      const int own_rows = own_kb * 1024 / 128, shared_rows = sh_kb * 1024 / 128;
      void *own_buf, *sh_buf;
      cudaMalloc(&own_buf, (size_t)own_rows * 128 * grid);
      cudaMalloc(&sh_buf, (size_t)shared_rows * 128);
      cudaMemset(own_buf, 1, (size_t)own_rows * 128 * grid);
      cudaMemset(sh_buf, 1, (size_t)shared_rows * 128);
      run<0>(enc, own_buf, sh_buf, own_rows, shared_rows, iters, grid);
      run<1>(enc, own_buf, sh_buf, own_rows, shared_rows, iters, grid);
      run<2>(enc, own_buf, sh_buf, own_rows, shared_rows, iters, grid);
      run<3>(enc, own_buf, sh_buf, own_rows, shared_rows, iters, grid);
      run<4>(enc, own_buf, sh_buf, own_rows, shared_rows, iters, grid);
      cudaFree(own_buf), cudaFree(sh_buf);
    }
  }
  return 0;
}
