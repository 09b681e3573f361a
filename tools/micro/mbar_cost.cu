// Microbenchmark: latency / issue cost of the barrier instructions in a pipeline round (one thread, one SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_cost mbar_cost.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
#define TRY(SEM)                                                                                                              \
  __device__ __forceinline__ uint32_t try_##SEM(uint32_t bar, uint32_t parity) {                                              \
    uint32_t ok;                                                                                                              \
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity." #SEM ".cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" \
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");                                                            \
    return ok;                                                                                                                \
  }
TRY(acquire)
TRY(relaxed)
__device__ __forceinline__ uint32_t test_acq(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void arrive_rel(uint32_t bar) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void arrive_rlx(uint32_t bar) { asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }

constexpr int N = 4096;
__global__ void k(long long* out) {
  __shared__ uint64_t done_bar, big_bar;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&done_bar), 1);
    mbar_init(smem_u32(&big_bar), (1u << 20) - 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory");  // phase 0 complete
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const uint32_t db = smem_u32(&done_bar), bb = smem_u32(&big_bar);
  uint32_t acc = 0;
  long long t;
  int o = 0;
#define TIME(body)                      \
  t = clock64();                        \
  for (int i = 0; i < N; ++i) { body; } \
  out[o++] = clock64() - t;
  TIME(acc += try_acquire(db, 0))                                   // 0 independent try_wait.acquire (ready)
  TIME(acc += try_relaxed(db, 0))                                   // 1 independent try_wait.relaxed
  TIME(while (!try_acquire(db, 0)) {})                              // 2 dependent (branch on the result)
  TIME(while (!try_relaxed(db, 0)) {})                              // 3
  TIME(acc += test_acq(db, 0))                                      // 4 test_wait
  TIME(arrive_rel(bb))                                              // 5 arrive.release
  TIME(arrive_rlx(bb))                                              // 6 arrive.relaxed
  TIME(commit(bb))                                                  // 7 tcgen05.commit
  TIME(while (!try_acquire(db, 0)) {} arrive_rel(bb))               // 8 wait + arrive (the producer's round)
  TIME(while (!try_relaxed(db, 0)) {} arrive_rlx(bb))               // 9
  TIME(while (!try_acquire(db, 0)) {} asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); commit(bb))  // 10 the MMA warp's round
  TIME(while (!try_acquire(db, 0)) {} acc += try_acquire(db, 0); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); commit(bb))  // 11 + test of the next stage
  TIME(asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"))  // 12
  TIME(while (!try_acquire(db, 0)) {} expect_tx(bb, 0))             // 13
  out[o++] = acc;
}

int main() {
  long long* d;
  cudaMalloc(&d, 32 * sizeof(long long));
  k<<<1, 64>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[32];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  const char* names[] = {"try_wait.acquire x1 (independent)", "try_wait.relaxed (independent)", "try_wait.acquire loop (dependent)", "try_wait.relaxed loop",
                         "test_wait", "arrive.release", "arrive.relaxed", "tcgen05.commit", "wait+arrive (acq/rel)", "wait+arrive (relaxed)",
                         "wait+fence+commit", "wait+test next+fence+commit", "tcgen05.fence::after", "wait+expect_tx"};
  for (int i = 0; i < 14; ++i) printf("%-40s %.1f clk\n", names[i], double(h[i]) / N);
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
