// Micro-benchmark: latency of an mbarrier hand-off between two warps of one CTA (B200, sm_100a).
// Warp 0 arrives on bar A and waits on bar B; warp 1 waits on A and arrives on B; N round trips.
// Variants of the wait: 0 = mbarrier.try_wait loop (what csrc/tc_ptx.cuh uses), 1 = mbarrier.test_wait spin,
// 2 = try_wait with a small suspendTimeHint, 3 = like 0 but only lane 0 waits (then __syncwarp), 4 = like 0 with
// `spin` other warps of the CTA polling a third barrier that never completes (the idle roles of k_gemm2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_pingpong tools/micro/mbar_pingpong.cu && ./mbar_pingpong
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <int MODE>
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  if (MODE == 1) { while (!test_wait(bar, parity)) {} }
  else if (MODE == 2) { while (!try_wait_hint(bar, parity, 32)) {} }
  else { while (!try_wait(bar, parity)) {} }
}

template <int MODE>
__global__ void k_pingpong(int n, int spin_warps, long long* out) {
  __shared__ uint64_t bars[3];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      if (lane == 0) mbar_arrive(&bars[0]);
      if (MODE == 3) { if (lane == 0) wait<0>(&bars[1], i & 1); __syncwarp(); }
      else wait<MODE>(&bars[1], i & 1);
    }
    const long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; stop = 1; mbar_arrive(&bars[2]); }
  } else if (warp == 1) {
    for (int i = 0; i < n; ++i) {
      if (MODE == 3) { if (lane == 0) wait<0>(&bars[0], i & 1); __syncwarp(); }
      else wait<MODE>(&bars[0], i & 1);
      if (lane == 0) mbar_arrive(&bars[1]);
    }
  } else if (warp < 2 + spin_warps) {
    wait<MODE == 3 ? 0 : MODE>(&bars[2], 0);  // idle roles: poll a barrier that completes only at the end
  }
}

template <int MODE>
void run(const char* name, int spin) {
  long long* d; cudaMalloc(&d, 8);
  const int n = 20000;
  k_pingpong<MODE><<<1, 32 * (2 + spin)>>>(n, spin, d);
  cudaDeviceSynchronize();
  k_pingpong<MODE><<<1, 32 * (2 + spin)>>>(n, spin, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-44s spin_warps=%d : %7.1f cycles per round trip (2 hand-offs)  [%s]\n", name, spin, (double)h / n, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int spin : {0, 9}) {
    run<0>("try_wait loop (all lanes)", spin);
    run<1>("test_wait spin (all lanes)", spin);
    run<2>("try_wait, suspendTimeHint 32 ns", spin);
    run<3>("try_wait loop, lane 0 only + __syncwarp", spin);
  }
  return 0;
}
