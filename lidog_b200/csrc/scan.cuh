// Device-wide exclusive scan over 0/1 flags produced by a functor, with a sink functor that
// receives (index, exclusive_prefix, flag).  Three launches: per-block sums, scan of the sums,
// apply.  Deterministic; used for first-occurrence compaction, pair lists and k-sorted plans.
#pragma once

#include "common.cuh"

namespace lg {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanBlock = kScanThreads * kScanItems;  // 2048 items per block

__device__ __forceinline__ int warp_inclusive_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the block; returns (exclusive, block_total)
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = warp_inclusive_scan(v);
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < (blockDim.x >> 5)) ? warp_sums[lane] : 0;
    int winc = warp_inclusive_scan(w);
    warp_sums[lane] = winc - w;
  }
  __syncthreads();
  int excl = inc - v + warp_sums[warp];
  // total = exclusive of the last thread + its value
  __shared__ int s_total;
  if (threadIdx.x == blockDim.x - 1) s_total = excl + v;
  __syncthreads();
  *total = s_total;
  return excl;
}

template <class Flag>
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(Flag flag, int64_t n, int* __restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
  int c = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    if (i < n) c += flag(i) ? 1 : 0;
  }
  int total;
  block_exclusive_scan(c, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums[0..nb) in place, total -> total_out (int64)
static __global__ void __launch_bounds__(1024) k_scan_partials(int* __restrict__ block_sums, int nb,
                                                         int64_t* __restrict__ total_out) {
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < nb) ? block_sums[i] : 0;
    int total;
    int excl = block_exclusive_scan(v, &total);
    int c = carry;
    if (i < nb) block_sums[i] = excl + c;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <class Flag, class Sink>
__global__ void __launch_bounds__(kScanThreads)
    k_scan_apply(Flag flag, Sink sink, int64_t n, const int* __restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
  bool f[kScanItems];
  int c = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    f[j] = (i < n) ? flag(i) : false;
    c += f[j] ? 1 : 0;
  }
  int total;
  int excl = block_exclusive_scan(c, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    if (i < n) sink(i, excl, f[j]);
    excl += f[j] ? 1 : 0;
  }
}

static inline size_t scan_workspace_bytes(int64_t n) { return (size_t)(ceil_div(n, kScanBlock) + 1) * sizeof(int); }

// Runs the three launches.  workspace >= scan_workspace_bytes(n).
template <class Flag, class Sink>
static inline int device_scan(Flag flag, Sink sink, int64_t n, int64_t* total_out, void* workspace,
                              cudaStream_t stream) {
  if (n <= 0) {
    if (total_out) LG_CUDA_OK(cudaMemsetAsync(total_out, 0, sizeof(int64_t), stream));
    return LG_OK;
  }
  if (n >= (int64_t)1 << 31) {
    set_error("scan of %lld items exceeds int32 range", (long long)n);
    return LG_ERR_INVALID;
  }
  int nb = (int)ceil_div(n, kScanBlock);
  int* block_sums = reinterpret_cast<int*>(workspace);
  k_scan_reduce<Flag><<<nb, kScanThreads, 0, stream>>>(flag, n, block_sums);
  LG_LAUNCH_OK();
  k_scan_partials<<<1, 1024, 0, stream>>>(block_sums, nb, total_out);
  LG_LAUNCH_OK();
  k_scan_apply<Flag, Sink><<<nb, kScanThreads, 0, stream>>>(flag, sink, n, block_sums);
  LG_LAUNCH_OK();
  return LG_OK;
}

}  // namespace lg
