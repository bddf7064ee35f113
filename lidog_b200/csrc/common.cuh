// Shared device/host helpers for liblidog_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lidog_b200.h"

namespace lg {

void set_error(const char* fmt, ...);

#define LG_CHECK_ARG(cond, ...)   \
  do {                            \
    if (!(cond)) {                \
      lg::set_error(__VA_ARGS__); \
      return LG_ERR_INVALID;      \
    }                             \
  } while (0)

#define LG_CUDA_OK(expr)                                                                    \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      lg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return LG_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

#define LG_LAUNCH_OK() LG_CUDA_OK(cudaPeekAtLastError())

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ------------------------------------------------------------------ voxel keys + hash table

constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kCoordBias = 1 << 15;

struct __align__(16) HashSlot {
  unsigned long long key;
  unsigned int val;  // row id; 0xFFFFFFFF while unset (table is memset to 0xFF)
  unsigned int pad;
};

__device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
  return (unsigned)b < 32768u && (unsigned)(x + kCoordBias) < 65536u && (unsigned)(y + kCoordBias) < 65536u &&
         (unsigned)(z + kCoordBias) < 65536u;
}

__device__ __forceinline__ unsigned long long pack_key(int b, int x, int y, int z) {
  return ((unsigned long long)(unsigned)b << 48) | ((unsigned long long)(unsigned)(x + kCoordBias) << 32) |
         ((unsigned long long)(unsigned)(y + kCoordBias) << 16) | (unsigned long long)(unsigned)(z + kCoordBias);
}

// 32-bit multiplicative mix of the packed key (3 IMUL + 4 shifts/xors).  The 64-bit murmur finaliser used at first
// cost ~40 instructions per probe (two 64-bit multiplies), a third of k_neighbors' issue-bound time; table contents
// and every result are independent of the hash function (first occurrence wins by atomicMin on the row).
__device__ __forceinline__ unsigned long long hash_key(unsigned long long k) {
  unsigned int h = (unsigned int)k * 0x9E3779B1u ^ (unsigned int)(k >> 32) * 0x85EBCA77u;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return (unsigned long long)h;
}

// Returns the row id stored for `key`, or -1.
__device__ __forceinline__ int hash_lookup(const HashSlot* __restrict__ table, unsigned long long mask,
                                           unsigned long long key) {
  unsigned long long s = hash_key(key) & mask;
#pragma unroll 1
  for (;;) {
    const ulonglong2 raw = __ldg(reinterpret_cast<const ulonglong2*>(table + s));
    if (raw.x == key) return (int)(unsigned)(raw.y & 0xFFFFFFFFull);
    if (raw.x == kEmptyKey) return -1;
    s = (s + 1) & mask;
  }
}

__device__ __forceinline__ int floor_div(int a, int s) {
  int q = a / s;
  return (a % s != 0 && ((a < 0) != (s < 0))) ? q - 1 : q;
}

}  // namespace lg
