// Exchange-buffer layout and memory-order helpers of the peer-memory SyncBN exchange (see peer.cu).
#pragma once

#include "common.cuh"

namespace lg {

constexpr int kPeerSlots = 4;
constexpr int kPeerMaxN = 3 * 1024 + 8;
constexpr int kPeerMaxWorld = 16;

struct PeerSlot {
  unsigned long long flag;
  unsigned long long pad;
  double data[kPeerMaxN];
};

struct PeerTable {
  PeerSlot* buf[kPeerMaxWorld];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}


}  // namespace lg
