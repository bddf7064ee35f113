// One-shot sum over the GPUs of a node through peer-mapped memory (NVLink 5 / NVSwitch), for the
// (2C+1)- and 3C-element exchanges of MinkowskiSyncBatchNorm (train_lidog.py:228 converts all 62 BN layers,
// which makes 124 latency-bound collectives per training step).  Instead of an NCCL launch per exchange,
// one small kernel publishes this rank's vector in its own exchange buffer, raises an epoch flag
// (release, system scope), waits for every peer's flag (acquire over NVLink) and adds the peers' vectors
// in RANK ORDER in double precision -- every rank obtains bit-identical sums, deterministically.
//
// Protocol.  Buffer = kSlots slots of {uint64 flag; double data[kMaxN]}.  Call number `epoch` (1, 2, ...;
// the same sequence on every rank) uses slot epoch % kSlots.  A rank can reach epoch e+1 only after it has
// seen every peer's flag >= e, and a peer raises flag e only after its kernel of epoch e-1 (which did its
// reads) has finished, so with kSlots >= 2 no slot is overwritten while a peer may still read it.
// The spin is bounded: a missing peer traps the kernel instead of hanging the GPU.
#include "common.cuh"
#include "peer.cuh"
#include "runtime.cuh"

namespace lg {

__global__ void __launch_bounds__(1024)
    k_peer_sum(const double* __restrict__ local, int n, PeerTable tab, int world, int rank, unsigned long long epoch,
               double* __restrict__ out, int* err) {
  PeerSlot* mine = tab.buf[rank] + (epoch % kPeerSlots);
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine->data[i] = local[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(&mine->flag, epoch);
  if (threadIdx.x < world && threadIdx.x != rank) {  // one waiting thread per peer
    const unsigned long long* f = &(tab.buf[threadIdx.x] + (epoch % kPeerSlots))->flag;
    unsigned long long spins = 0;
    while (ld_acquire_sys(f) < epoch) {
      if (++spins > (1ull << 28)) {  // ~ seconds: a peer never arrived
        if (err) atomicExch(err, 100 + (int)threadIdx.x);
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double a = 0.0;
    for (int r = 0; r < world; ++r) {
      const PeerSlot* s = tab.buf[r] + (epoch % kPeerSlots);
      a += (r == rank) ? local[i] : ld_relaxed_sys_f64(&s->data[i]);
    }
    out[i] = a;
  }
}

}  // namespace lg

using namespace lg;

/* bytes of one rank's exchange buffer (zero-initialised, mapped into every peer: symmetric memory / cudaIpc). */
extern "C" size_t lg_peer_exchange_bytes(void) { return sizeof(PeerSlot) * kPeerSlots; }

/* out[0..n) = sum over ranks of local[0..n), in rank order.  peer_bufs (HOST array, `world` entries) holds, for
 * every rank r, the device address in THIS process of rank r's exchange buffer; epoch = 1, 2, 3, ... must advance
 * identically on every rank (one increment per call). */
extern "C" int lg_peer_sum(const double* local, int32_t n, void* const* peer_bufs, int32_t world, int32_t rank,
                           uint64_t epoch, double* out, void* stream_) {
  LG_CHECK_ARG(local && out && peer_bufs && n >= 1 && n <= kPeerMaxN, "lg_peer_sum: n must be in [1, %d]", kPeerMaxN);
  LG_CHECK_ARG(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world && epoch >= 1,
               "lg_peer_sum: bad world / rank / epoch");
  PeerTable tab;
  for (int r = 0; r < kPeerMaxWorld; ++r) tab.buf[r] = (PeerSlot*)(r < world ? peer_bufs[r] : nullptr);
  int sm = 0;
  int* err = nullptr;
  int rc = tc_runtime(&sm, &err);
  if (rc) return rc;
  int threads = (n + 31) / 32 * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 32 * ((world + 31) / 32)) threads = 32 * ((world + 31) / 32);
  k_peer_sum<<<1, threads, 0, (cudaStream_t)stream_>>>(local, n, tab, world, rank, (unsigned long long)epoch, out, err);
  LG_LAUNCH_OK();
  return LG_OK;
}
