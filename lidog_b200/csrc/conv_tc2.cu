// Sparse convolution on tcgen05, second-generation pipeline (gather_mode 2, the default).
//
// Measured on B200 (profiles/r01_conv_bench_v1.txt): tile::gather4 TMA moves 256 B per instruction and
// sustains ~100 SM cycles per instruction, 2.7x slower than LDGSTS (cp.async) row gathers; and a
// producer that waits for its own copies keeps only two stages in flight.  This pipeline therefore
//   * gathers operand rows with cp.async (16 B per lane, 4 lanes per 64-byte row segment, so every
//     request is a full 32 B sector pair) and lets the copies themselves arrive on the stage's
//     mbarrier (cp.async.mbarrier.arrive.noinc): producers never wait for data, the whole ring is in
//     flight;
//   * gives each of the 4 producer warps whole stages (unit u -> warp u % 4), so the index fetch of
//     one stage overlaps the copies of three others;
//   * forward / dgrad (k_gemm2): a CTA owns a SUPER-TILE of T row tiles whose T accumulators live in
//     TMEM together (T * N <= 512 columns); the loop runs kernel offset outermost, so one TMA load of
//     the weight panel W[k] feeds up to T gathered tiles (weight traffic / T);
//   * wgrad (k_wgrad2): a CTA owns a group of G kernel offsets (G * Cout <= 512 TMEM columns) and a
//     chunk of tiles; the dY tile is staged once and reused by every offset of the group.
// Shared-memory operand tiles keep the SWIZZLE_64B layout of conv_tc.cu (64-byte rows = 32 channels).
// Results are deterministic: every output row / weight element is written once in a fixed order.
// Reference contract: MinkowskiConvolution(+Transpose) forward/backward, utils/models/minkunet_bev.py:57-123.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "runtime.cuh"
#include "tc_ptx.cuh"

namespace lg {

int launch_reduce_partials(const float* partial, int n_chunks, int64_t n_elems, const float* out_scale, float* dW,
                           cudaStream_t stream);

namespace v2 {
using namespace ptx;

// LIDOG_DBG & 8: CTA 0 accumulates the cycles each role spends waiting (read back by lg_debug_profile)
__device__ long long g_prof[16];
// LIDOG_DBG & 8: event trace of CTA 0, [event][unit] = clock64 for the first kTraceUnits units of the launch:
// 0 producer got the free stage, 1 producer armed the stage barrier, 2 MMA warp saw the stage full, 3 MMA warp
// committed the unit, 4 id copy issued (read back by lg_debug_trace)
constexpr int kTraceUnits = 512;
__device__ long long g_trace[5][kTraceUnits];
// The instrumentation exists only in the DBG instantiation of the kernel (LIDOG_DBG != 0 selects it): in the
// production instantiation every macro below compiles to nothing.  (The event trace showed the one-thread MMA role
// executing ~500 instructions per (offset, panel) step at ~5 cycles each -- it, not the data movement or the tensor
// pipe, set the pace -- and a third of those instructions were predicated-off instrumentation and switches.)
#define TRACE(ev, unit) do { if (DBG && prof && (unit) < kTraceUnits) g_trace[ev][unit] = clock64(); } while (0)
#define PROF_T0() long long _t0 = DBG ? clock64() : 0
#define PROF_ADD(slot) do { if (DBG && prof) { long long _t1 = clock64(); pacc[slot] += _t1 - _t0; _t0 = _t1; } } while (0)
#define PROF_FLUSH(lo, hi) do { if (DBG && prof) for (int _i = lo; _i <= hi; ++_i) g_prof[_i] += pacc[_i]; } while (0)

constexpr int kRowB = 64;            // bytes per smem operand row (32 x 16-bit)
constexpr int kSub = 128 * kRowB;    // one [128 rows x 32 channels] sub-tile = 8 KB
constexpr int kEpiWarps = 4;         // warps 0..3
constexpr int kMmaWarp = 4;
constexpr int kBWarp = 5;            // weight-panel TMA producer (forward) / unused (wgrad)
constexpr int kProdWarp0 = 6;
constexpr int kProdWarps = 4;        // warps 6..9
constexpr int kIdxWarp = kProdWarp0 + kProdWarps;  // warp 10: row-id ring (bulk copies of 512 B id rows)
constexpr int kThreads = (kIdxWarp + 1) * 32;
constexpr int kMma2Warp = kIdxWarp + 1;              // k_gemm2 only: second MMA issuer (SPLIT != 0)
constexpr int kThreadsG2 = (kMma2Warp + 1) * 32;
constexpr int kIdxSlotBytes = LG_TILE_ROWS * 4;
constexpr int kSchedSlots = 4;                                   // super-tile hand-out ring
constexpr int kSchedWords = 12;  // ring entry: [0] super-tile, [1] "tile has any offset" bits, [2..9] mask word 0 of its tiles
constexpr int kSchedConsumers = kEpiWarps + 2 + kProdWarps;      // warps that read every ring entry (+1 with two issuers)
constexpr int kStgPitch = 36;                                    // floats per staged row (32 + 4: conflict-free v4)
constexpr int kStgBytes = kEpiWarps * 32 * kStgPitch * 4;        // epilogue transposition buffers

__device__ __forceinline__ uint32_t sw64(uint32_t r, uint32_t j) { return r * kRowB + ((j ^ ((r >> 1) & 3u)) << 4); }

// 1D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// L2 prefetch of `bytes` (multiple of 16) at a 16-byte aligned global address: fire and forget
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ int lds32(uint32_t saddr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}

__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Dynamic super-tile schedule.  The id warp draws super-tiles from a global counter (largest index first:
// with mask-sorted plans the expensive tiles sit at the end, so the tail of the launch is made of cheap
// ones) and publishes them through a 4-entry shared-memory ring; every other warp reads each entry once.
struct SchedCursor {
  int slot;
  uint32_t phase;
};
// tile-mask words of the (<= 8) tiles of a super-tile, kept in registers; word 0 travels with the ring entry
// (every layer on the tensor-core path has K <= 27 offsets = one word), further words are reloaded every 32 offsets
struct TileMasks {
  uint32_t w[8];
};
struct SchedItem {
  int64_t st;    // super-tile index, < 0 = no more work
  uint32_t any;  // bit t = tile t has at least one offset
  TileMasks tm;  // mask word 0 of the tiles
};
// All helpers below are called by fully active warps and broadcast lane 0's value, so everything derived from
// them (unit sequence, stage counters, descriptors) is warp-uniform for ptxas.
__device__ __forceinline__ void sched_next(SchedItem& it, SchedCursor& c, uint64_t* fullS, uint64_t* emptyS,
                                           const int* ring, int lane, int* err) {
  mbar_wait(&fullS[c.slot], c.phase, err, 20);
  const volatile int* e = ring + c.slot * kSchedWords;
  it.st = (int)bcast0((uint32_t)e[0]);
  it.any = bcast0((uint32_t)e[1]);
#pragma unroll
  for (int t = 0; t < 8; ++t) it.tm.w[t] = bcast0((uint32_t)e[2 + t]);
  __syncwarp();
  if (lane == 0) mbar_arrive(&emptyS[c.slot]);
  if (++c.slot == kSchedSlots) {
    c.slot = 0;
    c.phase ^= 1;
  }
}

__device__ __forceinline__ void load_masks(TileMasks& tm, const lgConvPlan& p, int64_t tile0, int nt, int word) {
#pragma unroll
  for (int t = 0; t < 8; ++t)
    tm.w[t] = (t < nt) ? bcast0(__ldg(p.tile_mask + (tile0 + t) * p.mask_words + word)) : 0u;
}
__device__ __forceinline__ uint32_t present_bits(const TileMasks& tm, int k) {
  uint32_t m = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) m |= ((tm.w[t] >> (k & 31)) & 1u) << t;
  return m;
}
// offsets present in ANY tile of the super-tile, for mask word `word` (word 0 came with the ring entry)
__device__ __forceinline__ uint32_t offset_bits(TileMasks& tm, const lgConvPlan& p, int64_t tile0, int nt, int word) {
  if (word) load_masks(tm, p, tile0, nt, word);
  return ((tm.w[0] | tm.w[1]) | (tm.w[2] | tm.w[3])) | ((tm.w[4] | tm.w[5]) | (tm.w[6] | tm.w[7]));
}
// Every role walks the same (offset k, tile bits m) sequence: set bits only -- the one-thread roles used to spend
// ~100 cycles per ABSENT offset on the shift/or chains of a plain k = 0..K-1 scan.
// Lane j of the warp computes the tile bits of offset 32 w + j once per mask word (the 32 shift/or instructions of
// present_bits run in parallel across the lanes); inside the walk an offset's bits are one shuffle away.
#define LG_FOR_EACH_OFFSET(tm, plan, tile0, nt, k, m)                                                  \
  for (int _w = 0; _w < (plan).mask_words; ++_w)                                                       \
    for (uint32_t _kb = offset_bits(tm, plan, tile0, nt, _w), _mk = present_bits(tm, lane); _kb;       \
         _kb &= _kb - 1)                                                                               \
      if (const int _j = __ffs(_kb) - 1; true)                                                         \
        if (const int k = 32 * _w + _j; true)                                                          \
          if (const uint32_t m = __shfl_sync(0xffffffffu, _mk, _j); true)

__device__ __forceinline__ uint32_t any_mask(const lgConvPlan& p, int64_t tile0, int nt) {
  uint32_t m = 0;
  for (int t = 0; t < nt; ++t) {
    uint32_t w = 0;
    for (int i = 0; i < p.mask_words; ++i) w |= __ldg(p.tile_mask + (tile0 + t) * p.mask_words + i);
    m |= (w != 0 ? 1u : 0u) << t;
  }
  return bcast0(m);
}

// Gather 128 rows x (PC * 32) channels of a row-major 16-bit matrix into `dst` (PC sub-tiles of 8 KB,
// SWIZZLE_64B).  One warp does the whole stage: lane = (row % 8, 16-byte piece), 16 rows per lane
// (rows[i] = id of row 8*i + lane/4, negative = zero fill); PC copies per row are issued back to back;
// byte offsets stay in 32 bits (host checks rows * ld * 2 < 2^32).
template <int PC>
__device__ __forceinline__ void gather_rows(uint8_t* dst, const uint8_t* __restrict__ src, uint32_t ld_bytes,
                                            const int (&rows)[16], int lane) {
  const int piece = lane & 3, rsub = lane >> 2;
  const uint8_t* base = src + piece * 16;
  uint8_t* d0 = dst + sw64(rsub, piece);  // row 8*i + rsub sits 512*i bytes further (same swizzle phase)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int row = rows[i];
    const bool ok = row >= 0;
    const uint8_t* s = base + (size_t)((uint32_t)(ok ? row : 0) * ld_bytes);
#pragma unroll
    for (int c = 0; c < PC; ++c) cp_async16(d0 + i * 512 + c * kSub, s + c * 64, ok);
  }
}

__device__ __forceinline__ void gather_stage(uint8_t* dst, const uint16_t* src, int ld, int col0, int pc,
                                             const int (&rows)[16], int lane) {
  const uint8_t* s = reinterpret_cast<const uint8_t*>(src + col0);
  const uint32_t ldb = (uint32_t)ld * 2u;
  switch (pc) {
    case 1: gather_rows<1>(dst, s, ldb, rows, lane); break;
    case 2: gather_rows<2>(dst, s, ldb, rows, lane); break;
    case 3: gather_rows<3>(dst, s, ldb, rows, lane); break;
    default: gather_rows<4>(dst, s, ldb, rows, lane); break;
  }
}

// row ids of a stage from the shared-memory id ring (filled by the id warp's bulk copies)
__device__ __forceinline__ void ids_from_ring(int (&rows)[16], uint32_t slot_saddr, int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) rows[i] = lds32(slot_saddr + 4 * (8 * i + (lane >> 2)));
}

// ------------------------------------------------------------------------------------ forward / dgrad
struct Gemm2Args {
  lgConvPlan plan;
  const uint16_t* A;  // [n_in][Ck]
  float* Y;           // [n_out][N]
  float* stats;       // nullable: [n_tiles * 4][2N] per-(tile, epilogue warp) column sums / sums of squares of Y
  const float* out_scale;
  const float* bias;
  int Ck, N, n_blk, flip, umma_fmt;
  int dbg;  // experiment switches (LIDOG_DBG): 1 = no B loads, 2 = no A copies, 4 = no MMAs, 16 = no proxy fence,
            // 32 = no result stores, 64 = no row-id copies, 128 = stage release by plain arrive (only with 4);
            // 8 = per-role cycle counters of CTA 0
  int T, pc, n_panels, sa, sb, np;  // np = active producer warps (<= sa, see the ring-phase note)
  int ni;                           // row-id ring slots (a multiple of np)
  int pf;                           // id warp prefetches the operand rows of upcoming units into L2
  int bmax;                         // units the MMA warp waits for together (one proxy fence per batch), <= sa
  int sets;                         // accumulator sets in TMEM (sets * T * n_blk <= 512): 2 = the epilogue of one
                                    // super-tile overlaps the MMAs of the next
  int64_t n_tiles, n_super;
  int* err;
  int* sched;  // [0..3] next super-tile per blockIdx.y, [4] finished CTAs (self-resetting)
};

// DBG: instrumentation + experiment switches compiled in.  PC: 32-channel chunks per operand stage (= g.pc).
// SPLIT: MMA issuers.  The one-thread MMA role sets the pace of the kernel (section 4.1 of DESIGN.md), so two
// warps share it in a way that leaves every accumulator with ONE issuer (deterministic accumulation order):
//   0 = one issuer; 1 = by tile parity (super-tiles of T >= 2 tiles: warp j issues the units of tiles t = j mod 2);
//   2 = by column halves (T = 1: both warps issue every unit, warp j the MMAs of columns [j N/2, (j+1) N/2)).
template <bool DBG, int PC, int SPLIT>
__global__ void __launch_bounds__(kThreadsG2, 1) k_gemm2(const Gemm2Args g, const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stageA = PC * kSub, stageB = PC * g.n_blk * kRowB;
  uint8_t* smA = smem;
  uint8_t* smB = smem + (size_t)g.sa * stageA;
  uint8_t* smI = smB + (size_t)g.sb * stageB;  // id ring: ni slots of 128 row ids
  const int ni = g.ni;
  float* smStg = (float*)(smI + (size_t)ni * kIdxSlotBytes);  // epilogue transposition buffers
  uint64_t* fullA = (uint64_t*)((uint8_t*)smStg + kStgBytes);
  uint64_t* emptyA = fullA + g.sa;
  uint64_t* fullB = emptyA + g.sa;
  uint64_t* emptyB = fullB + g.sb;
  uint64_t* fullI = emptyB + g.sb;
  uint64_t* emptyI = fullI + ni;
  uint64_t* acc_full = emptyI + ni;
  uint64_t* acc_empty = acc_full + 2;  // one per accumulator: [set * T + tile slot]
  uint64_t* fullS = acc_empty + 8;
  uint64_t* emptyS = fullS + kSchedSlots;
  uint32_t* tmem_slot = (uint32_t*)(emptyS + kSchedSlots);
  int* sched_ring = (int*)(tmem_slot + 2);
  SchedCursor sc{0, 0};
  const int n0 = blockIdx.y * g.n_blk;
  const int K = g.plan.kernel_volume;
  const int dbg = DBG ? g.dbg : 0;  // experiment switches fold away in the production instantiation
  const bool prof = DBG && (dbg & 8) && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && warp != kMma2Warp;
  long long pacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long cta_t0 = DBG ? clock64() : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.sa; ++s) {
      mbar_init(&fullA[s], 32);
      mbar_init(&emptyA[s], SPLIT == 2 ? 2 : 1);  // column split: both issuers read the stage
    }
    for (int s = 0; s < g.sb; ++s) {
      mbar_init(&fullB[s], 1);
      mbar_init(&emptyB[s], SPLIT ? 2 : 1);  // every issuer releases the weight panel
    }
    for (int s = 0; s < ni; ++s) {
      mbar_init(&fullI[s], 1);
      mbar_init(&emptyI[s], 1);
    }
    mbar_init(&acc_full[0], SPLIT ? 2 : 1);
    mbar_init(&acc_full[1], SPLIT ? 2 : 1);
    for (int t = 0; t < 8; ++t) mbar_init(&acc_empty[t], kEpiWarps * 32);
    for (int t = 0; t < kSchedSlots; ++t) {
      mbar_init(&fullS[t], 1);
      mbar_init(&emptyS[t], kSchedConsumers + (SPLIT ? 1 : 0));
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  if (warp == kBWarp && lane == 0) prefetch_tmap(&tmB);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kIdxWarp) {
    // ===================================================================== id ring (runs ahead; one elected lane issues)
    int slot = 0;
    uint32_t iphase = 0;
    // Optional (LIDOG_G2_OPT & 4, OFF by default): L2 prefetch of the operand rows, kPfLag units behind the id copies
    // and up to ni - kPfLag units ahead of the gathers.  Measured on B200 (profiles/r01_s4_sweep_b.txt): 128
    // cp.async.bulk.prefetch.L2 per unit made every layer 1.2-6x SLOWER (0.229 -> 0.834 ms on the block8 shape);
    // the bulk-prefetch path serialises far below the rate the LDGSTS gathers sustain.  Kept as the record.
    constexpr int kPfLag = 8;
    int pslot = 0, pf_count = 0, iunit = 0;
    uint32_t pphase = 0, pf_hist = 0;
    const uint32_t row_bytes = (uint32_t)g.Ck * 2u;
    for (;;) {
      mbar_wait(&emptyS[sc.slot], sc.phase ^ 1, g.err, 21);
      int drawn = 0;
      if (lane == 0) drawn = atomicAdd(g.sched + blockIdx.y, 1);
      drawn = (int)bcast0((uint32_t)drawn);
      const int64_t st = drawn < g.n_super ? g.n_super - 1 - drawn : -1;
      const int64_t tile0 = st * g.T;
      const int nt = st < 0 ? 0 : (int)min((int64_t)g.T, g.n_tiles - tile0);
      // the entry carries the tiles' mask words, so the other nine warps never go to global memory for them
      uint32_t w0 = 0;
      if (lane < nt) w0 = __ldg(g.plan.tile_mask + (tile0 + lane) * g.plan.mask_words);
      uint32_t any = __ballot_sync(0xffffffffu, w0 != 0);
      if (g.plan.mask_words > 1 && st >= 0) any = any_mask(g.plan, tile0, nt);
      int* entry = sched_ring + sc.slot * kSchedWords;
      if (lane < 8) entry[2 + lane] = (int)w0;
      if (lane == 0) {
        entry[0] = (int)st;
        entry[1] = (int)any;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&fullS[sc.slot]);
      __syncwarp();
      if (++sc.slot == kSchedSlots) {
        sc.slot = 0;
        sc.phase ^= 1;
      }
      if (st < 0) break;
      TileMasks tm;
#pragma unroll
      for (int t = 0; t < 8; ++t) tm.w[t] = __shfl_sync(0xffffffffu, w0, t);
      LG_FOR_EACH_OFFSET(tm, g.plan, tile0, nt, k, m) {
        for (int p = 0; p < g.n_panels; ++p) {
          for (uint32_t mm = m; mm; mm &= mm - 1) {
            const int t = __ffs(mm) - 1;
            if (!(dbg & 512)) mbar_wait(&emptyI[slot], iphase ^ 1, g.err, 7);
            TRACE(4, iunit);
            ++iunit;
            if (elect_one()) {
              if (dbg & 64) {  // experiment: no id copies (the producers gather whatever the ring holds)
                mbar_arrive(&fullI[slot]);
              } else {
                mbar_arrive_expect_tx(&fullI[slot], kIdxSlotBytes);
                bulk_copy_g2s(smI + (size_t)slot * kIdxSlotBytes,
                              g.plan.nbr + (int64_t)k * g.plan.k_stride + (tile0 + t) * LG_TILE_ROWS, kIdxSlotBytes,
                              &fullI[slot]);
              }
            }
            __syncwarp();
            if (++slot == ni) {
              slot = 0;
              iphase ^= 1;
            }
            if (g.pf) {
              pf_hist = (pf_hist << 1) | (p == 0 ? 1u : 0u);  // later panels of a unit re-read the same rows
              if (++pf_count > kPfLag) {
                if ((pf_hist >> kPfLag) & 1u) {
                  mbar_wait(&fullI[pslot], pphase, g.err, 9);
                  const int4 ids = *reinterpret_cast<const int4*>(smI + (size_t)pslot * kIdxSlotBytes + lane * 16);
                  const uint8_t* a = reinterpret_cast<const uint8_t*>(g.A);
                  if (ids.x >= 0) bulk_prefetch_l2(a + (size_t)ids.x * row_bytes, row_bytes);
                  if (ids.y >= 0) bulk_prefetch_l2(a + (size_t)ids.y * row_bytes, row_bytes);
                  if (ids.z >= 0) bulk_prefetch_l2(a + (size_t)ids.z * row_bytes, row_bytes);
                  if (ids.w >= 0) bulk_prefetch_l2(a + (size_t)ids.w * row_bytes, row_bytes);
                }
                if (++pslot == ni) {
                  pslot = 0;
                  pphase ^= 1;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp >= kProdWarp0 && warp < kProdWarp0 + kProdWarps) {
    // ===================================================================== A producers (cp.async gather)
    const int pw = warp - kProdWarp0;
    int stage = 0, turn = 0, slot = 0, punit = 0;
    uint32_t phase = 0, iphase = 0;
    SchedItem item;
    for (;;) {
      sched_next(item, sc, fullS, emptyS, sched_ring, lane, g.err);
      const int64_t st = item.st;
      if (st < 0) break;
      const int64_t tile0 = st * g.T;
      const int nt = (int)min((int64_t)g.T, g.n_tiles - tile0);
      TileMasks& tm = item.tm;
      LG_FOR_EACH_OFFSET(tm, g.plan, tile0, nt, k, m) {
        for (int p = 0; p < g.n_panels; ++p) {
          for (int t = 0; t < nt; ++t) {
            if (!((m >> t) & 1u)) continue;
            if (turn == pw) {
              PROF_T0();
              int rows[16];
              if (!(dbg & 512)) {  // dbg 512: experiment without the id ring on the producer side
                mbar_wait(&fullI[slot], iphase, g.err, 8);
                ids_from_ring(rows, smem_u32(smI) + slot * kIdxSlotBytes, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyI[slot]);
              }
              if (pw == 0) PROF_ADD(15);
              if (lane == 0) mbar_wait(&emptyA[stage], phase ^ 1, g.err, 1);
              __syncwarp();
              TRACE(0, punit);
              if (pw == 0) PROF_ADD(0);
              if (!(dbg & 2))
                gather_rows<PC>(smA + (size_t)stage * stageA, reinterpret_cast<const uint8_t*>(g.A + p * PC * 32), (uint32_t)g.Ck * 2u, rows, lane);
              if (dbg & 256)  // experiment (with dbg 2 only): plain arrive instead of the cp.async-tracked one
                mbar_arrive(&fullA[stage]);
              else
                cp_async_arrive_noinc(&fullA[stage]);
              TRACE(1, punit);
              if (pw == 0) PROF_ADD(1);
            }
            ++punit;
            turn = (turn + 1 == g.np) ? 0 : turn + 1;
            if (++stage == g.sa) {
              stage = 0;
              phase ^= 1;
            }
            if (++slot == ni) {
              slot = 0;
              iphase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == kBWarp) {
    // ===================================================================== B producer (weight panels, TMA)
    int bs = 0;
    uint32_t bphase = 0;
    SchedItem item;
    for (;;) {
      sched_next(item, sc, fullS, emptyS, sched_ring, lane, g.err);
      const int64_t st = item.st;
      if (st < 0) break;
      const int64_t tile0 = st * g.T;
      const int nt = (int)min((int64_t)g.T, g.n_tiles - tile0);
      TileMasks& tm = item.tm;
      LG_FOR_EACH_OFFSET(tm, g.plan, tile0, nt, k, m) {
        (void)m;
        const int wk = g.flip ? (K - 1 - k) : k;
        for (int p = 0; p < g.n_panels; ++p) {
          PROF_T0();
          mbar_wait(&emptyB[bs], bphase ^ 1, g.err, 2);
          PROF_ADD(2);
          if (elect_one()) {
            if (dbg & 1) {
              mbar_arrive(&fullB[bs]);
            } else {
              mbar_arrive_expect_tx(&fullB[bs], (uint32_t)stageB);
              uint8_t* dst = smB + (size_t)bs * stageB;
#pragma unroll
              for (int c = 0; c < PC; ++c)
                tma_load_2d(dst + c * g.n_blk * kRowB, &tmB, (p * PC + c) * 32, wk * g.N + n0, &fullB[bs]);
            }
          }
          __syncwarp();
          if (++bs == g.sb) {
            bs = 0;
            bphase ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp || warp == kMma2Warp) {
    // ===================================================================== MMA issuer(s)
    const int mw = (warp == kMmaWarp) ? 0 : 1;  // issuer index
    if (SPLIT == 0 && mw) goto role_done;         // single-issuer instantiation: the extra warp has no role
    // The loop runs warp-uniformly (every lane computes the same unit sequence; the masks are broadcast), and
    // one elected lane issues the tcgen05 instructions: descriptors are a constant high word plus
    // (address >> 4) held in uniform registers, so a unit costs two barrier waits, a proxy fence and 2*pc
    // back-to-back UTCHMMAs.  Accumulators are handed back by the epilogue tile by tile (acc_empty[t]), so
    // the first offsets of the next super-tile overlap the drain of the previous one.
    // column split: this issuer's MMAs cover n_mma = N/2 columns starting at column col0 (weight rows col0..)
    const int n_mma = (SPLIT == 2) ? g.n_blk / 2 : g.n_blk;
    const uint32_t col0 = (SPLIT == 2 && mw) ? (uint32_t)(g.n_blk / 2) : 0u;
    const uint32_t idesc = make_idesc(g.umma_fmt, 0, 0, LG_TILE_ROWS, n_mma);
    const uint64_t desc_hi = make_smem_desc(0, 16, 8 * kRowB, kLayoutSw64);
    const uint32_t a_base = smem_u32(smA) >> 4, b_base = (smem_u32(smB) + col0 * kRowB) >> 4;
    const uint32_t a_stage16 = (uint32_t)stageA >> 4, b_stage16 = (uint32_t)stageB >> 4;
    const uint32_t b_sub16 = (uint32_t)(g.n_blk * kRowB) >> 4;
    // everything the loop needs lives in registers: no parameter loads, no 64-bit index arithmetic in the hot path
    const int sa = g.sa, sb = g.sb, T = g.T, n_panels = g.n_panels, n_blk = g.n_blk, two_sets = g.sets == 2;
    const uint32_t mine_bits = (SPLIT == 1) ? (mw ? 0xAAAAAAAAu : 0x55555555u) : 0xFFFFFFFFu;  // tiles this issuer owns
    int stage = 0, bs = 0, munit = 0;
    uint32_t phase = 0, bphase = 0, it = 0;
    SchedItem item;
    for (;;) {
      PROF_T0();
      sched_next(item, sc, fullS, emptyS, sched_ring, lane, g.err);
      const int64_t st = item.st;
      if (st < 0) break;
      const int64_t tile0 = st * T;
      const int nt = (int)min((int64_t)T, g.n_tiles - tile0);
      if (!item.any) continue;
      const int set = two_sets ? (int)(it & 1) : 0;
      const uint32_t eparity = ((two_sets ? (it >> 1) : it) & 1) ^ 1;
      const uint32_t d_base = tmem_base + (uint32_t)(set * T) * (uint32_t)n_blk;  // first accumulator of this set
      ++it;
      // all T accumulators of the set are taken up front (the epilogue hands them back tile by tile; the last one
      // arrives a few hundred cycles after the first) -- the inner loop then has no hand-off test per unit
      for (int t = 0; t < T; ++t) mbar_wait(&acc_empty[set * T + t], eparity, g.err, 3);
      uint32_t started = 0;
      TileMasks& tm = item.tm;
      PROF_ADD(11);
      LG_FOR_EACH_OFFSET(tm, g.plan, tile0, nt, k, m) {
        for (int p = 0; p < n_panels; ++p) {
          PROF_ADD(7);
          mbar_wait(&fullB[bs], bphase, g.err, 4);
          PROF_ADD(4);
          const uint32_t db_lo = b_base + bs * b_stage16;
          // The units of this (offset, panel) -- at most T <= ring depth (host-checked) -- are handled together:
          // wait for their stages, ONE proxy fence, issue their MMAs back to back, release the stages.
          const uint32_t mym = m & mine_bits;  // the units of this step that this issuer handles
          {
            int s = stage;
            uint32_t ph = phase;
            for (uint32_t bb = m; bb; bb &= bb - 1) {
              // Every issuer observes EVERY phase of every stage barrier, also for the units it does not issue: a
              // parity wait cannot tell phase n from phase n + 2, and an issuer that skipped the other issuer's
              // phase and ran >= one ring wrap ahead took the stale completion of its own previous use for the
              // new one (this is what trapped the tile-parity mode on the 5-stage 96->96 shape).
              mbar_wait(&fullA[s], ph, g.err, 5);
              TRACE(2, munit + __popc(m & ~bb));
              if (++s == sa) {
                s = 0;
                ph ^= 1;
              }
            }
          }
          PROF_ADD(5);
          if (!(dbg & 16)) fence_proxy_async();  // cp.async (generic proxy) writes -> tcgen05 (async proxy) reads
          tc_fence_after();
          PROF_ADD(10);
          if (elect_one()) {
            int s = stage;
            for (uint32_t bb = m; bb; bb &= bb - 1) {
              const int t = __ffs(bb) - 1;
              if (SPLIT == 1 && !((mym >> t) & 1u)) {  // the other issuer's unit: only the stage counter moves
                if (++s == sa) s = 0;
                continue;
              }
              const uint64_t da0 = desc_hi | (uint64_t)(a_base + s * a_stage16);
              const uint64_t db0 = desc_hi | (uint64_t)db_lo;
              const uint32_t d_tmem = d_base + (uint32_t)t * (uint32_t)n_blk + col0;
              const uint32_t acc0 = (started >> t) & 1u;
              if (!(dbg & 4)) {
#pragma unroll
                for (int c = 0; c < PC; ++c) {
                  umma_f16(d_tmem, da0 + c * (kSub >> 4), db0 + c * b_sub16, idesc, c == 0 ? acc0 : 1u);
                  umma_f16(d_tmem, da0 + c * (kSub >> 4) + 2, db0 + c * b_sub16 + 2, idesc, 1u);
                }
              }
              if (DBG && (dbg & 128))  // experiment (with dbg 4 only): plain arrive instead of tcgen05.commit
                mbar_arrive(&emptyA[s]);
              else
                umma_commit(&emptyA[s]);
              if (++s == sa) s = 0;
            }
            if (DBG && (dbg & 128))
              mbar_arrive(&emptyB[bs]);
            else
              umma_commit(&emptyB[bs]);
          }
          __syncwarp();
          if (DBG) {
            for (int i = 0; i < __popc(m); ++i) TRACE(3, munit + i);
            munit += __popc(m);
            pacc[13] += __popc(m);  // units
          }
          PROF_ADD(6);
          started |= mym;
          stage += __popc(m);
          if (stage >= sa) {
            stage -= sa;
            phase ^= 1;
          }
          if (++bs == sb) {
            bs = 0;
            bphase ^= 1;
          }
        }
      }
      if (elect_one()) umma_commit(&acc_full[set]);
      __syncwarp();
      PROF_ADD(3);
      pacc[14] += 1;  // super-tiles
    }
    if (DBG) pacc[12] = clock64() - cta_t0;  // lifetime of the MMA role of this CTA
  } else {
    // ===================================================================== epilogue (warps 0..3)
    const float scale = g.out_scale ? g.out_scale[0] : 1.f;
    uint32_t it = 0;
    SchedItem item;
    for (;;) {
      sched_next(item, sc, fullS, emptyS, sched_ring, lane, g.err);
      const int64_t st = item.st;
      if (st < 0) break;
      const int64_t tile0 = st * g.T;
      const int nt = (int)min((int64_t)g.T, g.n_tiles - tile0);
      const uint32_t am = item.any;
      PROF_T0();
      int a0 = 0;
      if (am) {
        const int set = (g.sets == 2) ? (int)(it & 1) : 0;
        a0 = set * g.T;
        mbar_wait(&acc_full[set], ((g.sets == 2) ? (it >> 1) : it) & 1, g.err, 6);
        ++it;
        tc_fence_after();
      }
      if (warp == 0) PROF_ADD(8);
      for (int t = 0; t < g.T; ++t) {
        if (t < nt) {
          const int64_t s = (tile0 + t) * LG_TILE_ROWS + warp * 32 + lane;
          int64_t row = g.plan.out_row ? (int64_t)g.plan.out_row[s] : s;
          const bool row_ok = row >= 0 && row < g.plan.n_out;
          float* yrow = g.Y + (row_ok ? row : 0) * g.N + n0;
          // BN statistics of the layer that follows: this warp's 32 rows contribute one partial
          // row [sum | sum of squares] per tile; a fixed shuffle order and a fixed (tile, warp) slot keep the later
          // reduction deterministic although tiles are handed out dynamically.
          float* strow = g.stats ? g.stats + ((tile0 + t) * kEpiWarps + warp) * 2 * (int64_t)g.N + n0 : nullptr;
          if (!((am >> t) & 1u)) {  // no neighbour at all: bias / zeros
            if (row_ok)
              for (int n = 0; n < g.n_blk; n += 4) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.bias) o = *reinterpret_cast<const float4*>(g.bias + n0 + n);
                *reinterpret_cast<float4*>(yrow + n) = o;
              }
            if (strow) {
              const int valid = __popc(__ballot_sync(0xffffffffu, row_ok));
              for (int n = 4 * lane; n < g.n_blk; n += 128) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.bias) b4 = *reinterpret_cast<const float4*>(g.bias + n0 + n);
                const float v = (float)valid;
                *reinterpret_cast<float4*>(strow + n) = make_float4(v * b4.x, v * b4.y, v * b4.z, v * b4.w);
                *reinterpret_cast<float4*>(strow + g.N + n) =
                    make_float4(v * b4.x * b4.x, v * b4.y * b4.y, v * b4.z * b4.z, v * b4.w * b4.w);
              }
            }
          } else {
            // TMEM lane = tile row, so a thread holds 32 consecutive columns of ONE row; storing them directly
            // makes every warp store touch 32 different lines (measured: the epilogue was store-bound).  The
            // 32x32 block goes through shared memory instead and leaves as 4 rows x 128 contiguous bytes per
            // warp instruction.
            const uint32_t taddr = tmem_base + (a0 + t) * g.n_blk + ((uint32_t)(warp * 32) << 16);
            float* stg = smStg + warp * 32 * kStgPitch;
            int srow[8];  // result rows this lane stores: tile rows (lane >> 3) + 4 j
#pragma unroll
            for (int j = 0; j < 8; ++j) srow[j] = __shfl_sync(0xffffffffu, row_ok ? (int)row : -1, (lane >> 3) + 4 * j);
            int n = 0;
            for (; n + 32 <= g.n_blk; n += 32) {
              uint32_t v[32];
              tmem_ld32(taddr + n, v);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 o;
                o.x = __uint_as_float(v[4 * q + 0]) * scale;
                o.y = __uint_as_float(v[4 * q + 1]) * scale;
                o.z = __uint_as_float(v[4 * q + 2]) * scale;
                o.w = __uint_as_float(v[4 * q + 3]) * scale;
                if (g.bias) {
                  const float4 bb = *reinterpret_cast<const float4*>(g.bias + n0 + n + 4 * q);
                  o.x += bb.x, o.y += bb.y, o.z += bb.z, o.w += bb.w;
                }
                *reinterpret_cast<float4*>(stg + lane * kStgPitch + 4 * q) = o;
              }
              __syncwarp();
              float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq = cs;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 o = *reinterpret_cast<const float4*>(stg + ((lane >> 3) + 4 * j) * kStgPitch + 4 * (lane & 7));
                if (srow[j] >= 0) {
                  // (Summing the gradient of a multi-consumer tensor here -- read-modify-write instead of a store --
                  // was measured: the 31 separate adds it removes cost 0.6 ms per step, the loads in this epilogue
                  // 2.8 ms (k_gemm2 8.39 -> 11.17 ms, profiles/r02_c_launch_summary_grad_accum_experiment.txt): the
                  // accumulators drain late and the MMA warp stalls.  Not kept.)
                  if (!(dbg & 32))  // dbg 32: experiment without the result stores
                    *reinterpret_cast<float4*>(g.Y + (int64_t)srow[j] * g.N + n0 + n + 4 * (lane & 7)) = o;
                  cs.x += o.x, cs.y += o.y, cs.z += o.z, cs.w += o.w;
                  cq.x = fmaf(o.x, o.x, cq.x), cq.y = fmaf(o.y, o.y, cq.y), cq.z = fmaf(o.z, o.z, cq.z),
                  cq.w = fmaf(o.w, o.w, cq.w);
                }
              }
              if (strow) {  // lanes l, l+8, l+16, l+24 hold the same 4 columns of different rows
#pragma unroll
                for (int d = 8; d <= 16; d <<= 1) {
                  cs.x += __shfl_xor_sync(0xffffffffu, cs.x, d), cs.y += __shfl_xor_sync(0xffffffffu, cs.y, d);
                  cs.z += __shfl_xor_sync(0xffffffffu, cs.z, d), cs.w += __shfl_xor_sync(0xffffffffu, cs.w, d);
                  cq.x += __shfl_xor_sync(0xffffffffu, cq.x, d), cq.y += __shfl_xor_sync(0xffffffffu, cq.y, d);
                  cq.z += __shfl_xor_sync(0xffffffffu, cq.z, d), cq.w += __shfl_xor_sync(0xffffffffu, cq.w, d);
                }
                if (lane < 8) {
                  *reinterpret_cast<float4*>(strow + n + 4 * lane) = cs;
                  *reinterpret_cast<float4*>(strow + g.N + n + 4 * lane) = cq;
                }
              }
              __syncwarp();
            }
            for (; n < g.n_blk; n += 16) {  // 16-column remainder (N % 32 == 16): direct stores
              uint32_t v[16];
              tmem_ld16(taddr + n, v);
              tmem_ld_wait();
              if (row_ok) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  float4 o;
                  o.x = __uint_as_float(v[4 * q + 0]) * scale;
                  o.y = __uint_as_float(v[4 * q + 1]) * scale;
                  o.z = __uint_as_float(v[4 * q + 2]) * scale;
                  o.w = __uint_as_float(v[4 * q + 3]) * scale;
                  if (g.bias) {
                    const float4 bb = *reinterpret_cast<const float4*>(g.bias + n0 + n + 4 * q);
                    o.x += bb.x, o.y += bb.y, o.z += bb.z, o.w += bb.w;
                  }
                  *reinterpret_cast<float4*>(yrow + n + 4 * q) = o;
                }
              }
            }
          }
        }
        if (am) {  // hand accumulator t back (tcgen05.ld of it have completed: wait::ld above)
          tc_fence_before();
          mbar_arrive(&acc_empty[a0 + t]);
        }
      }
      if (warp == 0) PROF_ADD(9);
    }
  }
role_done:
  if (warp == kProdWarp0) PROF_FLUSH(0, 1);
  if (warp == kProdWarp0) PROF_FLUSH(15, 15);
  if (warp == kBWarp) PROF_FLUSH(2, 2);
  if (warp == kMmaWarp) PROF_FLUSH(3, 7);
  if (warp == kMmaWarp) PROF_FLUSH(10, 14);  // 11 = schedule hand-out + super-tile prologue of the MMA warp
  if (warp == 0) PROF_FLUSH(8, 9);
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
  if (threadIdx.x == 0) {  // the last CTA to finish rewinds the counters for the launch that reuses this slot
    __threadfence();
    const int done = atomicAdd(g.sched + 4, 1);
    if (done == (int)(gridDim.x * gridDim.y) - 1) {
      for (int y = 0; y < 5; ++y) g.sched[y] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------ wgrad
// dW[k] (Cin x Cout) = sum_s X16[nbr[k][s], :]^T dY16[out_row[s] (or s), :].
// CTA = (group of G kernel offsets, 128-channel block of Cin, chunk of tiles).
// A stage: gathered X rows [128 pairs x 4 sub-blocks of 32 channels]; B stage: dY rows [128 pairs x Cout].
struct Wgrad2Args {
  lgConvPlan plan;
  const uint16_t* X;
  const uint16_t* dY;
  float* partial;  // [chunks][K][Cin][Cout]
  int Cin, Cout, m_blocks, G, n_groups, tiles_per_chunk, umma_fmt, sa, sb, np, bmax;
  int strided;  // offsets of a group are grp + j * n_groups (load balance) instead of a contiguous range
  int na_max;  // 32-channel sub-blocks per A stage (4; see the launcher for why not min(4, Cin / 32))
  int dbg;     // DBG instantiation only (LIDOG_WG_DBG): 1 = no dY gathers, 2 = no X gathers, 4 = no MMAs, 8 = no partial stores
  int64_t n_tiles;
  int* err;
};

template <bool DBG>
__global__ void __launch_bounds__(kThreads, 1) k_wgrad2(const Wgrad2Args g) {
  const int dbg = DBG ? g.dbg : 0;  // experiment switches fold away in the production instantiation
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.x / g.m_blocks, mb = blockIdx.x % g.m_blocks;
  const int chunk = blockIdx.y;
  const int K = g.plan.kernel_volume;
  // The offsets of a group are STRIDED: group grp owns offsets grp, grp + n_groups, grp + 2 n_groups, ...  Contiguous
  // ranges put the whole centre plane of a 3x3x3 kernel (offsets 9..17: 60 % of all pairs of a LiDAR map) into one or
  // two groups: on the bench batch the heaviest CTA had 2.8x the mean number of units (190 vs 69,
  // profiles/r02_d_wgrad_balance.txt); strided groups bring that to 1.4x for 1.5x more dY tile loads.
  const int ng = g.strided ? g.n_groups : 1;
  const int k0 = g.strided ? grp : grp * g.G;                                  // first offset of the group
  const int gcount = g.strided ? (K - grp + ng - 1) / ng : min(K, k0 + g.G) - k0;  // number of offsets
  const int m0 = mb * 128;
  const int na = min(4, (g.Cin - m0) / 32);  // real 32-channel sub-blocks of the A operand
  const int nb = g.Cout / 32;
  // (the MMA always reads 4 sub-blocks = 128 accumulator lanes; lanes of sub-blocks that are not staged hold whatever
  // the neighbouring stage holds and are never read back: the epilogue stops at Cin)
  const int stageA = g.na_max * kSub, stageB = nb * kSub;
  uint8_t* smA = smem;
  uint8_t* smB = smem + (size_t)g.sa * stageA;
  uint8_t* smI = smB + (size_t)g.sb * stageB;  // id ring: ni slots of 128 row ids
  const int ni = 8 * g.np;
  uint64_t* fullA = (uint64_t*)(smI + (size_t)ni * kIdxSlotBytes);
  uint64_t* emptyA = fullA + g.sa;
  uint64_t* fullB = emptyA + g.sa;
  uint64_t* emptyB = fullB + g.sb;
  uint64_t* fullI = emptyB + g.sb;
  uint64_t* emptyI = fullI + ni;
  uint64_t* done = emptyI + ni;
  uint32_t* tmem_slot = (uint32_t*)(done + 1);
  // chunk c owns tiles c, c + chunks, c + 2 chunks, ...: with mask-sorted plans neighbouring tiles have the
  // same offsets, so interleaving gives every CTA of an offset group the same amount of work
  const int64_t t0 = chunk, t1 = g.n_tiles, tstep = gridDim.y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.sa; ++s) {
      mbar_init(&fullA[s], 32);
      mbar_init(&emptyA[s], 1);
    }
    for (int s = 0; s < g.sb; ++s) {
      mbar_init(&fullB[s], 32);
      mbar_init(&emptyB[s], 1);
    }
    for (int s = 0; s < ni; ++s) {
      mbar_init(&fullI[s], 1);
      mbar_init(&emptyI[s], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // bit j of group_mask(tile) = offset k0 + j * ng present in the tile (K <= 64, host-checked)
  auto group_mask = [&](int64_t tile) -> uint32_t {
    const uint32_t* mw = g.plan.tile_mask + tile * g.plan.mask_words;
    uint64_t bits = (uint64_t)__ldg(mw) | (g.plan.mask_words > 1 ? (uint64_t)__ldg(mw + 1) << 32 : 0ull);
    bits >>= k0;
    uint32_t m;
    if (ng == 1) {
      m = (uint32_t)bits;
    } else {
      m = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < gcount) m |= (uint32_t)((bits >> (j * ng)) & 1ull) << j;
    }
    return bcast0(m & ((1u << gcount) - 1u));
  };

  if (warp == kIdxWarp) {
    // ---------------------------------------------------------------- id ring (runs ahead; one elected lane issues)
    int slot = 0;
    uint32_t iphase = 0;
    for (int64_t tile = t0; tile < t1; tile += tstep) {
      for (uint32_t mm = group_mask(tile); mm; mm &= mm - 1) {
        const int k = k0 + (__ffs(mm) - 1) * ng;
        mbar_wait(&emptyI[slot], iphase ^ 1, g.err, 16);
        if (elect_one()) {
          mbar_arrive_expect_tx(&fullI[slot], kIdxSlotBytes);
          bulk_copy_g2s(smI + (size_t)slot * kIdxSlotBytes,
                        g.plan.nbr + (int64_t)k * g.plan.k_stride + tile * LG_TILE_ROWS, kIdxSlotBytes, &fullI[slot]);
        }
        __syncwarp();
        if (++slot == ni) {
          slot = 0;
          iphase ^= 1;
        }
      }
    }
  } else if (warp >= kProdWarp0) {
    // ---------------------------------------------------------------- X producers
    const int pw = warp - kProdWarp0;
    int stage = 0, turn = 0, slot = 0;
    uint32_t phase = 0, iphase = 0;
    for (int64_t tile = t0; tile < t1; tile += tstep) {
      for (uint32_t mm = group_mask(tile); mm; mm &= mm - 1) {
        if (turn == pw) {
          int rows[16];
          mbar_wait(&fullI[slot], iphase, g.err, 17);
          ids_from_ring(rows, smem_u32(smI) + slot * kIdxSlotBytes, lane);
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&emptyI[slot]);
            mbar_wait(&emptyA[stage], phase ^ 1, g.err, 11);
          }
          __syncwarp();
          if (!(dbg & 2)) gather_stage(smA + (size_t)stage * stageA, g.X, g.Cin, m0, na, rows, lane);
          cp_async_arrive_noinc(&fullA[stage]);
        }
        turn = (turn + 1 == g.np) ? 0 : turn + 1;
        if (++stage == g.sa) {
          stage = 0;
          phase ^= 1;
        }
        if (++slot == ni) {
          slot = 0;
          iphase ^= 1;
        }
      }
    }
  } else if (warp == kBWarp) {
    // ---------------------------------------------------------------- dY producer (one warp, whole tiles)
    // ncu (profiles/r02_a_wgrad2_source_stalls.txt): the MMA warp polled this ring 41 times per tile -- two stages and
    // two dependent global latencies per tile (out_row, then the rows) could not keep it fed.  The row ids of the NEXT
    // tile are now loaded before the wait for a free stage, and the ring is as deep as shared memory allows.
    int bs = 0;
    uint32_t bphase = 0;
    const int lim = (int)g.plan.n_out;
    auto next_tile = [&](int64_t t) {
      while (t < t1 && !group_mask(t)) t += tstep;
      return t;
    };
    auto load_rows = [&](int (&rows)[16], int64_t tile) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int sl = (int)(tile * LG_TILE_ROWS) + 8 * i + (lane >> 2);
        const int r = g.plan.out_row ? __ldg(g.plan.out_row + sl) : sl;
        rows[i] = (r < lim) ? r : -1;
      }
    };
    int64_t tile = next_tile(t0);
    int rows[16], rows_n[16];
    if (tile < t1) load_rows(rows, tile);
    while (tile < t1) {
      const int64_t nxt = next_tile(tile + tstep);
      if (nxt < t1) load_rows(rows_n, nxt);
      if (lane == 0) mbar_wait(&emptyB[bs], bphase ^ 1, g.err, 12);
      __syncwarp();
      uint8_t* dst = smB + (size_t)bs * stageB;
      if (!(dbg & 1))
        for (int c0 = 0; c0 < nb; c0 += 4)
          gather_stage(dst + c0 * kSub, g.dY, g.Cout, c0 * 32, min(4, nb - c0), rows, lane);
      cp_async_arrive_noinc(&fullB[bs]);
      if (++bs == g.sb) {
        bs = 0;
        bphase ^= 1;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) rows[i] = rows_n[i];
      tile = nxt;
    }
  } else if (warp == kMmaWarp) {
    // warp-uniform loop, one elected lane issues (see k_gemm2)
    const uint32_t idesc = make_idesc(g.umma_fmt, 1, 1, 128, g.Cout);
    // MN-major SWIZZLE_64B: LBO = pitch between 32-channel sub-blocks, SBO = 8 rows * 64 B
    const uint64_t desc_hi = make_smem_desc(0, kSub, 8 * kRowB, kLayoutSw64);
    const uint32_t a_base = smem_u32(smA) >> 4, b_base = smem_u32(smB) >> 4;
    const uint32_t a_stage16 = (uint32_t)stageA >> 4, b_stage16 = (uint32_t)stageB >> 4;
    int stage = 0, bs = 0;
    uint32_t phase = 0, bphase = 0, started = 0;
    for (int64_t tile = t0; tile < t1; tile += tstep) {
      const uint32_t m = group_mask(tile);
      if (!m) continue;
      mbar_wait(&fullB[bs], bphase, g.err, 13);
      const uint64_t db0 = desc_hi | (uint64_t)(b_base + bs * b_stage16);
      // the offsets of this tile in batches of <= bmax (<= ring depth): all waits, one proxy fence, then the MMAs
      uint32_t mm = m;
      while (mm) {
        uint32_t batch = 0;
        {
          uint32_t x = mm;
          for (int i = 0; i < g.bmax && x; ++i) {
            batch |= x & (0u - x);
            x &= x - 1;
          }
        }
        mm &= ~batch;
        {
          int s = stage;
          uint32_t ph = phase;
          for (uint32_t bb = batch; bb; bb &= bb - 1) {
            mbar_wait(&fullA[s], ph, g.err, 14);
            if (++s == g.sa) {
              s = 0;
              ph ^= 1;
            }
          }
        }
        fence_proxy_async();
        tc_fence_after();
        if (elect_one()) {
          int s = stage;
          for (uint32_t bb = batch; bb; bb &= bb - 1) {
            const int j = __ffs(bb) - 1;
            const uint64_t da0 = desc_hi | (uint64_t)(a_base + s * a_stage16);
            const uint32_t d_tmem = tmem_base + j * g.Cout;
            const uint32_t acc0 = (started >> j) & 1u;
            if (!(dbg & 4)) {
              umma_f16(d_tmem, da0, db0, idesc, acc0);
#pragma unroll
              for (int q = 1; q < LG_TILE_ROWS / 16; ++q)  // K = 16 gathered rows per MMA
                umma_f16(d_tmem, da0 + q * (16 * kRowB >> 4), db0 + q * (16 * kRowB >> 4), idesc, 1u);
            }
            umma_commit(&emptyA[s]);
            if (++s == g.sa) s = 0;
          }
        }
        __syncwarp();
        started |= batch;
        for (int i = __popc(batch); i > 0; --i) {
          if (++stage == g.sa) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (elect_one()) umma_commit(&emptyB[bs]);
      __syncwarp();
      if (++bs == g.sb) {
        bs = 0;
        bphase ^= 1;
      }
    }
    if (elect_one()) {
      if (started)
        umma_commit(done);
      else
        mbar_arrive(done);
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue: lane = input channel
    uint32_t started = 0;
    for (int64_t tile = t0; tile < t1; tile += tstep) started |= group_mask(tile);
    const int ci = m0 + warp * 32 + lane;
    mbar_wait(done, 0, g.err, 15);
    tc_fence_after();
    for (int j = 0; j < gcount; ++j) {
      const int k = k0 + j * ng;
      const bool have = (started >> j) & 1u;
      float* prow = g.partial + (((int64_t)chunk * K + k) * g.Cin + (ci < g.Cin ? ci : 0)) * g.Cout;
      const uint32_t taddr = tmem_base + j * g.Cout + ((uint32_t)(warp * 32) << 16);
      for (int n = 0; n < g.Cout; n += 16) {
        uint32_t v[16];
        if (have) {
          tmem_ld16(taddr + n, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0u;
        }
        if (ci < g.Cin && !(dbg & 8)) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(prow + n + 4 * q) =
                make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                            __uint_as_float(v[4 * q + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

constexpr size_t kSmemBudget = 225 * 1024;

static inline size_t wgrad_tail_bytes(int sa, int sb) {  // id ring + barriers + tmem slot + alignment slack
  const int ni = 8 * kProdWarps;
  return (size_t)ni * kIdxSlotBytes + (size_t)(2 * sa + 2 * sb + 2 * ni + 2) * 8 + 16 + 1024;
}
static inline size_t tail_bytes(int sa, int sb, int ni = 8 * kProdWarps) {  // + epilogue staging + schedule ring
  return (size_t)ni * kIdxSlotBytes + kStgBytes + (size_t)(2 * sa + 2 * sb + 2 * ni + 11 + 2 * kSchedSlots) * 8 + 16 +
         4 * kSchedWords * kSchedSlots + 1024;
}

}  // namespace v2

int debug_profile(long long* out16, int reset) {
  if (out16) LG_CUDA_OK(cudaMemcpyFromSymbol(out16, v2::g_prof, sizeof(long long) * 16));
  if (reset) {
    long long z[16] = {0};
    LG_CUDA_OK(cudaMemcpyToSymbol(v2::g_prof, z, sizeof(z)));
  }
  return LG_OK;
}

// Experiment switches, read ONCE per process (they used to cost ~10 getenv calls per launch):
//   LIDOG_DBG        instrumented instantiation + experiment bits (see Gemm2Args::dbg)
//   LIDOG_ACC_SETS   1 = single accumulator set
//   LIDOG_G2_OPT     bit 2 = deep weight-panel ring, bit 4 = L2 row prefetch (measured slower, off)
//   LIDOG_G2_SB / LIDOG_G2_PC   pin the panel-ring depth / 32-channel chunks per operand stage
//   LIDOG_G2_RING    0 = the ring shape of the first validated lean kernel (32 id slots, 3 weight stages when they fit)
//   LIDOG_G2_T       pin the tiles per super-tile (parity tests sweep the multi-tile schedules on small inputs)
//   LIDOG_WG_CTAS / LIDOG_WG_BATCH / LIDOG_WG_SB   wgrad CTA target / MMA-warp batch / dY ring depth
//   LIDOG_WG_STRIDED 0 = contiguous offset groups (round 1)
struct Switches {
  int dbg, acc_sets, opt, force_sb, force_pc, ring_new, force_t, wg_ctas, wg_batch, mma2, wg_sb, wg_na4, wg_sa, wg_dbg, wg_strided;
};
static const Switches& switches() {
  static const Switches sw = {env_int("LIDOG_DBG", 0),     env_int("LIDOG_ACC_SETS", 2), env_int("LIDOG_G2_OPT", 3),
                              env_int("LIDOG_G2_SB", 0),   env_int("LIDOG_G2_PC", 0),    env_int("LIDOG_G2_RING", 1) != 0,
                              env_int("LIDOG_G2_T", 0),    env_int("LIDOG_WG_CTAS", 0),  env_int("LIDOG_WG_BATCH", 4),
                              env_int("LIDOG_G2_MMA2", 0), env_int("LIDOG_WG_SB", 0),
                              env_int("LIDOG_WG_NA4", 1),  env_int("LIDOG_WG_SA", 0),
                              env_int("LIDOG_WG_DBG", 0),  env_int("LIDOG_WG_STRIDED", 1)};
  return sw;
}

// host launcher: forward / dgrad
int launch_gemm_tc2(const lgConvPlan* plan, const void* A16, int Ck, const void* B16, int N, int flip_k, int fmt,
                    const float* out_scale, const float* bias, float* Y, float* stats, cudaStream_t stream) {
  using namespace v2;
  const Switches& sw = switches();
  int sm_count = 0;
  int* err = nullptr;
  int rc = tc_runtime(&sm_count, &err);
  if (rc) return rc;
  Gemm2Args g;
  g.plan = *plan;
  g.A = (const uint16_t*)A16;
  g.Y = Y;
  g.stats = stats;
  g.out_scale = out_scale;
  g.bias = bias;
  g.Ck = Ck;
  g.N = N;
  const int n_split = (N + 255) / 256;
  g.n_blk = N / n_split;
  if (g.n_blk % 16 != 0 || g.n_blk * n_split != N) {
    set_error("lg_conv_gemm_tc: N=%d cannot be split into equal multiples of 16", N);
    return LG_ERR_UNSUPPORTED;
  }
  if (stats && g.n_blk % 32 != 0) {
    set_error("lg_conv_layer_forward: epilogue statistics need column blocks that are multiples of 32 (N=%d)", N);
    return LG_ERR_UNSUPPORTED;
  }
  g.flip = flip_k;
  g.dbg = sw.dbg;
  g.umma_fmt = (fmt == LG_FMT_BF16) ? 1 : 0;
  g.n_tiles = plan->n_slots / LG_TILE_ROWS;
  g.err = err;
  rc = counter_slot(&g.sched);
  if (rc) return rc;
  if (n_split > 4) {
    set_error("lg_conv_gemm_tc: N=%d needs more than 4 column blocks", N);
    return LG_ERR_UNSUPPORTED;
  }
  const int n_chunks = Ck / 32;
  g.n_panels = (n_chunks + 3) / 4;
  while (n_chunks % g.n_panels != 0) ++g.n_panels;
  g.pc = n_chunks / g.n_panels;
  // tiles per super-tile: as many accumulators as TMEM holds (<= 8), but keep >= ~3 super-tiles per SM
  int T = 512 / g.n_blk;
  if (T > 8) T = 8;
  while (T > 1 && ceil_div(g.n_tiles, T) * n_split < 3 * (int64_t)sm_count) --T;
  if (sw.force_t > 0 && sw.force_t < T) T = sw.force_t;
  if (sw.force_t > T && sw.force_t <= 8 && sw.force_t * g.n_blk <= 512) T = sw.force_t;
  // two accumulator sets when that still leaves T >= 1 (LIDOG_ACC_SETS=1 forces the single-set schedule)
  g.sets = 1;
  if (sw.acc_sets == 2 && 2 * g.n_blk <= 512) {
    g.sets = 2;
    const int tmax = 512 / (2 * g.n_blk) < 4 ? 512 / (2 * g.n_blk) : 4;  // 2 * T accumulator barriers <= 8
    if (T > tmax) T = tmax;
  }
  g.T = T;
  g.n_super = ceil_div(g.n_tiles, T);
  // Pipeline shape.  The weight-panel ring gets 3 stages whenever 4 operand stages still fit next to it (with 2,
  // the TMA load of the next panel cannot start before the MMAs of the current one retire: the 256-channel
  // layers waited 22 % of the time for weights); the rest of the budget goes to operand (gather) stages.
  const int opt = sw.opt;
  const int force_sb = sw.force_sb, force_pc = sw.force_pc;
  if (force_pc > 0 && force_pc < g.pc) {
    int pc = force_pc;
    while (n_chunks % pc != 0) --pc;
    g.pc = pc;
    g.n_panels = n_chunks / pc;
  }
  // Ring shape (LIDOG_G2_RING=0 restores 32 id slots / 3 weight stages).  The ncu source view of the lean kernel
  // (profiles/r01_s4_gemm2_source_stalls.txt) shows the MMA warp polling the stage-full barrier ~17 times per unit
  // and the producers waiting for free stages a quarter of the time -- a latency-bound ring -- so shared memory goes
  // to operand stages first: 16 id slots (8 KB instead of 16) and 2 weight stages whenever that buys another stage.
  const int ring_new = sw.ring_new;
  const int ni_slots = ring_new ? 4 * kProdWarps : 8 * kProdWarps;
  size_t stageA, stageB;
  for (;;) {
    stageA = (size_t)g.pc * kSub, stageB = (size_t)g.pc * g.n_blk * kRowB;
    if (force_sb >= 2)
      g.sb = force_sb;
    else if (opt & 2)
      g.sb = (3 * stageB + 4 * stageA + tail_bytes(4, 3, ni_slots) <= kSmemBudget) ? 3 : 2;
    else
      g.sb = (3 * stageB <= 80 * 1024) ? 3 : 2;
    g.sa = 12;
    while (g.sa > 2 && g.sa * stageA + g.sb * stageB + tail_bytes(g.sa, g.sb, ni_slots) > kSmemBudget) --g.sa;
    // would 2 weight stages buy an operand stage?  Only where a weight stage is no bigger than an operand stage:
    // measured (profiles/r01_s4_sweep_n_ring.txt) -17 % / -12 % on the 128-channel layers at tensor stride 8 / 4,
    // but +2 ... +5 % on the 256-channel single-tile layers, whose 32 KB panels want the third stage.
    if (ring_new && force_sb < 2 && g.sb == 3 && g.sa < 12 && stageB <= stageA) {
      int sa2 = 12;
      while (sa2 > 2 && sa2 * stageA + 2 * stageB + tail_bytes(sa2, 2, ni_slots) > kSmemBudget) --sa2;
      if (sa2 > g.sa) {
        g.sa = sa2;
        g.sb = 2;
      }
    }
    if (g.sa >= kProdWarps || g.pc == 1) break;
    // deeper ring with smaller panels (pc must divide the chunk count)
    int pc = g.pc - 1;
    while (n_chunks % pc != 0) --pc;
    g.pc = pc;
    g.n_panels = n_chunks / pc;
  }
  g.bmax = 0;  // (unused since the MMA warp takes all units of an (offset, panel) step together)
  g.pf = (opt & 4) ? 1 : 0;
  if (g.T > g.sa) {  // the MMA warp waits for up to T stages at once: they must all fit in the ring
    g.T = g.sa;
    g.n_super = ceil_div(g.n_tiles, g.T);
  }
  // Ring-phase rule: a producer warp revisits a stage only after the consumer freed it once, which the
  // parity wait can tell only when consecutive units of one warp are < one ring wrap apart: np <= sa.
  g.np = g.sa < kProdWarps ? g.sa : kProdWarps;
  g.ni = ni_slots;
  const size_t smem = g.sa * stageA + g.sb * stageB + tail_bytes(g.sa, g.sb, ni_slots);
  if (smem > kSmemBudget) {
    set_error("lg_conv_gemm_tc: shared memory %zu exceeds the budget (Ck=%d N=%d)", smem, Ck, N);
    return LG_ERR_UNSUPPORTED;
  }
  CUtensorMap tmB;
  memset(&tmB, 0, sizeof(tmB));
  rc = tc_make_tmap(&tmB, B16, (int64_t)plan->kernel_volume * N, Ck, g.n_blk);
  if (rc) return rc;
  dim3 grid((unsigned)(g.n_super < sm_count ? g.n_super : sm_count), (unsigned)n_split);
  // One production instantiation per stage width.  The two-issuer modes (SPLIT 1 / 2) are EXPERIMENTAL and compiled
  // only with -DLIDOG_EXPERIMENTAL (then LIDOG_G2_MMA2=1 selects them): measured <= 6 % where they ran, and the
  // tile-parity mode trapped on one shape before the every-phase wait fix, which is unverified on hardware.
  int split = 0;
#ifdef LIDOG_EXPERIMENTAL
  if (sw.mma2) split = g.T >= 2 ? 1 : ((g.n_blk % 32 == 0) ? 2 : 0);
#endif
#define LG_LAUNCH_GEMM2(DBGV, PCV, SPV)                                                                                   \
  do {                                                                                                                     \
    LG_CUDA_OK(cudaFuncSetAttribute(k_gemm2<DBGV, PCV, SPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    k_gemm2<DBGV, PCV, SPV><<<grid, kThreadsG2, smem, stream>>>(g, tmB);                                                   \
  } while (0)
#ifdef LIDOG_EXPERIMENTAL
#define LG_LAUNCH_GEMM2_SP(DBGV, PCV)                                                                                      \
  switch (split) {                                                                                                         \
    case 0: LG_LAUNCH_GEMM2(DBGV, PCV, 0); break;                                                                          \
    case 1: LG_LAUNCH_GEMM2(DBGV, PCV, 1); break;                                                                          \
    default: LG_LAUNCH_GEMM2(DBGV, PCV, 2); break;                                                                         \
  }
#else
#define LG_LAUNCH_GEMM2_SP(DBGV, PCV) LG_LAUNCH_GEMM2(DBGV, PCV, 0);
#endif
#define LG_LAUNCH_GEMM2_PC(DBGV)                                                                                           \
  switch (g.pc) {                                                                                                          \
    case 1: LG_LAUNCH_GEMM2_SP(DBGV, 1); break;                                                                            \
    case 2: LG_LAUNCH_GEMM2_SP(DBGV, 2); break;                                                                            \
    case 3: LG_LAUNCH_GEMM2_SP(DBGV, 3); break;                                                                            \
    default: LG_LAUNCH_GEMM2_SP(DBGV, 4); break;                                                                           \
  }
  (void)split;
  if (g.dbg) {
    LG_LAUNCH_GEMM2_PC(true)
  } else {
    LG_LAUNCH_GEMM2_PC(false)
  }
#undef LG_LAUNCH_GEMM2_PC
#undef LG_LAUNCH_GEMM2_SP
#undef LG_LAUNCH_GEMM2
  LG_LAUNCH_OK();
  return LG_OK;
}

static int wgrad2_shape(const lgConvPlan* plan, int Cin, int Cout, int* G, int* n_groups, int* tiles_per_chunk) {
  const int K = plan->kernel_volume;
  int gmax = 512 / Cout;
  if (gmax > 8) gmax = 8;
  if (gmax > K) gmax = K;
  *n_groups = (K + gmax - 1) / gmax;
  *G = (K + *n_groups - 1) / *n_groups;
  const int m_blocks = (Cin + 127) / 128;
  const int64_t n_tiles = plan->n_slots / LG_TILE_ROWS;
  // Chunks of tiles per (offset group, channel block).  Every chunk costs one partial dW (written by the epilogue,
  // re-read by the reduction) and one pipeline ramp, so a chunk should own >= ~32 tiles; but the launch wants at
  // least one wave of 148 CTAs and gains nothing beyond four.  Measured (profiles/r01_s4_sweep_a.txt): the layers at
  // tensor stride >= 4 ran 1.3-1.8x faster with 148 CTAs than with 592, the stride-1 layers 1.3x slower.
  const int target_ctas = switches().wg_ctas;
  const int64_t per = (int64_t)*n_groups * m_blocks;
  int64_t want;
  if (target_ctas > 0) {
    want = target_ctas / per;
  } else {
    const int64_t hi = 592 / per, lo = ceil_div((int64_t)148, per);
    want = n_tiles / 32;
    if (want < lo) want = lo;
    if (want > hi) want = hi;
  }
  if (want < 1) want = 1;
  int64_t chunks = n_tiles < want ? n_tiles : want;
  if (chunks < 1) chunks = 1;
  *tiles_per_chunk = (int)ceil_div(n_tiles > 0 ? n_tiles : 1, chunks);
  return (int)ceil_div(n_tiles > 0 ? n_tiles : 1, *tiles_per_chunk);
}

size_t wgrad_tc2_workspace(const lgConvPlan* plan, int Cin, int Cout) {
  int G, ng, tpc;
  const int chunks = wgrad2_shape(plan, Cin, Cout, &G, &ng, &tpc);
  return (size_t)chunks * plan->kernel_volume * Cin * Cout * sizeof(float) + 256;
}

int launch_wgrad_tc2(const lgConvPlan* plan, const void* X16, int Cin, const void* dY16, int Cout, int fmt,
                     const float* out_scale, float* dW, void* workspace, cudaStream_t stream) {
  using namespace v2;
  int sm_count = 0;
  int* err = nullptr;
  int rc = tc_runtime(&sm_count, &err);
  if (rc) return rc;
  Wgrad2Args g;
  g.plan = *plan;
  g.X = (const uint16_t*)X16;
  g.dY = (const uint16_t*)dY16;
  g.partial = (float*)workspace;
  g.Cin = Cin;
  g.Cout = Cout;
  g.m_blocks = (Cin + 127) / 128;
  const int chunks = wgrad2_shape(plan, Cin, Cout, &g.G, &g.n_groups, &g.tiles_per_chunk);
  g.umma_fmt = (fmt == LG_FMT_BF16) ? 1 : 0;
  g.n_tiles = plan->n_slots / LG_TILE_ROWS;
  g.err = err;
  // An X stage always holds 4 sub-blocks (32 KB) although a 96-channel layer fills 3 and a 32-channel layer 1.
  // Packing the stages tighter (LIDOG_WG_NA4=0) is WRONG: the M = 128 MMA then reads its unused sub-blocks out of
  // the neighbouring stages, and the full-scale parity test (tests/test_gpu_fullscale.py) measured 0.3-2.5 % error in
  // dW -- bisected on hardware (profiles/r02_b_wgrad_bisect.log); the switch stays for that record only.
  g.na_max = switches().wg_na4 ? 4 : (Cin / 32 < 4 ? Cin / 32 : 4);
  // strided groups where there are enough groups for the imbalance to matter: measured -7 ... -20 % on the 96- / 128- /
  // 256-channel layers (5 ... 14 groups), +15 % on the 32-channel ones (4 groups of 7 offsets, already balanced, and
  // every extra dY tile load is pure HBM cost there) -- profiles/r02_d_layer_table_strided.txt
  g.strided = (switches().wg_strided && g.n_groups >= 5) ? 1 : 0;
  if (plan->kernel_volume > 64) {
    set_error("lg_conv_wgrad_tc: kernel volume %d > 64 is not on the tensor-core path", plan->kernel_volume);
    return LG_ERR_UNSUPPORTED;
  }
  const size_t stageA = (size_t)g.na_max * kSub, stageB = (size_t)(Cout / 32) * kSub;
  // dY ring: 3 stages when 4 X stages still fit next to them (the MMA warp waited on this ring, see the dY producer)
  g.sb = (4 * stageA + 3 * stageB + wgrad_tail_bytes(4, 3) <= kSmemBudget) ? 3 : 2;
  if (switches().wg_sb >= 2) g.sb = switches().wg_sb;
  g.sa = switches().wg_sa >= 2 ? switches().wg_sa : 6;
  while (g.sa > 2 && g.sa * stageA + g.sb * stageB + wgrad_tail_bytes(g.sa, g.sb) > kSmemBudget) --g.sa;
  g.np = g.sa < kProdWarps ? g.sa : kProdWarps;
  {
    const int want = switches().wg_batch;
    g.bmax = want < 1 ? 1 : (want < g.sa ? want : g.sa);
  }
  const size_t smem = g.sa * stageA + g.sb * stageB + wgrad_tail_bytes(g.sa, g.sb);
  if (smem > kSmemBudget) {
    set_error("lg_conv_wgrad_tc: shared memory %zu exceeds the budget (Cin=%d Cout=%d)", smem, Cin, Cout);
    return LG_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)(g.n_groups * g.m_blocks), (unsigned)chunks);
  g.dbg = switches().wg_dbg;
  if (g.dbg) {
    LG_CUDA_OK(cudaFuncSetAttribute(k_wgrad2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_wgrad2<true><<<grid, kThreads, smem, stream>>>(g);
  } else {
    LG_CUDA_OK(cudaFuncSetAttribute(k_wgrad2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_wgrad2<false><<<grid, kThreads, smem, stream>>>(g);
  }
  LG_LAUNCH_OK();
  return launch_reduce_partials((const float*)workspace, chunks, (int64_t)plan->kernel_volume * Cin * Cout, out_scale,
                                dW, stream);
}

static int check_tc_plan(const lgConvPlan* p, const char* who) {
  LG_CHECK_ARG(p != nullptr, "%s: null plan", who);
  LG_CHECK_ARG(p->n_slots > 0 && p->n_slots % LG_TILE_ROWS == 0, "%s: n_slots must be a positive multiple of 128", who);
  LG_CHECK_ARG(p->kernel_volume >= 1 && p->kernel_volume <= 128 && p->mask_words == (p->kernel_volume + 31) / 32,
               "%s: bad kernel_volume/mask_words", who);
  LG_CHECK_ARG(p->k_stride == 0 || p->k_stride == p->n_slots, "%s: k_stride must be 0 or n_slots", who);
  LG_CHECK_ARG(p->nbr && p->tile_mask, "%s: null plan arrays", who);
  LG_CHECK_ARG(p->n_in < ((int64_t)1 << 31) && p->n_out < ((int64_t)1 << 31), "%s: more than 2^31 rows", who);
  return LG_OK;
}

static int check_gemm(const lgConvPlan* plan, int Ck, int N, int fmt, const char* who) {
  int rc = check_tc_plan(plan, who);
  if (rc) return rc;
  if (Ck % 32 != 0 || N % 16 != 0 || N < 16 || N > 512) {
    set_error("%s: Ck=%d must be a multiple of 32 and N=%d a multiple of 16 in [16,512]", who, Ck, N);
    return LG_ERR_UNSUPPORTED;
  }
  // the gather kernels keep row byte offsets in 32 bits
  if ((uint64_t)plan->n_in * (uint64_t)Ck * 2ull >= ((uint64_t)1 << 32)) {
    set_error("%s: operand matrix of %lld x %d 16-bit elements exceeds the 4 GiB row-offset range", who,
              (long long)plan->n_in, Ck);
    return LG_ERR_UNSUPPORTED;
  }
  LG_CHECK_ARG(fmt == LG_FMT_BF16 || fmt == LG_FMT_FP16, "%s: bad format", who);
  return LG_OK;
}

static int check_wgrad(const lgConvPlan* plan, int Cin, int Cout, int fmt, const char* who) {
  int rc = check_tc_plan(plan, who);
  if (rc) return rc;
  if (Cin % 32 != 0 || Cout % 32 != 0 || Cout > 256 || Cin < 32) {
    set_error("%s: Cin=%d, Cout=%d must be multiples of 32 with Cout <= 256", who, Cin, Cout);
    return LG_ERR_UNSUPPORTED;
  }
  if ((uint64_t)plan->n_in * (uint64_t)Cin * 2ull >= ((uint64_t)1 << 32) ||
      (uint64_t)plan->n_out * (uint64_t)Cout * 2ull >= ((uint64_t)1 << 32)) {
    set_error("%s: operand matrix exceeds the 4 GiB row-offset range", who);
    return LG_ERR_UNSUPPORTED;
  }
  LG_CHECK_ARG(fmt == LG_FMT_BF16 || fmt == LG_FMT_FP16, "%s: bad format", who);
  return LG_OK;
}

}  // namespace lg

using namespace lg;

extern "C" int lg_conv_gemm_tc(const lgConvPlan* plan, const void* A16, int32_t Ck, const void* B16, int32_t N,
                               int32_t flip_k, int32_t fmt, const float* out_scale, const float* bias, float* Y,
                               int32_t gather_mode, void* stream_) {
  int rc = check_gemm(plan, Ck, N, fmt, "lg_conv_gemm_tc");
  if (rc) return rc;
  LG_CHECK_ARG(A16 && B16 && Y, "lg_conv_gemm_tc: null pointer");
  if (gather_mode != 2) {
    set_error("lg_conv_gemm_tc: gather_mode %d is not in the library (the first-generation kernels live in "
              "tools/legacy/conv_tc_gen1.cu as the record)", gather_mode);
    return LG_ERR_UNSUPPORTED;
  }
  return launch_gemm_tc2(plan, A16, Ck, B16, N, flip_k, fmt, out_scale, bias, Y, nullptr, (cudaStream_t)stream_);
}

extern "C" size_t lg_conv_wgrad_tc_workspace(const lgConvPlan* plan, int32_t Cin, int32_t Cout) {
  if (!plan || Cout < 32 || Cout > 256) return 0;
  return wgrad_tc2_workspace(plan, Cin, Cout);
}

extern "C" int lg_conv_wgrad_tc(const lgConvPlan* plan, const void* X16, int32_t Cin, const void* dY16, int32_t Cout,
                                int32_t fmt, const float* out_scale, float* dW, int32_t gather_mode, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  int rc = check_wgrad(plan, Cin, Cout, fmt, "lg_conv_wgrad_tc");
  if (rc) return rc;
  LG_CHECK_ARG(X16 && dY16 && dW && workspace, "lg_conv_wgrad_tc: null pointer");
  LG_CHECK_ARG(workspace_bytes >= lg_conv_wgrad_tc_workspace(plan, Cin, Cout), "lg_conv_wgrad_tc: workspace too small");
  if (gather_mode != 2) {
    set_error("lg_conv_wgrad_tc: gather_mode %d is not in the library", gather_mode);
    return LG_ERR_UNSUPPORTED;
  }
  return launch_wgrad_tc2(plan, X16, Cin, dY16, Cout, fmt, out_scale, dW, workspace, (cudaStream_t)stream_);
}

/* One MinkowskiConvolution forward on the tensor-core path in one host call (see include/lidog_b200.h). */
extern "C" int lg_conv_layer_forward(const lgConvPlan* plan, const void* X16, int32_t Cin, const float* W, int32_t Cout,
                                     void* w16, void* w16t, int32_t prep, int32_t fmt, const float* bias, float* Y,
                                     float* stat_partials, void* stream_) {
  int rc = check_gemm(plan, Cin, Cout, fmt, "lg_conv_layer_forward");
  if (rc) return rc;
  LG_CHECK_ARG(X16 && w16t && Y && (!prep || W), "lg_conv_layer_forward: null pointer");
  if (prep) {
    rc = lg_prep_weights(W, plan->kernel_volume, Cin, Cout, w16, w16t, fmt, stream_);
    if (rc) return rc;
  }
  return launch_gemm_tc2(plan, X16, Cin, w16t, Cout, 0, fmt, nullptr, bias, Y, stat_partials, (cudaStream_t)stream_);
}

/* dgrad and wgrad of one MinkowskiConvolution in one host call; the split-K partials live in the library arena. */
extern "C" int lg_conv_layer_backward(const lgConvPlan* plan_dgrad, const lgConvPlan* plan_wgrad, int32_t flip_dgrad,
                                      const void* X16, int32_t Cin, const void* dY16, int32_t Cout, const void* w16,
                                      int32_t fmt, const float* inv_scale, float* dX, float* dW, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc;
  LG_CHECK_ARG(dY16, "lg_conv_layer_backward: null gradient");
  if (dX) {
    rc = check_gemm(plan_dgrad, Cout, Cin, fmt, "lg_conv_layer_backward(dgrad)");
    if (rc) return rc;
    LG_CHECK_ARG(w16, "lg_conv_layer_backward: null weights");
    rc = launch_gemm_tc2(plan_dgrad, dY16, Cout, w16, Cin, flip_dgrad, fmt, inv_scale, nullptr, dX, nullptr, stream);
    if (rc) return rc;
  }
  if (dW) {
    rc = check_wgrad(plan_wgrad, Cin, Cout, fmt, "lg_conv_layer_backward(wgrad)");
    if (rc) return rc;
    LG_CHECK_ARG(X16, "lg_conv_layer_backward: null operand");
    ArenaCursor ar;
    const size_t ws = wgrad_tc2_workspace(plan_wgrad, Cin, Cout);
    rc = arena_begin(stream, ws, &ar);
    if (rc) return rc;
    rc = launch_wgrad_tc2(plan_wgrad, X16, Cin, dY16, Cout, fmt, inv_scale, dW, arena_take(&ar, ws), stream);
    if (rc) return rc;
  }
  return LG_OK;
}

// experiment hook (not part of the reference-facing ABI): per-role wait cycles of CTA 0, LIDOG_DBG & 8
extern "C" int lg_debug_profile(long long* out16, int reset) { return lg::debug_profile(out16, reset); }
// event trace of CTA 0 (LIDOG_DBG & 8): out[5][512] clock64 values, see g_trace
extern "C" int lg_debug_trace(long long* out) {
  if (cudaMemcpyFromSymbol(out, lg::v2::g_trace, sizeof(long long) * 5 * lg::v2::kTraceUnits) != cudaSuccess) return LG_ERR_CUDA;
  return LG_OK;
}
