// Sparse convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Forward / dgrad (k_gemm_tc):  Y[out_row[s], :] = sum_k A16[nbr[k][s], :] @ B16[wk]^T
//   output-stationary implicit GEMM: a CTA owns a tile of 128 result rows; for every kernel offset
//   present in the tile (tile_mask) the 128 operand rows are gathered into shared memory
//   (TMA tile::gather4, 4 rows per instruction, missing neighbours = out-of-bounds rows = zero fill;
//   or cp.async with zero-fill), the [N x 32] weight slab arrives by a tiled TMA load, and one elected
//   thread issues tcgen05.mma (M=128, N<=256, K=16, fp16/bf16 -> fp32) into a TMEM accumulator.
//   Two accumulators (2 x 256 TMEM columns) let the epilogue of tile t overlap the MMAs of tile t+1.
//   Every result row is written exactly once by its owner: deterministic, no atomics, no memset.
// Wgrad (k_wgrad_tc):  dW[k] = sum_s X16[nbr[k][s], :]^T dY16[out_row[s], :]
//   both operands are MN-major (the reduction runs over the gathered rows); a CTA owns
//   (k, 128-channel block of Cin, chunk of tiles) and keeps the [128 x Cout] fp32 accumulator in TMEM;
//   partials are reduced in a fixed order by a second kernel (deterministic split-K).
// Shared-memory operand tiles use 64-byte rows (32 x 16-bit) in the SWIZZLE_64B canonical layout, so
// all LiDOG channel counts (multiples of 32, including 96 and 192) need no padding.
// Reference contract: MinkowskiConvolution / MinkowskiConvolutionTranspose forward + backward,
// utils/models/minkunet_bev.py:57-123 (fp32 SIMT gather-GEMM-scatter in MinkowskiEngine 0.5.4).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lg {

int launch_reduce_partials(const float* partial, int n_chunks, int64_t n_elems, const float* out_scale, float* dW,
                           cudaStream_t stream);
// second-generation pipeline (conv_tc2.cu), gather_mode 2
int launch_gemm_tc2(const lgConvPlan* plan, const void* A16, int Ck, const void* B16, int N, int flip_k, int fmt,
                    const float* out_scale, const float* bias, float* Y, cudaStream_t stream);
size_t wgrad_tc2_workspace(const lgConvPlan* plan, int Cin, int Cout);
int launch_wgrad_tc2(const lgConvPlan* plan, const void* X16, int Cin, const void* dY16, int Cout, int fmt,
                     const float* out_scale, float* dW, void* workspace, cudaStream_t stream);

using namespace ptx;

constexpr int kRowBytes = 64;    // one smem operand row = 32 x 16-bit
constexpr int kChunkElems = 32;  // K elements per pipeline stage
constexpr int kEpiWarps = 4;     // warps 0..3 (TMEM lane quadrant = warp id)
constexpr int kMmaWarp = 4;
constexpr int kProdWarp0 = 5;
constexpr int kProdWarps = 4;  // warps 5..8 (only warp 5 works in TMA mode)
constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;
constexpr int kAccStride = 256;  // TMEM columns per accumulator buffer

// physical byte offset of 16-byte chunk j of row r inside a SWIZZLE_64B tile (64-byte rows)
__device__ __forceinline__ uint32_t sw64(uint32_t r, uint32_t j) { return r * kRowBytes + ((j ^ ((r >> 1) & 3u)) << 4); }

struct GemmArgs {
  lgConvPlan plan;
  const uint16_t* A;  // [n_in][Ck]
  const uint16_t* B;  // [K][N][Ck]
  float* Y;           // [n_out][N]
  const float* out_scale;
  const float* bias;
  int Ck, N, n_blk, flip, umma_fmt, use_tma, stages;
  int64_t n_tiles;
  int* err;
};

__device__ __forceinline__ bool tile_has(const lgConvPlan& p, int64_t tile, int k) {
  return (p.tile_mask[tile * p.mask_words + (k >> 5)] >> (k & 31)) & 1u;
}
__device__ __forceinline__ bool tile_any(const lgConvPlan& p, int64_t tile) {
  uint32_t m = 0;
  for (int w = 0; w < p.mask_words; ++w) m |= p.tile_mask[tile * p.mask_words + w];
  return m != 0;
}

__global__ void __launch_bounds__(kThreads, 1)
    k_gemm_tc(const GemmArgs g, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = LG_TILE_ROWS * kRowBytes, b_bytes = g.n_blk * kRowBytes;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* full = (uint64_t*)(smem + (size_t)g.stages * stage_bytes);
  uint64_t* empty = full + g.stages;
  uint64_t* acc_full = empty + g.stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
  const int n0 = blockIdx.y * g.n_blk;
  const int n_chunks = g.Ck / kChunkElems;
  const int K = g.plan.kernel_volume;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full[s], g.use_tma ? 1 : kProdWarps * 32);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  if (warp == kProdWarp0 && lane == 0 && g.use_tma) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kProdWarp0) {
    // ===================================================================== producers
    int stage = 0;
    uint32_t phase = 0;
    if (g.use_tma) {
      if (warp == kProdWarp0) {
        for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
          for (int k = 0; k < K; ++k) {
            if (!tile_has(g.plan, tile, k)) continue;
            const int wk = g.flip ? (K - 1 - k) : k;
            int4 idx = *reinterpret_cast<const int4*>(g.plan.nbr + (int64_t)k * g.plan.k_stride + tile * LG_TILE_ROWS +
                                                      4 * lane);
            const int oob = (int)g.plan.n_in;  // rows >= n_in are out of bounds -> TMA zero fill
            idx.x = idx.x < 0 ? oob : idx.x;
            idx.y = idx.y < 0 ? oob : idx.y;
            idx.z = idx.z < 0 ? oob : idx.z;
            idx.w = idx.w < 0 ? oob : idx.w;
            for (int c = 0; c < n_chunks; ++c) {
              if (lane == 0) {
                mbar_wait(&empty[stage], phase ^ 1, g.err, 1);
                mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
              }
              __syncwarp();
              uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
              tma_gather4(a_dst + lane * 4 * kRowBytes, &tmA, c * kChunkElems, idx.x, idx.y, idx.z, idx.w,
                          &full[stage]);
              if (lane == 0) tma_load_2d(a_dst + a_bytes, &tmB, c * kChunkElems, wk * g.N + n0, &full[stage]);
              if (++stage == g.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    } else {
      const int t = threadIdx.x - kProdWarp0 * 32;  // 0..127 = tile row
      int pending = -1;
      for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        for (int k = 0; k < K; ++k) {
          if (!tile_has(g.plan, tile, k)) continue;
          const int wk = g.flip ? (K - 1 - k) : k;
          const int row = g.plan.nbr[(int64_t)k * g.plan.k_stride + tile * LG_TILE_ROWS + t];
          const uint16_t* a_src = g.A + (int64_t)(row < 0 ? 0 : row) * g.Ck;
          const uint16_t* b_src = g.B + ((int64_t)wk * g.N + n0) * g.Ck;
          for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&empty[stage], phase ^ 1, g.err, 1);
            uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
            uint8_t* b_dst = a_dst + a_bytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) cp_async16(a_dst + sw64(t, j), a_src + c * kChunkElems + j * 8, row >= 0);
            for (int e = t; e < g.n_blk * 4; e += kProdWarps * 32) {
              const int r = e >> 2, j = e & 3;
              cp_async16(b_dst + sw64(r, j), b_src + (int64_t)r * g.Ck + c * kChunkElems + j * 8, true);
            }
            cp_async_commit();
            if (pending >= 0) {  // complete the previous stage while this one is in flight
              cp_async_wait<1>();
              fence_proxy_async();
              mbar_arrive(&full[pending]);
            }
            pending = stage;
            if (++stage == g.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      if (pending >= 0) {
        cp_async_wait<0>();
        fence_proxy_async();
        mbar_arrive(&full[pending]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================================================== MMA issuer
    const uint32_t idesc = make_idesc(g.umma_fmt, 0, 0, LG_TILE_ROWS, g.n_blk);
    int stage = 0;
    uint32_t phase = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
      if (!tile_any(g.plan, tile)) continue;
      const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
      ++it;
      mbar_wait(&acc_empty[buf], acc_phase ^ 1, g.err, 2);
      tc_fence_after();
      uint32_t accumulate = 0;
      for (int k = 0; k < K; ++k) {
        if (!tile_has(g.plan, tile, k)) continue;
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&full[stage], phase, g.err, 3);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {  // two K=16 steps inside the 64-byte row
              const uint64_t da = make_smem_desc(a_addr + kk * 32, 16, 8 * kRowBytes, kLayoutSw64);
              const uint64_t db = make_smem_desc(b_addr + kk * 32, 16, 8 * kRowBytes, kLayoutSw64);
              umma_f16(tmem_base + buf * kAccStride, da, db, idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(&empty[stage]);  // frees the smem stage when these MMAs retire
          }
          __syncwarp();
          if (++stage == g.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (lane == 0) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ===================================================================== epilogue (warps 0..3)
    const float scale = g.out_scale ? g.out_scale[0] : 1.f;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
      const int64_t s = tile * LG_TILE_ROWS + warp * 32 + lane;
      int64_t row = g.plan.out_row ? (int64_t)g.plan.out_row[s] : s;
      const bool row_ok = row >= 0 && row < g.plan.n_out;
      float* yrow = g.Y + (row_ok ? row : 0) * g.N + n0;
      if (!tile_any(g.plan, tile)) {  // no neighbour at all: bias / zeros
        if (row_ok)
          for (int n = 0; n < g.n_blk; ++n) yrow[n] = g.bias ? g.bias[n0 + n] : 0.f;
        continue;
      }
      const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
      ++it;
      mbar_wait(&acc_full[buf], acc_phase, g.err, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * kAccStride + ((uint32_t)(warp * 32) << 16);
      for (int n = 0; n < g.n_blk; n += 16) {
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
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------ wgrad
constexpr int kWgPairs = 64;  // gathered rows (reduction length) per pipeline stage
constexpr int kWgSub = kWgPairs * kRowBytes;  // one 32-channel sub-block of a stage = 4 KB
constexpr int kWgStages = 4;

struct WgradArgs {
  lgConvPlan plan;
  const uint16_t* X;   // [n_in][Cin]   gathered through nbr
  const uint16_t* dY;  // [n_out][Cout] read at out_row[s] (or s)
  float* partial;      // [chunks][K][Cin][Cout]
  int Cin, Cout, m_blocks, tiles_per_chunk, umma_fmt, use_tma, tmem_cols;
  int64_t n_tiles;
  int* err;
};

__global__ void __launch_bounds__(kThreads, 1)
    k_wgrad_tc(const WgradArgs g, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x / g.m_blocks, mb = blockIdx.x % g.m_blocks;
  const int chunk = blockIdx.y;
  const int m0 = mb * 128;
  const int na = min(4, (g.Cin - m0) / kChunkElems);  // real 32-channel sub-blocks of the A operand
  const int nb = g.Cout / kChunkElems;
  const int a_bytes = 4 * kWgSub, stage_bytes = a_bytes + nb * kWgSub;
  uint64_t* full = (uint64_t*)(smem + (size_t)kWgStages * stage_bytes);
  uint64_t* empty = full + kWgStages;
  uint64_t* done = empty + kWgStages;
  uint32_t* tmem_slot = (uint32_t*)(done + 1);
  const int64_t t0 = (int64_t)chunk * g.tiles_per_chunk;
  const int64_t t1 = (t0 + g.tiles_per_chunk < g.n_tiles) ? t0 + g.tiles_per_chunk : g.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full[s], g.use_tma ? 1 : kProdWarps * 32);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  bool any = false;
  for (int64_t tile = t0; tile < t1; ++tile) any |= tile_has(g.plan, tile, k);

  if (warp >= kProdWarp0) {
    int stage = 0;
    uint32_t phase = 0;
    int pending = -1;
    const int t = threadIdx.x - kProdWarp0 * 32;
    for (int64_t tile = t0; tile < t1; ++tile) {
      if (!tile_has(g.plan, tile, k)) continue;
      for (int half = 0; half < LG_TILE_ROWS / kWgPairs; ++half) {
        const int64_t s0 = tile * LG_TILE_ROWS + half * kWgPairs;
        const int32_t* ia = g.plan.nbr + (int64_t)k * g.plan.k_stride + s0;
        uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
        uint8_t* b_dst = a_dst + a_bytes;
        if (g.use_tma) {
          if (warp == kProdWarp0) {
            if (lane == 0) {
              mbar_wait(&empty[stage], phase ^ 1, g.err, 11);
              mbar_arrive_expect_tx(&full[stage], (uint32_t)((na + nb) * kWgSub));
            }
            __syncwarp();
            for (int e = lane; e < (na + nb) * 16; e += 32) {
              const int sb = e >> 4, grp = e & 15;  // sub-block, group of 4 rows
              int4 idx;
              if (sb < na) {
                idx = *reinterpret_cast<const int4*>(ia + 4 * grp);
                const int oob = (int)g.plan.n_in;
                idx.x = idx.x < 0 ? oob : idx.x, idx.y = idx.y < 0 ? oob : idx.y;
                idx.z = idx.z < 0 ? oob : idx.z, idx.w = idx.w < 0 ? oob : idx.w;
                tma_gather4(a_dst + sb * kWgSub + grp * 4 * kRowBytes, &tmX, m0 + sb * kChunkElems, idx.x, idx.y,
                            idx.z, idx.w, &full[stage]);
              } else {
                const int oob = (int)g.plan.n_out;
                if (g.plan.out_row) {
                  idx = *reinterpret_cast<const int4*>(g.plan.out_row + s0 + 4 * grp);
                  idx.x = idx.x < 0 ? oob : idx.x, idx.y = idx.y < 0 ? oob : idx.y;
                  idx.z = idx.z < 0 ? oob : idx.z, idx.w = idx.w < 0 ? oob : idx.w;
                } else {
                  const int r = (int)(s0 + 4 * grp);
                  idx = make_int4(min(r, oob), min(r + 1, oob), min(r + 2, oob), min(r + 3, oob));
                }
                tma_gather4(b_dst + (sb - na) * kWgSub + grp * 4 * kRowBytes, &tmD, (sb - na) * kChunkElems, idx.x,
                            idx.y, idx.z, idx.w, &full[stage]);
              }
            }
          }
        } else {
          mbar_wait(&empty[stage], phase ^ 1, g.err, 11);
          for (int e = t; e < (na + nb) * kWgPairs * 4; e += kProdWarps * 32) {
            const int sb = e / (kWgPairs * 4), rem = e % (kWgPairs * 4);
            const int r = rem >> 2, j = rem & 3;
            if (sb < na) {
              const int row = ia[r];
              cp_async16(a_dst + sb * kWgSub + sw64(r, j),
                         g.X + (int64_t)(row < 0 ? 0 : row) * g.Cin + m0 + sb * kChunkElems + j * 8, row >= 0);
            } else {
              int64_t row = g.plan.out_row ? (int64_t)g.plan.out_row[s0 + r] : s0 + r;
              const bool ok = row >= 0 && row < g.plan.n_out;
              cp_async16(b_dst + (sb - na) * kWgSub + sw64(r, j),
                         g.dY + (ok ? row : 0) * g.Cout + (sb - na) * kChunkElems + j * 8, ok);
            }
          }
          cp_async_commit();
          if (pending >= 0) {
            cp_async_wait<1>();
            fence_proxy_async();
            mbar_arrive(&full[pending]);
          }
          pending = stage;
        }
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if (!g.use_tma && pending >= 0) {
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(&full[pending]);
    }
  } else if (warp == kMmaWarp) {
    const uint32_t idesc = make_idesc(g.umma_fmt, 1, 1, 128, g.Cout);
    int stage = 0;
    uint32_t phase = 0, accumulate = 0;
    for (int64_t tile = t0; tile < t1; ++tile) {
      if (!tile_has(g.plan, tile, k)) continue;
      for (int half = 0; half < LG_TILE_ROWS / kWgPairs; ++half) {
        mbar_wait(&full[stage], phase, g.err, 13);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
          for (int j = 0; j < kWgPairs / 16; ++j) {  // K = 16 gathered rows per MMA
            // MN-major SWIZZLE_64B: LBO = pitch between 32-channel sub-blocks, SBO = 8 rows * 64 B
            const uint64_t da = make_smem_desc(a_addr + j * 16 * kRowBytes, kWgSub, 8 * kRowBytes, kLayoutSw64);
            const uint64_t db = make_smem_desc(b_addr + j * 16 * kRowBytes, kWgSub, 8 * kRowBytes, kLayoutSw64);
            umma_f16(tmem_base, da, db, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if (lane == 0 && any) umma_commit(done);
    __syncwarp();
  } else {
    // epilogue: lane = input channel m0 + warp*32 + lane, columns = output channels
    const int ci = m0 + warp * 32 + lane;
    float* prow = g.partial + (((int64_t)chunk * g.plan.kernel_volume + k) * g.Cin + (ci < g.Cin ? ci : 0)) * g.Cout;
    if (any) {
      mbar_wait(done, 0, g.err, 14);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int n = 0; n < g.Cout; n += 16) {
      uint32_t v[16];
      if (any) {
        tmem_ld16(taddr + n, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = 0u;
      }
      if (ci < g.Cin) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(prow + n + 4 * q) =
              make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                          __uint_as_float(v[4 * q + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, g.tmem_cols);
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2D row-major 16-bit matrix [rows][cols]; box = box_rows x 32 columns, SWIZZLE_64B
static int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return LG_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kChunkElems, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d", (int)r, (long long)rows,
              (long long)cols, box_rows);
    return LG_ERR_CUDA;
  }
  return LG_OK;
}

static int g_sm_count = 0;
static int* g_err_word = nullptr;

static int tc_runtime_init() {
  if (!g_sm_count) {
    int dev = 0;
    LG_CUDA_OK(cudaGetDevice(&dev));
    LG_CUDA_OK(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  if (!g_err_word) {
    LG_CUDA_OK(cudaMalloc(&g_err_word, sizeof(int)));
    LG_CUDA_OK(cudaMemset(g_err_word, 0, sizeof(int)));
  }
  return LG_OK;
}

int tc_make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int box_rows) {
  return make_tmap(m, base, rows, cols, box_rows);
}
int tc_runtime(int* sm_count, int** err_word) {
  int rc = tc_runtime_init();
  if (rc) return rc;
  *sm_count = g_sm_count;
  *err_word = g_err_word;
  return LG_OK;
}

int wgrad_tc_chunks(const lgConvPlan* plan, int m_blocks, int* tiles_per_chunk) {
  const int64_t n_tiles = plan->n_slots / LG_TILE_ROWS;
  int64_t want = 592 / ((int64_t)plan->kernel_volume * m_blocks);
  if (want < 1) want = 1;
  int64_t chunks = n_tiles < want ? n_tiles : want;
  if (chunks < 1) chunks = 1;
  *tiles_per_chunk = (int)ceil_div(n_tiles > 0 ? n_tiles : 1, chunks);
  chunks = ceil_div(n_tiles > 0 ? n_tiles : 1, *tiles_per_chunk);
  return (int)chunks;
}

}  // namespace lg

using namespace lg;

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

extern "C" int lg_conv_gemm_tc(const lgConvPlan* plan, const void* A16, int32_t Ck, const void* B16, int32_t N,
                               int32_t flip_k, int32_t fmt, const float* out_scale, const float* bias, float* Y,
                               int32_t gather_mode, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_tc_plan(plan, "lg_conv_gemm_tc");
  if (rc) return rc;
  if (Ck % kChunkElems != 0 || N % 16 != 0 || N < 16 || N > 512) {
    set_error("lg_conv_gemm_tc: Ck=%d must be a multiple of 32 and N=%d a multiple of 16 in [16,512]", Ck, N);
    return LG_ERR_UNSUPPORTED;
  }
  LG_CHECK_ARG(fmt == LG_FMT_BF16 || fmt == LG_FMT_FP16, "lg_conv_gemm_tc: bad format");
  LG_CHECK_ARG(A16 && B16 && Y, "lg_conv_gemm_tc: null pointer");
  if (gather_mode == 2)
    return launch_gemm_tc2(plan, A16, Ck, B16, N, flip_k, fmt, out_scale, bias, Y, stream);
  rc = tc_runtime_init();
  if (rc) return rc;
  GemmArgs g;
  g.plan = *plan;
  g.A = (const uint16_t*)A16;
  g.B = (const uint16_t*)B16;
  g.Y = Y;
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
  g.flip = flip_k;
  g.umma_fmt = (fmt == LG_FMT_BF16) ? 1 : 0;
  g.use_tma = gather_mode & 1;
  g.n_tiles = plan->n_slots / LG_TILE_ROWS;
  g.err = g_err_word;
  const int stage_bytes = LG_TILE_ROWS * kRowBytes + g.n_blk * kRowBytes;
  g.stages = 8;
  while (g.stages > 2 && (size_t)g.stages * stage_bytes > 180 * 1024) --g.stages;
  const size_t smem = (size_t)g.stages * stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
  CUtensorMap tmA, tmB;
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmB, 0, sizeof(tmB));
  if (g.use_tma) {
    rc = make_tmap(&tmA, A16, plan->n_in, Ck, 1);
    if (rc) return rc;
    rc = make_tmap(&tmB, B16, (int64_t)plan->kernel_volume * N, Ck, g.n_blk);
    if (rc) return rc;
  }
  LG_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(g.n_tiles < g_sm_count ? g.n_tiles : g_sm_count), (unsigned)n_split);
  k_gemm_tc<<<grid, kThreads, smem, stream>>>(g, tmA, tmB);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" size_t lg_conv_wgrad_tc_workspace(const lgConvPlan* plan, int32_t Cin, int32_t Cout) {
  if (!plan) return 0;
  int tpc;
  const int chunks = wgrad_tc_chunks(plan, (Cin + 127) / 128, &tpc);
  const size_t v1 = (size_t)chunks * plan->kernel_volume * Cin * Cout * sizeof(float) + 256;
  const size_t v2 = (Cout >= 32 && Cout <= 256) ? wgrad_tc2_workspace(plan, Cin, Cout) : 0;
  return v1 > v2 ? v1 : v2;
}

extern "C" int lg_conv_wgrad_tc(const lgConvPlan* plan, const void* X16, int32_t Cin, const void* dY16, int32_t Cout,
                                int32_t fmt, const float* out_scale, float* dW, int32_t gather_mode, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_tc_plan(plan, "lg_conv_wgrad_tc");
  if (rc) return rc;
  if (Cin % kChunkElems != 0 || Cout % kChunkElems != 0 || Cout > 256 || Cin < 32) {
    set_error("lg_conv_wgrad_tc: Cin=%d, Cout=%d must be multiples of 32 with Cout <= 256", Cin, Cout);
    return LG_ERR_UNSUPPORTED;
  }
  LG_CHECK_ARG(fmt == LG_FMT_BF16 || fmt == LG_FMT_FP16, "lg_conv_wgrad_tc: bad format");
  LG_CHECK_ARG(X16 && dY16 && dW && workspace, "lg_conv_wgrad_tc: null pointer");
  LG_CHECK_ARG(workspace_bytes >= lg_conv_wgrad_tc_workspace(plan, Cin, Cout), "lg_conv_wgrad_tc: workspace too small");
  if (gather_mode == 2) return launch_wgrad_tc2(plan, X16, Cin, dY16, Cout, fmt, out_scale, dW, workspace, stream);
  rc = tc_runtime_init();
  if (rc) return rc;
  WgradArgs g;
  g.plan = *plan;
  g.X = (const uint16_t*)X16;
  g.dY = (const uint16_t*)dY16;
  g.partial = (float*)workspace;
  g.Cin = Cin;
  g.Cout = Cout;
  g.m_blocks = (Cin + 127) / 128;
  const int chunks = wgrad_tc_chunks(plan, g.m_blocks, &g.tiles_per_chunk);
  g.umma_fmt = (fmt == LG_FMT_BF16) ? 1 : 0;
  g.use_tma = gather_mode & 1;
  g.tmem_cols = 32;
  while (g.tmem_cols < Cout) g.tmem_cols <<= 1;
  g.n_tiles = plan->n_slots / LG_TILE_ROWS;
  g.err = g_err_word;
  const int stage_bytes = 4 * kWgSub + (Cout / kChunkElems) * kWgSub;
  const size_t smem = (size_t)kWgStages * stage_bytes + 1024 + 256;
  CUtensorMap tmX, tmD;
  memset(&tmX, 0, sizeof(tmX));
  memset(&tmD, 0, sizeof(tmD));
  if (g.use_tma) {
    rc = make_tmap(&tmX, X16, plan->n_in, Cin, 1);
    if (rc) return rc;
    rc = make_tmap(&tmD, dY16, plan->n_out, Cout, 1);
    if (rc) return rc;
  }
  LG_CUDA_OK(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(plan->kernel_volume * g.m_blocks), (unsigned)chunks);
  k_wgrad_tc<<<grid, kThreads, smem, stream>>>(g, tmX, tmD);
  LG_LAUNCH_OK();
  return launch_reduce_partials((const float*)workspace, chunks, (int64_t)plan->kernel_volume * Cin * Cout, out_scale,
                                dW, stream);
}
