// fp32 SIMT gather-GEMM and wgrad over a gather plan.  These are the exact-fp32 kernels used for the
// layer shapes the tensor-core path does not take (1-channel stem, 7-class head) and as the on-device
// fp32 cross-check of the tcgen05 kernels.  Output-stationary: every output row is produced by one
// CTA, so there are no atomics and the summation order is fixed (k ascending, channel ascending).
// Reference contract: MinkowskiConvolution forward/backward, utils/models/minkunet_bev.py:57-123.
#include "common.cuh"

namespace lg {

constexpr int TM = LG_TILE_ROWS;  // 128 slots per CTA
constexpr int TN = 64;            // output columns per CTA
constexpr int TK = 16;            // reduction chunk

// Y[out_row[s], n0:n0+64] = sum_k A[nbr[k][s], :] @ Wk[:, n0:n0+64]
__global__ void __launch_bounds__(256)
    k_gemm_simt(lgConvPlan plan, const float* __restrict__ A, int Ca, const float* __restrict__ W, int N,
                int w_transposed, int flip_k, const float* __restrict__ bias, float* __restrict__ Y) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN];
  __shared__ int s_row[TM];
  const int tid = threadIdx.x;
  const int64_t tile = blockIdx.x;
  const int n0 = blockIdx.y * TN;
  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, 8 rows x 4 cols each
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int K = plan.kernel_volume;
  for (int k = 0; k < K; ++k) {
    const uint32_t m = plan.tile_mask[tile * plan.mask_words + (k >> 5)];
    if (!((m >> (k & 31)) & 1u)) continue;
    const int wk = flip_k ? (K - 1 - k) : k;
    __syncthreads();
    if (tid < TM) s_row[tid] = plan.nbr[(int64_t)k * plan.k_stride + tile * TM + tid];
    for (int c0 = 0; c0 < Ca; c0 += TK) {
      const int kc = min(TK, Ca - c0);
      __syncthreads();
      // A chunk: 128 rows x kc
      for (int e = tid; e < TM * TK; e += 256) {
        const int r = e / TK, c = e % TK;
        const int row = s_row[r];
        As[c][r] = (row >= 0 && c < kc) ? __ldg(A + (int64_t)row * Ca + c0 + c) : 0.f;
      }
      // B chunk: kc x 64
      for (int e = tid; e < TK * TN; e += 256) {
        const int c = e / TN, n = e % TN;
        float v = 0.f;
        if (c < kc && n0 + n < N) {
          v = w_transposed ? __ldg(W + ((int64_t)wk * N + (n0 + n)) * Ca + c0 + c)
                           : __ldg(W + ((int64_t)wk * Ca + (c0 + c)) * N + n0 + n);
        }
        Bs[c][n] = v;
      }
      __syncthreads();
      for (int c = 0; c < kc; ++c) {
        float a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = As[c][ty * 8 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[c][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t s = tile * TM + ty * 8 + i;
    int64_t row = plan.out_row ? (int64_t)plan.out_row[s] : s;
    if (row < 0 || row >= plan.n_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) Y[row * N + n] = acc[i][j] + (bias ? bias[n] : 0.f);
    }
  }
}

// Partial dW for one (k, slot chunk, 64x64 tile of Cin x Cout):
//   P[chunk][k][ci][co] = sum_{s in chunk} X[nbr[k][s], ci] * dY[out_row[s], co]
constexpr int WG_T = 64;
constexpr int WG_S = 16;  // slots per smem step

__global__ void __launch_bounds__(256)
    k_wgrad_simt(lgConvPlan plan, const float* __restrict__ X, int Cin, const float* __restrict__ dY, int Cout,
                 int tiles_per_chunk, float* __restrict__ partial) {
  __shared__ float Xs[WG_S][WG_T + 1];
  __shared__ float Ds[WG_S][WG_T + 1];
  __shared__ int s_in[WG_S], s_out[WG_S];
  const int tid = threadIdx.x;
  const int k = blockIdx.x;
  const int chunk = blockIdx.y;
  const int tiles_b = (Cout + WG_T - 1) / WG_T;
  const int a0 = (blockIdx.z / tiles_b) * WG_T, b0 = (blockIdx.z % tiles_b) * WG_T;
  const int ty = tid >> 4, tx = tid & 15;  // 4 x 4 outputs per thread
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int64_t n_tiles = plan.n_slots / TM;
  const int64_t t_begin = (int64_t)chunk * tiles_per_chunk;
  const int64_t t_end = (n_tiles < t_begin + tiles_per_chunk) ? n_tiles : t_begin + tiles_per_chunk;
  for (int64_t tile = t_begin; tile < t_end; ++tile) {
    const uint32_t m = plan.tile_mask[tile * plan.mask_words + (k >> 5)];
    if (!((m >> (k & 31)) & 1u)) continue;
    for (int s0 = 0; s0 < TM; s0 += WG_S) {
      __syncthreads();
      if (tid < WG_S) {
        const int64_t s = tile * TM + s0 + tid;
        int in = plan.nbr[(int64_t)k * plan.k_stride + s];
        int64_t out = plan.out_row ? (int64_t)plan.out_row[s] : s;
        if (out < 0 || out >= plan.n_out) in = -1;
        s_in[tid] = in;
        s_out[tid] = (int)out;
      }
      __syncthreads();
      for (int e = tid; e < WG_S * WG_T; e += 256) {
        const int r = e / WG_T, c = e % WG_T;
        const int in = s_in[r];
        Xs[r][c] = (in >= 0 && a0 + c < Cin) ? __ldg(X + (int64_t)in * Cin + a0 + c) : 0.f;
        Ds[r][c] = (in >= 0 && b0 + c < Cout) ? __ldg(dY + (int64_t)s_out[r] * Cout + b0 + c) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < WG_S; ++r) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = Xs[r][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Ds[r][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
  float* P = partial + ((int64_t)chunk * plan.kernel_volume + k) * Cin * Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int a = a0 + ty * 4 + i, b = b0 + tx * 4 + j;
      if (a < Cin && b < Cout) P[(int64_t)a * Cout + b] = acc[i][j];
    }
}

__global__ void __launch_bounds__(256)
    k_reduce_partials(const float* __restrict__ partial, int n_chunks, int64_t n_elems, float scale_mode,
                      const float* __restrict__ out_scale, float* __restrict__ dW) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_elems) return;
  float s = 0.f;
  for (int c = 0; c < n_chunks; ++c) s += partial[(int64_t)c * n_elems + i];  // fixed order
  dW[i] = out_scale ? s * out_scale[0] : s;
}

// float4 form: block = 64 element-quads x 4 parts; part p folds the chunks of its contiguous quarter (8 independent
// 16-byte loads in flight per thread), then the four part sums are added in part order through shared memory:
// fixed order, 2.5 GB of partials per step read at HBM speed instead of one 4-byte load per thread and chunk.
constexpr int kRpX = 64, kRpParts = 4;
__global__ void __launch_bounds__(kRpX* kRpParts)
    k_reduce_partials4(const float4* __restrict__ partial, int n_chunks, int64_t n4, const float* __restrict__ out_scale,
                       float4* __restrict__ dW) {
  __shared__ float4 sm[kRpParts][kRpX];
  const int64_t i = (int64_t)blockIdx.x * kRpX + threadIdx.x;
  const int per = (n_chunks + kRpParts - 1) / kRpParts;
  const int c0 = threadIdx.y * per, c1 = min(n_chunks, c0 + per);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(partial + (int64_t)(c + u) * n4 + i);
#pragma unroll
      for (int u = 0; u < 8; ++u) s.x += v[u].x, s.y += v[u].y, s.z += v[u].z, s.w += v[u].w;
    }
    for (; c < c1; ++c) {
      const float4 v = __ldcs(partial + (int64_t)c * n4 + i);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
  }
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < n4) {
#pragma unroll
    for (int p = 1; p < kRpParts; ++p) {
      const float4 v = sm[p][threadIdx.x];
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    if (out_scale) {
      const float k = out_scale[0];
      s.x *= k, s.y *= k, s.z *= k, s.w *= k;
    }
    dW[i] = s;
  }
}


// ------------------------------------------------------------------------------------ Cin == 1 stem
// conv0p1s1 of MinkUNet34 (utils/models/minkunet_bev.py:57: kernel 5, 1 -> 32 channels, 125 offsets) is a
// scalar gather, not a GEMM: Y[o][:] = sum_k x[nbr[k][o]] * W[k][:].  Thread = output row, W in shared
// memory (broadcast reads), result rows transposed through shared memory for coalesced stores.
constexpr int C1_N = 32;

__global__ void __launch_bounds__(128)
    k_conv_c1_fwd(lgConvPlan plan, const float* __restrict__ X, const float* __restrict__ W, int flip_k,
                  const float* __restrict__ bias, float* __restrict__ Y) {
  extern __shared__ float c1_sm[];
  float* Ws = c1_sm;                                   // [K][32]
  float* Ts = c1_sm + plan.kernel_volume * C1_N;       // [4 warps][32][33]
  const int K = plan.kernel_volume;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < K * C1_N; e += 128) {
    const int k = e / C1_N;
    Ws[e] = W[(flip_k ? (K - 1 - k) : k) * C1_N + (e - k * C1_N)];
  }
  __syncthreads();
  const int64_t n_tiles = plan.n_slots / TM;
  float* Tw = Ts + warp * 32 * 33;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s = tile * TM + threadIdx.x;
    float acc[C1_N];
#pragma unroll
    for (int j = 0; j < C1_N; ++j) acc[j] = 0.f;
    uint32_t mw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < plan.mask_words) mw[i] = __ldg(plan.tile_mask + tile * plan.mask_words + i);
    // 8 offsets per round: the 8 neighbour ids, then the 8 gathered scalars, are independent loads in flight
    // together (the k-at-a-time loop was a chain of 2 * 125 dependent global latencies per row)
    for (int k0 = 0; k0 < K; k0 += 8) {
      const int wi = k0 >> 5;  // k0 is a multiple of 8: the 8 bits never straddle a word
      const uint32_t word = wi == 0 ? mw[0] : wi == 1 ? mw[1] : wi == 2 ? mw[2] : mw[3];
      const uint32_t bits = (word >> (k0 & 31)) & 0xffu;
      if (!bits) continue;
      int nb[8];
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        nb[j] = ((bits >> j) & 1u) ? __ldg(plan.nbr + (int64_t)(k0 + j) * plan.k_stride + s) : -1;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = nb[j] >= 0 ? __ldg(X + nb[j]) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (nb[j] >= 0) {
          const float4* w4 = reinterpret_cast<const float4*>(Ws + (k0 + j) * C1_N);
#pragma unroll
          for (int q = 0; q < C1_N / 4; ++q) {
            const float4 w = w4[q];
            acc[4 * q + 0] = fmaf(x[j], w.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(x[j], w.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(x[j], w.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(x[j], w.w, acc[4 * q + 3]);
          }
        }
      }
    }
    // transpose through shared memory: lane = row -> lane = channel
#pragma unroll
    for (int j = 0; j < C1_N; ++j) Tw[lane * 33 + j] = acc[j];
    __syncwarp();
    const float b = bias ? bias[lane] : 0.f;
    const int64_t s0 = tile * TM + warp * 32;
    for (int r = 0; r < 32; ++r) {
      int64_t row = plan.out_row ? (int64_t)plan.out_row[s0 + r] : s0 + r;
      if (row >= 0 && row < plan.n_out) Y[row * C1_N + lane] = Tw[r * 33 + lane] + b;
    }
    __syncwarp();
  }
}

// dW[k][co] = sum_o x[nbr[k][o]] * dY[o][co]: lane = output channel, acc[k] in registers, the gathered
// scalars of a 32-row window staged in shared memory; one partial [K][32] per CTA, reduced in fixed order.
constexpr int C1_KMAX = 128;
constexpr int C1_XPITCH = C1_KMAX + 4;  // row pitch of the staged scalars: 16-byte aligned rows, 4-way instead of 32-way store conflicts

__global__ void __launch_bounds__(256, 1)
    k_conv_c1_wgrad(lgConvPlan plan, const float* __restrict__ X, const float* __restrict__ dY,
                    float* __restrict__ partial) {
  __shared__ __align__(16) float xs[32][C1_XPITCH];
  __shared__ float red[C1_KMAX][C1_N];
  const int K = plan.kernel_volume;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[C1_KMAX];
#pragma unroll
  for (int k = 0; k < C1_KMAX; ++k) acc[k] = 0.f;
  const int64_t n_win = plan.n_slots / 32;
  for (int64_t win = blockIdx.x; win < n_win; win += gridDim.x) {
    const int64_t s0 = win * 32;
    __syncthreads();
    {  // stage x[nbr[k][s0 + r]] for the 32 rows of the window: all 16 id loads, then all 16 gathers, in flight together
      constexpr int kPer = C1_KMAX * 32 / 256;
      int nb[kPer];
      float v[kPer];
#pragma unroll
      for (int it = 0; it < kPer; ++it) {
        const int k = (threadIdx.x >> 5) + 8 * it;
        nb[it] = (k < K) ? __ldg(plan.nbr + (int64_t)k * plan.k_stride + s0 + lane) : -1;
      }
#pragma unroll
      for (int it = 0; it < kPer; ++it) v[it] = nb[it] >= 0 ? __ldg(X + nb[it]) : 0.f;
#pragma unroll
      for (int it = 0; it < kPer; ++it) xs[lane][(threadIdx.x >> 5) + 8 * it] = v[it];
    }
    // the warp's 4 rows of the window: row ids and dY values loaded up front
    int64_t rows[4];
    float dys[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t s = s0 + warp + 8 * q;
      rows[q] = plan.out_row ? (int64_t)__ldg(plan.out_row + s) : s;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dys[q] = (rows[q] >= 0 && rows[q] < plan.n_out) ? __ldg(dY + rows[q] * C1_N + lane) : 0.f;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (rows[q] < 0 || rows[q] >= plan.n_out) continue;
      const float dy = dys[q];
      const float4* x4 = reinterpret_cast<const float4*>(&xs[warp + 8 * q][0]);
#pragma unroll
      for (int j = 0; j < C1_KMAX / 4; ++j) {
        const float4 x = x4[j];
        acc[4 * j + 0] = fmaf(x.x, dy, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(x.y, dy, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(x.z, dy, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(x.w, dy, acc[4 * j + 3]);
      }
    }
  }
  // fixed-order reduction over the 8 warps
  for (int w = 0; w < 8; ++w) {
    __syncthreads();
    if (warp == w) {
#pragma unroll
      for (int k = 0; k < C1_KMAX; ++k) red[k][lane] = (w == 0 ? 0.f : red[k][lane]) + acc[k];
    }
  }
  __syncthreads();
  float* P = partial + (int64_t)blockIdx.x * K * C1_N;
  for (int e = threadIdx.x; e < K * C1_N; e += 256) P[e] = red[e >> 5][e & 31];
}

// ------------------------------------------------------------------------------------ narrow head (K = 1)
// `final` of MinkUNet34 (utils/models/minkunet_bev.py:118-123: kernel 1, 96 -> num_classes with bias) is a
// plain [rows x C] x [C x n] product with n <= 8: one HBM pass over the wide matrix, nothing else.  The generic
// 64-column tile kernels above spent 9x the arithmetic and staged everything through shared memory; these three
// read / write the wide matrix once, coalesced (8 lanes per 128 bytes of a row).
constexpr int HD_N = 8;      // padded narrow width
constexpr int HD_CMAX = 256; // widest wide matrix
// Shared-memory slot of channel c of the narrow weight matrix.  The 8 lanes of a row read channels 4*(l8 + 8t) + i
// at the same time; stored at [channel] they all fell into the same 4 banks (a 128-bit load is served a quarter warp
// at a time: 8-way conflict, 32 wavefronts per instruction -- the first version ran at 0.5 TB/s because of it).
// Slot ((t * 4 + i) * 8 + l8) makes those 8 lanes read 8 consecutive 16-byte words: conflict free.
__device__ __forceinline__ int hd_slot(int c) {
  const int c4 = c >> 2, i = c & 3;
  return (((c4 >> 3) * 4 + i) << 3) + (c4 & 7);
}

// Y[s][0..n) = A[nbr[s]][:] @ W (C x n, row-major, n <= 8) + bias.  8 lanes per row, 4 rows per warp.
__global__ void __launch_bounds__(256)
    k_head_fwd(lgConvPlan plan, const float* __restrict__ A, int C, const float* __restrict__ W, int n,
               const float* __restrict__ bias, float* __restrict__ Y) {
  __shared__ __align__(16) float Wlo[HD_CMAX][4], Whi[HD_CMAX][4];
  for (int e = threadIdx.x; e < C * HD_N; e += 256) {
    const int c = e / HD_N, j = e % HD_N;
    const float w = j < n ? W[c * n + j] : 0.f;
    if (j < 4) Wlo[hd_slot(c)][j] = w; else Whi[hd_slot(c)][j - 4] = w;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, l8 = lane & 7;
  const int64_t rows_per_blk = 32 * 4;  // 8 warps x 4 rows x 4 rounds
  for (int64_t base = (int64_t)blockIdx.x * rows_per_blk; base < plan.n_out; base += (int64_t)gridDim.x * rows_per_blk) {
#pragma unroll
    for (int rnd = 0; rnd < 4; ++rnd) {
      const int64_t s = base + rnd * 32 + (threadIdx.x >> 3);
      const bool live = s < plan.n_out;
      const int row = live ? __ldg(plan.nbr + s) : -1;
      float acc[HD_N];
#pragma unroll
      for (int j = 0; j < HD_N; ++j) acc[j] = 0.f;
      if (row >= 0) {
        const float4* a4 = reinterpret_cast<const float4*>(A + (int64_t)row * C);
        const int nt = C >> 5;  // float4 per lane
        float4 a[HD_CMAX / 32];
#pragma unroll
        for (int t = 0; t < HD_CMAX / 32; ++t)  // all loads of the row before the first use
          if (t < nt) a[t] = __ldg(a4 + l8 + 8 * t);
#pragma unroll
        for (int t = 0; t < HD_CMAX / 32; ++t) {
          if (t < nt) {
            const float av[4] = {a[t].x, a[t].y, a[t].z, a[t].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int slot = ((t * 4 + i) << 3) + l8;  // = hd_slot(4 * (l8 + 8 t) + i)
              const float4 w0 = *reinterpret_cast<const float4*>(&Wlo[slot][0]);
              const float4 w1 = *reinterpret_cast<const float4*>(&Whi[slot][0]);
              acc[0] = fmaf(av[i], w0.x, acc[0]); acc[1] = fmaf(av[i], w0.y, acc[1]);
              acc[2] = fmaf(av[i], w0.z, acc[2]); acc[3] = fmaf(av[i], w0.w, acc[3]);
              acc[4] = fmaf(av[i], w1.x, acc[4]); acc[5] = fmaf(av[i], w1.y, acc[5]);
              acc[6] = fmaf(av[i], w1.z, acc[6]); acc[7] = fmaf(av[i], w1.w, acc[7]);
            }
          }
        }
      }
      // fixed-order butterfly over the 8 lanes of the row (every lane ends with the same sums)
#pragma unroll
      for (int off = 1; off < 8; off <<= 1)
#pragma unroll
        for (int j = 0; j < HD_N; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
      float mine = 0.f;
#pragma unroll
      for (int j = 0; j < HD_N; ++j) mine = (l8 == j) ? acc[j] : mine;
      if (live && l8 < n) Y[s * n + l8] = mine + (bias ? bias[l8] : 0.f);
    }
  }
}

// dX[s][0..C) = dY[nbr[s]][0..n) @ W^T, W stored (C x n) row-major.  8 lanes per row, float4 stores.
__global__ void __launch_bounds__(256)
    k_head_dgrad(lgConvPlan plan, const float* __restrict__ dY, int n, const float* __restrict__ W, int C,
                 float* __restrict__ dX) {
  __shared__ __align__(16) float Wlo[HD_CMAX][4], Whi[HD_CMAX][4];
  for (int e = threadIdx.x; e < C * HD_N; e += 256) {
    const int c = e / HD_N, j = e % HD_N;
    const float w = j < n ? W[c * n + j] : 0.f;
    if (j < 4) Wlo[hd_slot(c)][j] = w; else Whi[hd_slot(c)][j - 4] = w;
  }
  __syncthreads();
  const int l8 = threadIdx.x & 7;
  for (int64_t s = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3); s < plan.n_out; s += (int64_t)gridDim.x * 32) {
    const int row = __ldg(plan.nbr + s);
    float g[HD_N];
#pragma unroll
    for (int j = 0; j < HD_N; ++j) g[j] = (row >= 0 && j < n) ? __ldg(dY + (int64_t)row * n + j) : 0.f;
    float4* o4 = reinterpret_cast<float4*>(dX + s * C);
    for (int c4 = l8; c4 < C / 4; c4 += 8) {
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w0 = *reinterpret_cast<const float4*>(&Wlo[hd_slot(4 * c4 + i)][0]);
        const float4 w1 = *reinterpret_cast<const float4*>(&Whi[hd_slot(4 * c4 + i)][0]);
        float a = g[0] * w0.x;
        a = fmaf(g[1], w0.y, a); a = fmaf(g[2], w0.z, a); a = fmaf(g[3], w0.w, a);
        a = fmaf(g[4], w1.x, a); a = fmaf(g[5], w1.y, a); a = fmaf(g[6], w1.z, a); a = fmaf(g[7], w1.w, a);
        o[i] = a;
      }
      o4[c4] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// partial[blk][c][0..n) = sum over the block's rows of X[nbr[s]][c] * dY[s][0..n): lane = channel (coalesced
// 128-byte row segments), warp = (row stripe, 32-channel group); warps of a block are combined in a fixed order.
constexpr int HD_WG_BLOCKS = 592;
__global__ void __launch_bounds__(256)
    k_head_wgrad(lgConvPlan plan, const float* __restrict__ X, int C, const float* __restrict__ dY, int n,
                 float* __restrict__ partial) {
  __shared__ float red[8][32][HD_N + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = C / 32;
  const int64_t rows_per_blk = (plan.n_out + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_blk;
  const int64_t r1 = r0 + rows_per_blk < plan.n_out ? r0 + rows_per_blk : plan.n_out;
  for (int gch = 0; gch < groups; ++gch) {
    float acc[HD_N];
#pragma unroll
    for (int j = 0; j < HD_N; ++j) acc[j] = 0.f;
    for (int64_t s = r0 + warp; s < r1; s += 32) {  // 4 rows in flight per warp
      float x[4], g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t su = s + 8 * u;
        const int row = su < r1 ? __ldg(plan.nbr + su) : -1;
        x[u] = row >= 0 ? __ldg(X + (int64_t)row * C + gch * 32 + lane) : 0.f;
        g[u] = (su < r1 && lane < n) ? __ldg(dY + su * n + lane) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < HD_N; ++j) acc[j] = fmaf(x[u], __shfl_sync(0xffffffffu, g[u], j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < HD_N; ++j) red[warp][lane][j] = acc[j];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int j = 0; j < HD_N; ++j) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][lane][j];
        if (j < n) partial[((int64_t)blockIdx.x * C + gch * 32 + lane) * n + j] = t;
      }
    }
    __syncthreads();
  }
}

static bool head_shape(const lgConvPlan* plan, int wide, int narrow) {
  return plan->kernel_volume == 1 && plan->out_row == nullptr && narrow >= 1 && narrow <= HD_N && wide % 32 == 0 &&
         wide >= 32 && wide <= HD_CMAX && plan->n_out > 0;
}
static int head_wgrad_blocks(const lgConvPlan* plan) {
  const int64_t want = ceil_div(plan->n_out, (int64_t)256);
  return (int)(want < HD_WG_BLOCKS ? (want > 0 ? want : 1) : HD_WG_BLOCKS);
}

static int c1_wgrad_blocks(const lgConvPlan* plan) {
  const int64_t n_win = plan->n_slots / 32;
  return (int)(n_win < 296 ? (n_win > 0 ? n_win : 1) : 296);
}

int wgrad_chunks(const lgConvPlan* plan, int* tiles_per_chunk) {
  const int64_t n_tiles = plan->n_slots / TM;
  int64_t want = 592 / (plan->kernel_volume > 0 ? plan->kernel_volume : 1);
  if (want < 1) want = 1;
  int chunks = (int)(n_tiles < want ? n_tiles : want);
  if (chunks < 1) chunks = 1;
  *tiles_per_chunk = (int)ceil_div(n_tiles, chunks);
  chunks = (int)ceil_div(n_tiles, *tiles_per_chunk);
  return chunks < 1 ? 1 : chunks;
}

int launch_reduce_partials(const float* partial, int n_chunks, int64_t n_elems, const float* out_scale, float* dW,
                           cudaStream_t stream) {
  if (n_elems % 4 == 0 && (((uintptr_t)partial | (uintptr_t)dW) & 15) == 0) {
    const int64_t n4 = n_elems / 4;
    k_reduce_partials4<<<(unsigned)ceil_div(n4, kRpX), dim3(kRpX, kRpParts), 0, stream>>>(
        reinterpret_cast<const float4*>(partial), n_chunks, n4, out_scale, reinterpret_cast<float4*>(dW));
  } else {
    k_reduce_partials<<<(unsigned)ceil_div(n_elems, 256), 256, 0, stream>>>(partial, n_chunks, n_elems, 0.f, out_scale,
                                                                            dW);
  }
  LG_LAUNCH_OK();
  return LG_OK;
}

}  // namespace lg

using namespace lg;

static int check_plan(const lgConvPlan* p, const char* who) {
  LG_CHECK_ARG(p != nullptr, "%s: null plan", who);
  LG_CHECK_ARG(p->n_slots >= 0 && p->n_slots % LG_TILE_ROWS == 0, "%s: n_slots must be a multiple of 128", who);
  LG_CHECK_ARG(p->kernel_volume >= 1 && p->kernel_volume <= 128 && p->mask_words == (p->kernel_volume + 31) / 32,
               "%s: bad kernel_volume/mask_words", who);
  LG_CHECK_ARG(p->k_stride == 0 || p->k_stride == p->n_slots, "%s: k_stride must be 0 or n_slots", who);
  if (p->n_slots > 0) LG_CHECK_ARG(p->nbr && p->tile_mask, "%s: null plan arrays", who);
  return LG_OK;
}

extern "C" int lg_conv_gemm_simt(const lgConvPlan* plan, const float* A, int32_t Ca, const float* W, int32_t N,
                                 int32_t w_transposed, int32_t flip_k, const float* bias, float* Y, void* stream) {
  int rc = check_plan(plan, "lg_conv_gemm_simt");
  if (rc) return rc;
  LG_CHECK_ARG(Ca >= 1 && N >= 1, "lg_conv_gemm_simt: bad channel counts");
  if (plan->n_slots == 0) return LG_OK;
  LG_CHECK_ARG(A && W && Y, "lg_conv_gemm_simt: null pointer");
  if (Ca == 1 && N == C1_N && !w_transposed && plan->kernel_volume <= C1_KMAX) {
    const int64_t n_tiles = plan->n_slots / TM;
    const size_t smem = ((size_t)plan->kernel_volume * C1_N + 4 * 32 * 33) * sizeof(float);
    const unsigned blocks = (unsigned)(n_tiles < 148 * 8 ? n_tiles : 148 * 8);
    k_conv_c1_fwd<<<blocks, 128, smem, (cudaStream_t)stream>>>(*plan, A, W, flip_k, bias, Y);
    LG_LAUNCH_OK();
    return LG_OK;
  }
  if (!w_transposed && head_shape(plan, Ca, N)) {  // 96 -> 7 head, forward
    const int64_t want = ceil_div(plan->n_out, (int64_t)128);
    k_head_fwd<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, (cudaStream_t)stream>>>(*plan, A, Ca, W, N, bias, Y);
    LG_LAUNCH_OK();
    return LG_OK;
  }
  if (w_transposed && !bias && head_shape(plan, N, Ca)) {  // its dgrad: 7 -> 96 through W^T
    const int64_t want = ceil_div(plan->n_out, (int64_t)32);
    k_head_dgrad<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), 256, 0, (cudaStream_t)stream>>>(*plan, A, Ca, W, N, Y);
    LG_LAUNCH_OK();
    return LG_OK;
  }
  dim3 grid((unsigned)(plan->n_slots / TM), (unsigned)ceil_div(N, TN));
  k_gemm_simt<<<grid, 256, 0, (cudaStream_t)stream>>>(*plan, A, Ca, W, N, w_transposed, flip_k, bias, Y);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" size_t lg_conv_wgrad_workspace(const lgConvPlan* plan, int32_t Ca, int32_t Cb) {
  if (!plan) return 0;
  int tpc;
  int chunks = wgrad_chunks(plan, &tpc);
  if (Ca == 1 && Cb == C1_N && chunks < c1_wgrad_blocks(plan)) chunks = c1_wgrad_blocks(plan);
  if (head_shape(plan, Ca, Cb) && chunks < head_wgrad_blocks(plan)) chunks = head_wgrad_blocks(plan);
  return (size_t)chunks * plan->kernel_volume * Ca * Cb * sizeof(float) + 256;
}

extern "C" int lg_conv_wgrad_simt(const lgConvPlan* plan, const float* X_in, int32_t Cin, const float* dY_out,
                                  int32_t Cout, float* dW, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_plan(plan, "lg_conv_wgrad_simt");
  if (rc) return rc;
  LG_CHECK_ARG(Cin >= 1 && Cout >= 1 && dW, "lg_conv_wgrad_simt: bad arguments");
  const int64_t n_elems = (int64_t)plan->kernel_volume * Cin * Cout;
  if (plan->n_slots == 0) {
    LG_CUDA_OK(cudaMemsetAsync(dW, 0, sizeof(float) * n_elems, stream));
    return LG_OK;
  }
  LG_CHECK_ARG(workspace && workspace_bytes >= lg_conv_wgrad_workspace(plan, Cin, Cout),
               "lg_conv_wgrad_simt: workspace too small");
  if (Cin == 1 && Cout == C1_N && plan->kernel_volume <= C1_KMAX) {
    const int blocks = c1_wgrad_blocks(plan);
    k_conv_c1_wgrad<<<blocks, 256, 0, stream>>>(*plan, X_in, dY_out, (float*)workspace);
    LG_LAUNCH_OK();
    return launch_reduce_partials((const float*)workspace, blocks, n_elems, nullptr, dW, stream);
  }
  if (head_shape(plan, Cin, Cout)) {
    const int blocks = head_wgrad_blocks(plan);
    k_head_wgrad<<<blocks, 256, 0, stream>>>(*plan, X_in, Cin, dY_out, Cout, (float*)workspace);
    LG_LAUNCH_OK();
    return launch_reduce_partials((const float*)workspace, blocks, n_elems, nullptr, dW, stream);
  }
  int tpc;
  const int chunks = wgrad_chunks(plan, &tpc);
  dim3 grid((unsigned)plan->kernel_volume, (unsigned)chunks,
            (unsigned)(ceil_div(Cin, WG_T) * ceil_div(Cout, WG_T)));
  k_wgrad_simt<<<grid, 256, 0, stream>>>(*plan, X_in, Cin, dY_out, Cout, tpc, (float*)workspace);
  LG_LAUNCH_OK();
  return launch_reduce_partials((const float*)workspace, chunks, n_elems, nullptr, dW, stream);
}
