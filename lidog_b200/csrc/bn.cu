// Fused batch-norm passes over the N x C feature matrix of a sparse tensor.
// Reference contract: ME.MinkowskiBatchNorm / MinkowskiSyncBatchNorm (= torch BatchNorm1d on the feature
// matrix) followed by ME.MinkowskiReLU and the residual `out += x` of ME's BasicBlock, as chained at
// utils/models/minkunet_bev.py:308-368 and converted to SyncBN at train_lidog.py:228.
//
// torch runs that chain as 4-5 elementwise passes per layer forward (statistics, normalise, add, ReLU,
// operand cast) and 6-7 backward.  Here a layer is two passes each way:
//   forward   lg_bn_stats     per-channel sum / sum of squares          (read x)
//             lg_bn_finalize  mean, invstd, scale/shift, running stats  (C threads)
//             lg_bn_apply     y = act(x*scale+shift [+ x2*scale2+shift2] [+ res]) -> fp32 y and the 16-bit
//                             operand copy the next convolution gathers   (read x, write y, y16)
//   backward  lg_bn_bwd_stats     g = dy*[y>0]; sum g, sum g*(x-mean), max|g|, max|x-mean| (read dy, y, x)
//             lg_bn_bwd_finalize  dx coefficients, dgamma, dbeta, power-of-two fp16 scale from a bound
//             lg_bn_bwd_apply     dx = a*g + b*(x-mean) + c -> fp32 dx and the scaled 16-bit copy the
//                                 convolution's dgrad / wgrad kernels read; residual gradient g
// The sums cross ranks between the stats and the finalize step (SyncBN: one all-reduce of 2C+1 doubles).
// All reductions run in a fixed order (per-block partials, then one thread per channel): deterministic.
// Bound: HBM.  Algorithmic bytes per element: forward 4 (stats) + 4+4+2 (apply), backward 12 + 12+4+2.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>

#include "common.cuh"
#include "peer.cuh"
#include "runtime.cuh"

namespace lg {

constexpr int kBnThreads = 256;
constexpr int kBnMaxBlocks = 592;  // 4 per SM

__device__ __forceinline__ unsigned short bn_to16(float v, int fmt) {
  if (fmt == LG_FMT_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  return __half_as_ushort(__float2half_rn(v));
}
__device__ __forceinline__ void store16x4(unsigned short* dst, float4 v, int fmt) {
  uint2 o;
  o.x = bn_to16(v.x, fmt) | ((unsigned)bn_to16(v.y, fmt) << 16);
  o.y = bn_to16(v.z, fmt) | ((unsigned)bn_to16(v.w, fmt) << 16);
  *reinterpret_cast<uint2*>(dst) = o;
}

// Thread layout of the reduction kernels: a thread owns 4 consecutive channels (one float4) and every
// rpp-th row of its block's row range; Cq = C / 4 threads cover a row.
struct RedGeom {
  int cq, ry, rpp;
  bool active;
  int64_t r0, r1;
};
__device__ __forceinline__ RedGeom red_geom(int64_t n, int C) {
  RedGeom g;
  const int Cq = C >> 2;
  g.rpp = kBnThreads / Cq;
  g.cq = threadIdx.x % Cq;
  g.ry = threadIdx.x / Cq;
  g.active = g.ry < g.rpp;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  g.r0 = (int64_t)blockIdx.x * per;
  g.r1 = min(n, g.r0 + per);
  return g;
}

// sum over ry (fixed order) of one float4 per thread -> out[0..C) of this block's partial row
__device__ __forceinline__ void block_sum4(float4 v, const RedGeom& g, int C, float* sm, float* out, bool is_max) {
  __syncthreads();
  if (g.active) *reinterpret_cast<float4*>(sm + (size_t)g.ry * C + g.cq * 4) = v;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    float a = sm[c];
    for (int r = 1; r < g.rpp; ++r) a = is_max ? fmaxf(a, sm[(size_t)r * C + c]) : a + sm[(size_t)r * C + c];
    out[c] = a;
  }
}

__global__ void __launch_bounds__(kBnThreads)
    k_bn_stats(const float* __restrict__ x, int64_t n, int C, float* __restrict__ partial) {
  extern __shared__ float sm[];
  const RedGeom g = red_geom(n, C);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  if (g.active) {
    int64_t r = g.r0 + g.ry;
    const int64_t step = g.rpp;
    for (; r + 3 * step < g.r1; r += 4 * step) {  // four independent loads in flight per thread
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + r * C) + g.cq);
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (r + step) * C) + g.cq);
      const float4 v2 = __ldg(reinterpret_cast<const float4*>(x + (r + 2 * step) * C) + g.cq);
      const float4 v3 = __ldg(reinterpret_cast<const float4*>(x + (r + 3 * step) * C) + g.cq);
      s.x += (v0.x + v1.x) + (v2.x + v3.x), s.y += (v0.y + v1.y) + (v2.y + v3.y);
      s.z += (v0.z + v1.z) + (v2.z + v3.z), s.w += (v0.w + v1.w) + (v2.w + v3.w);
      ss.x += (v0.x * v0.x + v1.x * v1.x) + (v2.x * v2.x + v3.x * v3.x);
      ss.y += (v0.y * v0.y + v1.y * v1.y) + (v2.y * v2.y + v3.y * v3.y);
      ss.z += (v0.z * v0.z + v1.z * v1.z) + (v2.z * v2.z + v3.z * v3.z);
      ss.w += (v0.w * v0.w + v1.w * v1.w) + (v2.w * v2.w + v3.w * v3.w);
    }
    for (; r < g.r1; r += step) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C) + g.cq);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
      ss.x += v.x * v.x, ss.y += v.y * v.y, ss.z += v.z * v.z, ss.w += v.w * v.w;
    }
  }
  float* out = partial + (size_t)blockIdx.x * 2 * C;
  block_sum4(s, g, C, sm, out, false);
  block_sum4(ss, g, C, sm, out + C, false);
}

// Second stage of the reductions.  Block (128 channels, 8 partial groups): group y folds partials y, y+8, ...
// and the 8 group results are combined in order by group 0 -- a fixed order, so deterministic.
constexpr int kRedX = 128, kRedY = 8;

// partial [n_partials][width] (float) -> sums [width] (double)
__global__ void __launch_bounds__(kRedX * kRedY)
    k_bn_reduce(const float* __restrict__ partial, int n_partials, int width, double* __restrict__ sums, double tail) {
  __shared__ double sm[kRedY][kRedX];
  const int c = blockIdx.x * kRedX + threadIdx.x;
  if (c == 0 && threadIdx.y == 0) sums[width] = tail;  // the row count rides along (SyncBN sums it too)
  double a = 0.0;
  if (c < width)
    for (int p = threadIdx.y; p < n_partials; p += kRedY) a += (double)partial[(size_t)p * width + c];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < width) {
    for (int y = 1; y < kRedY; ++y) a += sm[y][threadIdx.x];
    sums[c] = a;
  }
}
// backward partial rows are [3C sums | 3C maxes]
__global__ void __launch_bounds__(kRedX * kRedY)
    k_bn_bwd_reduce(const float* __restrict__ partial, int n_partials, int C, double* __restrict__ sums,
                    float* __restrict__ maxes) {
  __shared__ double sm[kRedY][kRedX];
  const int c = blockIdx.x * kRedX + threadIdx.x;
  const bool is_sum = c < 3 * C;  // uniform per block when 3C % 128 == 0, harmless otherwise
  double a = 0.0;
  if (c < 6 * C)
    for (int p = threadIdx.y; p < n_partials; p += kRedY) {
      const double v = (double)partial[(size_t)p * 6 * C + c];
      a = is_sum ? a + v : fmax(a, v);
    }
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < 6 * C) {
    for (int y = 1; y < kRedY; ++y) a = is_sum ? a + sm[y][threadIdx.x] : fmax(a, sm[y][threadIdx.x]);
    if (is_sum)
      sums[c] = a;
    else
      maxes[c - 3 * C] = (float)a;
  }
}

// sums = [sum x (C), sum x^2 (C)], count = rows over all ranks.  stats_out = [mean, invstd, scale, shift] x C.
__global__ void k_bn_finalize(const double* __restrict__ sums, double count, const double* __restrict__ count_dev, int C,
                              const float* __restrict__ gamma,
                              const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                              float* running_var, long long* num_batches_tracked, float* __restrict__ stats_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
  if (c >= C) return;
  if (count_dev) count = *count_dev;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float gm = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  const float scale = gm * invstd;
  stats_out[c] = (float)mean;
  stats_out[C + c] = invstd;
  stats_out[2 * C + c] = scale;
  stats_out[3 * C + c] = bt - (float)mean * scale;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// The pre-activation value of a layer WITHOUT a residual, exactly as k_bn_apply computes it (same fmaf, same operand
// order): the backward recomputes the ReLU mask from x (and x2) instead of reading y -- 8 of 30 bytes per element.
__device__ __forceinline__ float4 preact(float4 v, float4 sc, float4 sh) {
  return make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
}
__device__ __forceinline__ float4 preact_add(float4 o, float4 v2, float4 s2, float4 h2) {
  o.x += fmaf(v2.x, s2.x, h2.x), o.y += fmaf(v2.y, s2.y, h2.y), o.z += fmaf(v2.z, s2.z, h2.z), o.w += fmaf(v2.w, s2.w, h2.w);
  return o;
}

// y = act(x * scale + shift [+ x2 * scale2 + shift2] [+ res]); st = [mean, invstd, scale, shift] x C.
// A thread handles kApplyU float4 a block-width apart: all loads are issued before the first use, so a CTA
// keeps 4x the bytes in flight of the one-element-per-thread form (HBM-latency bound otherwise).
constexpr int kApplyU = 4;
__global__ void __launch_bounds__(256)
    k_bn_apply(const float* __restrict__ x, const float* __restrict__ st, const float* __restrict__ x2,
               const float* __restrict__ st2, const float* __restrict__ res, int relu, int64_t n4, int C,
               float* __restrict__ y, unsigned short* __restrict__ y16, int fmt) {
  const int64_t base = (int64_t)blockIdx.x * (256 * kApplyU) + threadIdx.x;
  const int Cq = C >> 2;
  float4 v[kApplyU], v2[kApplyU], r[kApplyU];
#pragma unroll
  for (int u = 0; u < kApplyU; ++u) {
    const int64_t i = base + u * 256;
    if (i < n4) {
      v[u] = __ldg(reinterpret_cast<const float4*>(x) + i);
      if (x2) v2[u] = __ldg(reinterpret_cast<const float4*>(x2) + i);
      if (res) r[u] = __ldg(reinterpret_cast<const float4*>(res) + i);
    }
  }
#pragma unroll
  for (int u = 0; u < kApplyU; ++u) {
    const int64_t i = base + u * 256;
    if (i >= n4) break;
    const int c = (int)(i % Cq) * 4;
    const float4 sc = *reinterpret_cast<const float4*>(st + 2 * C + c), sh = *reinterpret_cast<const float4*>(st + 3 * C + c);
    float4 o = preact(v[u], sc, sh);
    if (x2) {
      const float4 s2 = *reinterpret_cast<const float4*>(st2 + 2 * C + c), h2 = *reinterpret_cast<const float4*>(st2 + 3 * C + c);
      o = preact_add(o, v2[u], s2, h2);
    }
    if (res) o.x += r[u].x, o.y += r[u].y, o.z += r[u].z, o.w += r[u].w;
    if (relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
    reinterpret_cast<float4*>(y)[i] = o;
    if (y16) store16x4(y16 + i * 4, o, fmt);
  }
}

__device__ __forceinline__ float4 masked(float4 dy, float4 y, int relu) {
  if (!relu) return dy;
  return make_float4(y.x > 0.f ? dy.x : 0.f, y.y > 0.f ? dy.y : 0.f, y.z > 0.f ? dy.z : 0.f, y.w > 0.f ? dy.w : 0.f);
}

// partial row of a block: [sum g | sum g*(x-mean) | sum g*(x2-mean2) | max|g| | max|x-mean| | max|x2-mean2|] x C
// RECOMP: ReLU mask from x (y is not read; layers without a residual).  HAS2: second BN branch (x2) present.
template <bool DENSE, bool RECOMP, bool HAS2>
__global__ void __launch_bounds__(kBnThreads, 4)  // the grid is 4 blocks per SM (kBnMaxBlocks): all of them must be resident
    k_bn_bwd_stats(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                   const float* __restrict__ st, const float* __restrict__ x2, const float* __restrict__ st2, int relu,
                   int64_t n, int C, float* __restrict__ partial, int64_t dy_ld4) {
  // dy_ld4: row pitch of dy in float4 (C / 4 when dense; larger for a column slice of a wider matrix -- the gradient
  // of one input of ME.cat arrives as such a view, and copying it first cost a pass)
  extern __shared__ float sm[];
  const RedGeom g = red_geom(n, C);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sg = z, sgx = z, sgx2 = z, mg = z, mx = z, mx2 = z;
  if (g.active) {
    const float4 mu = *reinterpret_cast<const float4*>(st + g.cq * 4);
    const float4 mu2 = HAS2 ? *reinterpret_cast<const float4*>(st2 + g.cq * 4) : z;
    float4 sc = z, sh = z, sc2 = z, sh2 = z;
    if (RECOMP) {
      sc = *reinterpret_cast<const float4*>(st + 2 * C + g.cq * 4), sh = *reinterpret_cast<const float4*>(st + 3 * C + g.cq * 4);
      if (HAS2)
        sc2 = *reinterpret_cast<const float4*>(st2 + 2 * C + g.cq * 4), sh2 = *reinterpret_cast<const float4*>(st2 + 3 * C + g.cq * 4);
    }
#pragma unroll 4
    for (int64_t r = g.r0 + g.ry; r < g.r1; r += g.rpp) {
      const int64_t i = r * (C >> 2) + g.cq;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      float4 yv = z;
      if (RECOMP) {
        yv = preact(v, sc, sh);
        if (HAS2) yv = preact_add(yv, __ldg(reinterpret_cast<const float4*>(x2) + i), sc2, sh2);
      } else if (relu) {
        yv = __ldg(reinterpret_cast<const float4*>(y) + i);
      }
      const float4 gy = masked(__ldg(reinterpret_cast<const float4*>(dy) + (DENSE ? i : r * dy_ld4 + g.cq)), yv, relu);
      const float4 d = make_float4(v.x - mu.x, v.y - mu.y, v.z - mu.z, v.w - mu.w);
      sg.x += gy.x, sg.y += gy.y, sg.z += gy.z, sg.w += gy.w;
      sgx.x += gy.x * d.x, sgx.y += gy.y * d.y, sgx.z += gy.z * d.z, sgx.w += gy.w * d.w;
      mg = make_float4(fmaxf(mg.x, fabsf(gy.x)), fmaxf(mg.y, fabsf(gy.y)), fmaxf(mg.z, fabsf(gy.z)), fmaxf(mg.w, fabsf(gy.w)));
      mx = make_float4(fmaxf(mx.x, fabsf(d.x)), fmaxf(mx.y, fabsf(d.y)), fmaxf(mx.z, fabsf(d.z)), fmaxf(mx.w, fabsf(d.w)));
      if (HAS2) {
        const float4 v2 = __ldg(reinterpret_cast<const float4*>(x2) + i);
        const float4 d2 = make_float4(v2.x - mu2.x, v2.y - mu2.y, v2.z - mu2.z, v2.w - mu2.w);
        sgx2.x += gy.x * d2.x, sgx2.y += gy.y * d2.y, sgx2.z += gy.z * d2.z, sgx2.w += gy.w * d2.w;
        mx2 = make_float4(fmaxf(mx2.x, fabsf(d2.x)), fmaxf(mx2.y, fabsf(d2.y)), fmaxf(mx2.z, fabsf(d2.z)),
                          fmaxf(mx2.w, fabsf(d2.w)));
      }
    }
  }
  float* out = partial + (size_t)blockIdx.x * 6 * C;
  block_sum4(sg, g, C, sm, out, false);
  block_sum4(sgx, g, C, sm, out + C, false);
  block_sum4(sgx2, g, C, sm, out + 2 * C, false);
  block_sum4(mg, g, C, sm, out + 3 * C, true);
  block_sum4(mx, g, C, sm, out + 4 * C, true);
  block_sum4(mx2, g, C, sm, out + 5 * C, true);
}

// sums_local = this rank's [sum g, sum g*xm, sum g*xm2] (parameter gradients), sums_global = the same over all
// ranks (dx coefficients; = sums_local without SyncBN), maxes = [max|g|, max|xm|, max|xm2|] x C (this rank).
// coef = [a, b, c] x C per branch:  dx = a*g + b*(x-mean) + c.
// scale_io: in  [2] = running max of the bounds written by earlier branches (0 to start);
//           out [0] = power of two bringing that bound into [2^11, 2^12), [1] = its inverse, [2] = bound.
__global__ void k_bn_bwd_finalize(const double* __restrict__ sums_local, const double* __restrict__ sums_global,
                                  const float* __restrict__ maxes, double count, const double* __restrict__ count_dev,
                                  int C, int branch,
                                  const float* __restrict__ gamma, const float* __restrict__ st,
                                  float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                  float* __restrict__ scale_io) {
  __shared__ float s_bound[32];
  const int c = threadIdx.x;
  float bound = 0.f;
  if (count_dev) count = *count_dev;
  if (c < C) {
    const float invstd = st[C + c], gm = gamma ? gamma[c] : 1.f;
    const double sg = sums_global[c], sgx = sums_global[(1 + branch) * C + c];
    const float a = gm * invstd;
    const float b = (float)(-(double)gm * invstd * invstd * invstd * sgx / count);
    const float c0 = (float)(-(double)gm * invstd * sg / count);
    coef[c] = a, coef[C + c] = b, coef[2 * C + c] = c0;
    if (dgamma) dgamma[c] = (float)(sums_local[(1 + branch) * C + c] * invstd);
    if (dbeta) dbeta[c] = (float)sums_local[c];
    bound = fabsf(a) * maxes[c] + fabsf(b) * maxes[(1 + branch) * C + c] + fabsf(c0);
  }
  for (int d = 16; d > 0; d >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, d));
  if ((threadIdx.x & 31) == 0) s_bound[threadIdx.x >> 5] = bound;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) m = fmaxf(m, s_bound[w]);
    float s = 1.f;
    if (m > 0.f && m <= 3.0e38f) {
      int e;
      frexpf(m, &e);
      s = ldexpf(1.f, 12 - e);
    }
    scale_io[0] = s, scale_io[1] = 1.f / s, scale_io[2] = m;
  }
}

// residual-only scale (no BN on that branch): bound = max|g|
__global__ void k_bn_gscale(const float* __restrict__ maxes, int C, float* __restrict__ scale_out) {
  float m = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) m = fmaxf(m, maxes[c]);
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  __shared__ float s_m[32];
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    m = 0.f;
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) m = fmaxf(m, s_m[w]);
    float s = 1.f;
    if (m > 0.f && m <= 3.0e38f) {
      int e;
      frexpf(m, &e);
      s = ldexpf(1.f, 12 - e);
    }
    scale_out[0] = s, scale_out[1] = 1.f / s, scale_out[2] = m;
  }
}

constexpr int kBwdU = 2;
__global__ void __launch_bounds__(256)
    k_bn_bwd_apply(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                   const float* __restrict__ st, const float* __restrict__ coef, const float* __restrict__ x2,
                   const float* __restrict__ st2, const float* __restrict__ coef2, int relu, int64_t n4, int C,
                   float* __restrict__ dx, unsigned short* __restrict__ dx16, const float* __restrict__ scale,
                   float* __restrict__ dx2, unsigned short* __restrict__ dx2_16, const float* __restrict__ scale2,
                   float* __restrict__ dres, unsigned short* __restrict__ dres16, const float* __restrict__ scale_r,
                   int fmt, int64_t dy_ld4) {
  const int64_t base = (int64_t)blockIdx.x * (256 * kBwdU) + threadIdx.x;
  const int Cq = C >> 2;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gd[kBwdU], gy[kBwdU], vx[kBwdU], vx2[kBwdU];
#pragma unroll
  for (int u = 0; u < kBwdU; ++u) {
    const int64_t i = base + u * 256;
    if (i < n4) {
      gd[u] = __ldg(reinterpret_cast<const float4*>(dy) + (dy_ld4 == Cq ? i : (i / Cq) * dy_ld4 + i % Cq));
      gy[u] = (relu && y) ? __ldg(reinterpret_cast<const float4*>(y) + i) : z;
      vx[u] = __ldg(reinterpret_cast<const float4*>(x) + i);
      if (x2) vx2[u] = __ldg(reinterpret_cast<const float4*>(x2) + i);
    }
  }
  const bool recomp = relu && !y;  // no residual in the forward: the mask is the sign of the recomputed pre-activation
  const float s1 = (dx16 && scale) ? scale[0] : 1.f, s2 = (dx2_16 && scale2) ? scale2[0] : 1.f,
              sr = (dres16 && scale_r) ? scale_r[0] : 1.f;
#pragma unroll
  for (int u = 0; u < kBwdU; ++u) {
    const int64_t i = base + u * 256;
    if (i >= n4) break;
    const int c = (int)(i % Cq) * 4;
    if (recomp) {
      gy[u] = preact(vx[u], *reinterpret_cast<const float4*>(st + 2 * C + c), *reinterpret_cast<const float4*>(st + 3 * C + c));
      if (x2)
        gy[u] = preact_add(gy[u], vx2[u], *reinterpret_cast<const float4*>(st2 + 2 * C + c),
                           *reinterpret_cast<const float4*>(st2 + 3 * C + c));
    }
    const float4 g = masked(gd[u], gy[u], relu);
    {
      const float4 v = vx[u];
      const float4 mu = *reinterpret_cast<const float4*>(st + c);
      const float4 a = *reinterpret_cast<const float4*>(coef + c), b = *reinterpret_cast<const float4*>(coef + C + c),
                   k = *reinterpret_cast<const float4*>(coef + 2 * C + c);
      const float4 o = make_float4(fmaf(a.x, g.x, fmaf(b.x, v.x - mu.x, k.x)), fmaf(a.y, g.y, fmaf(b.y, v.y - mu.y, k.y)),
                                   fmaf(a.z, g.z, fmaf(b.z, v.z - mu.z, k.z)), fmaf(a.w, g.w, fmaf(b.w, v.w - mu.w, k.w)));
      if (dx) reinterpret_cast<float4*>(dx)[i] = o;  // NULL: only the 16-bit copy is consumed (a tensor-core convolution)
      if (dx16) store16x4(dx16 + i * 4, make_float4(o.x * s1, o.y * s1, o.z * s1, o.w * s1), fmt);
    }
    if (x2) {
      const float4 v = vx2[u];
      const float4 mu = *reinterpret_cast<const float4*>(st2 + c);
      const float4 a = *reinterpret_cast<const float4*>(coef2 + c), b = *reinterpret_cast<const float4*>(coef2 + C + c),
                   k = *reinterpret_cast<const float4*>(coef2 + 2 * C + c);
      const float4 o = make_float4(fmaf(a.x, g.x, fmaf(b.x, v.x - mu.x, k.x)), fmaf(a.y, g.y, fmaf(b.y, v.y - mu.y, k.y)),
                                   fmaf(a.z, g.z, fmaf(b.z, v.z - mu.z, k.z)), fmaf(a.w, g.w, fmaf(b.w, v.w - mu.w, k.w)));
      if (dx2) reinterpret_cast<float4*>(dx2)[i] = o;
      if (dx2_16) store16x4(dx2_16 + i * 4, make_float4(o.x * s2, o.y * s2, o.z * s2, o.w * s2), fmt);
    }
    if (dres) {
      reinterpret_cast<float4*>(dres)[i] = g;
      if (dres16) store16x4(dres16 + i * 4, make_float4(g.x * sr, g.y * sr, g.z * sr, g.w * sr), fmt);
    }
  }
}


// ------------------------------------------------------------------------------------ fused layer tails
// One kernel between the statistics pass and the apply pass of a layer: second-stage reduction of the block
// partials (fixed order), the SyncBN exchange over NVLink peer memory (the protocol of peer.cu, inline: the LAST
// block of the grid publishes the sums, raises the epoch flag, waits for the peers' flags and adds their vectors in
// rank order), and the per-channel finalisation.  It replaces k_bn_reduce + k_peer_sum + k_bn_finalize (forward) and
// k_bn_bwd_reduce + k_peer_sum + 2 x k_bn_bwd_finalize + k_bn_gscale (backward): 62 x 3 + 55 x 5 launches per step.
struct PeerCtx {
  PeerSlot* buf[kPeerMaxWorld];
  int world, rank;
  unsigned long long epoch;
};

// grid (ceil(width / 128), S), block (128, 8): block (bx, by) folds the partial rows of slice by for 128 columns
// into slice_out[by][c] (double).  Returns true in the block that finished last (its threads then see every slice).
__device__ __forceinline__ bool tail_reduce(const float* __restrict__ partial, int n_partials, int width, int n_sum,
                                            double* __restrict__ slice_out, int* ticket) {
  __shared__ double sm[kRedY][kRedX];
  __shared__ int s_last;
  const int c = blockIdx.x * kRedX + threadIdx.x;
  const int S = gridDim.y;
  const int per = (n_partials + S - 1) / S;
  const int p0 = blockIdx.y * per, p1 = min(n_partials, p0 + per);
  const bool is_sum = c < n_sum;  // columns >= n_sum are maxima
  double a = 0.0;
  if (c < width) {
    // 8 loads in flight per thread: one L2 round trip per 8 rows instead of one per row (the adds stay in row order)
    int p = p0 + threadIdx.y;
    for (; p + 7 * kRedY < p1; p += 8 * kRedY) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(partial + (size_t)(p + u * kRedY) * width + c);
#pragma unroll
      for (int u = 0; u < 8; ++u) a = is_sum ? a + (double)v[u] : fmax(a, (double)v[u]);
    }
    for (; p < p1; p += kRedY) {
      const double v = (double)__ldg(partial + (size_t)p * width + c);
      a = is_sum ? a + v : fmax(a, v);
    }
  }
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < width) {
    for (int y = 1; y < kRedY; ++y) a = is_sum ? a + sm[y][threadIdx.x] : fmax(a, sm[y][threadIdx.x]);
    slice_out[(size_t)blockIdx.y * width + c] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const int t = atomicAdd(ticket, 1);
    s_last = (t == (int)(gridDim.x * gridDim.y) - 1);
    if (s_last) *ticket = 0;  // the slot is zero again for the next launch that draws it
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// vec[0..n) (global, written by this block) -> sum over ranks in rank order, in place.  All 1024 threads call it.
__device__ __forceinline__ void tail_exchange(double* vec, int n, const PeerCtx& pc, int* err) {
  if (pc.world <= 1) return;
  const int tid = threadIdx.y * kRedX + threadIdx.x, nt = kRedX * kRedY;
  PeerSlot* mine = pc.buf[pc.rank] + (pc.epoch % kPeerSlots);
  for (int i = tid; i < n; i += nt) mine->data[i] = vec[i];
  __threadfence_system();
  __syncthreads();
  if (tid == 0) st_release_sys(&mine->flag, pc.epoch);
  if (tid < pc.world && tid != pc.rank) {
    const unsigned long long* f = &(pc.buf[tid] + (pc.epoch % kPeerSlots))->flag;
    unsigned long long spins = 0;
    while (ld_acquire_sys(f) < pc.epoch) {
      if (++spins > (1ull << 28)) {
        if (err) atomicExch(err, 100 + tid);
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    double a = 0.0;
    for (int r = 0; r < pc.world; ++r) {
      const PeerSlot* s = pc.buf[r] + (pc.epoch % kPeerSlots);
      a += (r == pc.rank) ? vec[i] : ld_relaxed_sys_f64(&s->data[i]);
    }
    vec[i] = a;  // each thread rewrites only what it read itself; the peers read the published copy
  }
}

struct BnFwdTail {
  const float* partial;  // [n_partials][2C]
  int n_partials, C;
  double n_local;
  double* slices;  // [S][2C] scratch
  double* sums;    // [2C + 1] scratch
  const float *gamma, *beta;
  float eps, momentum;
  float *running_mean, *running_var;
  long long* nbt;
  float* stats;  // [4C + 2]: mean, invstd, scale, shift, then the (global) row count as one double
  int* ticket;
  int* err;
  PeerCtx peer;
};

__global__ void __launch_bounds__(kRedX* kRedY) k_bn_tail_fwd(const BnFwdTail t) {
  const int C = t.C, W = 2 * C;
  if (!tail_reduce(t.partial, t.n_partials, W, W, t.slices, t.ticket)) return;
  const int tid = threadIdx.y * kRedX + threadIdx.x, nt = kRedX * kRedY, S = gridDim.y;
  for (int c = tid; c < W; c += nt) {
    double a = 0.0;
#pragma unroll 8
    for (int s = 0; s < S; ++s) a += __ldcg(t.slices + (size_t)s * W + c);
    t.sums[c] = a;
  }
  if (tid == 0) t.sums[W] = t.n_local;
  __syncthreads();
  tail_exchange(t.sums, W + 1, t.peer, t.err);
  __syncthreads();
  const double count = t.sums[W];
  if (tid == 0) {
    *reinterpret_cast<double*>(t.stats + 4 * C) = count;
    if (t.nbt) *t.nbt += 1;
  }
  for (int c = tid; c < C; c += nt) {
    const double mean = t.sums[c] / count;
    double var = t.sums[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)t.eps));
    const float gm = t.gamma ? t.gamma[c] : 1.f, bt = t.beta ? t.beta[c] : 0.f;
    const float scale = gm * invstd;
    t.stats[c] = (float)mean;
    t.stats[C + c] = invstd;
    t.stats[2 * C + c] = scale;
    t.stats[3 * C + c] = bt - (float)mean * scale;
    if (t.running_mean) t.running_mean[c] = (1.f - t.momentum) * t.running_mean[c] + t.momentum * (float)mean;
    if (t.running_var) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      t.running_var[c] = (1.f - t.momentum) * t.running_var[c] + t.momentum * (float)unbiased;
    }
  }
}

struct BnBwdBranch {
  const float* gamma;
  const float* stats;  // [4C + 2] of the forward
  float *coef, *dgamma, *dbeta, *scale;  // coef [3C] scratch; scale float[4]
};
struct BnBwdTail {
  const float* partial;  // [n_partials][6C]: sums 3C | maxes 3C
  int n_partials, C, n_branches, want_gscale;
  double* slices;  // [S][6C]
  double* sums;    // [6C]: local sums 3C | global sums 3C
  float* maxes;    // [3C]
  BnBwdBranch br[2];
  float* scale_r;  // float[4] for the plain residual gradient g (want_gscale)
  int* ticket;
  int* err;
  PeerCtx peer;
};

__device__ __forceinline__ float block_max_1024(float v, float* s_w) {
  const int tid = threadIdx.y * kRedX + threadIdx.x;
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  __syncthreads();
  if ((tid & 31) == 0) s_w[tid >> 5] = v;
  __syncthreads();
  float m = 0.f;
  for (int w = 0; w < 32; ++w) m = fmaxf(m, s_w[w]);
  return m;
}
__device__ __forceinline__ void write_scale(float m, float* out) {
  float s = 1.f;
  if (m > 0.f && m <= 3.0e38f) {
    int e;
    frexpf(m, &e);
    s = ldexpf(1.f, 12 - e);
  }
  out[0] = s, out[1] = 1.f / s, out[2] = m;
}

__global__ void __launch_bounds__(kRedX* kRedY) k_bn_tail_bwd(const BnBwdTail t) {
  __shared__ float s_w[32];
  const int C = t.C, W = 6 * C;
  if (!tail_reduce(t.partial, t.n_partials, W, 3 * C, t.slices, t.ticket)) return;
  const int tid = threadIdx.y * kRedX + threadIdx.x, nt = kRedX * kRedY, S = gridDim.y;
  for (int c = tid; c < W; c += nt) {
    const bool is_sum = c < 3 * C;
    double a = 0.0;
#pragma unroll 8
    for (int s = 0; s < S; ++s) {
      const double v = __ldcg(t.slices + (size_t)s * W + c);
      a = is_sum ? a + v : fmax(a, v);
    }
    if (is_sum)
      t.sums[c] = a, t.sums[3 * C + c] = a;
    else
      t.maxes[c - 3 * C] = (float)a;
  }
  __syncthreads();
  tail_exchange(t.sums + 3 * C, 3 * C, t.peer, t.err);  // dx needs the sums over every rank; dgamma / dbeta stay local
  __syncthreads();
  for (int b = 0; b < t.n_branches; ++b) {
    const BnBwdBranch& br = t.br[b];
    const double count = *reinterpret_cast<const double*>(br.stats + 4 * C);
    float bound = 0.f;
    for (int c = tid; c < C; c += nt) {
      const float invstd = br.stats[C + c], gm = br.gamma ? br.gamma[c] : 1.f;
      const double sg = t.sums[3 * C + c], sgx = t.sums[3 * C + (1 + b) * C + c];
      const float a = gm * invstd;
      const float bb = (float)(-(double)gm * invstd * invstd * invstd * sgx / count);
      const float c0 = (float)(-(double)gm * invstd * sg / count);
      br.coef[c] = a, br.coef[C + c] = bb, br.coef[2 * C + c] = c0;
      if (br.dgamma) br.dgamma[c] = (float)(t.sums[(1 + b) * C + c] * invstd);
      if (br.dbeta) br.dbeta[c] = (float)t.sums[c];
      bound = fmaxf(bound, fabsf(a) * t.maxes[c] + fabsf(bb) * t.maxes[(1 + b) * C + c] + fabsf(c0));
    }
    const float m = block_max_1024(bound, s_w);
    if (tid == 0) write_scale(m, br.scale);
  }
  if (t.want_gscale) {
    float m = 0.f;
    for (int c = tid; c < C; c += nt) m = fmaxf(m, t.maxes[c]);
    m = block_max_1024(m, s_w);
    if (tid == 0) write_scale(m, t.scale_r);
  }
}

static int bn_blocks(int64_t n) {
  int64_t b = n / 128;
  if (b < 1) b = 1;
  if (b > kBnMaxBlocks) b = kBnMaxBlocks;
  return (int)b;
}
static int bn_check(int64_t n, int C, const char* who) {
  LG_CHECK_ARG(n >= 1 && C >= 4 && C % 4 == 0 && C <= 1024, "%s: needs n >= 1 and C a multiple of 4 in [4, 1024]", who);
  return LG_OK;
}
static size_t red_smem(int C) { return sizeof(float) * (size_t)(kBnThreads / (C >> 2)) * C; }

}  // namespace lg

using namespace lg;

// workspace: per-block partials of the widest reduction (6 C floats per block)
extern "C" size_t lg_bn_workspace(int64_t n, int32_t C) { return sizeof(float) * (size_t)kBnMaxBlocks * 6 * C + 256; }

/* sums double[2C + 1] = [sum x, sum x^2, n] over the n rows (this rank). */
extern "C" int lg_bn_stats(const float* x, int64_t n, int32_t C, double* sums, void* workspace, size_t workspace_bytes,
                           void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = bn_check(n, C, "lg_bn_stats");
  if (rc) return rc;
  LG_CHECK_ARG(x && sums && workspace && workspace_bytes >= lg_bn_workspace(n, C), "lg_bn_stats: bad pointers / workspace");
  const int nb = bn_blocks(n);
  k_bn_stats<<<nb, kBnThreads, red_smem(C), stream>>>(x, n, C, (float*)workspace);
  LG_LAUNCH_OK();
  k_bn_reduce<<<ceil_div(2 * C, kRedX), dim3(kRedX, kRedY), 0, stream>>>((const float*)workspace, nb, 2 * C, sums,
                                                                          (double)n);
  LG_LAUNCH_OK();
  return LG_OK;
}

/* stats_out float[4C] = [mean, invstd, scale, shift]; running stats / num_batches_tracked updated when non-NULL. */
extern "C" int lg_bn_finalize(const double* sums, double count, const double* count_dev, int32_t C, const float* gamma,
                              const float* beta,
                              float eps, float momentum, float* running_mean, float* running_var,
                              int64_t* num_batches_tracked, float* stats_out, void* stream_) {
  LG_CHECK_ARG(sums && stats_out && (count_dev || count >= 1.0) && C >= 1, "lg_bn_finalize: bad arguments");
  k_bn_finalize<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream_>>>(sums, count, count_dev, C, gamma, beta, eps, momentum,
                                                                     running_mean, running_var,
                                                                     (long long*)num_batches_tracked, stats_out);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" int lg_bn_apply(const float* x, const float* stats, const float* x2, const float* stats2, const float* res,
                           int32_t relu, int64_t n, int32_t C, float* y, void* y16, int32_t fmt, void* stream_) {
  int rc = bn_check(n, C, "lg_bn_apply");
  if (rc) return rc;
  LG_CHECK_ARG(x && stats && y && (!x2 || stats2), "lg_bn_apply: null pointer");
  const int64_t n4 = n * (C >> 2);
  k_bn_apply<<<(unsigned)ceil_div(n4, 256 * kApplyU), 256, 0, (cudaStream_t)stream_>>>(x, stats, x2, stats2, res, relu, n4, C, y,
                                                                             (unsigned short*)y16, fmt);
  LG_LAUNCH_OK();
  return LG_OK;
}

/* sums double[3C] = [sum g, sum g*(x-mean), sum g*(x2-mean2)], maxes float[3C] = [max|g|, max|x-mean|, max|x2-mean2|]. */
extern "C" int lg_bn_bwd_stats(const float* dy, const float* y, const float* x, const float* stats, const float* x2,
                               const float* stats2, int32_t relu, int64_t n, int32_t C, double* sums, float* maxes,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = bn_check(n, C, "lg_bn_bwd_stats");
  if (rc) return rc;
  LG_CHECK_ARG(dy && x && stats && sums && maxes && (!relu || y) && (!x2 || stats2), "lg_bn_bwd_stats: null pointer");
  LG_CHECK_ARG(workspace && workspace_bytes >= lg_bn_workspace(n, C), "lg_bn_bwd_stats: workspace too small");
  const int nb = bn_blocks(n);
  float* partial = (float*)workspace;
  if (x2)
    k_bn_bwd_stats<true, false, true><<<nb, kBnThreads, red_smem(C), stream>>>(dy, y, x, stats, x2, stats2, relu, n, C, partial, C >> 2);
  else
    k_bn_bwd_stats<true, false, false><<<nb, kBnThreads, red_smem(C), stream>>>(dy, y, x, stats, x2, stats2, relu, n, C, partial, C >> 2);
  LG_LAUNCH_OK();
  k_bn_bwd_reduce<<<ceil_div(6 * C, kRedX), dim3(kRedX, kRedY), 0, stream>>>(partial, nb, C, sums, maxes);
  LG_LAUNCH_OK();
  return LG_OK;
}

/* One branch (0: x, 1: x2) of the backward: coef float[3C] = [a, b, c] with dx = a*g + b*(x-mean) + c;
 * dgamma / dbeta from THIS rank's sums (DDP reduces parameter gradients), coefficients from the all-rank
 * sums; scale_out float[3] = power-of-two fp16 scale for dx from an upper bound of max|dx|, its inverse, the bound. */
extern "C" int lg_bn_bwd_finalize(const double* sums_local, const double* sums_global, const float* maxes, double count,
                                  const double* count_dev, int32_t C, int32_t branch, const float* gamma, const float* stats, float* coef,
                                  float* dgamma, float* dbeta, float* scale_out, void* stream_) {
  LG_CHECK_ARG(sums_local && sums_global && maxes && stats && coef && scale_out && (count_dev || count >= 1.0) && C >= 1 &&
                   C <= 1024 &&
                   (branch == 0 || branch == 1),
               "lg_bn_bwd_finalize: bad arguments");
  k_bn_bwd_finalize<<<1, (unsigned)round_up(C, 32), 0, (cudaStream_t)stream_>>>(sums_local, sums_global, maxes, count,
                                                                                  count_dev, C, branch, gamma, stats, coef, dgamma,
                                                                                  dbeta, scale_out);
  LG_LAUNCH_OK();
  return LG_OK;
}

/* fp16 scale of the residual gradient g itself (bound = max|g| = maxes[0..C)). */
extern "C" int lg_bn_bwd_gscale(const float* maxes, int32_t C, float* scale_out, void* stream_) {
  LG_CHECK_ARG(maxes && scale_out && C >= 1, "lg_bn_bwd_gscale: bad arguments");
  k_bn_gscale<<<1, 256, 0, (cudaStream_t)stream_>>>(maxes, C, scale_out);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" int lg_bn_bwd_apply(const float* dy, const float* y, const float* x, const float* stats, const float* coef,
                               const float* x2, const float* stats2, const float* coef2, int32_t relu, int64_t n,
                               int32_t C, float* dx, void* dx16, const float* scale, float* dx2, void* dx2_16,
                               const float* scale2, float* dres, void* dres16, const float* scale_r, int32_t fmt,
                               void* stream_) {
  int rc = bn_check(n, C, "lg_bn_bwd_apply");
  if (rc) return rc;
  LG_CHECK_ARG(dy && x && stats && coef && dx && (!relu || y) && (!x2 || (stats2 && coef2 && dx2)),
               "lg_bn_bwd_apply: null pointer");
  const int64_t n4 = n * (C >> 2);
  k_bn_bwd_apply<<<(unsigned)ceil_div(n4, 256 * kBwdU), 256, 0, (cudaStream_t)stream_>>>(
      dy, y, x, stats, coef, x2, stats2, coef2, relu, n4, C, dx, (unsigned short*)dx16, scale, dx2,
      (unsigned short*)dx2_16, scale2, dres, (unsigned short*)dres16, scale_r, fmt, C >> 2);
  LG_LAUNCH_OK();
  return LG_OK;
}

// ------------------------------------------------------------------------------------ fused layer entry points
namespace lg {

static int peer_ctx(const lgPeerCtx* p, PeerCtx* out, int extra_epoch, const char* who) {
  memset(out, 0, sizeof(*out));
  out->world = 1;
  if (!p || p->world <= 1) return LG_OK;
  LG_CHECK_ARG(p->bufs && p->world <= kPeerMaxWorld && p->rank >= 0 && p->rank < p->world && p->epoch >= 1,
               "%s: bad peer exchange context", who);
  for (int r = 0; r < p->world; ++r) out->buf[r] = (PeerSlot*)p->bufs[r];
  out->world = p->world;
  out->rank = p->rank;
  out->epoch = p->epoch + (unsigned long long)extra_epoch;
  return LG_OK;
}

constexpr int kTailSlicesMax = 64;
// slices of ~32 partial rows: the epilogue statistics of a 648 k-row layer are 20 k partial rows, and with 32 slices
// every thread of the tail walked ~80 of them one after the other (12 us per layer, profiles/r02_a_launch_summary.txt)
static int tail_slices(int64_t n_partials) {
  int64_t s = (n_partials + 31) / 32;
  return (int)(s < 1 ? 1 : (s > kTailSlicesMax ? kTailSlicesMax : s));
}

static size_t fwd_scratch(int64_t n, int C, int64_t n_stat_rows) {
  size_t b = 0;
  if (!n_stat_rows) b += arena_pad(sizeof(float) * (size_t)kBnMaxBlocks * 2 * C);
  b += arena_pad(sizeof(double) * kTailSlicesMax * 2 * C) + arena_pad(sizeof(double) * (2 * C + 1));
  return b;
}

// statistics of one branch: block partials (or the partials a convolution epilogue already wrote) -> tail kernel
static int bn_branch_forward(const lgBnBranch* b, int64_t n, int C, ArenaCursor* ar, const lgPeerCtx* peer, int peer_k,
                             int* err, cudaStream_t stream) {
  const float* partial = b->stat_partials;
  int64_t n_partials = b->n_stat_rows;
  if (!partial) {
    const int nb = bn_blocks(n);
    float* p = (float*)arena_take(ar, sizeof(float) * (size_t)kBnMaxBlocks * 2 * C);
    k_bn_stats<<<nb, kBnThreads, red_smem(C), stream>>>(b->x, n, C, p);
    LG_LAUNCH_OK();
    partial = p;
    n_partials = nb;
  }
  BnFwdTail t;
  t.partial = partial;
  t.n_partials = (int)n_partials;
  t.C = C;
  t.n_local = (double)n;
  const int S = tail_slices(n_partials);
  t.slices = (double*)arena_take(ar, sizeof(double) * kTailSlicesMax * 2 * C);
  t.sums = (double*)arena_take(ar, sizeof(double) * (2 * C + 1));
  t.gamma = b->gamma, t.beta = b->beta, t.eps = b->eps, t.momentum = b->momentum;
  t.running_mean = b->running_mean, t.running_var = b->running_var, t.nbt = (long long*)b->num_batches_tracked;
  t.stats = b->stats;
  t.err = err;
  int rc = counter_slot(&t.ticket);
  if (rc) return rc;
  rc = peer_ctx(peer, &t.peer, peer_k, "lg_bn_layer_forward");
  if (rc) return rc;
  k_bn_tail_fwd<<<dim3((unsigned)ceil_div(2 * C, kRedX), (unsigned)S), dim3(kRedX, kRedY), 0, stream>>>(t);
  LG_LAUNCH_OK();
  return LG_OK;
}

}  // namespace lg

extern "C" int lg_bn_layer_forward(const lgBnBranch* a, const lgBnBranch* b, const float* res, int32_t relu, int64_t n,
                                   int32_t C, float* y, void* y16, int32_t fmt, const lgPeerCtx* peer, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  // n = 0 is legal: under SyncBN a rank without voxels at this stride still takes part in the exchange
  int rc = bn_check(n > 0 ? n : 1, C, "lg_bn_layer_forward");
  if (rc) return rc;
  LG_CHECK_ARG(n >= 0 && a && a->stats && (!b || b->stats) && (n == 0 || (a->x && y && (!b || b->x))),
               "lg_bn_layer_forward: null pointer");
  LG_CHECK_ARG(n < ((int64_t)1 << 31) * 32, "lg_bn_layer_forward: too many rows");
  int sm = 0;
  int* err = nullptr;
  rc = tc_runtime(&sm, &err);
  if (rc) return rc;
  ArenaCursor ar;
  rc = arena_begin(stream, fwd_scratch(n, C, a->stat_partials ? a->n_stat_rows : 0) +
                               (b ? fwd_scratch(n, C, b->stat_partials ? b->n_stat_rows : 0) : 0),
                   &ar);
  if (rc) return rc;
  rc = bn_branch_forward(a, n, C, &ar, peer, 0, err, stream);
  if (rc) return rc;
  if (b) {
    rc = bn_branch_forward(b, n, C, &ar, peer, 1, err, stream);
    if (rc) return rc;
  }
  const int64_t n4 = n * (C >> 2);
  if (n4 > 0) {
    k_bn_apply<<<(unsigned)ceil_div(n4, 256 * kApplyU), 256, 0, stream>>>(a->x, a->stats, b ? b->x : nullptr,
                                                                          b ? b->stats : nullptr, res, relu, n4, C, y,
                                                                          (unsigned short*)y16, fmt);
    LG_LAUNCH_OK();
  }
  return LG_OK;
}

extern "C" int lg_bn_layer_backward(const float* dy, int64_t dy_ld, const float* y, int32_t relu, int64_t n, int32_t C,
                                    const lgBnBwdBranch* a, const lgBnBwdBranch* b, float* dres, int32_t fmt,
                                    float* scales /* [12]: a, b, residual */, const lgPeerCtx* peer, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = bn_check(n > 0 ? n : 1, C, "lg_bn_layer_backward");
  if (rc) return rc;
  LG_CHECK_ARG(n >= 0 && a && a->stats && scales && (!b || b->stats) &&
                   (n == 0 || (dy && a->x && (a->dx || a->dx16) && (!b || (b->x && (b->dx || b->dx16))))),
               "lg_bn_layer_backward: null pointer");
  LG_CHECK_ARG(!(relu && !y && (dres || b)),
               "lg_bn_layer_backward: a layer with a residual or a second branch needs y for its ReLU mask");
  LG_CHECK_ARG(dy_ld >= C && dy_ld % 4 == 0 && ((uintptr_t)dy & 15) == 0,
               "lg_bn_layer_backward: dy needs a row pitch >= C that is a multiple of 4 floats and 16-byte alignment");
  int sm = 0;
  int* err = nullptr;
  rc = tc_runtime(&sm, &err);
  if (rc) return rc;
  ArenaCursor ar;
  rc = arena_begin(stream,
                   arena_pad(sizeof(float) * (size_t)kBnMaxBlocks * 6 * C) + arena_pad(sizeof(double) * kTailSlicesMax * 6 * C) +
                       arena_pad(sizeof(double) * 6 * C) + arena_pad(sizeof(float) * 3 * C) +
                       2 * arena_pad(sizeof(float) * 3 * C),
                   &ar);
  if (rc) return rc;
  const int nb = bn_blocks(n);
  float* partial = (float*)arena_take(&ar, sizeof(float) * (size_t)kBnMaxBlocks * 6 * C);
  const bool recomp = relu && !y;  // see lidog_b200.h: y == NULL says "the forward had no residual"
  const float* x2p = b ? b->x : nullptr;
  const float* st2p = b ? b->stats : nullptr;
  const int64_t ld4 = dy_ld >> 2;
#define LG_BWD_STATS(D, R, H) \
  k_bn_bwd_stats<D, R, H><<<nb, kBnThreads, red_smem(C), stream>>>(dy, y, a->x, a->stats, x2p, st2p, relu, n, C, partial, ld4)
  const int variant = (dy_ld == C ? 4 : 0) | (recomp ? 2 : 0) | (b ? 1 : 0);
  switch (variant) {
    case 0: LG_BWD_STATS(false, false, false); break;
    case 1: LG_BWD_STATS(false, false, true); break;
    case 2: LG_BWD_STATS(false, true, false); break;
    case 4: LG_BWD_STATS(true, false, false); break;
    case 5: LG_BWD_STATS(true, false, true); break;
    case 6: LG_BWD_STATS(true, true, false); break;
    default: break;  // recomputation with two branches is refused above
  }
#undef LG_BWD_STATS
  LG_LAUNCH_OK();
  BnBwdTail t;
  memset(&t, 0, sizeof(t));
  t.partial = partial;
  t.n_partials = nb;
  t.C = C;
  t.n_branches = b ? 2 : 1;
  t.want_gscale = 0;  // the residual gradient feeds an elementwise backward, never a 16-bit convolution operand
  t.slices = (double*)arena_take(&ar, sizeof(double) * kTailSlicesMax * 6 * C);
  t.sums = (double*)arena_take(&ar, sizeof(double) * 6 * C);
  t.maxes = (float*)arena_take(&ar, sizeof(float) * 3 * C);
  const lgBnBwdBranch* brs[2] = {a, b};
  for (int i = 0; i < t.n_branches; ++i) {
    t.br[i].gamma = brs[i]->gamma;
    t.br[i].stats = brs[i]->stats;
    t.br[i].coef = (float*)arena_take(&ar, sizeof(float) * 3 * C);
    t.br[i].dgamma = brs[i]->dgamma;
    t.br[i].dbeta = brs[i]->dbeta;
    t.br[i].scale = scales + 4 * i;
  }
  t.scale_r = scales + 8;
  t.err = err;
  rc = counter_slot(&t.ticket);
  if (rc) return rc;
  rc = peer_ctx(peer, &t.peer, 0, "lg_bn_layer_backward");
  if (rc) return rc;
  k_bn_tail_bwd<<<dim3((unsigned)ceil_div(6 * C, kRedX), (unsigned)tail_slices(nb)), dim3(kRedX, kRedY), 0, stream>>>(t);
  LG_LAUNCH_OK();
  const int64_t n4 = n * (C >> 2);
  const bool use16 = a->dx16 != nullptr;
  if (n4 > 0)
    k_bn_bwd_apply<<<(unsigned)ceil_div(n4, 256 * kBwdU), 256, 0, stream>>>(
      dy, y, a->x, a->stats, t.br[0].coef, b ? b->x : nullptr, b ? b->stats : nullptr, b ? t.br[1].coef : nullptr, relu,
      n4, C, a->dx, (unsigned short*)a->dx16, use16 ? scales : nullptr, b ? b->dx : nullptr,
      b ? (unsigned short*)b->dx16 : nullptr, (b && b->dx16) ? scales + 4 : nullptr, dres, nullptr, nullptr, fmt,
      dy_ld >> 2);
  LG_LAUNCH_OK();
  return LG_OK;
}
