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

#include "common.cuh"

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
    float4 o = make_float4(fmaf(v[u].x, sc.x, sh.x), fmaf(v[u].y, sc.y, sh.y), fmaf(v[u].z, sc.z, sh.z),
                           fmaf(v[u].w, sc.w, sh.w));
    if (x2) {
      const float4 s2 = *reinterpret_cast<const float4*>(st2 + 2 * C + c), h2 = *reinterpret_cast<const float4*>(st2 + 3 * C + c);
      o.x += fmaf(v2[u].x, s2.x, h2.x), o.y += fmaf(v2[u].y, s2.y, h2.y), o.z += fmaf(v2[u].z, s2.z, h2.z),
          o.w += fmaf(v2[u].w, s2.w, h2.w);
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
__global__ void __launch_bounds__(kBnThreads)
    k_bn_bwd_stats(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                   const float* __restrict__ st, const float* __restrict__ x2, const float* __restrict__ st2, int relu,
                   int64_t n, int C, float* __restrict__ partial) {
  extern __shared__ float sm[];
  const RedGeom g = red_geom(n, C);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sg = z, sgx = z, sgx2 = z, mg = z, mx = z, mx2 = z;
  if (g.active) {
    const float4 mu = *reinterpret_cast<const float4*>(st + g.cq * 4);
    const float4 mu2 = x2 ? *reinterpret_cast<const float4*>(st2 + g.cq * 4) : z;
#pragma unroll 4
    for (int64_t r = g.r0 + g.ry; r < g.r1; r += g.rpp) {
      const int64_t i = r * (C >> 2) + g.cq;
      const float4 gy = masked(__ldg(reinterpret_cast<const float4*>(dy) + i),
                               relu ? __ldg(reinterpret_cast<const float4*>(y) + i) : z, relu);
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      const float4 d = make_float4(v.x - mu.x, v.y - mu.y, v.z - mu.z, v.w - mu.w);
      sg.x += gy.x, sg.y += gy.y, sg.z += gy.z, sg.w += gy.w;
      sgx.x += gy.x * d.x, sgx.y += gy.y * d.y, sgx.z += gy.z * d.z, sgx.w += gy.w * d.w;
      mg = make_float4(fmaxf(mg.x, fabsf(gy.x)), fmaxf(mg.y, fabsf(gy.y)), fmaxf(mg.z, fabsf(gy.z)), fmaxf(mg.w, fabsf(gy.w)));
      mx = make_float4(fmaxf(mx.x, fabsf(d.x)), fmaxf(mx.y, fabsf(d.y)), fmaxf(mx.z, fabsf(d.z)), fmaxf(mx.w, fabsf(d.w)));
      if (x2) {
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
                   int fmt) {
  const int64_t base = (int64_t)blockIdx.x * (256 * kBwdU) + threadIdx.x;
  const int Cq = C >> 2;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gd[kBwdU], gy[kBwdU], vx[kBwdU], vx2[kBwdU];
#pragma unroll
  for (int u = 0; u < kBwdU; ++u) {
    const int64_t i = base + u * 256;
    if (i < n4) {
      gd[u] = __ldg(reinterpret_cast<const float4*>(dy) + i);
      gy[u] = relu ? __ldg(reinterpret_cast<const float4*>(y) + i) : z;
      vx[u] = __ldg(reinterpret_cast<const float4*>(x) + i);
      if (x2) vx2[u] = __ldg(reinterpret_cast<const float4*>(x2) + i);
    }
  }
  const float s1 = (dx16 && scale) ? scale[0] : 1.f, s2 = (dx2_16 && scale2) ? scale2[0] : 1.f,
              sr = (dres16 && scale_r) ? scale_r[0] : 1.f;
#pragma unroll
  for (int u = 0; u < kBwdU; ++u) {
    const int64_t i = base + u * 256;
    if (i >= n4) break;
    const int c = (int)(i % Cq) * 4;
    const float4 g = masked(gd[u], gy[u], relu);
    {
      const float4 v = vx[u];
      const float4 mu = *reinterpret_cast<const float4*>(st + c);
      const float4 a = *reinterpret_cast<const float4*>(coef + c), b = *reinterpret_cast<const float4*>(coef + C + c),
                   k = *reinterpret_cast<const float4*>(coef + 2 * C + c);
      const float4 o = make_float4(fmaf(a.x, g.x, fmaf(b.x, v.x - mu.x, k.x)), fmaf(a.y, g.y, fmaf(b.y, v.y - mu.y, k.y)),
                                   fmaf(a.z, g.z, fmaf(b.z, v.z - mu.z, k.z)), fmaf(a.w, g.w, fmaf(b.w, v.w - mu.w, k.w)));
      reinterpret_cast<float4*>(dx)[i] = o;
      if (dx16) store16x4(dx16 + i * 4, make_float4(o.x * s1, o.y * s1, o.z * s1, o.w * s1), fmt);
    }
    if (x2) {
      const float4 v = vx2[u];
      const float4 mu = *reinterpret_cast<const float4*>(st2 + c);
      const float4 a = *reinterpret_cast<const float4*>(coef2 + c), b = *reinterpret_cast<const float4*>(coef2 + C + c),
                   k = *reinterpret_cast<const float4*>(coef2 + 2 * C + c);
      const float4 o = make_float4(fmaf(a.x, g.x, fmaf(b.x, v.x - mu.x, k.x)), fmaf(a.y, g.y, fmaf(b.y, v.y - mu.y, k.y)),
                                   fmaf(a.z, g.z, fmaf(b.z, v.z - mu.z, k.z)), fmaf(a.w, g.w, fmaf(b.w, v.w - mu.w, k.w)));
      reinterpret_cast<float4*>(dx2)[i] = o;
      if (dx2_16) store16x4(dx2_16 + i * 4, make_float4(o.x * s2, o.y * s2, o.z * s2, o.w * s2), fmt);
    }
    if (dres) {
      reinterpret_cast<float4*>(dres)[i] = g;
      if (dres16) store16x4(dres16 + i * 4, make_float4(g.x * sr, g.y * sr, g.z * sr, g.w * sr), fmt);
    }
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
  k_bn_bwd_stats<<<nb, kBnThreads, red_smem(C), stream>>>(dy, y, x, stats, x2, stats2, relu, n, C, partial);
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
      (unsigned short*)dx2_16, scale2, dres, (unsigned short*)dres16, scale_r, fmt);
  LG_LAUNCH_OK();
  return LG_OK;
}
