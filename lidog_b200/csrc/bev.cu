// Fused point-to-BEV projection (forward + backward).
// Reference: MinkUNetBaseBEV.sparse2super + filter_bounds, utils/models/minkunet_bev.py:158-230 --
// scatter of voxel features into a dense (H, W, C) image, a raw re-view of that memory as (C, H, W)
// and MaxPool2d(5, 3, 1).  The dense tensor (1.5 GB per sample at H = W = 2000, C = 96) is never
// materialised: a pixel -> row map (4 B per pixel) is built by one scatter, then one CTA per output
// row (b, c', i) rebuilds the pk scrambled input rows it needs in shared memory -- in flat terms
// they are ONE contiguous run of pk*W floats of the dense buffer, i.e. a run of ~pk*W/C consecutive
// pixels -- and reduces them.  Output rows are written exactly once, coalesced, without atomics;
// rows whose pixel run is empty are stored as zeros without touching the features.
// Pixel maths replicates the reference's three separately rounded float32 operations.
#include "common.cuh"

namespace lg {

struct BevWorkspace {
  int* pix_of_row;  // [n]      flat pixel (b*H*W + py*W + px) or -1
  int* next;        // [n]      chain of rows sharing a pixel (LG_BEV_MAX)
  int* pixmap;      // [B*H*W]  winner row (LAST) / chain head (MAX) / -1
  size_t total;
};

static BevWorkspace carve_bev(void* base, int64_t n, int64_t pixels) {
  BevWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  w.pix_of_row = (int*)take(sizeof(int) * (size_t)(n > 0 ? n : 1));
  w.next = (int*)take(sizeof(int) * (size_t)(n > 0 ? n : 1));
  w.pixmap = (int*)take(sizeof(int) * (size_t)pixels);
  w.total = off;
  return w;
}

__global__ void __launch_bounds__(256)
    k_bev_pixels(const int4* __restrict__ coords, int64_t n, int batch_size, float bound, float vs, int H, int W,
                 int policy, int* __restrict__ pix_of_row, int* __restrict__ next, int* __restrict__ pixmap) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int4 c = coords[r];
  const float lo = -bound, hi = bound;
  const float x = __fmul_rn((float)c.y, vs), y = __fmul_rn((float)c.z, vs);
  int pix = -1;
  if ((unsigned)c.x < (unsigned)batch_size && lo < x && x < hi && lo < y && y < hi) {
    int px = (int)floorf(__fdiv_rn(__fsub_rn(x, lo), vs));
    int py = (int)floorf(__fsub_rn((float)H, __fdiv_rn(__fsub_rn(y, lo), vs))) - 1;
    if (py < 0) py += H;  // python negative-index wrap
    if (px >= 0 && px < W && py >= 0 && py < H) pix = (c.x * H + py) * W + px;
  }
  pix_of_row[r] = pix;
  next[r] = -1;
  if (pix >= 0) {
    if (policy == LG_BEV_LAST)
      atomicMax(&pixmap[pix], (int)r);
    else
      next[r] = atomicExch(&pixmap[pix], (int)r);
  }
}

// value of dense cell (pixel p, channel ch): LAST -> winner row, MAX -> max over the chain
__device__ __forceinline__ float cell_value(int head, int ch, int C, int policy, const float* __restrict__ feats,
                                            const int* __restrict__ next) {
  if (head < 0) return 0.f;
  float v = __ldg(feats + (int64_t)head * C + ch);
  if (policy == LG_BEV_MAX) {
    for (int r = next[head]; r >= 0; r = next[r]) v = fmaxf(v, __ldg(feats + (int64_t)r * C + ch));
  }
  return v;
}

// first (smallest) row of the chain attaining the max for channel ch
__device__ __forceinline__ int cell_argrow(int head, int ch, int C, int policy, const float* __restrict__ feats,
                                           const int* __restrict__ next) {
  if (policy == LG_BEV_LAST || head < 0) return head;
  float best = __ldg(feats + (int64_t)head * C + ch);
  int arg = head;
  for (int r = next[head]; r >= 0; r = next[r]) {
    const float v = __ldg(feats + (int64_t)r * C + ch);
    if (v > best || (v == best && r < arg)) {
      best = v;
      arg = r;
    }
  }
  return arg;
}

struct RowGeom {
  int b, cp, i;        // sample, scrambled channel, output row
  int h_lo, rows;      // first valid scrambled row of the window, number of valid rows
  int64_t m_lo;        // flat offset (within the sample's H*W*C buffer) of scrambled row h_lo
  int64_t cells;       // rows * W
};

__device__ __forceinline__ RowGeom row_geom(int C, int H, int W, int h_out, int pk, int ps, int pp) {
  RowGeom g;
  int idx = blockIdx.x;
  g.i = idx % h_out;
  idx /= h_out;
  g.cp = idx % C;
  g.b = idx / C;
  int h0 = g.i * ps - pp;
  g.h_lo = max(h0, 0);
  int h_hi = min(h0 + pk, H);
  g.rows = h_hi - g.h_lo;
  g.m_lo = ((int64_t)g.cp * H + g.h_lo) * W;
  g.cells = (int64_t)g.rows * W;
  return g;
}

// fills sm[0..cells) with the scrambled rows; returns false (uniformly) when the pixel run is empty
__device__ __forceinline__ bool load_rows(const RowGeom& g, int C, int H, int W, int policy,
                                          const float* __restrict__ feats, const int* __restrict__ next,
                                          const int* __restrict__ pixmap, float* sm) {
  const int* pm = pixmap + (int64_t)g.b * H * W;
  const int64_t p_lo = g.m_lo / C, p_hi = (g.m_lo + g.cells - 1) / C;
  int any = 0;
  for (int64_t p = p_lo + threadIdx.x; p <= p_hi; p += blockDim.x) any |= (__ldg(pm + p) >= 0);
  if (!__syncthreads_or(any)) return false;
  for (int64_t e = threadIdx.x; e < g.cells; e += blockDim.x) {
    const int64_t m = g.m_lo + e;
    const int64_t p = m / C;
    const int ch = (int)(m - p * C);
    sm[e] = cell_value(__ldg(pm + p), ch, C, policy, feats, next);
  }
  __syncthreads();
  return true;
}

__global__ void __launch_bounds__(256)
    k_bev_pool_fwd(const float* __restrict__ feats, const int* __restrict__ next, const int* __restrict__ pixmap, int C,
                   int H, int W, int h_out, int w_out, int pk, int ps, int pp, int policy, float* __restrict__ out) {
  extern __shared__ float sm[];
  const RowGeom g = row_geom(C, H, W, h_out, pk, ps, pp);
  float* orow = out + (((int64_t)g.b * C + g.cp) * h_out + g.i) * w_out;
  if (!load_rows(g, C, H, W, policy, feats, next, pixmap, sm)) {
    for (int j = threadIdx.x; j < w_out; j += blockDim.x) orow[j] = 0.f;
    return;
  }
  for (int j = threadIdx.x; j < w_out; j += blockDim.x) {
    const int w0 = max(j * ps - pp, 0), w1 = min(j * ps - pp + pk, W);
    float best = -INFINITY;
    for (int r = 0; r < g.rows; ++r)
      for (int w = w0; w < w1; ++w) best = fmaxf(best, sm[r * W + w]);
    orow[j] = best;
  }
}

__global__ void __launch_bounds__(256)
    k_bev_pool_bwd(const float* __restrict__ feats, const int* __restrict__ next, const int* __restrict__ pixmap, int C,
                   int H, int W, int h_out, int w_out, int pk, int ps, int pp, int policy,
                   const float* __restrict__ grad_out, float* __restrict__ grad_feats) {
  extern __shared__ float sm[];
  const RowGeom g = row_geom(C, H, W, h_out, pk, ps, pp);
  if (!load_rows(g, C, H, W, policy, feats, next, pixmap, sm)) return;
  const float* grow = grad_out + (((int64_t)g.b * C + g.cp) * h_out + g.i) * w_out;
  // first arg-max (scan order: row, then column; strictly greater) of each window
  constexpr int kMaxPerThread = 8;
  int arg[kMaxPerThread];
  float gv[kMaxPerThread];
  int cnt = 0;
  for (int j = threadIdx.x; j < w_out && cnt < kMaxPerThread; j += blockDim.x, ++cnt) {
    const int w0 = max(j * ps - pp, 0), w1 = min(j * ps - pp + pk, W);
    float best = -INFINITY;
    int a = -1;
    for (int r = 0; r < g.rows; ++r)
      for (int w = w0; w < w1; ++w) {
        const float v = sm[r * W + w];
        if (v > best) {
          best = v;
          a = r * W + w;
        }
      }
    arg[cnt] = a;
    gv[cnt] = grow[j];
  }
  __syncthreads();
  for (int64_t e = threadIdx.x; e < g.cells; e += blockDim.x) sm[e] = 0.f;
  __syncthreads();
  for (int q = 0; q < cnt; ++q)
    if (arg[q] >= 0 && gv[q] != 0.f) atomicAdd(&sm[arg[q]], gv[q]);
  __syncthreads();
  const int* pm = pixmap + (int64_t)g.b * H * W;
  for (int64_t e = threadIdx.x; e < g.cells; e += blockDim.x) {
    const float gsum = sm[e];
    if (gsum == 0.f) continue;
    const int64_t m = g.m_lo + e;
    const int64_t p = m / C;
    const int ch = (int)(m - p * C);
    const int row = cell_argrow(__ldg(pm + p), ch, C, policy, feats, next);
    if (row >= 0) atomicAdd(grad_feats + (int64_t)row * C + ch, gsum);
  }
}

// LG_BEV_LAST: every row that wrote the pixel receives the pixel's gradient (index_put_ backward)
__global__ void __launch_bounds__(256)
    k_bev_dup_grad(const int* __restrict__ pix_of_row, const int* __restrict__ pixmap, int64_t n, int C,
                   float* __restrict__ grad_feats) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * C) return;
  const int64_t r = e / C;
  const int pix = pix_of_row[r];
  if (pix < 0) return;
  const int w = pixmap[pix];
  if (w != (int)r) grad_feats[e] = grad_feats[(int64_t)w * C + (e - r * C)];
}

__global__ void k_fill_i32_bev(int* p, int64_t n, int v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace lg

using namespace lg;

static int bev_check(int64_t n, int C, int B, int H, int W, int pk, int ps, int pp, int policy, const char* who) {
  LG_CHECK_ARG(n >= 0 && C >= 1 && B >= 1 && H >= 1 && W >= 1, "%s: bad sizes", who);
  LG_CHECK_ARG(pk >= 1 && ps >= 1 && pp >= 0 && 2 * pp <= pk, "%s: bad pooling parameters", who);
  LG_CHECK_ARG(policy == LG_BEV_LAST || policy == LG_BEV_MAX, "%s: bad policy", who);
  LG_CHECK_ARG((int64_t)B * H * W < ((int64_t)1 << 31), "%s: batch*H*W exceeds int32", who);
  LG_CHECK_ARG((size_t)pk * W * sizeof(float) <= 200 * 1024, "%s: pk*W too large for shared memory", who);
  return LG_OK;
}

extern "C" size_t lg_bev_workspace(int64_t n, int32_t batch_size, int32_t H, int32_t W) {
  return carve_bev(nullptr, n, (int64_t)batch_size * H * W).total;
}

extern "C" int lg_bev_forward(const int32_t* coords4, const float* feats, int64_t n, int32_t C, int32_t batch_size,
                              float bound, float voxel_size, int32_t H, int32_t W, int32_t pk, int32_t ps, int32_t pp,
                              int32_t policy, float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = bev_check(n, C, batch_size, H, W, pk, ps, pp, policy, "lg_bev_forward");
  if (rc) return rc;
  const int64_t pixels = (int64_t)batch_size * H * W;
  BevWorkspace w = carve_bev(workspace, n, pixels);
  LG_CHECK_ARG(workspace && workspace_bytes >= w.total, "lg_bev_forward: workspace too small");
  LG_CHECK_ARG(out && (n == 0 || (coords4 && feats)), "lg_bev_forward: null pointer");
  const int h_out = (H + 2 * pp - pk) / ps + 1, w_out = (W + 2 * pp - pk) / ps + 1;
  LG_CUDA_OK(cudaMemsetAsync(w.pixmap, 0xFF, sizeof(int) * (size_t)pixels, stream));
  if (n > 0) {
    k_bev_pixels<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>((const int4*)coords4, n, batch_size, bound,
                                                                 voxel_size, H, W, policy, w.pix_of_row, w.next,
                                                                 w.pixmap);
    LG_LAUNCH_OK();
  }
  const size_t smem = (size_t)pk * W * sizeof(float);
  LG_CUDA_OK(cudaFuncSetAttribute(k_bev_pool_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (int64_t)batch_size * C * h_out;
  LG_CHECK_ARG(blocks < ((int64_t)1 << 31), "lg_bev_forward: too many output rows");
  k_bev_pool_fwd<<<(unsigned)blocks, 256, smem, stream>>>(feats, w.next, w.pixmap, C, H, W, h_out, w_out, pk, ps, pp,
                                                          policy, out);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" int lg_bev_backward(const int32_t* coords4, const float* feats, int64_t n, int32_t C, int32_t batch_size,
                               int32_t H, int32_t W, int32_t pk, int32_t ps, int32_t pp, int32_t policy,
                               const float* grad_out, float* grad_feats, const void* workspace, size_t workspace_bytes,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = bev_check(n, C, batch_size, H, W, pk, ps, pp, policy, "lg_bev_backward");
  if (rc) return rc;
  const int64_t pixels = (int64_t)batch_size * H * W;
  BevWorkspace w = carve_bev(const_cast<void*>(workspace), n, pixels);
  LG_CHECK_ARG(workspace && workspace_bytes >= w.total, "lg_bev_backward: workspace too small");
  if (n == 0) return LG_OK;
  LG_CHECK_ARG(feats && grad_out && grad_feats, "lg_bev_backward: null pointer");
  const int h_out = (H + 2 * pp - pk) / ps + 1, w_out = (W + 2 * pp - pk) / ps + 1;
  LG_CHECK_ARG(w_out <= 8 * 256, "lg_bev_backward: output rows wider than 2048 are not supported");
  LG_CUDA_OK(cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)n * C, stream));
  const size_t smem = (size_t)pk * W * sizeof(float);
  LG_CUDA_OK(cudaFuncSetAttribute(k_bev_pool_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (int64_t)batch_size * C * h_out;
  k_bev_pool_bwd<<<(unsigned)blocks, 256, smem, stream>>>(feats, w.next, w.pixmap, C, H, W, h_out, w_out, pk, ps, pp,
                                                          policy, grad_out, grad_feats);
  LG_LAUNCH_OK();
  if (policy == LG_BEV_LAST) {
    k_bev_dup_grad<<<(unsigned)ceil_div(n * C, 256), 256, 0, stream>>>(w.pix_of_row, w.pixmap, n, C, grad_feats);
    LG_LAUNCH_OK();
  }
  return LG_OK;
}
