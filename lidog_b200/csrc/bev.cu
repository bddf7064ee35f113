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

// ---------------------------------------------------------------------------------------------
// Tiled pooling.  A CTA owns the output tile (b, 32 scrambled channels c', kIB rows i, kJC columns j).
// Per (c', h') the cells of its windows are one contiguous run of <= (kJC-1)*ps+pk floats of the
// never-materialised dense buffer, i.e. a handful of consecutive PIXELS, and only ~2 % of the pixels
// hold a voxel.  Phase 1 looks those pixels up once (pixel -> row heads in shared memory, plus one
// "row has a voxel" bit per (c', h')).  Phase 2 evaluates every window in the reference's scan order:
// windows whose rows are all empty cost one shared-memory read, empty cells are zeros, occupied cells
// read their feature value (L1/L2 hits: neighbouring windows share pixels).  Results are staged in
// shared memory and leave coalesced in NCHW or NHWC (channels_last, what cuDNN's tensor-op
// convolutions consume).  No atomics in the forward; every output is written exactly once.
constexpr int kCG = 32;   // scrambled channels per CTA
constexpr int kIB = 8;    // output rows per CTA
constexpr int kJC = 32;   // output columns per CTA
constexpr int kPitch = kCG + 1;

struct TileGeom {
  int b, c0, i0, j0;  // sample, first scrambled channel, first output row / column
  int ncg, nib, njc;  // valid extents of the tile
  int h_lo, h_hi, w_lo, w_hi;  // cell region (clipped to the image), hi exclusive
};

__device__ __forceinline__ TileGeom tile_geom(int C, int H, int W, int h_out, int w_out, int pk, int ps, int pp) {
  TileGeom g;
  const int n_jc = (w_out + kJC - 1) / kJC, n_ib = (h_out + kIB - 1) / kIB, n_cg = (C + kCG - 1) / kCG;
  int idx = blockIdx.x;
  g.j0 = (idx % n_jc) * kJC;
  idx /= n_jc;
  g.i0 = (idx % n_ib) * kIB;
  idx /= n_ib;
  g.c0 = (idx % n_cg) * kCG;
  g.b = idx / n_cg;
  g.ncg = min(kCG, C - g.c0);
  g.nib = min(kIB, h_out - g.i0);
  g.njc = min(kJC, w_out - g.j0);
  g.h_lo = max(g.i0 * ps - pp, 0);
  g.h_hi = min((g.i0 + g.nib - 1) * ps - pp + pk, H);
  g.w_lo = max(g.j0 * ps - pp, 0);
  g.w_hi = min((g.j0 + g.njc - 1) * ps - pp + pk, W);
  return g;
}

struct TileSmem {
  float* out;         // [kIB][kJC][kPitch] staged results (forward) / gradients (backward)
  int* heads;         // [pairs][max_pix] pixel -> row head (or -1)
  int* plo;           // [pairs] first pixel of the run
  uint32_t* rowmask;  // [kCG] bit (h - h_lo): scrambled row h of channel c' holds a voxel
};
__device__ __forceinline__ TileSmem carve_tile(unsigned char* base, int pairs_cap, int max_pix) {
  TileSmem t;
  t.out = reinterpret_cast<float*>(base);
  t.heads = reinterpret_cast<int*>(t.out + kIB * kJC * kPitch);
  t.plo = t.heads + pairs_cap * max_pix;
  t.rowmask = reinterpret_cast<uint32_t*>(t.plo + pairs_cap);
  return t;
}

// Phase 1.  Returns false (uniformly) when no pixel of the tile is occupied.
__device__ __forceinline__ bool tile_lookup(const TileGeom& g, const TileSmem& t, int C, int H, int W, int max_pix,
                                            const int* __restrict__ pixmap) {
  const int nrows = g.h_hi - g.h_lo;
  const int* pm = pixmap + (int64_t)g.b * H * W;
  int any = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int cl = warp; cl < g.ncg; cl += nwarps) {  // lane = scrambled row (nrows <= 32, host-checked)
    const int hr = lane;
    int row_any = 0;
    if (hr < nrows) {
      const int q = cl * nrows + hr;
      const int m_lo = ((g.c0 + cl) * H + g.h_lo + hr) * W + g.w_lo, m_hi = m_lo + (g.w_hi - g.w_lo);
      const int p_lo = m_lo / C, p_hi = (m_hi - 1) / C;
      t.plo[q] = p_lo;
      for (int d = 0; d < max_pix; ++d) {
        const int head = (p_lo + d <= p_hi) ? __ldg(pm + p_lo + d) : -1;
        t.heads[q * max_pix + d] = head;
        row_any |= (head >= 0);
      }
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, row_any);
    if (lane == 0) t.rowmask[cl] = bits;
    any |= (bits != 0);
  }
  return __syncthreads_or(any) != 0;
}

// Scan of one window in the reference's order (rows, then columns; strict > keeps the first maximum).
// Returns the pooled value; *arg_m = flat dense index of the winning cell, *arg_row = its voxel row
// (-1 when an empty cell wins).  WANT_ARG = false skips the bookkeeping.
template <bool WANT_ARG>
__device__ __forceinline__ float window_scan(const TileGeom& g, const TileSmem& t, int cl, int i, int j, int C, int H,
                                             int W, int pk, int ps, int pp, int max_pix, int policy,
                                             const float* __restrict__ feats, const int* __restrict__ next,
                                             int* arg_row, int* arg_ch) {
  const int nrows = g.h_hi - g.h_lo;
  const int h0 = i * ps - pp, w0 = j * ps - pp;
  const int ha = max(h0, 0), hb = min(h0 + pk, H), wa = max(w0, 0), wb = min(w0 + pk, W);
  const uint32_t rbits = (t.rowmask[cl] >> (ha - g.h_lo)) & ((1u << (hb - ha)) - 1u);
  if (WANT_ARG) *arg_row = -1;
  if (!rbits) return 0.f;  // every cell is an empty zero; the first one wins
  float best = -INFINITY;
  for (int h = ha; h < hb; ++h) {
    if (!((rbits >> (h - ha)) & 1u)) {  // an all-empty row: zeros
      if (0.f > best) {
        best = 0.f;
        if (WANT_ARG) *arg_row = -1;
      }
      continue;
    }
    const int q = cl * nrows + (h - g.h_lo);
    const int p_lo = t.plo[q];
    const int x0 = ((g.c0 + cl) * H + h) * W - p_lo * C;  // flat index of column 0 relative to pixel p_lo
    int ch = x0 + wa, d = 0;
    while (ch >= C) {  // <= max_pix - 1 steps, once per row
      ch -= C;
      ++d;
    }
    for (int w = wa; w < wb;) {  // the <= pk cells of this row lie in one or two pixels: one head test per pixel
      const int span = min(wb - w, C - ch);
      const int head = t.heads[q * max_pix + d];
      if (head < 0) {  // a run of empty zeros; only its first cell can take the lead
        if (0.f > best) {
          best = 0.f;
          if (WANT_ARG) *arg_row = -1;
        }
      } else {
        for (int k = 0; k < span; ++k) {
          const float v = cell_value(head, ch + k, C, policy, feats, next);
          if (v > best) {
            best = v;
            if (WANT_ARG) {
              *arg_row = (policy == LG_BEV_LAST) ? head : -2 - head;  // MAX: resolve the chain lazily
              *arg_ch = ch + k;
            }
          }
        }
      }
      w += span;
      ch = 0;
      ++d;
    }
  }
  return best;
}

__device__ __forceinline__ int64_t out_index(int layout, int b, int c, int i, int j, int C, int h_out, int w_out) {
  return layout ? (((int64_t)b * h_out + i) * w_out + j) * C + c : (((int64_t)b * C + c) * h_out + i) * w_out + j;
}

// staged tile <-> global, coalesced for the layout.  A warp moves one 32-element line per iteration (NHWC:
// the 32 scrambled channels of one output pixel, NCHW: 32 columns of one (channel, row)), lane = position
// in the line: one shared-memory access (pitch 33: conflict-free both ways), one global access and a handful
// of integer instructions per element -- the flat-index form with its three divisions and 64-bit
// multiplications per element cost as much as the pooling itself.
template <bool STORE>
__device__ __forceinline__ void tile_transfer(const TileGeom& g, float* s_out, float* gptr, int layout, int C, int h_out,
                                              int w_out, bool zero) {
  static_assert(kCG == 32 && kJC == 32 && kIB == 8, "index decomposition below assumes 32 x 8 x 32 tiles");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (layout) {  // NHWC: line = (il, jl), lane = c'
    if (lane >= g.ncg) return;
    float* base = gptr + (((int64_t)g.b * h_out + g.i0) * w_out + g.j0) * C + g.c0 + lane;
    for (int line = warp; line < kIB * kJC; line += nwarps) {
      const int il = line >> 5, jl = line & 31;
      if (il >= g.nib || jl >= g.njc) continue;
      float* ge = base + ((int64_t)il * w_out + jl) * C;
      float* se = s_out + (il * kJC + jl) * kPitch + lane;
      if (STORE)
        *ge = zero ? 0.f : *se;
      else
        *se = *ge;
    }
  } else {  // NCHW: line = (c', il), lane = jl
    if (lane >= g.njc) return;
    float* base = gptr + (((int64_t)g.b * C + g.c0) * h_out + g.i0) * w_out + g.j0 + lane;
    for (int line = warp; line < kCG * kIB; line += nwarps) {
      const int cl = line >> 3, il = line & 7;
      if (cl >= g.ncg || il >= g.nib) continue;
      float* ge = base + ((int64_t)cl * h_out + il) * w_out;
      float* se = s_out + (il * kJC + lane) * kPitch + cl;
      if (STORE)
        *ge = zero ? 0.f : *se;
      else
        *se = *ge;
    }
  }
}

// bit h - ha of the result: scrambled row h of window row i (channel c') holds a voxel; 0 = every window of this
// (c', i) line pools zeros only.  Independent of the column, so a warp tests it once for its 32 windows.
__device__ __forceinline__ uint32_t line_row_bits(const TileGeom& g, const TileSmem& t, int cl, int i, int H, int pk,
                                                  int ps, int pp) {
  const int h0 = i * ps - pp;
  const int ha = max(h0, 0), hb = min(h0 + pk, H);
  return (t.rowmask[cl] >> (ha - g.h_lo)) & ((1u << (hb - ha)) - 1u);
}

__global__ void __launch_bounds__(256)
    k_bev_pool_fwd(const float* __restrict__ feats, const int* __restrict__ next, const int* __restrict__ pixmap, int C,
                   int H, int W, int h_out, int w_out, int pk, int ps, int pp, int policy, int layout, int max_pix,
                   int pairs_cap, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const TileSmem t = carve_tile(s_dyn, pairs_cap, max_pix);
  const TileGeom g = tile_geom(C, H, W, h_out, w_out, pk, ps, pp);
  if (!tile_lookup(g, t, C, H, W, max_pix, pixmap)) {
    tile_transfer<true>(g, t.out, out, layout, C, h_out, w_out, true);
    return;
  }
  // a warp takes one (c', i) row of windows at a time, lane = column: the 32 windows share their scrambled
  // rows (an all-empty row set is detected once per warp) and neighbouring lanes read neighbouring cells
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < kCG * kIB; r += 8) {
    const int cl = r >> 3, il = r & 7;
    if (cl >= g.ncg || il >= g.nib) continue;
    float v = 0.f;
    if (lane < g.njc && line_row_bits(g, t, cl, g.i0 + il, H, pk, ps, pp))  // warp-uniform: most lines are all zeros
      v = window_scan<false>(g, t, cl, g.i0 + il, g.j0 + lane, C, H, W, pk, ps, pp, max_pix, policy, feats, next,
                             nullptr, nullptr);
    t.out[(il * kJC + lane) * kPitch + cl] = v;
  }
  __syncthreads();
  tile_transfer<true>(g, t.out, out, layout, C, h_out, w_out, false);
}

// Backward: the gradient of each window goes to its first arg-max in scan order (max_pool2d backward) if
// that cell holds a voxel.  A cell sits in <= ceil(pk/ps)^2 windows, possibly of different CTAs, so the
// <= 4 contributions meet in grad_feats through float atomics (summation order of <= 4 terms: 1e-6
// relative, the tolerance of the parity test).
__global__ void __launch_bounds__(256)
    k_bev_pool_bwd(const float* __restrict__ feats, const int* __restrict__ next, const int* __restrict__ pixmap, int C,
                   int H, int W, int h_out, int w_out, int pk, int ps, int pp, int policy, int layout, int max_pix,
                   int pairs_cap, const float* __restrict__ grad_out, float* __restrict__ grad_feats) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const TileSmem t = carve_tile(s_dyn, pairs_cap, max_pix);
  const TileGeom g = tile_geom(C, H, W, h_out, w_out, pk, ps, pp);
  if (!tile_lookup(g, t, C, H, W, max_pix, pixmap)) return;
  tile_transfer<false>(g, t.out, const_cast<float*>(grad_out), layout, C, h_out, w_out, false);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < kCG * kIB; r += 8) {
    const int cl = r >> 3, il = r & 7, jl = lane;
    if (cl >= g.ncg || il >= g.nib || jl >= g.njc) continue;
    if (!line_row_bits(g, t, cl, g.i0 + il, H, pk, ps, pp)) continue;  // warp-uniform: only empty zeros can win
    const float gv = t.out[(il * kJC + jl) * kPitch + cl];
    if (gv == 0.f) continue;
    int row = -1, ch = 0;
    window_scan<true>(g, t, cl, g.i0 + il, g.j0 + jl, C, H, W, pk, ps, pp, max_pix, policy, feats, next, &row, &ch);
    if (row == -1) continue;
    if (row < -1) row = cell_argrow(-2 - row, ch, C, policy, feats, next);
    atomicAdd(grad_feats + (int64_t)row * C + ch, gv);
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

// a (c', h') run of a tile spans (kJC-1)*ps+pk floats = at most this many pixels of C channels
static int bev_max_pix(int C, int pk, int ps) { return ((kJC - 1) * ps + pk + C - 2) / C + 1; }
static int bev_pairs_cap(int C, int pk, int ps) { return (C < kCG ? C : kCG) * ((kIB - 1) * ps + pk); }
static size_t bev_smem_bytes(int C, int pk, int ps) {  // staged tile + heads + first pixels + row masks
  const size_t pairs = bev_pairs_cap(C, pk, ps);
  return sizeof(float) * kIB * kJC * kPitch + sizeof(int) * pairs * (bev_max_pix(C, pk, ps) + 1) + sizeof(int) * kCG + 16;
}
static int64_t bev_tiles(int B, int C, int h_out, int w_out) {
  return (int64_t)B * ceil_div(C, kCG) * ceil_div(h_out, kIB) * ceil_div(w_out, kJC);
}

static int bev_check(int64_t n, int C, int B, int H, int W, int pk, int ps, int pp, int policy, const char* who) {
  LG_CHECK_ARG(n >= 0 && C >= 1 && B >= 1 && H >= 1 && W >= 1, "%s: bad sizes", who);
  LG_CHECK_ARG(pk >= 1 && ps >= 1 && pp >= 0 && 2 * pp <= pk, "%s: bad pooling parameters", who);
  LG_CHECK_ARG(policy == LG_BEV_LAST || policy == LG_BEV_MAX, "%s: bad policy", who);
  LG_CHECK_ARG((int64_t)B * H * W < ((int64_t)1 << 31), "%s: batch*H*W exceeds int32", who);
  LG_CHECK_ARG((kIB - 1) * ps + pk <= 32, "%s: pool stride/kernel too large for the row mask", who);
  LG_CHECK_ARG((int64_t)C * H * W < ((int64_t)1 << 31), "%s: C*H*W exceeds int32", who);
  LG_CHECK_ARG(bev_smem_bytes(C, pk, ps) <= 200 * 1024, "%s: C=%d too small for pool stride %d (shared memory)", who, C,
               ps);
  return LG_OK;
}

extern "C" size_t lg_bev_workspace(int64_t n, int32_t batch_size, int32_t H, int32_t W) {
  return carve_bev(nullptr, n, (int64_t)batch_size * H * W).total;
}

extern "C" int lg_bev_forward(const int32_t* coords4, const float* feats, int64_t n, int32_t C, int32_t batch_size,
                              float bound, float voxel_size, int32_t H, int32_t W, int32_t pk, int32_t ps, int32_t pp,
                              int32_t policy, int32_t layout, float* out, void* workspace, size_t workspace_bytes,
                              void* stream_) {
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
  LG_CHECK_ARG(layout == 0 || layout == 1, "lg_bev_forward: layout must be 0 (NCHW) or 1 (NHWC)");
  const size_t smem = bev_smem_bytes(C, pk, ps);
  const int64_t blocks = bev_tiles(batch_size, C, h_out, w_out);
  LG_CHECK_ARG(blocks < ((int64_t)1 << 31), "lg_bev_forward: too many output tiles");
  LG_CUDA_OK(cudaFuncSetAttribute(k_bev_pool_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bev_pool_fwd<<<(unsigned)blocks, 256, smem, stream>>>(feats, w.next, w.pixmap, C, H, W, h_out, w_out, pk, ps, pp,
                                                          policy, layout, bev_max_pix(C, pk, ps), bev_pairs_cap(C, pk, ps),
                                                          out);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" int lg_bev_backward(const int32_t* coords4, const float* feats, int64_t n, int32_t C, int32_t batch_size,
                               int32_t H, int32_t W, int32_t pk, int32_t ps, int32_t pp, int32_t policy,
                               int32_t layout, const float* grad_out, float* grad_feats, const void* workspace,
                               size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = bev_check(n, C, batch_size, H, W, pk, ps, pp, policy, "lg_bev_backward");
  if (rc) return rc;
  const int64_t pixels = (int64_t)batch_size * H * W;
  BevWorkspace w = carve_bev(const_cast<void*>(workspace), n, pixels);
  LG_CHECK_ARG(workspace && workspace_bytes >= w.total, "lg_bev_backward: workspace too small");
  if (n == 0) return LG_OK;
  LG_CHECK_ARG(feats && grad_out && grad_feats, "lg_bev_backward: null pointer");
  const int h_out = (H + 2 * pp - pk) / ps + 1, w_out = (W + 2 * pp - pk) / ps + 1;
  LG_CHECK_ARG(layout == 0 || layout == 1, "lg_bev_backward: layout must be 0 (NCHW) or 1 (NHWC)");
  LG_CUDA_OK(cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)n * C, stream));
  const size_t smem = bev_smem_bytes(C, pk, ps);
  const int64_t blocks = bev_tiles(batch_size, C, h_out, w_out);
  LG_CHECK_ARG(blocks < ((int64_t)1 << 31), "lg_bev_backward: too many output tiles");
  LG_CUDA_OK(cudaFuncSetAttribute(k_bev_pool_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bev_pool_bwd<<<(unsigned)blocks, 256, smem, stream>>>(feats, w.next, w.pixmap, C, H, W, h_out, w_out, pk, ps, pp,
                                                          policy, layout, bev_max_pix(C, pk, ps), bev_pairs_cap(C, pk, ps),
                                                          grad_out, grad_feats);
  LG_LAUNCH_OK();
  if (policy == LG_BEV_LAST) {
    k_bev_dup_grad<<<(unsigned)ceil_div(n * C, 256), 256, 0, stream>>>(w.pix_of_row, w.pixmap, n, C, grad_feats);
    LG_LAUNCH_OK();
  }
  return LG_OK;
}
