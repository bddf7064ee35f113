// LiDOG's two training criteria on the device, two passes each way.
// Reference: DICELoss (utils/losses/losses.py:56-97; the BEV criterion, powerize = False, no target mask) and
// SoftDICELoss (:129-187 with get_soft / get_kitti_soft :100-126; the 3D criterion: eps-smoothed targets, squared
// probabilities in the union, only classes present in the batch count).  The reference moves logits and labels to
// the CPU and evaluates ~12 torch ops there; the torch-on-device form still costs ~80 small kernels per step.
//
//   forward   k_dice_sums      per-class sums over the rows whose label is not ignored (w_n = 1):
//                                I_c = sum p_nc t_nc,  P_c = sum p_nc^q (q = 2 powerize, else 1),
//                                T_c = sum t_nc,       O_c = sum onehot_nc          (p = softmax(logits_n))
//             k_dice_finalize  loss = 1 - sum_c m_c 2 I_c / (P_c + T_c + 1e-12) / (sum_c m_c + 1e-12)
//                              (m_c = [O_c > 0] with the target mask, else 1) and the two coefficient vectors
//                              a_c = dL/dI_c, u_c = dL/dP_c the backward needs
//   backward  k_dice_backward  g_nc = a_c t_nc + u_c q p_nc^(q-1);  dlogits_nc = s * p_nc (g_nc - sum_j p_nj g_nj)
// Reductions run in a fixed order (registers -> warp shuffles -> per-block partials -> one block): deterministic.
// Bound: HBM, one read of the logits each way (+ one write backward); C <= 32.
#include "common.cuh"
#include "runtime.cuh"

namespace lg {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 592;

struct DiceCfg {
  int C, ignore, soft, is_kitti, powerize, use_tmask;
  float hi, lo;  // soft targets: 1 - eps and eps / (C - 1); hard: 1 and 0
};

// soft target of class c for a row with label y (get_soft / get_kitti_soft)
__device__ __forceinline__ float target_of(const DiceCfg& k, int y, int c) {
  if (k.soft && k.is_kitti && (y == 1 || y == 6) && (c == 1 || c == 6)) return 0.5f * k.hi;
  return c == y ? k.hi : k.lo;
}

// softmax of one row (3 passes over <= 32 floats that sit in L1): returns max and 1 / sum exp
__device__ __forceinline__ void row_softmax(const float* __restrict__ z, int C, float* mx, float* inv) {
  float m = z[0];
  for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(z[c] - m);
  *mx = m;
  *inv = 1.f / s;
}

template <int CMAX>
__global__ void __launch_bounds__(kLossThreads)
    k_dice_sums(const float* __restrict__ logits, const long long* __restrict__ target, int64_t n, DiceCfg k,
                float* __restrict__ partial /* [blocks][4][C] */) {
  __shared__ float sm[kLossThreads / 32][4 * CMAX];
  float acc[4][CMAX];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[a][c] = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)target[r];
    if (y == k.ignore) continue;
    const float* z = logits + r * k.C;
    float m, inv;
    row_softmax(z, k.C, &m, &inv);
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < k.C) {
        const float p = expf(z[c] - m) * inv;
        const float t = target_of(k, y, c);
        acc[0][c] += p * t;
        acc[1][c] += k.powerize ? p * p : p;
        acc[2][c] += t;
        acc[3][c] += (c == y) ? 1.f : 0.f;
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      float v = acc[a][c];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) sm[warp][a * CMAX + c] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * k.C; i += blockDim.x) {
    const int a = i / k.C, c = i - a * k.C;
    float v = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) v += sm[w][a * CMAX + c];
    partial[(size_t)blockIdx.x * 4 * k.C + i] = v;
  }
}

// one block: sums of the block partials in block order (double), loss and backward coefficients
__global__ void k_dice_finalize(const float* __restrict__ partial, int n_blocks, DiceCfg k, float* __restrict__ loss,
                                float* __restrict__ coef /* [2][C]: a_c, u_c */) {
  __shared__ double s[4 * 32];
  const int i = threadIdx.x;
  if (i < 4 * k.C) {
    double v = 0.0;
    for (int b = 0; b < n_blocks; ++b) v += (double)partial[(size_t)b * 4 * k.C + i];
    s[i] = v;
  }
  __syncthreads();
  if (i == 0) {
    const int C = k.C;
    double M = 0.0, acc = 0.0;
    for (int c = 0; c < C; ++c) {
      const double m = k.use_tmask ? (s[3 * C + c] > 0.0 ? 1.0 : 0.0) : 1.0;
      const double U = s[C + c] + s[2 * C + c] + 1e-12;
      M += m;
      acc += m * 2.0 * s[c] / U;
    }
    M += 1e-12;
    *loss = (float)(1.0 - acc / M);
    for (int c = 0; c < C; ++c) {
      const double m = k.use_tmask ? (s[3 * C + c] > 0.0 ? 1.0 : 0.0) : 1.0;
      const double U = s[C + c] + s[2 * C + c] + 1e-12;
      coef[c] = (float)(-2.0 * m / (U * M));
      coef[C + c] = (float)(2.0 * m * s[c] / (U * U * M));
    }
  }
}

template <int CMAX>
__global__ void __launch_bounds__(kLossThreads)
    k_dice_backward(const float* __restrict__ logits, const long long* __restrict__ target, int64_t n, DiceCfg k,
                    const float* __restrict__ coef, const float* __restrict__ grad_scale, float* __restrict__ dlogits) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float* dz = dlogits + r * k.C;
  const int y = (int)target[r];
  if (y == k.ignore) {
    for (int c = 0; c < k.C; ++c) dz[c] = 0.f;
    return;
  }
  const float s = grad_scale ? grad_scale[0] : 1.f;
  const float* z = logits + r * k.C;
  float m, inv;
  row_softmax(z, k.C, &m, &inv);
  float p[CMAX], g[CMAX];
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    if (c < k.C) {
      p[c] = expf(z[c] - m) * inv;
      g[c] = coef[c] * target_of(k, y, c) + coef[k.C + c] * (k.powerize ? 2.f * p[c] : 1.f);
      dot += p[c] * g[c];
    }
  }
#pragma unroll
  for (int c = 0; c < CMAX; ++c)
    if (c < k.C) dz[c] = s * p[c] * (g[c] - dot);
}

static int dice_cfg(DiceCfg* k, int C, int ignore, int soft, float eps, int is_kitti, const char* who) {
  LG_CHECK_ARG(C >= 2 && C <= 32, "%s: 2 <= C <= 32 classes", who);
  LG_CHECK_ARG(!is_kitti || C > 6, "%s: the kitti soft-target rule needs classes 1 and 6", who);
  k->C = C;
  k->ignore = ignore;
  k->soft = soft ? 1 : 0;
  k->is_kitti = (soft && is_kitti) ? 1 : 0;
  k->powerize = soft ? 1 : 0;   // LiDOG's configurations: SoftDICELoss(powerize=True, use_tmask=True),
  k->use_tmask = soft ? 1 : 0;  // DICELoss(powerize=False, use_tmask=False)
  k->hi = soft ? 1.f - eps : 1.f;
  k->lo = soft ? eps / (float)(C - 1) : 0.f;
  return LG_OK;
}

static int loss_blocks(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, (int64_t)kLossThreads * 4);
  return (int)(b > kLossMaxBlocks ? kLossMaxBlocks : b);
}

}  // namespace lg

using namespace lg;

extern "C" int lg_dice_forward(const float* logits, const int64_t* target, int64_t n, int32_t C, int32_t ignore_label,
                               int32_t soft, float eps, int32_t is_kitti, float* loss, float* coef, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DiceCfg k;
  int rc = dice_cfg(&k, C, ignore_label, soft, eps, is_kitti, "lg_dice_forward");
  if (rc) return rc;
  LG_CHECK_ARG(n >= 0 && loss && coef && (n == 0 || (logits && target)), "lg_dice_forward: null pointer");
  const int nb = loss_blocks(n);
  ArenaCursor ar;
  rc = arena_begin(stream, sizeof(float) * (size_t)kLossMaxBlocks * 4 * 32, &ar);
  if (rc) return rc;
  float* partial = (float*)arena_take(&ar, sizeof(float) * (size_t)kLossMaxBlocks * 4 * 32);
  if (C <= 8)
    k_dice_sums<8><<<nb, kLossThreads, 0, stream>>>(logits, (const long long*)target, n, k, partial);
  else
    k_dice_sums<32><<<nb, kLossThreads, 0, stream>>>(logits, (const long long*)target, n, k, partial);
  LG_LAUNCH_OK();
  k_dice_finalize<<<1, 128, 0, stream>>>(partial, nb, k, loss, coef);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" int lg_dice_backward(const float* logits, const int64_t* target, int64_t n, int32_t C, int32_t ignore_label,
                                int32_t soft, float eps, int32_t is_kitti, const float* coef, const float* grad_scale,
                                float* dlogits, void* stream_) {
  DiceCfg k;
  int rc = dice_cfg(&k, C, ignore_label, soft, eps, is_kitti, "lg_dice_backward");
  if (rc) return rc;
  if (n == 0) return LG_OK;
  LG_CHECK_ARG(logits && target && coef && dlogits, "lg_dice_backward: null pointer");
  const unsigned grid = (unsigned)ceil_div(n, kLossThreads);
  if (C <= 8)
    k_dice_backward<8><<<grid, kLossThreads, 0, (cudaStream_t)stream_>>>(logits, (const long long*)target, n, k, coef,
                                                                         grad_scale, dlogits);
  else
    k_dice_backward<32><<<grid, kLossThreads, 0, (cudaStream_t)stream_>>>(logits, (const long long*)target, n, k, coef,
                                                                          grad_scale, dlogits);
  LG_LAUNCH_OK();
  return LG_OK;
}
