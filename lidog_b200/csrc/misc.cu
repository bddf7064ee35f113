// Library plumbing (error string, device info) and the small operand-preparation kernels of the
// tensor-core convolution path: fp32 -> bf16/fp16 row casts, power-of-two gradient scaling for the
// fp16 format, weight re-layout.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace lg {

static thread_local char g_error[512] = "no error";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

__device__ __forceinline__ unsigned short to16(float v, int fmt) {
  if (fmt == LG_FMT_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  // saturating: fp16 has no room above 65504
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  return __half_as_ushort(__float2half_rn(v));
}

__global__ void __launch_bounds__(256)
    k_cast_rows(const float* __restrict__ src, unsigned short* __restrict__ dst, int64_t n, int fmt,
                const float* __restrict__ scale) {
  const float s = scale ? scale[0] : 1.f;
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i);
    const float4 b = *reinterpret_cast<const float4*>(src + i + 4);
    uint4 o;
    o.x = to16(a.x * s, fmt) | ((unsigned)to16(a.y * s, fmt) << 16);
    o.y = to16(a.z * s, fmt) | ((unsigned)to16(a.w * s, fmt) << 16);
    o.z = to16(b.x * s, fmt) | ((unsigned)to16(b.y * s, fmt) << 16);
    o.w = to16(b.z * s, fmt) | ((unsigned)to16(b.w * s, fmt) << 16);
    *reinterpret_cast<uint4*>(dst + i) = o;
  } else {
    for (; i < n; ++i) dst[i] = to16(src[i] * s, fmt);
  }
}

__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ src, int64_t n, unsigned* __restrict__ bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = fabsf(src[i]);
    if (v <= 3.0e38f) m = fmaxf(m, v);  // ignores inf / nan
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));  // positive floats order as uints
}

__global__ void k_scale_from_absmax(const unsigned* bits, float* scale_out) {
  const float m = __uint_as_float(*bits);
  float s = 1.f;
  if (m > 0.f) {
    int e;
    frexpf(m, &e);               // m = f * 2^e, f in [0.5, 1)
    s = ldexpf(1.f, 12 - e);     // absmax lands in [2^11, 2^12)
  }
  scale_out[0] = s;
  scale_out[1] = 1.f / s;
}

// W [K][Cin][Cout] fp32 -> w16 (same layout) and w16t [K][Cout][Cin]
__global__ void __launch_bounds__(256)
    k_prep_weights(const float* __restrict__ W, int K, int Cin, int Cout, unsigned short* __restrict__ w16,
                   unsigned short* __restrict__ w16t, int fmt) {
  const int64_t n = (int64_t)K * Cin * Cout;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned short v = to16(W[i], fmt);
  if (w16) w16[i] = v;
  if (w16t) {
    const int co = (int)(i % Cout);
    const int64_t t = i / Cout;
    const int ci = (int)(t % Cin);
    const int64_t k = t / Cin;
    w16t[(k * Cout + co) * Cin + ci] = v;
  }
}

}  // namespace lg

using namespace lg;

extern "C" int lg_version(void) { return 100; }
extern "C" const char* lg_last_error_string(void) { return g_error; }

extern "C" int lg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  LG_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  LG_CUDA_OK(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return LG_OK;
}

extern "C" int lg_cast_rows(const float* src, void* dst16, int64_t n_elems, int32_t fmt, const float* scale,
                            void* stream) {
  LG_CHECK_ARG(n_elems >= 0 && (fmt == LG_FMT_BF16 || fmt == LG_FMT_FP16), "lg_cast_rows: bad arguments");
  if (n_elems == 0) return LG_OK;
  LG_CHECK_ARG(src && dst16, "lg_cast_rows: null pointer");
  LG_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst16 & 15) == 0, "lg_cast_rows: pointers must be 16B aligned");
  k_cast_rows<<<(unsigned)ceil_div(n_elems, 256 * 8), 256, 0, (cudaStream_t)stream>>>(src, (unsigned short*)dst16,
                                                                                     n_elems, fmt, scale);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" int lg_absmax_scale(const float* src, int64_t n_elems, float* scale_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LG_CHECK_ARG(n_elems >= 0 && scale_out, "lg_absmax_scale: bad arguments");
  // scale_out[2] is used as the uint scratch word for the max
  unsigned* bits = reinterpret_cast<unsigned*>(scale_out + 2);
  LG_CUDA_OK(cudaMemsetAsync(bits, 0, sizeof(unsigned), stream));
  if (n_elems > 0) {
    int64_t want = ceil_div(n_elems, 256 * 8);
    int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
    k_absmax<<<blocks, 256, 0, stream>>>(src, n_elems, bits);
    LG_LAUNCH_OK();
  }
  k_scale_from_absmax<<<1, 1, 0, stream>>>(bits, scale_out);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" int lg_prep_weights(const float* W, int32_t K, int32_t Cin, int32_t Cout, void* w16, void* w16t,
                               int32_t fmt, void* stream) {
  LG_CHECK_ARG(W && K >= 1 && Cin >= 1 && Cout >= 1 && (fmt == LG_FMT_BF16 || fmt == LG_FMT_FP16),
               "lg_prep_weights: bad arguments");
  const int64_t n = (int64_t)K * Cin * Cout;
  k_prep_weights<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(W, K, Cin, Cout, (unsigned short*)w16,
                                                                              (unsigned short*)w16t, fmt);
  LG_LAUNCH_OK();
  return LG_OK;
}
