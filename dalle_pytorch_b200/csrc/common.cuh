// Shared helpers for the libdalle_b200.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/dalle_b200.h"

namespace db200 {

// ---- error slot (thread-local; SURVEY.md §8b error convention) ------------------------------------
std::string& last_error_slot();
int set_error(int code, const char* fmt, ...);

#define DB200_CHECK_ARG(cond, ...)                                                   \
  do {                                                                               \
    if (!(cond)) return ::db200::set_error(DB200_ERR_BAD_ARG, __VA_ARGS__);          \
  } while (0)

#define DB200_CUDA_OK(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess)                                                           \
      return ::db200::set_error(DB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,      \
                                cudaGetErrorString(_e), __FILE__, __LINE__);         \
  } while (0)

#define DB200_LAUNCH_OK(what)                                                        \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess)                                                           \
      return ::db200::set_error(DB200_ERR_CUDA, "launch of %s failed: %s", what,     \
                                cudaGetErrorString(_e));                             \
  } while (0)

// ---- storage-type conversions ------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> __device__ __forceinline__ void store2(T* p, float a, float b);
template <> __device__ __forceinline__ void store2<float>(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
template <> __device__ __forceinline__ void store2<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

template <typename T> __device__ __forceinline__ float2 load2(const T* p);
template <> __device__ __forceinline__ float2 load2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 load2<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

// exact erf GELU, as F.gelu default (transformer.py:109)
__device__ __forceinline__ float gelu_erf(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float g) {
  const float cdf = 0.5f * (1.0f + erff(g * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * g * g);
  return cdf + g * pdf;
}

// GELU and its derivative from ONE exponential, for the bf16 epilogues (4 epilogue warps-per-quarter do this for every
// element of the [M, 4d] hidden tensor, so instruction count matters): erf by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, far below bf16 resolution), whose exp(-x^2) term with x = g/sqrt(2) is also the Gaussian pdf.
__device__ __forceinline__ void gelu_fast(float g, float& gelu, float& dgelu) {
  const float x = g * 0.70710678118654752440f;
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  const float ex = __expf(-ax * ax);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, ex, 1.0f);
  const float cdf = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
  gelu = g * cdf;
  dgelu = fmaf(g * 0.39894228040143267794f, ex, cdf);
}
template <typename T> __device__ __forceinline__ void gelu_pair(float g, float& gelu, float& dgelu);
template <> __device__ __forceinline__ void gelu_pair<float>(float g, float& gelu, float& dgelu) { gelu = gelu_erf(g); dgelu = gelu_erf_grad(g); }
template <> __device__ __forceinline__ void gelu_pair<__nv_bfloat16>(float g, float& gelu, float& dgelu) { gelu_fast(g, gelu, dgelu); }
template <typename T> __device__ __forceinline__ float gelu_fwd(float g);
template <> __device__ __forceinline__ float gelu_fwd<float>(float g) { return gelu_erf(g); }
template <> __device__ __forceinline__ float gelu_fwd<__nv_bfloat16>(float g) { float a, b; gelu_fast(g, a, b); return a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();   // cached multiprocessor count of the current device

}  // namespace db200
