// Shared helpers for the libdalle_b200.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
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
// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): one instruction per two elements on the fp32 pipe.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// gelu_fast for two elements at once: the same Abramowitz-Stegun evaluation with the polynomial and the products on packed
// fp32 instructions (13 packed + 4 MUFU + 4 scalar per pair instead of ~19 scalar per element).
__device__ __forceinline__ void gelu_fast2(float2 g, float2& gelu, float2& dgelu) {
  const float2 x = mul2(g, splat2(0.70710678118654752440f));
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = fma2(ax, splat2(0.3275911f), splat2(1.0f));
  float2 t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  const float2 q = mul2(mul2(ax, ax), splat2(-1.4426950408889634f));        // -x^2 * log2(e)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(q.y));
  float2 poly = fma2(t, splat2(1.061405429f), splat2(-1.453152027f));
  poly = fma2(t, poly, splat2(1.421413741f));
  poly = fma2(t, poly, splat2(-0.284496736f));
  poly = fma2(t, poly, splat2(0.254829592f));
  const float2 m = mul2(mul2(poly, t), ex);
  const float2 erf_abs = fma2(m, splat2(-1.0f), splat2(1.0f));
  const float2 cs = make_float2(copysignf(erf_abs.x, x.x), copysignf(erf_abs.y, x.y));
  const float2 cdf = fma2(cs, splat2(0.5f), splat2(0.5f));
  gelu = mul2(g, cdf);
  dgelu = fma2(mul2(g, splat2(0.39894228040143267794f)), ex, cdf);
}

template <typename T> __device__ __forceinline__ void gelu_pair(float g, float& gelu, float& dgelu);
template <> __device__ __forceinline__ void gelu_pair<float>(float g, float& gelu, float& dgelu) { gelu = gelu_erf(g); dgelu = gelu_erf_grad(g); }
template <> __device__ __forceinline__ void gelu_pair<__nv_bfloat16>(float g, float& gelu, float& dgelu) { gelu_fast(g, gelu, dgelu); }
template <typename T> __device__ __forceinline__ float gelu_fwd(float g);
template <> __device__ __forceinline__ float gelu_fwd<float>(float g) { return gelu_erf(g); }
template <> __device__ __forceinline__ float gelu_fwd<__nv_bfloat16>(float g) { float a, b; gelu_fast(g, a, b); return a; }
// two elements: fp32 mode stays on the exact scalar erf, bf16 mode uses the packed evaluation
template <typename T> __device__ __forceinline__ void gelu_pair2(float2 g, float2& gelu, float2& dgelu);
template <> __device__ __forceinline__ void gelu_pair2<float>(float2 g, float2& gelu, float2& dgelu) {
  gelu = make_float2(gelu_erf(g.x), gelu_erf(g.y)); dgelu = make_float2(gelu_erf_grad(g.x), gelu_erf_grad(g.y));
}
template <> __device__ __forceinline__ void gelu_pair2<__nv_bfloat16>(float2 g, float2& gelu, float2& dgelu) { gelu_fast2(g, gelu, dgelu); }

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

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------
// A kernel launched through launch_pdl() may be scheduled while its predecessor in the stream is still draining; it must
// execute pdl_wait() before its first global-memory access (the wait returns once the predecessor grid has completed and its
// writes are visible; it is a no-op for a normally launched kernel).  pdl_launch() lets OUR successor be scheduled early.
// Net effect: the ~2 us launch gap and the kernel prologue (barrier init, TMEM allocation) overlap the previous kernel's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();      // DALLE_B200_PDL=0 turns the launch attribute off (kernels then serialise as usual)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();   // cached multiprocessor count of the current device

}  // namespace db200
