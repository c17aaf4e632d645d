// Fused GEMM epilogues shared by the SIMT (fp32) and tcgen05 (bf16) GEMM kernels.
// Each function consumes one PAIR of adjacent accumulator columns (n even) of one output row m, which is
// the natural granule for the rotary pair rotation and for packed bf16x2 stores.
#pragma once
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace db200 {

struct EpiArgs {
  int M, N;
  // STORE
  void* C; long long ldc; int c_is_f32; const float* bias; int atomic_c;      // atomic_c: 0 store, 1 fp32 red.add (split-K), 2 multimem.red
  float c_scale;                                                             // multiplies the accumulator in the multimem mode
  int tma_store;                                                             // STORE, bf16 C: stage 32 x 64 boxes in smem, cp.async.bulk.tensor store
  // QKV
  void* q; void* k; void* v; const float* cos_t; const float* sin_t;
  int seq_n, heads, dim_head, pos_offset; float q_scale;
  // RESID
  const float* resid; const float* scale; float sign; void* y_out; float* out;
  // GEGLU / GEGLU_BWD
  void* u_out; void* h_out; int hidden; const void* u_in; void* du_out;
};

inline EpiArgs make_epi_args(const db200_gemm_params& p) {
  EpiArgs e;
  e.M = p.M; e.N = p.N;
  e.C = p.C; e.ldc = p.ldc; e.c_is_f32 = (p.c_dtype == DB200_F32); e.bias = p.bias; e.atomic_c = 0; e.c_scale = 1.0f; e.tma_store = 0;
  if (p.C_multicast) { e.C = p.C_multicast; e.atomic_c = 2; e.c_scale = p.c_scale; }   // weight gradient reduced across the GPUs in the NVSwitch
  // diagnosis only (tools/gemm_probe.py): DALLE_B200_GEMM_DBG=nostore drops the global stores of the STORE epilogue (3),
  // =noepi skips the whole epilogue body (4) -- isolates the mainloop from the epilogue in the roofline gap
  static const int dbg = [] { const char* v = getenv("DALLE_B200_GEMM_DBG"); return !v ? 0 : !strcmp(v, "nostore") ? 3 : !strcmp(v, "noepi") ? 4 : 0; }();
  if (dbg && p.epilogue == DB200_EPI_STORE && !p.C_multicast) e.atomic_c = dbg;
  e.q = p.q; e.k = p.k; e.v = p.v; e.cos_t = p.cos_t; e.sin_t = p.sin_t;
  e.seq_n = p.seq_n; e.heads = p.heads; e.dim_head = p.dim_head; e.pos_offset = p.pos_offset; e.q_scale = p.q_scale;
  e.resid = p.resid; e.scale = p.scale; e.sign = p.sign; e.y_out = p.y_out; e.out = p.out;
  e.u_out = p.u_out; e.h_out = p.h_out; e.hidden = p.hidden; e.u_in = p.u_in; e.du_out = p.du_out;
  return e;
}

// Reduction into the same offset of EVERY GPU's copy of a symmetric buffer through its NVLink multicast address: the NVSwitch adds
// the value into all replicas (NVLS).  Data-parallel weight gradients are summed across the GPUs by the wgrad GEMM's own
// epilogue -- no collective kernel, no extra pass over the gradient (distributed.py, multimem mode).
__device__ __forceinline__ void mc_red_add_v4(float* mc, float a, float b, float c, float d) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mc_red_add_v2(float* mc, float a, float b) {
  asm volatile("multimem.red.relaxed.sys.global.add.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(a), "f"(b) : "memory");
}

// EPI_STORE ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void epi_store_pair(const EpiArgs& e, int m, int n, float v0, float v1) {
  if (e.bias) { const float2 bb = __ldg(reinterpret_cast<const float2*>(e.bias + n)); v0 += bb.x; v1 += bb.y; }
  const long long off = (long long)m * e.ldc + n;
  if (e.c_is_f32) store2<float>(reinterpret_cast<float*>(e.C) + off, v0, v1);
  else            store2<T>(reinterpret_cast<T*>(e.C) + off, v0, v1);
}

// EPI_QKV: 'b n (h d) -> b h n d', rotary on q,k,v (interleaved pairs), q *= dh^-0.5 -------------------
// (attention.py:63-69 ; rotary_embedding_torch.apply_rotary_emb: t*cos + rotate_half(t)*sin)
template <typename T>
__device__ __forceinline__ void epi_qkv_pair(const EpiArgs& e, int m, int n, float v0, float v1) {
  const int inner = e.heads * e.dim_head;
  const int which = n / inner;
  const int rem = n - which * inner;
  const int head = rem / e.dim_head;
  const int d = rem - head * e.dim_head;           // even
  const int b = m / e.seq_n;
  const int p = m - b * e.seq_n;
  if (e.cos_t) {
    const int ti = (p + e.pos_offset) * (e.dim_head >> 1) + (d >> 1);
    const float c = __ldg(e.cos_t + ti), s = __ldg(e.sin_t + ti);
    const float r0 = v0 * c + (-v1) * s;
    const float r1 = v1 * c + v0 * s;
    v0 = r0; v1 = r1;
  }
  if (which == 0) { v0 *= e.q_scale; v1 *= e.q_scale; }
  T* base = reinterpret_cast<T*>(which == 0 ? e.q : (which == 1 ? e.k : e.v));
  const long long off = (((long long)b * e.heads + head) * e.seq_n + p) * e.dim_head + d;
  store2<T>(base + off, v0, v1);
}

// EPI_RESID: out = resid + sign*scale*(acc+bias) ; optionally keep y = acc+bias for the LayerScale grad ----
template <typename T>
__device__ __forceinline__ void epi_resid_pair(const EpiArgs& e, int m, int n, float v0, float v1) {
  if (e.bias) { const float2 bb = __ldg(reinterpret_cast<const float2*>(e.bias + n)); v0 += bb.x; v1 += bb.y; }
  const long long off = (long long)m * e.N + n;
  if (e.y_out) store2<T>(reinterpret_cast<T*>(e.y_out) + off, v0, v1);
  float s0 = e.sign, s1 = e.sign;
  if (e.scale) { const float2 sc = __ldg(reinterpret_cast<const float2*>(e.scale + n)); s0 *= sc.x; s1 *= sc.y; }
  float r0 = 0.f, r1 = 0.f;
  if (e.resid) { const float2 r = *reinterpret_cast<const float2*>(e.resid + off); r0 = r.x; r1 = r.y; }
  *reinterpret_cast<float2*>(e.out + off) = make_float2(r0 + s0 * v0, r1 + s1 * v1);
}

// EPI_GEGLU: j indexes the hidden dim; a = acc[:, j], g = acc[:, hidden + j] -------------------------------
template <typename T>
__device__ __forceinline__ void epi_geglu_pair(const EpiArgs& e, int m, int j, float a0, float a1, float g0, float g1) {
  if (e.bias) {
    const float2 ba = __ldg(reinterpret_cast<const float2*>(e.bias + j));
    const float2 bg = __ldg(reinterpret_cast<const float2*>(e.bias + e.hidden + j));
    a0 += ba.x; a1 += ba.y;
    g0 += bg.x; g1 += bg.y;
  }
  if (e.u_out) {
    T* u = reinterpret_cast<T*>(e.u_out) + (long long)m * (2 * e.hidden);
    store2<T>(u + j, a0, a1);
    store2<T>(u + e.hidden + j, g0, g1);
  }
  T* h = reinterpret_cast<T*>(e.h_out) + (long long)m * e.hidden;
  store2<T>(h + j, a0 * gelu_fwd<T>(g0), a1 * gelu_fwd<T>(g1));
}

// EPI_GEGLU_BWD: acc = dh[m, j..j+1] --------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void epi_geglu_bwd_pair(const EpiArgs& e, int m, int j, float d0, float d1) {
  const T* u = reinterpret_cast<const T*>(e.u_in) + (long long)m * (2 * e.hidden);
  const float2 a = load2<T>(u + j);
  const float2 g = load2<T>(u + e.hidden + j);
  T* du = reinterpret_cast<T*>(e.du_out) + (long long)m * (2 * e.hidden);
  float f0, f1, df0, df1;
  gelu_pair<T>(g.x, f0, df0);
  gelu_pair<T>(g.y, f1, df1);
  store2<T>(du + j, d0 * f0, d1 * f1);
  store2<T>(du + e.hidden + j, d0 * a.x * df0, d1 * a.y * df1);
}

template <int EPI, typename T>
__device__ __forceinline__ void epi_pair(const EpiArgs& e, int m, int n, float v0, float v1) {
  if constexpr (EPI == DB200_EPI_STORE) epi_store_pair<T>(e, m, n, v0, v1);
  else if constexpr (EPI == DB200_EPI_QKV) epi_qkv_pair<T>(e, m, n, v0, v1);
  else if constexpr (EPI == DB200_EPI_RESID) epi_resid_pair<T>(e, m, n, v0, v1);
  else if constexpr (EPI == DB200_EPI_GEGLU_BWD) epi_geglu_bwd_pair<T>(e, m, n, v0, v1);
}


// =====================================================================================================
// 8-column granules (one 16-byte bf16 store per output tensor) for the tcgen05 epilogue, where a thread
// owns one accumulator row and walks along N.  v[8] are columns n .. n+7 (n % 8 == 0) of row m.
// =====================================================================================================
template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
    uint4 u;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = u;
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.z));
    const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.w));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void load(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// 32-byte (one full L2 sector) stores: sm_100 has 256-bit global stores (st.global.v8.b32).  The row-mode STORE epilogue writes one
// sector per lane and instruction instead of two 16-byte halves of it: measured with the stores removed the K = 1024 GEMMs run at
// 1.55 PFLOP/s against 1.14 with 16-byte stores (tools/gemm_gap.py) -- partial-sector writes to 32 different lines per warp
// instruction were the whole roofline gap of the short-K shapes.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]),
               "r"(w[6]), "r"(w[7])
               : "memory");
}
// 16 accumulator columns n .. n+15 of row m (n % 16 == 0) of a plain STORE epilogue without atomics; returns false if the
// alignment does not allow the 32-byte path (the caller then uses two epi_vec8 granules)
template <typename T>
__device__ __forceinline__ bool epi_store16(const EpiArgs& e, int m, int n, const uint32_t* acc) {
  if (e.atomic_c != 0) return false;
  const long long off = (long long)m * e.ldc + n;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + n + i));
      v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
    }
  }
  if (e.c_is_f32) {
    float* c = reinterpret_cast<float*>(e.C) + off;
    if (reinterpret_cast<uintptr_t>(c) & 31) return false;
    uint32_t w[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(v[8 * h + i]);
      st_global_v8(c + 8 * h, w);
    }
  } else {
    T* c = reinterpret_cast<T*>(e.C) + off;
    if (reinterpret_cast<uintptr_t>(c) & 31) return false;
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    st_global_v8(c, w);
  }
  return true;
}

__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t* w) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
// EPI_RESID, row mode with full-sector accesses: 16 accumulator columns n .. n+15 of row m (n % 16 == 0, N % 8 == 0 so that every
// row of out / resid / y starts on a 32-byte boundary).  Per lane and 16 columns: 2 x 32-byte loads of the residual, 2 x 32-byte
// stores of out, one 32-byte store of y -- no shared-memory transpose.
template <typename T>
__device__ __forceinline__ void epi_resid16(const EpiArgs& e, int m, int n, const uint32_t* acc) {
  const long long off = (long long)m * e.N + n;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + n + i));
      v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
    }
  }
  uint32_t rr[16];
  if (e.resid) { ld_global_v8(e.resid + off, rr); ld_global_v8(e.resid + off + 8, rr + 8); }
  if (e.y_out) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    st_global_v8(reinterpret_cast<T*>(e.y_out) + off, w);
  }
  uint32_t o[16];
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (e.scale) sc = __ldg(reinterpret_cast<const float4*>(e.scale + n + i));
    const float s4[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float r = e.resid ? __uint_as_float(rr[i + k]) : 0.f;
      o[i + k] = __float_as_uint(r + e.sign * s4[k] * v[i + k]);
    }
  }
  st_global_v8(e.out + off, o);
  st_global_v8(e.out + off + 8, o + 8);
}

// EPI_GEGLU, row mode with full-sector stores: hidden indices j .. j+15 of row m (j % 16 == 0, hidden % 16 == 0)
template <typename T>
__device__ __forceinline__ void epi_geglu16(const EpiArgs& e, int m, int j, const uint32_t* acc_a, const uint32_t* acc_g) {
  float a[16], g[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(acc_a[i]); g[i] = __uint_as_float(acc_g[i]); }
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 ba = __ldg(reinterpret_cast<const float4*>(e.bias + j + i));
      const float4 bg = __ldg(reinterpret_cast<const float4*>(e.bias + e.hidden + j + i));
      a[i] += ba.x; a[i + 1] += ba.y; a[i + 2] += ba.z; a[i + 3] += ba.w;
      g[i] += bg.x; g[i + 1] += bg.y; g[i + 2] += bg.z; g[i + 3] += bg.w;
    }
  }
  uint32_t w[8];
  if (e.u_out) {
    T* u = reinterpret_cast<T*>(e.u_out) + (long long)m * (2 * e.hidden);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const __nv_bfloat162 b2 = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&b2); }
    st_global_v8(u + j, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const __nv_bfloat162 b2 = __floats2bfloat162_rn(g[2 * i], g[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&b2); }
    st_global_v8(u + e.hidden + j, w);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 f, df;
    gelu_pair2<T>(make_float2(g[2 * i], g[2 * i + 1]), f, df);
    const __nv_bfloat162 b2 = __floats2bfloat162_rn(a[2 * i] * f.x, a[2 * i + 1] * f.y);
    w[i] = *reinterpret_cast<const uint32_t*>(&b2);
  }
  st_global_v8(reinterpret_cast<T*>(e.h_out) + (long long)m * e.hidden + j, w);
}

template <int EPI, typename T>
__device__ __forceinline__ void epi_vec8(const EpiArgs& e, int m, int n, float* v) {
  if constexpr (EPI == DB200_EPI_STORE) {
    if (e.bias) {
      float bb[8]; Vec8<float>::load(e.bias + n, bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += bb[i];
    }
    const long long off = (long long)m * e.ldc + n;
    if (e.atomic_c >= 3) return;                       // (diagnosis switch, see make_epi_args)
    if (e.atomic_c == 2) {                             // cross-GPU sum through the multicast address (16-byte aligned granules)
      float* c = reinterpret_cast<float*>(e.C) + off;
      mc_red_add_v4(c, v[0] * e.c_scale, v[1] * e.c_scale, v[2] * e.c_scale, v[3] * e.c_scale);
      mc_red_add_v4(c + 4, v[4] * e.c_scale, v[5] * e.c_scale, v[6] * e.c_scale, v[7] * e.c_scale);
    } else if (e.atomic_c) {                           // split-K partial sum (fp32 C, zero-initialised by the caller)
      float* c = reinterpret_cast<float*>(e.C) + off;
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(c + i, v[i]);
    } else if (e.c_is_f32) Vec8<float>::store(reinterpret_cast<float*>(e.C) + off, v);
    else Vec8<T>::store(reinterpret_cast<T*>(e.C) + off, v);
  } else if constexpr (EPI == DB200_EPI_QKV) {
    const int inner = e.heads * e.dim_head;
    const int which = n / inner;
    const int rem = n - which * inner;
    const int head = rem / e.dim_head;
    const int d = rem - head * e.dim_head;          // multiple of 8
    const int b = m / e.seq_n;
    const int p = m - b * e.seq_n;
    if (e.cos_t) {
      const int ti = (p + e.pos_offset) * (e.dim_head >> 1) + (d >> 1);
      const float4 c = *reinterpret_cast<const float4*>(e.cos_t + ti);
      const float4 s = *reinterpret_cast<const float4*>(e.sin_t + ti);
      const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x0 = v[2 * i], x1 = v[2 * i + 1];
        v[2 * i] = x0 * cc[i] + (-x1) * ss[i];
        v[2 * i + 1] = x1 * cc[i] + x0 * ss[i];
      }
    }
    if (which == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= e.q_scale;
    }
    T* base = reinterpret_cast<T*>(which == 0 ? e.q : (which == 1 ? e.k : e.v));
    Vec8<T>::store(base + (((long long)b * e.heads + head) * e.seq_n + p) * e.dim_head + d, v);
  } else if constexpr (EPI == DB200_EPI_RESID) {
    if (e.bias) {
      float bb[8]; Vec8<float>::load(e.bias + n, bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += bb[i];
    }
    const long long off = (long long)m * e.N + n;
    if (e.y_out) Vec8<T>::store(reinterpret_cast<T*>(e.y_out) + off, v);
    float sc[8], rr[8];
    if (e.scale) Vec8<float>::load(e.scale + n, sc);
    if (e.resid) Vec8<float>::load(e.resid + off, rr);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float s = e.scale ? e.sign * sc[i] : e.sign;
      const float r = e.resid ? rr[i] : 0.f;
      v[i] = r + s * v[i];
    }
    Vec8<float>::store(e.out + off, v);
  } else if constexpr (EPI == DB200_EPI_GEGLU_BWD) {
    const T* u = reinterpret_cast<const T*>(e.u_in) + (long long)m * (2 * e.hidden);
    float a[8], g[8], da[8], dg[8];
    Vec8<T>::load(u + n, a);
    Vec8<T>::load(u + e.hidden + n, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float f, df;
      gelu_pair<T>(g[i], f, df);
      da[i] = v[i] * f;
      dg[i] = v[i] * a[i] * df;
    }
    T* du = reinterpret_cast<T*>(e.du_out) + (long long)m * (2 * e.hidden);
    Vec8<T>::store(du + n, da);
    Vec8<T>::store(du + e.hidden + n, dg);
  }
}

// GEGLU forward: a[8], g[8] are hidden indices j .. j+7 of row m
template <typename T>
__device__ __forceinline__ void epi_geglu_vec8(const EpiArgs& e, int m, int j, float* a, float* g) {
  if (e.bias) {
    float ba[8], bg[8];
    Vec8<float>::load(e.bias + j, ba);
    Vec8<float>::load(e.bias + e.hidden + j, bg);
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] += ba[i]; g[i] += bg[i]; }
  }
  if (e.u_out) {
    T* u = reinterpret_cast<T*>(e.u_out) + (long long)m * (2 * e.hidden);
    Vec8<T>::store(u + j, a);
    Vec8<T>::store(u + e.hidden + j, g);
  }
  float h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = a[i] * gelu_fwd<T>(g[i]);
  Vec8<T>::store(reinterpret_cast<T*>(e.h_out) + (long long)m * e.hidden + j, h);
}


// =====================================================================================================================
// "cols" epilogue granule for the tcgen05 kernel: after the smem transpose a lane owns TWO adjacent columns (n, n+1) of 16
// rows (m = mbase + 2*it, it = 0..15; t[2*it], t[2*it+1] are the accumulators).  Per-column constants (bias, LayerScale,
// head/offset arithmetic) are hoisted out of the row loop and every global INPUT of the chunk is fetched up front (batched
// read-only loads), so the 16 rows do not serialise on memory latency.
// =====================================================================================================================
template <int EPI, typename T>
__device__ __forceinline__ void epi_chunk_cols(const EpiArgs& e, int mbase, int n, const float* t, int M) {
  if constexpr (EPI == DB200_EPI_STORE) {
    float b0 = 0.f, b1 = 0.f;
    if (e.bias) { const float2 bb = __ldg(reinterpret_cast<const float2*>(e.bias + n)); b0 = bb.x; b1 = bb.y; }
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int m = mbase + 2 * it;
      if (m < M) {
        const long long off = (long long)m * e.ldc + n;
        if (e.atomic_c == 2) mc_red_add_v2(reinterpret_cast<float*>(e.C) + off, t[2 * it] * e.c_scale, t[2 * it + 1] * e.c_scale);
        else if (e.atomic_c) { atomicAdd(reinterpret_cast<float*>(e.C) + off, t[2 * it]); atomicAdd(reinterpret_cast<float*>(e.C) + off + 1, t[2 * it + 1]); }
        else if (e.c_is_f32) store2<float>(reinterpret_cast<float*>(e.C) + off, t[2 * it] + b0, t[2 * it + 1] + b1);
        else store2<T>(reinterpret_cast<T*>(e.C) + off, t[2 * it] + b0, t[2 * it + 1] + b1);
      }
    }
  } else if constexpr (EPI == DB200_EPI_RESID) {
    float b0 = 0.f, b1 = 0.f, s0 = e.sign, s1 = e.sign;
    if (e.bias) { const float2 bb = __ldg(reinterpret_cast<const float2*>(e.bias + n)); b0 = bb.x; b1 = bb.y; }
    if (e.scale) { const float2 sc = __ldg(reinterpret_cast<const float2*>(e.scale + n)); s0 *= sc.x; s1 *= sc.y; }
    float2 rr[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int m = mbase + 2 * it;
      rr[it] = make_float2(0.f, 0.f);
      if (e.resid && m < M) rr[it] = __ldg(reinterpret_cast<const float2*>(e.resid + (long long)m * e.N + n));
    }
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int m = mbase + 2 * it;
      if (m < M) {
        const long long off = (long long)m * e.N + n;
        const float v0 = t[2 * it] + b0, v1 = t[2 * it + 1] + b1;
        if (e.y_out) store2<T>(reinterpret_cast<T*>(e.y_out) + off, v0, v1);
        *reinterpret_cast<float2*>(e.out + off) = make_float2(rr[it].x + s0 * v0, rr[it].y + s1 * v1);
      }
    }
  } else if constexpr (EPI == DB200_EPI_QKV) {
    const int inner = e.heads * e.dim_head;
    const int which = n / inner;
    const int rem = n - which * inner;
    const int head = rem / e.dim_head;
    const int d = rem - head * e.dim_head;
    const float qs = which == 0 ? e.q_scale : 1.0f;
    T* base = reinterpret_cast<T*>(which == 0 ? e.q : (which == 1 ? e.k : e.v));
    const int b0 = mbase / e.seq_n;
    const int p0 = mbase - b0 * e.seq_n;
    float cs[16], sn[16];
    {
      int b = b0, p = p0;
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        while (p >= e.seq_n) { p -= e.seq_n; ++b; }
        cs[it] = 1.f; sn[it] = 0.f;
        if (e.cos_t && mbase + 2 * it < M) {
          const int ti = (p + e.pos_offset) * (e.dim_head >> 1) + (d >> 1);
          cs[it] = __ldg(e.cos_t + ti);
          sn[it] = __ldg(e.sin_t + ti);
        }
        p += 2;
      }
    }
    int b = b0, p = p0;
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      while (p >= e.seq_n) { p -= e.seq_n; ++b; }
      if (mbase + 2 * it < M) {
        const float x0 = t[2 * it], x1 = t[2 * it + 1];
        T* dst = base + (((long long)b * e.heads + head) * e.seq_n + p) * e.dim_head + d;
        store2<T>(dst, (x0 * cs[it] + (-x1) * sn[it]) * qs, (x1 * cs[it] + x0 * sn[it]) * qs);
      }
      p += 2;
    }
  } else if constexpr (EPI == DB200_EPI_GEGLU_BWD) {
    float2 a[16], g[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int m = mbase + 2 * it;
      a[it] = make_float2(0.f, 0.f); g[it] = make_float2(0.f, 0.f);
      if (m < M) {
        const T* u = reinterpret_cast<const T*>(e.u_in) + (long long)m * (2 * e.hidden);
        a[it] = load2<T>(u + n);
        g[it] = load2<T>(u + e.hidden + n);
      }
    }
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int m = mbase + 2 * it;
      if (m < M) {
        float f0, f1, df0, df1;
        gelu_pair<T>(g[it].x, f0, df0);
        gelu_pair<T>(g[it].y, f1, df1);
        T* du = reinterpret_cast<T*>(e.du_out) + (long long)m * (2 * e.hidden);
        store2<T>(du + n, t[2 * it] * f0, t[2 * it + 1] * f1);
        store2<T>(du + e.hidden + n, t[2 * it] * a[it].x * df0, t[2 * it + 1] * a[it].y * df1);
      }
    }
  }
}

// GEGLU forward, cols granule: ta / tg = a and gate accumulators of hidden columns (j, j+1) for the 16 rows
template <typename T>
__device__ __forceinline__ void epi_chunk_cols_geglu(const EpiArgs& e, int mbase, int j, const float* ta, const float* tg, int M) {
  float ba0 = 0.f, ba1 = 0.f, bg0 = 0.f, bg1 = 0.f;
  if (e.bias) {
    const float2 ba = __ldg(reinterpret_cast<const float2*>(e.bias + j));
    const float2 bg = __ldg(reinterpret_cast<const float2*>(e.bias + e.hidden + j));
    ba0 = ba.x; ba1 = ba.y; bg0 = bg.x; bg1 = bg.y;
  }
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int m = mbase + 2 * it;
    if (m < M) {
      const float a0 = ta[2 * it] + ba0, a1 = ta[2 * it + 1] + ba1, g0 = tg[2 * it] + bg0, g1 = tg[2 * it + 1] + bg1;
      if (e.u_out) {
        T* u = reinterpret_cast<T*>(e.u_out) + (long long)m * (2 * e.hidden);
        store2<T>(u + j, a0, a1);
        store2<T>(u + e.hidden + j, g0, g1);
      }
      float2 f, df;
      gelu_pair2<T>(make_float2(g0, g1), f, df);
      store2<T>(reinterpret_cast<T*>(e.h_out) + (long long)m * e.hidden + j, a0 * f.x, a1 * f.y);
    }
  }
}

}  // namespace db200
