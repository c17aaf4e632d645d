// Flash-style fused attention on CUDA cores with fp32 arithmetic, for every sparsity pattern of the reference.
//
// Role: the fp32 PARITY path (softmax statistics, QK^T and PV all in fp32 FFMA so logits meet rtol 1e-3 /
// atol 1e-5 against the reference's CPU arithmetic) and the always-available cross-check for the tensor-core
// attention.  The [n,n] score matrix is never materialised (the reference materialises and saves it,
// attention.py:78-94); only out and the per-row log-sum-exp are written.  64x64 tiles, 256 threads, each thread
// owns a 4x4 patch of the score tile; tiles with no allowed (query,key) pair are skipped.
#include "attn_common.cuh"

namespace db200 {

namespace {

constexpr int TQ = 64, TK = 64, DH = 64;
constexpr int LD = 68;   // padded leading dimension (floats) of every smem tile

// ---- tile loaders ---------------------------------------------------------------------------------------
// rows [r0, r0+64) of src[rows, 64] (row stride `ld` elements) -> S[row][d]
template <typename T>
__device__ __forceinline__ void load_rows(float (*S)[LD], const T* __restrict__ src, long long ld, int r0, int nrows) {
  for (int idx = threadIdx.x; idx < 64 * 32; idx += blockDim.x) {
    const int r = idx >> 5, c = (idx & 31) << 1;
    float2 v = make_float2(0.f, 0.f);
    if (r0 + r < nrows) v = load2<T>(src + (long long)(r0 + r) * ld + c);
    S[r][c] = v.x; S[r][c + 1] = v.y;
  }
}
// same, transposed: S[d][row]
template <typename T>
__device__ __forceinline__ void load_rows_t(float (*S)[LD], const T* __restrict__ src, long long ld, int r0, int nrows) {
  for (int idx = threadIdx.x; idx < 64 * 32; idx += blockDim.x) {
    const int r = idx & 63, c = (idx >> 6) << 1;
    float2 v = make_float2(0.f, 0.f);
    if (r0 + r < nrows) v = load2<T>(src + (long long)(r0 + r) * ld + c);
    S[c][r] = v.x; S[c + 1][r] = v.y;
  }
}

// acc[i][j] += sum_d A[rowA(i)][d] * Bt[d][colB(j)] ; A given non-transposed (broadcast reads), B transposed
__device__ __forceinline__ void mma_abt(float acc[4][4], const float (*A)[LD], const float (*Bt)[LD], int ty, int tx) {
#pragma unroll 16
  for (int d = 0; d < DH; ++d) {
    const float a0 = A[ty * 4 + 0][d], a1 = A[ty * 4 + 1][d], a2 = A[ty * 4 + 2][d], a3 = A[ty * 4 + 3][d];
    const float4 b = *reinterpret_cast<const float4*>(&Bt[d][tx * 4]);
    const float av[4] = {a0, a1, a2, a3};
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}
// acc[i][j] += sum_r At[r][ty*4+i] * B[r][tx*4+j]   (both read as float4 along the fast index)
__device__ __forceinline__ void mma_atb(float acc[4][4], const float (*At)[LD], const float (*B)[LD], int ty, int tx, int nr) {
#pragma unroll 8
  for (int r = 0; r < nr; ++r) {
    const float4 a = *reinterpret_cast<const float4*>(&At[r][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&B[r][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct AttnPtrs {
  const void* q; const void* k; const void* v; void* out; float* lse;
  const uint8_t* key_mask;
  int batch, heads;
};

constexpr float NEG_BIG = -1.0e30f;

// ---- forward ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) attn_fwd_simt_kernel(AttnPtrs P, AttnGeom g) {
  extern __shared__ float smem[];
  float (*Qs)[LD] = reinterpret_cast<float (*)[LD]>(smem);
  float (*Kt)[LD] = reinterpret_cast<float (*)[LD]>(smem + 64 * LD);
  float (*Vs)[LD] = reinterpret_cast<float (*)[LD]>(smem + 2 * 64 * LD);
  float (*Pt)[LD] = reinterpret_cast<float (*)[LD]>(smem + 3 * 64 * LD);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  const int q0 = blockIdx.x * TQ;
  const int off = g.n_k - g.n_q;                       // absolute position of query 0
  const T* Q = reinterpret_cast<const T*>(P.q) + (long long)bh * g.n_q * DH;     // q: [b,h,n_q,64]
  const T* K = reinterpret_cast<const T*>(P.k) + (long long)bh * g.kv_rows * DH;      // kv_rows >= n_k rows are allocated per (b,h)
  const T* V = reinterpret_cast<const T*>(P.v) + (long long)bh * g.kv_rows * DH;
  const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * g.n_k : nullptr;

  load_rows<T>(Qs, Q, DH, q0, g.n_q);
  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = NEG_BIG; l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  const int q_last = min(q0 + TQ, g.n_q) - 1;
  const int nkt = (g.n_k + TK - 1) / TK;
  for (int kt = 0; kt < nkt; ++kt) {
    const int k0 = kt * TK;
    const int k1 = min(k0 + TK, g.n_k) - 1;
    if (!attn_tile_needed(g, q0 + off, q_last + off, k0, k1)) continue;   // uniform across the CTA
    __syncthreads();                                    // previous tile's readers of Kt/Vs/Pt are done
    load_rows_t<T>(Kt, K, DH, k0, g.n_k);
    load_rows<T>(Vs, V, DH, k0, g.n_k);
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    mma_abt(s, Qs, Kt, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      float mx = NEG_BIG;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kj = k0 + tx * 4 + j;
        bool ok = (qi < g.n_q) && (kj < g.n_k) && attn_allowed(g, qi + off, kj);
        if (ok && km) ok = km[kj] != 0;
        s[i][j] = ok ? s[i][j] : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
      mx = group16_max(mx);
      const float m_new = fmaxf(m_run[i], mx);
      const float corr = expf(m_run[i] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pv = expf(s[i][j] - m_new);        // exp(-inf) = 0 for masked entries
        s[i][j] = pv;
        rs += pv;
      }
      rs = group16_sum(rs);
      l_run[i] = l_run[i] * corr + rs;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
#pragma unroll
      for (int j = 0; j < 4; ++j) Pt[tx * 4 + j][ty * 4 + i] = s[i][j];
    }
    __syncthreads();
    mma_atb(o, Pt, Vs, ty, tx, TK);
  }
  T* O = reinterpret_cast<T*>(P.out);
  const int inner = P.heads * DH;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    if (qi >= g.n_q) continue;
    const float inv = l_run[i] > 0.f ? 1.0f / l_run[i] : 0.f;
    T* orow = O + ((long long)b * g.n_q + qi) * inner + h * DH + tx * 4;
    store2<T>(orow, o[i][0] * inv, o[i][1] * inv);
    store2<T>(orow + 2, o[i][2] * inv, o[i][3] * inv);
    if (tx == 0) P.lse[(long long)bh * g.n_q + qi] = l_run[i] > 0.f ? m_run[i] + logf(l_run[i]) : 0.f;
  }
}

// ---- backward ---------------------------------------------------------------------------------------------
// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]
template <typename T>
__global__ void attn_delta_kernel(const T* __restrict__ O, const T* __restrict__ dO, float* __restrict__ delta,
                                  int batch, int heads, int n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = batch * heads * n;
  if (warp >= total) return;
  const int i = warp % n, bh = warp / n, h = bh % heads, b = bh / heads;
  const long long off = ((long long)b * n + i) * heads * DH + h * DH + lane * 2;
  const float2 o = load2<T>(O + off), d = load2<T>(dO + off);
  const float s = warp_sum(o.x * d.x + o.y * d.y);
  if (lane == 0) delta[(long long)bh * n + i] = s;
}

struct AttnBwdPtrs {
  const void* q; const void* k; const void* v; const void* d_out; const float* lse; const float* delta;
  const uint8_t* key_mask;
  const float* cos_t; const float* sin_t; float q_scale;
  void* dqkv;
  int batch, heads;
};

// Recompute P and dS for one (query tile, key tile); p[i][j] / ds[i][j] for rows ty*4+i, keys tx*4+j.
__device__ __forceinline__ void recompute_p_ds(float p[4][4], float ds[4][4], const float (*Qs)[LD], const float (*Kt)[LD],
                                               const float (*dOs)[LD], const float (*Vt)[LD], const float* lse_s,
                                               const float* delta_s, const AttnGeom& g, const uint8_t* km, int q0, int k0,
                                               int ty, int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { p[i][j] = 0.f; ds[i][j] = 0.f; }
  mma_abt(p, Qs, Kt, ty, tx);      // S = Q K^T
  mma_abt(ds, dOs, Vt, ty, tx);    // dP = dO V^T
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    const float L = lse_s[ty * 4 + i], D = delta_s[ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kj = k0 + tx * 4 + j;
      bool ok = (qi < g.n_q) && (kj < g.n_k) && attn_allowed(g, qi, kj);
      if (ok && km) ok = km[kj] != 0;
      const float pv = ok ? expf(p[i][j] - L) : 0.f;
      p[i][j] = pv;
      ds[i][j] = pv * (ds[i][j] - D);
    }
  }
}

// dK, dV: one CTA per key tile, loops over the query tiles that can see it.
template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_dkv_simt_kernel(AttnBwdPtrs P, AttnGeom g) {
  extern __shared__ float smem[];
  float (*Qs)[LD] = reinterpret_cast<float (*)[LD]>(smem);
  float (*dOs)[LD] = reinterpret_cast<float (*)[LD]>(smem + 1 * 64 * LD);
  float (*Kt)[LD] = reinterpret_cast<float (*)[LD]>(smem + 2 * 64 * LD);
  float (*Vt)[LD] = reinterpret_cast<float (*)[LD]>(smem + 3 * 64 * LD);
  float (*Ps)[LD] = reinterpret_cast<float (*)[LD]>(smem + 4 * 64 * LD);
  float (*dSs)[LD] = reinterpret_cast<float (*)[LD]>(smem + 5 * 64 * LD);
  float* lse_s = smem + 6 * 64 * LD;
  float* delta_s = lse_s + 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  const int n = g.n_k, inner = P.heads * DH;
  const int k0 = blockIdx.x * TK, k1 = min(k0 + TK, n) - 1;
  const T* Q = reinterpret_cast<const T*>(P.q) + (long long)bh * n * DH;
  const T* K = reinterpret_cast<const T*>(P.k) + (long long)bh * n * DH;
  const T* V = reinterpret_cast<const T*>(P.v) + (long long)bh * n * DH;
  const T* dO = reinterpret_cast<const T*>(P.d_out) + (long long)b * n * inner + h * DH;
  const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * n : nullptr;
  load_rows_t<T>(Kt, K, DH, k0, n);
  load_rows_t<T>(Vt, V, DH, k0, n);
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { dk[i][j] = 0.f; dv[i][j] = 0.f; }
  const int nqt = (n + TQ - 1) / TQ;
  for (int qt = 0; qt < nqt; ++qt) {
    const int q0 = qt * TQ, q1 = min(q0 + TQ, n) - 1;
    if (!attn_tile_needed(g, q0, q1, k0, k1)) continue;
    __syncthreads();
    load_rows<T>(Qs, Q, DH, q0, n);
    load_rows<T>(dOs, dO, inner, q0, n);
    if (threadIdx.x < 64) {
      const int qi = q0 + threadIdx.x;
      lse_s[threadIdx.x] = qi < n ? P.lse[(long long)bh * n + qi] : 0.f;
      delta_s[threadIdx.x] = qi < n ? P.delta[(long long)bh * n + qi] : 0.f;
    }
    __syncthreads();
    float p[4][4], ds[4][4];
    recompute_p_ds(p, ds, Qs, Kt, dOs, Vt, lse_s, delta_s, g, km, q0, k0, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { Ps[ty * 4 + i][tx * 4 + j] = p[i][j]; dSs[ty * 4 + i][tx * 4 + j] = ds[i][j]; }
    __syncthreads();
    mma_atb(dv, Ps, dOs, ty, tx, TQ);     // dV[key][d] += sum_row P[row][key] dO[row][d]
    mma_atb(dk, dSs, Qs, ty, tx, TQ);     // dK[key][d] += sum_row dS[row][key] Q[row][d]   (Q carries dh^-0.5)
  }
  T* dqkv = reinterpret_cast<T*>(P.dqkv);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kj = k0 + ty * 4 + i;
    if (kj >= n) continue;
    const int d = tx * 4;
    if (P.cos_t) {
      const int ti = kj * (DH / 2) + (d >> 1);
      const float c0 = P.cos_t[ti], s0 = P.sin_t[ti], c1 = P.cos_t[ti + 1], s1 = P.sin_t[ti + 1];
      rotary_adjoint(c0, s0, dk[i][0], dk[i][1]); rotary_adjoint(c1, s1, dk[i][2], dk[i][3]);
      rotary_adjoint(c0, s0, dv[i][0], dv[i][1]); rotary_adjoint(c1, s1, dv[i][2], dv[i][3]);
    }
    T* row = dqkv + ((long long)b * n + kj) * (3 * inner) + h * DH + d;
    store2<T>(row + inner, dk[i][0], dk[i][1]); store2<T>(row + inner + 2, dk[i][2], dk[i][3]);
    store2<T>(row + 2 * inner, dv[i][0], dv[i][1]); store2<T>(row + 2 * inner + 2, dv[i][2], dv[i][3]);
  }
}

// dQ: one CTA per query tile, loops over the key tiles it can see.
template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_dq_simt_kernel(AttnBwdPtrs P, AttnGeom g) {
  extern __shared__ float smem[];
  float (*Qs)[LD] = reinterpret_cast<float (*)[LD]>(smem);
  float (*dOs)[LD] = reinterpret_cast<float (*)[LD]>(smem + 1 * 64 * LD);
  float (*Kt)[LD] = reinterpret_cast<float (*)[LD]>(smem + 2 * 64 * LD);
  float (*Vt)[LD] = reinterpret_cast<float (*)[LD]>(smem + 3 * 64 * LD);
  float (*Ks)[LD] = reinterpret_cast<float (*)[LD]>(smem + 4 * 64 * LD);
  float (*dSt)[LD] = reinterpret_cast<float (*)[LD]>(smem + 5 * 64 * LD);
  float* lse_s = smem + 6 * 64 * LD;
  float* delta_s = lse_s + 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  const int n = g.n_k, inner = P.heads * DH;
  const int q0 = blockIdx.x * TQ, q1 = min(q0 + TQ, n) - 1;
  const T* Q = reinterpret_cast<const T*>(P.q) + (long long)bh * n * DH;
  const T* K = reinterpret_cast<const T*>(P.k) + (long long)bh * n * DH;
  const T* V = reinterpret_cast<const T*>(P.v) + (long long)bh * n * DH;
  const T* dO = reinterpret_cast<const T*>(P.d_out) + (long long)b * n * inner + h * DH;
  const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * n : nullptr;
  load_rows<T>(Qs, Q, DH, q0, n);
  load_rows<T>(dOs, dO, inner, q0, n);
  if (threadIdx.x < 64) {
    const int qi = q0 + threadIdx.x;
    lse_s[threadIdx.x] = qi < n ? P.lse[(long long)bh * n + qi] : 0.f;
    delta_s[threadIdx.x] = qi < n ? P.delta[(long long)bh * n + qi] : 0.f;
  }
  float dq[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
  const int nkt = (n + TK - 1) / TK;
  for (int kt = 0; kt < nkt; ++kt) {
    const int k0 = kt * TK, k1 = min(k0 + TK, n) - 1;
    if (!attn_tile_needed(g, q0, q1, k0, k1)) continue;
    __syncthreads();
    load_rows_t<T>(Kt, K, DH, k0, n);
    load_rows_t<T>(Vt, V, DH, k0, n);
    load_rows<T>(Ks, K, DH, k0, n);
    __syncthreads();
    float p[4][4], ds[4][4];
    recompute_p_ds(p, ds, Qs, Kt, dOs, Vt, lse_s, delta_s, g, km, q0, k0, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dSt[tx * 4 + j][ty * 4 + i] = ds[i][j];
    __syncthreads();
    mma_atb(dq, dSt, Ks, ty, tx, TK);     // dQ[row][d] += sum_key dS[row][key] K[key][d]
  }
  T* dqkv = reinterpret_cast<T*>(P.dqkv);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    if (qi >= n) continue;
    const int d = tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] *= P.q_scale;        // q = rot(x) * scale  (attention.py:69)
    if (P.cos_t) {
      const int ti = qi * (DH / 2) + (d >> 1);
      rotary_adjoint(P.cos_t[ti], P.sin_t[ti], dq[i][0], dq[i][1]);
      rotary_adjoint(P.cos_t[ti + 1], P.sin_t[ti + 1], dq[i][2], dq[i][3]);
    }
    T* row = dqkv + ((long long)b * n + qi) * (3 * inner) + h * DH + d;
    store2<T>(row, dq[i][0], dq[i][1]); store2<T>(row + 2, dq[i][2], dq[i][3]);
  }
}

template <typename T>
int attn_fwd_simt_t(const db200_attn_fwd_params& p, cudaStream_t st) {
  AttnPtrs P{p.q, p.k, p.v, p.out, p.lse, p.key_mask, p.batch, p.heads};
  const AttnGeom g = make_geom(p);
  const size_t smem = (size_t)4 * 64 * LD * sizeof(float);
  static std::atomic<bool> attr_done{false};   // idempotent set-up; atomic because forward and autograd threads both launch
  if (!attr_done.load(std::memory_order_acquire)) {
    DB200_CUDA_OK(cudaFuncSetAttribute(attn_fwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done.store(true, std::memory_order_release);
  }
  dim3 grid(ceil_div(p.n_q, TQ), p.batch * p.heads);
  attn_fwd_simt_kernel<T><<<grid, 256, smem, st>>>(P, g);
  DB200_LAUNCH_OK("attn_fwd_simt_kernel");
  return DB200_OK;
}

template <typename T>
int attn_bwd_simt_t(const db200_attn_bwd_params& p, cudaStream_t st) {
  const db200_attn_fwd_params& f = p.f;
  const int n = f.n_k;
  AttnBwdPtrs P{f.q, f.k, f.v, p.d_out, f.lse, p.delta, f.key_mask, p.cos_t, p.sin_t, p.q_scale, p.dqkv, f.batch, f.heads};
  const AttnGeom g = make_geom(f);
  const size_t smem = ((size_t)6 * 64 * LD + 128) * sizeof(float);
  static std::atomic<bool> attr_done{false};   // idempotent set-up; atomic because forward and autograd threads both launch
  if (!attr_done.load(std::memory_order_acquire)) {
    DB200_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DB200_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done.store(true, std::memory_order_release);
  }
  const int total_rows = f.batch * f.heads * n;
  attn_delta_kernel<T><<<ceil_div(total_rows * 32, 256), 256, 0, st>>>(reinterpret_cast<const T*>(f.out),
                                                                       reinterpret_cast<const T*>(p.d_out), p.delta, f.batch,
                                                                       f.heads, n);
  DB200_LAUNCH_OK("attn_delta_kernel");
  dim3 grid(ceil_div(n, TQ), f.batch * f.heads);
  attn_bwd_dkv_simt_kernel<T><<<grid, 256, smem, st>>>(P, g);
  DB200_LAUNCH_OK("attn_bwd_dkv_simt_kernel");
  attn_bwd_dq_simt_kernel<T><<<grid, 256, smem, st>>>(P, g);
  DB200_LAUNCH_OK("attn_bwd_dq_simt_kernel");
  return DB200_OK;
}

}  // namespace

int attn_fwd_simt_launch(const db200_attn_fwd_params& p, cudaStream_t st) {
  if (p.dtype == DB200_F32) return attn_fwd_simt_t<float>(p, st);
  return attn_fwd_simt_t<__nv_bfloat16>(p, st);
}
int attn_bwd_simt_launch(const db200_attn_bwd_params& p, cudaStream_t st) {
  if (p.f.dtype == DB200_F32) return attn_bwd_simt_t<float>(p, st);
  return attn_bwd_simt_t<__nv_bfloat16>(p, st);
}

}  // namespace db200
