// Tensor-core (bf16) fused attention, forward and backward, for every sparsity pattern of the reference.
//
// Flash-style: scores never leave registers, softmax statistics in fp32, P/dS rounded to bf16 only as MMA operands.
// 64-query x 64-key tiles, 4 warps (each warp owns 16 rows of the tile), operands staged with cp.async into
// XOR-swizzled shared memory and fed to mma.sync.m16n8k16 through ldmatrix.  Tiles with no allowed (query,key) pair
// are skipped, tiles that are entirely allowed skip the predicate.
//
//   forward   : one CTA per (query tile, batch*head); K/V tiles double-buffered.
//   backward  : delta = rowsum(dO * O); then two kernels that each recompute P and dS from the saved log-sum-exp:
//               dK/dV (one CTA per key tile, loops over query tiles) and dQ (one CTA per query tile, loops over key
//               tiles) — no atomics, deterministic.  The adjoint of the rotary rotation and the q scale are applied in
//               the final store, which writes straight into the [rows, 3*h*64] gradient of the to_qkv output.
//
// This is the legacy-MMA (HMMA) tensor path; attention is ~7 % of the step FLOPs — the tcgen05/TMEM
// version is future work, the GEMMs that carry 92 % of the FLOPs already run on tcgen05 (gemm_tcgen05.cu).
#include "attn_common.cuh"

namespace db200 {

namespace {

constexpr int T64 = 64;           // tile edge (queries and keys)
constexpr int DH = 64;
constexpr float LOG2E = 1.4426950408889634f;
using bf16 = __nv_bfloat16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 64 x 64 bf16 tile, 128 B per row, 16-byte chunks XOR-swizzled with the row index
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows [r0, r0+64) of a [nrows, 64] bf16 matrix with row stride `ld` elements -> swizzled tile (zero fill past nrows)
__device__ __forceinline__ void load_tile_async(uint32_t sbase, const bf16* __restrict__ src, long long ld, int r0, int nrows) {
  for (int idx = threadIdx.x; idx < 64 * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    const bool ok = (r0 + r) < nrows;
    const bf16* p = src + (long long)(ok ? r0 + r : 0) * ld + c * 8;
    cp_async16(tile_addr(sbase, r, c), p, ok);
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// A fragments (16 rows x 64 k) of rows [row0, row0+16) of a swizzled tile: frag[ks][4], ks = k16 step
__device__ __forceinline__ void load_a_frags(uint32_t (*frag)[4], uint32_t sbase, int row0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(frag[ks], tile_addr(sbase, row0 + (lane & 15), ks * 2 + (lane >> 4)));
}

// acc[nt][4] (16 x 64 output, nt = n8 tile) += A(16 x 64, frags) * B, with B(k, n) = X[n][k], X a swizzled row-major tile
// (rows = n).  Used for S = Q K^T (X = K), dP = dO V^T (X = V) and the transposed variants.
__device__ __forceinline__ void mma_a_xt(float (*acc)[4], const uint32_t (*a)[4], uint32_t xbase, int lane) {
  const int mat = lane >> 3, r = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, tile_addr(xbase, np * 16 + (mat >> 1) * 8 + r, ks * 2 + (mat & 1)));
      mma_bf16(acc[2 * np], a[ks], b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}

// acc[nt][4] (16 x 64) += A(16 x 64 k, frags) * X, with X[k][n] a swizzled row-major tile (rows = k).
// Used for O = P V, dV = P^T dO, dK = dS^T Q, dQ = dS K.
__device__ __forceinline__ void mma_a_x(float (*acc)[4], const uint32_t (*a)[4], uint32_t xbase, int lane) {
  const int mat = lane >> 3, r = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, tile_addr(xbase, ks * 16 + (mat & 1) * 8 + r, np * 2 + (mat >> 1)));
      mma_bf16(acc[2 * np], a[ks], b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}

// C-fragment (16 x 64 fp32) -> A fragments (bf16) of the same 16 x 64 matrix
__device__ __forceinline__ void c_to_a(uint32_t (*a)[4], const float (*c)[4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_bf16(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_bf16(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_bf16(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_bf16(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct FwdPtrs {
  const bf16* q; const bf16* k; const bf16* v; bf16* out; float* lse; const uint8_t* key_mask; int heads;
};

// =============================================== forward ===================================================
__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(FwdPtrs P, AttnGeom g) {
  __shared__ __align__(128) bf16 sQ[T64 * DH];
  __shared__ __align__(128) bf16 sK[2][T64 * DH];
  __shared__ __align__(128) bf16 sV[2][T64 * DH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  const int q0 = blockIdx.x * T64;
  const int off = g.n_k - g.n_q;
  const bf16* Q = P.q + (long long)bh * g.n_q * DH;
  const bf16* K = P.k + (long long)bh * g.kv_rows * DH;      // kv_rows >= n_k rows are allocated per (b,h)
  const bf16* V = P.v + (long long)bh * g.kv_rows * DH;
  const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * g.n_k : nullptr;
  const uint32_t sq = smem_u32(sQ), sk[2] = {smem_u32(sK[0]), smem_u32(sK[1])}, sv[2] = {smem_u32(sV[0]), smem_u32(sV[1])};

  const int q_last = min(q0 + T64, g.n_q) - 1;
  const int nkt = (g.n_k + T64 - 1) / T64;
  auto needed = [&](int kt) { return attn_tile_needed(g, q0 + off, q_last + off, kt * T64, min(kt * T64 + T64, g.n_k) - 1); };
  auto next_needed = [&](int kt) { while (kt < nkt && !needed(kt)) ++kt; return kt; };

  load_tile_async(sq, Q, DH, q0, g.n_q);
  int kt = next_needed(0);
  if (kt < nkt) { load_tile_async(sk[0], K, DH, kt * T64, g.n_k); load_tile_async(sv[0], V, DH, kt * T64, g.n_k); }
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-1.0e30f, -1.0e30f}, l_run[2] = {0.f, 0.f};
  const int row_lo = q0 + warp * 16 + (lane >> 2);       // this thread's rows: row_lo and row_lo + 8
  bool have_q = false;
  int buf = 0;
  while (kt < nkt) {
    const int kt_next = next_needed(kt + 1);
    if (kt_next < nkt) {
      load_tile_async(sk[buf ^ 1], K, DH, kt_next * T64, g.n_k);
      load_tile_async(sv[buf ^ 1], V, DH, kt_next * T64, g.n_k);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (!have_q) { load_a_frags(qf, sq, warp * 16, lane); have_q = true; }
    const int k0 = kt * T64, k1 = min(k0 + T64, g.n_k) - 1;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_a_xt(s, qf, sk[buf], lane);
    const bool full = (k1 - k0 == T64 - 1) && km == nullptr && attn_tile_full(g, q0 + off, q_last + off, k0, k1);
    if (!full) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = row_lo + (e >> 1) * 8;
          const int kj = k0 + nt * 8 + (lane & 3) * 2 + (e & 1);
          bool ok = (qi < g.n_q) && (kj < g.n_k) && attn_allowed(g, qi + off, kj);
          if (ok && km) ok = km[kj] != 0;
          if (!ok) s[nt][e] = -INFINITY;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mx = fmaxf(mx, fmaxf(s[nt][2 * r], s[nt][2 * r + 1]));
      mx = quad_max(mx);
      const float m_new = fmaxf(m_run[r], mx);
      const float corr = exp2f((m_run[r] - m_new) * LOG2E);
      const float mb = m_new * LOG2E;
      float rs = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = exp2f(fmaf(s[nt][2 * r], LOG2E, -mb));
        const float p1 = exp2f(fmaf(s[nt][2 * r + 1], LOG2E, -mb));
        s[nt][2 * r] = p0; s[nt][2 * r + 1] = p1;
        rs += p0 + p1;
      }
      l_run[r] = l_run[r] * corr + rs;
      m_run[r] = m_new;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { o[nt][2 * r] *= corr; o[nt][2 * r + 1] *= corr; }
    }
    uint32_t pa[4][4];
    c_to_a(pa, s);
    mma_a_x(o, pa, sv[buf], lane);
    __syncthreads();                     // everyone is done with buffer `buf` before it is refilled
    kt = kt_next;
    buf ^= 1;
  }
  cp_async_wait<0>();
  const int inner = P.heads * DH;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = row_lo + r * 8;
    const float l = quad_sum(l_run[r]);
    if (qi >= g.n_q) continue;
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    bf16* orow = P.out + ((long long)b * g.n_q + qi) * inner + h * DH + (lane & 3) * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<uint32_t*>(orow + nt * 8) = pack_bf16(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
    if ((lane & 3) == 0) P.lse[(long long)bh * g.n_q + qi] = l > 0.f ? m_run[r] + logf(l) : 0.f;
  }
}

// =============================================== backward ==================================================
__global__ void attn_delta_bf16_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ delta, int batch,
                                       int heads, int n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = batch * heads * n;
  if (warp >= total) return;
  const int i = warp % n, bh = warp / n, h = bh % heads, b = bh / heads;
  const long long off = ((long long)b * n + i) * heads * DH + h * DH + lane * 2;
  const float2 o = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(O + off));
  const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dO + off));
  const float s = warp_sum(o.x * d.x + o.y * d.y);
  if (lane == 0) delta[(long long)bh * n + i] = s;
}

struct BwdPtrs {
  const bf16* q; const bf16* k; const bf16* v; const bf16* d_out; const float* lse; const float* delta;
  const uint8_t* key_mask; const float* cos_t; const float* sin_t; float q_scale; bf16* dqkv; int heads;
};

// dK, dV.  Works on the TRANSPOSED score tile: warp w owns keys [k0+16w, k0+16w+16) as MMA rows, queries are columns.
__global__ void __launch_bounds__(128) attn_bwd_dkv_mma_kernel(BwdPtrs P, AttnGeom g) {
  __shared__ __align__(128) bf16 sQ[2][T64 * DH];
  __shared__ __align__(128) bf16 sdO[2][T64 * DH];
  __shared__ float s_lse[2][T64], s_delta[2][T64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  const int n = g.n_k, inner = P.heads * DH;
  const int k0 = blockIdx.x * T64, k1 = min(k0 + T64, n) - 1;
  const bf16* Q = P.q + (long long)bh * n * DH;
  const bf16* K = P.k + (long long)bh * n * DH;
  const bf16* V = P.v + (long long)bh * n * DH;
  const bf16* dO = P.d_out + (long long)b * n * inner + h * DH;
  const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * n : nullptr;
  const uint32_t sq[2] = {smem_u32(sQ[0]), smem_u32(sQ[1])}, sdo[2] = {smem_u32(sdO[0]), smem_u32(sdO[1])};
  const int nqt = (n + T64 - 1) / T64;
  auto needed = [&](int qt) { return attn_tile_needed(g, qt * T64, min(qt * T64 + T64, n) - 1, k0, k1); };
  auto next_needed = [&](int qt) { while (qt < nqt && !needed(qt)) ++qt; return qt; };
  auto load_q = [&](int qt, int bf) {
    load_tile_async(sq[bf], Q, DH, qt * T64, n);
    load_tile_async(sdo[bf], dO, inner, qt * T64, n);
    if (threadIdx.x < T64) {
      const int qi = qt * T64 + threadIdx.x;
      s_lse[bf][threadIdx.x] = qi < n ? P.lse[(long long)bh * n + qi] : 0.f;
      s_delta[bf][threadIdx.x] = qi < n ? P.delta[(long long)bh * n + qi] : 0.f;
    }
  };
  // this CTA's K and V tiles live in registers (A fragments) for the whole kernel; they are staged through the
  // second pipeline buffer before the query loop starts using it
  load_tile_async(sq[1], K, DH, k0, n);
  load_tile_async(sdo[1], V, DH, k0, n);
  int qt = next_needed(0);
  if (qt < nqt) load_q(qt, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
  load_a_frags(kf, sq[1], warp * 16, lane);
  load_a_frags(vf, sdo[1], warp * 16, lane);
  __syncthreads();

  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
  const int key_lo = k0 + warp * 16 + (lane >> 2);
  int buf = 0;
  while (qt < nqt) {
    const int qt_next = next_needed(qt + 1);
    if (qt_next < nqt) load_q(qt_next, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int qq0 = qt * T64, qq1 = min(qq0 + T64, n) - 1;
    float st[8][4], dpt[8][4];            // S^T and dP^T : rows = keys, cols = queries
#pragma unroll
    for (int i = 0; i < 8; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
    mma_a_xt(st, kf, sq[buf], lane);      // S^T = K Q^T
    mma_a_xt(dpt, vf, sdo[buf], lane);    // dP^T = V dO^T
    const bool full = (k1 - k0 == T64 - 1) && (qq1 - qq0 == T64 - 1) && km == nullptr && attn_tile_full(g, qq0, qq1, k0, k1);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kj = key_lo + (e >> 1) * 8;
        const int ql = nt * 8 + (lane & 3) * 2 + (e & 1);
        const int qi = qq0 + ql;
        bool ok = true;
        if (!full) {
          ok = (qi < n) && (kj < n) && attn_allowed(g, qi, kj);
          if (ok && km) ok = km[kj] != 0;
        }
        const float p = ok ? exp2f((st[nt][e] - s_lse[buf][ql]) * LOG2E) : 0.f;
        st[nt][e] = p;
        dpt[nt][e] = p * (dpt[nt][e] - s_delta[buf][ql]);
      }
    }
    uint32_t pa[4][4], dsa[4][4];
    c_to_a(pa, st);
    c_to_a(dsa, dpt);
    mma_a_x(dv, pa, sdo[buf], lane);      // dV += P^T dO
    mma_a_x(dk, dsa, sq[buf], lane);      // dK += dS^T Q     (Q already carries dh^-0.5)
    __syncthreads();
    qt = qt_next;
    buf ^= 1;
  }
  cp_async_wait<0>();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kj = key_lo + r * 8;
    if (kj >= n) continue;
    bf16* row = P.dqkv + ((long long)b * n + kj) * (3 * inner) + h * DH;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = nt * 8 + (lane & 3) * 2;
      float a0 = dk[nt][2 * r], a1 = dk[nt][2 * r + 1], c0 = dv[nt][2 * r], c1 = dv[nt][2 * r + 1];
      if (P.cos_t) {
        const int ti = kj * (DH / 2) + (d >> 1);
        const float c = P.cos_t[ti], s = P.sin_t[ti];
        rotary_adjoint(c, s, a0, a1);
        rotary_adjoint(c, s, c0, c1);
      }
      *reinterpret_cast<uint32_t*>(row + inner + d) = pack_bf16(a0, a1);
      *reinterpret_cast<uint32_t*>(row + 2 * inner + d) = pack_bf16(c0, c1);
    }
  }
}

// dQ.  Warp w owns queries [q0+16w, q0+16w+16).
__global__ void __launch_bounds__(128) attn_bwd_dq_mma_kernel(BwdPtrs P, AttnGeom g) {
  __shared__ __align__(128) bf16 sQ[T64 * DH];
  __shared__ __align__(128) bf16 sdO[T64 * DH];
  __shared__ __align__(128) bf16 sK[2][T64 * DH];
  __shared__ __align__(128) bf16 sV[2][T64 * DH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  const int n = g.n_k, inner = P.heads * DH;
  const int q0 = blockIdx.x * T64, q1 = min(q0 + T64, n) - 1;
  const bf16* Q = P.q + (long long)bh * n * DH;
  const bf16* K = P.k + (long long)bh * n * DH;
  const bf16* V = P.v + (long long)bh * n * DH;
  const bf16* dO = P.d_out + (long long)b * n * inner + h * DH;
  const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * n : nullptr;
  const uint32_t sq = smem_u32(sQ), sdo = smem_u32(sdO);
  const uint32_t sk[2] = {smem_u32(sK[0]), smem_u32(sK[1])}, sv[2] = {smem_u32(sV[0]), smem_u32(sV[1])};
  const int nkt = (n + T64 - 1) / T64;
  auto needed = [&](int kt) { return attn_tile_needed(g, q0, q1, kt * T64, min(kt * T64 + T64, n) - 1); };
  auto next_needed = [&](int kt) { while (kt < nkt && !needed(kt)) ++kt; return kt; };
  load_tile_async(sq, Q, DH, q0, n);
  load_tile_async(sdo, dO, inner, q0, n);
  int kt = next_needed(0);
  if (kt < nkt) { load_tile_async(sk[0], K, DH, kt * T64, n); load_tile_async(sv[0], V, DH, kt * T64, n); }
  cp_async_commit();
  const int row_lo = q0 + warp * 16 + (lane >> 2);
  float lse_r[2], delta_r[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = row_lo + r * 8;
    lse_r[r] = qi < n ? P.lse[(long long)bh * n + qi] : 0.f;
    delta_r[r] = qi < n ? P.delta[(long long)bh * n + qi] : 0.f;
  }
  uint32_t qf[4][4], dof[4][4];
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  bool have_q = false;
  int buf = 0;
  while (kt < nkt) {
    const int kt_next = next_needed(kt + 1);
    if (kt_next < nkt) { load_tile_async(sk[buf ^ 1], K, DH, kt_next * T64, n); load_tile_async(sv[buf ^ 1], V, DH, kt_next * T64, n); }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (!have_q) { load_a_frags(qf, sq, warp * 16, lane); load_a_frags(dof, sdo, warp * 16, lane); have_q = true; }
    const int k0 = kt * T64, k1 = min(k0 + T64, n) - 1;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
    mma_a_xt(s, qf, sk[buf], lane);       // S = Q K^T
    mma_a_xt(dp, dof, sv[buf], lane);     // dP = dO V^T
    const bool full = (k1 - k0 == T64 - 1) && km == nullptr && attn_tile_full(g, q0, q1, k0, k1);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int qi = row_lo + r * 8;
        const int kj = k0 + nt * 8 + (lane & 3) * 2 + (e & 1);
        bool ok = true;
        if (!full) {
          ok = (qi < n) && (kj < n) && attn_allowed(g, qi, kj);
          if (ok && km) ok = km[kj] != 0;
        }
        const float p = ok ? exp2f((s[nt][e] - lse_r[r]) * LOG2E) : 0.f;
        dp[nt][e] = p * (dp[nt][e] - delta_r[r]);
      }
    }
    uint32_t dsa[4][4];
    c_to_a(dsa, dp);
    mma_a_x(dq, dsa, sk[buf], lane);      // dQ += dS K
    __syncthreads();
    kt = kt_next;
    buf ^= 1;
  }
  cp_async_wait<0>();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = row_lo + r * 8;
    if (qi >= n) continue;
    bf16* row = P.dqkv + ((long long)b * n + qi) * (3 * inner) + h * DH;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = nt * 8 + (lane & 3) * 2;
      float a0 = dq[nt][2 * r] * P.q_scale, a1 = dq[nt][2 * r + 1] * P.q_scale;   // q = rot(x) * scale (attention.py:69)
      if (P.cos_t) {
        const int ti = qi * (DH / 2) + (d >> 1);
        rotary_adjoint(P.cos_t[ti], P.sin_t[ti], a0, a1);
      }
      *reinterpret_cast<uint32_t*>(row + d) = pack_bf16(a0, a1);
    }
  }
}

}  // namespace

bool attn_mma_supported(const db200_attn_fwd_params& p) {
  return p.dtype == DB200_BF16 && p.dim_head == 64 && (reinterpret_cast<uintptr_t>(p.q) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.k) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.v) & 15) == 0;
}

int attn_fwd_mma_launch(const db200_attn_fwd_params& p, cudaStream_t st) {
  FwdPtrs P{reinterpret_cast<const bf16*>(p.q), reinterpret_cast<const bf16*>(p.k), reinterpret_cast<const bf16*>(p.v),
            reinterpret_cast<bf16*>(p.out), p.lse, p.key_mask, p.heads};
  const AttnGeom g = make_geom(p);
  dim3 grid(ceil_div(p.n_q, T64), p.batch * p.heads);
  attn_fwd_mma_kernel<<<grid, 128, 0, st>>>(P, g);
  DB200_LAUNCH_OK("attn_fwd_mma_kernel");
  return DB200_OK;
}

int attn_bwd_mma_launch(const db200_attn_bwd_params& p, cudaStream_t st) {
  const db200_attn_fwd_params& f = p.f;
  const int n = f.n_k;
  if ((reinterpret_cast<uintptr_t>(p.d_out) & 15) != 0) return set_error(DB200_ERR_BAD_ARG, "attn_bwd: d_out must be 16-byte aligned");
  BwdPtrs P{reinterpret_cast<const bf16*>(f.q), reinterpret_cast<const bf16*>(f.k), reinterpret_cast<const bf16*>(f.v),
            reinterpret_cast<const bf16*>(p.d_out), f.lse, p.delta, f.key_mask, p.cos_t, p.sin_t, p.q_scale,
            reinterpret_cast<bf16*>(p.dqkv), f.heads};
  const AttnGeom g = make_geom(f);
  const int total_rows = f.batch * f.heads * n;
  attn_delta_bf16_kernel<<<ceil_div(total_rows * 32, 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(f.out), P.d_out, p.delta, f.batch,
                                                                         f.heads, n);
  DB200_LAUNCH_OK("attn_delta_bf16_kernel");
  dim3 grid(ceil_div(n, T64), f.batch * f.heads);
  attn_bwd_dkv_mma_kernel<<<grid, 128, 0, st>>>(P, g);
  DB200_LAUNCH_OK("attn_bwd_dkv_mma_kernel");
  attn_bwd_dq_mma_kernel<<<grid, 128, 0, st>>>(P, g);
  DB200_LAUNCH_OK("attn_bwd_dq_mma_kernel");
  return DB200_OK;
}

}  // namespace db200
