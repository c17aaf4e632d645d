// Kernels of the graph-replayed decoding step (dalle_pytorch_b200/decode.py): one new token per sequence, the position read from
// DEVICE memory so that the captured launch sequence is valid for every token.
//
//   decode_shift_kernel       PreShiftToken's cache branch (transformer.py:155-170) for one token: exchanges the first half of the
//                             normalised token with the ring of the last `fmap` image tokens and casts to the compute type
//   decode_kv_append_kernel   writes the token's key / value rows at row *pos of the in-place KV cache (attention.py:71-76)
//   attn_decode_kernel        attention of ONE query per (batch, head) over the cached keys (attention.py:78-96 with n = 1):
//                             HBM-bound streaming of K and V (2 * n_k * 128 B per head) instead of a 128-query tensor-core tile
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "attn_common.cuh"

namespace db200 {

// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) decode_shift_kernel(const float* __restrict__ h, T* __restrict__ y, int d, int batch,
                                                          float* __restrict__ ring_top, float* __restrict__ ring_left,
                                                          const long long* __restrict__ pos, int text_len, int fmap) {
  const int b = blockIdx.x;
  const int quarter = d >> 2, half = d >> 1, lw = half - quarter;
  long long s = (*pos - text_len) % fmap;
  if (s < 0) s += fmap;
  const int slot = (int)s, prev = (slot + fmap - 1) % fmap;
  const float* hr = h + (long long)b * d;
  T* yr = y + (long long)b * d;
  float* top_w = ring_top + ((long long)slot * batch + b) * quarter;        // this token's slot = the token one row up (read, then replaced)
  float* left_w = ring_left + ((long long)slot * batch + b) * lw;
  const float* left_r = ring_left + ((long long)prev * batch + b) * lw;     // the previous token
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float v = hr[c];
    float o = v;
    if (c < quarter) {
      o = top_w[c];
      top_w[c] = v;
    } else if (c < half) {
      o = slot != 0 ? left_r[c - quarter] : 0.f;       // first token of an image row: nothing to its left (transformer.py:166-167)
      left_w[c - quarter] = v;                         // (prev != slot whenever slot != 0, so the read above never sees this store)
    }
    yr[c] = from_f32<T>(o);
  }
}

int decode_shift_launch(const float* h, void* y, int out_dtype, int batch, int d, float* ring_top, float* ring_left, const long long* pos,
                        int text_len, int fmap, cudaStream_t st) {
  if (batch == 0) return DB200_OK;
  if (out_dtype == DB200_F32)
    decode_shift_kernel<float><<<batch, 256, 0, st>>>(h, reinterpret_cast<float*>(y), d, batch, ring_top, ring_left, pos, text_len, fmap);
  else
    decode_shift_kernel<__nv_bfloat16><<<batch, 256, 0, st>>>(h, reinterpret_cast<__nv_bfloat16*>(y), d, batch, ring_top, ring_left, pos, text_len, fmap);
  DB200_LAUNCH_OK("decode_shift_kernel");
  return DB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) decode_kv_append_kernel(const T* __restrict__ k_new, const T* __restrict__ v_new, T* __restrict__ k_cache,
                                                              T* __restrict__ v_cache, int bh, int dh, int kv_rows,
                                                              const long long* __restrict__ pos) {
  const long long row = *pos;
  if (row < 0 || row >= kv_rows) return;                // (the host never schedules a step past the cache capacity)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bh * dh) return;
  const int head = i / dh, c = i - head * dh;
  const long long dst = ((long long)head * kv_rows + row) * dh + c;
  k_cache[dst] = k_new[i];
  v_cache[dst] = v_new[i];
}

int decode_kv_append_launch(const void* k_new, const void* v_new, void* k_cache, void* v_cache, int dtype, int bh, int dh, int kv_rows,
                            const long long* pos, cudaStream_t st) {
  const int total = bh * dh;
  if (total == 0) return DB200_OK;
  const int grid = (total + 255) / 256;
  if (dtype == DB200_F32)
    decode_kv_append_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(k_new), reinterpret_cast<const float*>(v_new),
                                                         reinterpret_cast<float*>(k_cache), reinterpret_cast<float*>(v_cache), bh, dh, kv_rows, pos);
  else
    decode_kv_append_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(k_new), reinterpret_cast<const __nv_bfloat16*>(v_new),
                                                                 reinterpret_cast<__nv_bfloat16*>(k_cache), reinterpret_cast<__nv_bfloat16*>(v_cache), bh, dh,
                                                                 kv_rows, pos);
  DB200_LAUNCH_OK("decode_kv_append_kernel");
  return DB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// One query per (batch, head): CTA = 4 warps, 8 lanes per key (16 bytes = 8 of the 64 head channels each), 16 keys per CTA pass.
// Keys are processed in blocks of 512 with an online softmax across blocks (running max m, running sum l, running output o):
//   1. scores of the block -> shared memory (dot product reduced over the 8 lanes of a key), block maximum
//   2. p = exp(s - m_new) in place, block sum
//   3. o = o * exp(m_old - m_new) + sum_j p_j * V_j  (the same lane <-> channel mapping, so V loads are 16 bytes per lane as well)
// Masked keys carry the reference's -FLT_MAX (attention.py:82-90), so a row with no allowed key attends uniformly like the reference.
// q is already scaled (QKV epilogue).  All loads of K and V are unconditional (clamped row index) so that the unrolled loop keeps
// several 16-byte loads per lane in flight: the kernel is a stream over 2 * n_k * 128 bytes per head.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int DA_THREADS = 128, DA_BLOCK = 512, DA_DH = 64;
constexpr int DA_U = 8;                       // 16-byte loads per lane issued back to back (bytes in flight per CTA: 128 * 8 * 16 = 16 KB)

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float* f) {      // bf16 -> fp32 is a 16-bit shift
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

__global__ void __launch_bounds__(DA_THREADS) attn_decode_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                                 const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out,
                                                                 float* __restrict__ lse, const uint8_t* __restrict__ key_mask, AttnGeom g,
                                                                 int heads) {
  __shared__ float sc[DA_BLOCK];
  __shared__ float red_max[DA_THREADS / 32], red_sum[DA_THREADS / 32];
  __shared__ float osum[DA_THREADS / 32][DA_DH];
  const int bh = blockIdx.x, b = bh / heads, hd = bh - b * heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane & 7, kq = lane >> 3;
  const int n_k = g.n_k, qi = n_k - 1;                 // the single query is the last of the n_k positions
  float qf[8];
  bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(q + (long long)bh * DA_DH + sub * 8)), qf);
  const uint4* kb = reinterpret_cast<const uint4*>(k + (long long)bh * g.kv_rows * DA_DH + sub * 8);      // row pitch = 8 uint4
  const uint4* vb = reinterpret_cast<const uint4*>(v + (long long)bh * g.kv_rows * DA_DH + sub * 8);
  const uint8_t* km = key_mask ? key_mask + (long long)b * n_k : nullptr;
  const int kofs = warp * 4 + kq;                      // this lane's key within a CTA pass of 16 keys
  float m_run = -FLT_MAX, l_run = 0.f;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;

  for (int j0 = 0; j0 < n_k; j0 += DA_BLOCK) {
    const int nb = n_k - j0 < DA_BLOCK ? n_k - j0 : DA_BLOCK;
    const int iters = (nb + 15) >> 4;
    // ---- 1a. raw scores (pure streaming loop: DA_U loads per lane in flight, no divergent code) ----
    for (int it0 = 0; it0 < iters; it0 += DA_U) {
      uint4 kv[DA_U];
#pragma unroll
      for (int u = 0; u < DA_U; ++u) {
        const int jj = (it0 + u) * 16 + kofs;
        kv[u] = __ldg(kb + (long long)(j0 + (jj < nb ? jj : nb - 1)) * (DA_DH / 8));
      }
#pragma unroll
      for (int u = 0; u < DA_U; ++u) {
        const int jj = (it0 + u) * 16 + kofs;
        float kf[8];
        bf16x8_to_f32(kv[u], kf);
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) part = fmaf(qf[i], kf[i], part);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        if (sub == 0 && jj < nb) sc[jj] = part;
      }
    }
    __syncthreads();
    // ---- 1b. pattern / key mask, block maximum ----
    float lmax = -FLT_MAX;
    for (int jj = tid; jj < nb; jj += DA_THREADS) {
      const int j = j0 + jj;
      const bool ok = (km == nullptr || km[j] != 0) && attn_allowed(g, qi, j);
      const float s = ok ? sc[jj] : -FLT_MAX;
      sc[jj] = s;
      lmax = fmaxf(lmax, s);
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, w));
    if (lane == 0) red_max[warp] = lmax;
    __syncthreads();
    const float m_new = fmaxf(fmaxf(m_run, fmaxf(red_max[0], red_max[1])), fmaxf(red_max[2], red_max[3]));
    const float alpha = __expf(m_run - m_new);
    // ---- 2. probabilities (each thread rewrites the entries it masked above) ----
    float lsum = 0.f;
    for (int jj = tid; jj < nb; jj += DA_THREADS) {
      const float p = __expf(sc[jj] - m_new);
      sc[jj] = p;
      lsum += p;
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, w);
    if (lane == 0) red_sum[warp] = lsum;
    __syncthreads();
    l_run = l_run * alpha + ((red_sum[0] + red_sum[1]) + (red_sum[2] + red_sum[3]));
    m_run = m_new;
    // ---- 3. output ----
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] *= alpha;
    for (int it0 = 0; it0 < iters; it0 += DA_U) {
      uint4 vv[DA_U];
#pragma unroll
      for (int u = 0; u < DA_U; ++u) {
        const int jj = (it0 + u) * 16 + kofs;
        vv[u] = __ldg(vb + (long long)(j0 + (jj < nb ? jj : nb - 1)) * (DA_DH / 8));
      }
#pragma unroll
      for (int u = 0; u < DA_U; ++u) {
        const int jj = (it0 + u) * 16 + kofs;
        const float p = jj < nb ? sc[jj] : 0.f;
        float vf[8];
        bf16x8_to_f32(vv[u], vf);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(p, vf[i], o[i]);
      }
    }
    __syncthreads();                                   // sc / red_* are rewritten by the next block
  }
  // the 4 keys of a warp pass, then the 4 warps
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) osum[warp][sub * 8 + i] = o[i];
  }
  __syncthreads();
  if (tid < DA_DH) {
    const float tot = (osum[0][tid] + osum[1][tid]) + (osum[2][tid] + osum[3][tid]);
    out[((long long)b * heads + hd) * DA_DH + tid] = __float2bfloat16_rn(tot / l_run);
    if (tid == 0) lse[bh] = m_run + logf(l_run);
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// DALLE_B200_DECODE_ATTN=0 sends single-query attention back to the general kernels (A/B timing, tests)
bool attn_decode_supported(const db200_attn_fwd_params& p) {
  if (!(p.n_q == 1 && p.n_k >= 1 && p.dtype == DB200_BF16 && p.dim_head == DA_DH && !p.gather && al16(p.q) && al16(p.k) && al16(p.v)))
    return false;
  const char* e = getenv("DALLE_B200_DECODE_ATTN");
  return !(e && e[0] == '0');
}

int attn_decode_launch(const db200_attn_fwd_params& p, cudaStream_t st) {
  const AttnGeom g = make_geom(p);
  attn_decode_kernel<<<p.batch * p.heads, DA_THREADS, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(p.q), reinterpret_cast<const __nv_bfloat16*>(p.k),
                                                                reinterpret_cast<const __nv_bfloat16*>(p.v), reinterpret_cast<__nv_bfloat16*>(p.out), p.lse,
                                                                p.key_mask, g, p.heads);
  DB200_LAUNCH_OK("attn_decode_kernel");
  return DB200_OK;
}

}  // namespace db200
