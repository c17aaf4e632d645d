// Attention sparsity patterns of the reference expressed as an allowed(query, key) predicate plus a
// conservative tile-skip test, shared by every attention kernel (SIMT fp32 and tensor-core bf16).
#pragma once
#include "common.cuh"

namespace db200 {

struct AttnGeom {
  int pattern, causal, text_len, fmap, ksize, dil;
  int n_q, n_k;                       // queries are the last n_q of the n_k key positions
  const uint8_t* static_mask; long long static_ld;
  // Gathered axial mode (tcgen05 kernels only, see "gathered axial tiling" below): `pattern` is AXIAL_ROW in VIRTUAL
  // coordinates, `col` says the virtual image order is column-major (axis = 1), rows per (b,h) of q/k/v = n_alloc,
  // entries per (b,h) of lse / delta = n_stat, t_pad = text_len rounded up to 64.
  int gather, col, n_alloc, n_stat, t_pad;
  int kv_rows;                        // rows allocated per (b,h) in k / v (>= n_k: in-place KV cache of the decoding path)
};

// layout constants of the gathered mode, shared with the host (ops.py mirrors them)
__host__ __device__ inline int gather_n_alloc(int text_len, int fmap) { return text_len + fmap * fmap; }
__host__ __device__ inline int gather_t_pad(int text_len) { return (text_len + 63) / 64 * 64; }
__host__ __device__ inline int gather_n_stat(int text_len, int fmap) { return gather_t_pad(text_len) + fmap * fmap; }

inline AttnGeom make_geom(const db200_attn_fwd_params& p) {
  AttnGeom g;
  g.pattern = p.pattern; g.causal = p.causal; g.text_len = p.text_len; g.fmap = p.fmap > 0 ? p.fmap : 1;
  g.ksize = p.kernel_size; g.dil = p.dilation > 0 ? p.dilation : 1;
  g.n_q = p.n_q; g.n_k = p.n_k; g.static_mask = p.static_mask; g.static_ld = p.static_ld;
  g.gather = p.gather != 0; g.col = 0; g.n_alloc = p.n_k; g.n_stat = p.n_q; g.t_pad = 0;
  g.kv_rows = p.kv_rows > p.n_k ? p.kv_rows : p.n_k;
  if (g.gather) {
    g.col = p.pattern == DB200_ATTN_AXIAL_COL;
    g.pattern = DB200_ATTN_AXIAL_ROW;           // column attention IS row attention in the column-major virtual order
    g.n_alloc = gather_n_alloc(p.text_len, g.fmap);
    g.n_stat = gather_n_stat(p.text_len, g.fmap);
    g.t_pad = gather_t_pad(p.text_len);
  }
  return g;
}

// ---------------------------------------------------------------------------------------------------------------------
// Gathered axial tiling (reference attention.py:287-323 reshapes the image keys by rows / columns and attends text + own axis).
// Virtual sequence: u in [0,T) = text position u; u = T + v = image token number v in AXIS-MAJOR order (row-major for
// axis 0 = the natural order, column-major for axis 1).  In this order both axial patterns are "same line of fm tokens, up to
// itself", i.e. the AXIAL_ROW predicate, and the token that does not exist in training (the last image token) is the last
// in either order.  Tiles are cut per SEGMENT: text tiles start at multiples of W from 0, image tiles at T + multiples of W, so
// an image tile is a whole number of lines and one TMA box: a contiguous run of rows for axis 0, a strided 4-D box
// {dh, fm rows of the grid, W/fm columns, 1} (row pitch fm tokens) for axis 1.
// ---------------------------------------------------------------------------------------------------------------------
struct SegTiles {
  int gather, T, n, W, nt_text, nt;
  __device__ SegTiles(const AttnGeom& g, int W_, int n_) : gather(g.gather), T(g.text_len), n(n_), W(W_) {
    if (gather) { nt_text = (T + W - 1) / W; nt = nt_text + g.fmap * g.fmap / W; }
    else { nt_text = 0; nt = (n + W - 1) / W; }
  }
  __device__ int count() const { return nt; }
  __device__ bool is_img(int t) const { return gather && t >= nt_text; }
  __device__ int origin(int t) const { return (gather && t >= nt_text) ? T + (t - nt_text) * W : t * W; }
  __device__ int limit(int t) const {           // exclusive end of the valid positions of tile t
    const int e = origin(t) + W;
    const int cap = (gather && t < nt_text) ? T : n;
    return e < cap ? e : cap;
  }
};
__host__ inline int seg_tile_count(const AttnGeom& g, int W, int n) {
  return g.gather ? (g.text_len + W - 1) / W + g.fmap * g.fmap / W : (n + W - 1) / W;
}

// virtual position -> position in the token sequence (identity unless column-major)
__device__ __forceinline__ int attn_nat(const AttnGeom& g, int u) {
  if (!g.col || u < g.text_len) return u;
  const int v = u - g.text_len, c = v / g.fmap, r = v - c * g.fmap;
  return g.text_len + r * g.fmap + c;
}
// token position -> virtual position
__device__ __forceinline__ int attn_virt(const AttnGeom& g, int p) {
  if (!g.col || p < g.text_len) return p;
  const int i = p - g.text_len, r = i / g.fmap, c = i - r * g.fmap;
  return g.text_len + c * g.fmap + r;
}
// virtual position -> index into the per-(b,h) lse / delta arrays (image entries start at t_pad so that every tile's
// statistics are one 16-byte aligned run)
__device__ __forceinline__ int attn_sidx(const AttnGeom& g, int u) {
  return (g.gather && u >= g.text_len) ? g.t_pad + (u - g.text_len) : u;
}

// i = absolute position of the query, j = key position.
//  FULL      attention.py:84-87
//  AXIAL_*   attention.py:271-314  (text: causal over text; image (r,c): all text + same row/col up to itself)
//  CONV_LIKE attention.py:147-207  (image (r,c): all text + k x k dilated window whose bottom-right corner is (r,c))
//  STATIC    attention.py:89-90 with transformer.py:333-350 masks
__device__ __forceinline__ bool attn_allowed(const AttnGeom& g, int i, int j) {
  switch (g.pattern) {
    case DB200_ATTN_FULL: return !g.causal || j <= i;
    case DB200_ATTN_STATIC: return (!g.causal || j <= i) && g.static_mask[(long long)i * g.static_ld + j] != 0;
    default: break;
  }
  const int T = g.text_len;
  if (j < T) return (i >= T) || (j <= i);
  if (i < T) return false;
  const int qi = i - T, kj = j - T;
  const int qr = qi / g.fmap, qc = qi - qr * g.fmap;
  const int kr = kj / g.fmap, kc = kj - kr * g.fmap;
  if (g.pattern == DB200_ATTN_AXIAL_ROW) return kr == qr && kc <= qc;
  if (g.pattern == DB200_ATTN_AXIAL_COL) return kc == qc && kr <= qr;
  // conv_like
  const int dr = qr - kr, dc = qc - kc, span = (g.ksize - 1) * g.dil;
  return dr >= 0 && dc >= 0 && dr <= span && dc <= span && (dr % g.dil) == 0 && (dc % g.dil) == 0;
}

// May any (i in [q0,q1], j in [k0,k1]) pair be allowed?  (absolute positions, inclusive)  Conservative: a
// `true` for a tile with no allowed pair only costs time, the in-tile predicate keeps the result exact.
__device__ __forceinline__ bool attn_tile_needed(const AttnGeom& g, int q0, int q1, int k0, int k1) {
  const bool causal_like = g.causal || (g.pattern != DB200_ATTN_FULL && g.pattern != DB200_ATTN_STATIC);
  if (causal_like && k0 > q1) return false;
  if (g.pattern == DB200_ATTN_FULL || g.pattern == DB200_ATTN_STATIC || g.pattern == DB200_ATTN_AXIAL_COL) return true;
  const int T = g.text_len;
  if (k0 < T) return true;                       // tile contains text keys
  if (q1 < T) return false;                      // image keys only, text queries only
  const int qs = (q0 > T ? q0 : T) - T;          // first image query of the tile
  int lo;
  if (g.pattern == DB200_ATTN_AXIAL_ROW) lo = (qs / g.fmap) * g.fmap;       // start of its row
  else {                                                                     // conv_like: top-left of its window
    const int span = (g.ksize - 1) * g.dil;
    lo = qs - span * g.fmap - span;
    if (lo < 0) lo = 0;
  }
  return k1 - T >= lo;
}

// Is EVERY (i in [q0,q1], j in [k0,k1]) pair allowed (so the per-element predicate can be skipped)?  Exact `true`s only.
__device__ __forceinline__ bool attn_tile_full(const AttnGeom& g, int q0, int q1, int k0, int k1) {
  switch (g.pattern) {
    case DB200_ATTN_FULL: return !g.causal || k1 <= q0;
    case DB200_ATTN_STATIC: return false;
    default: break;
  }
  return k1 < g.text_len && (q0 >= g.text_len || k1 <= q0);
}

// ---------------------------------------------------------------------------------------------------------------------
// 128-wide allowed-bit masks of one query row / one key column inside a tile.  For the reference's patterns the allowed set of
// a row is a union of a few index ranges (or an arithmetic progression for axial columns), so the mask costs O(1) per tile
// instead of 128 predicate evaluations with integer divisions; the softmax loops then only test bits.
// ---------------------------------------------------------------------------------------------------------------------
struct Mask128 { uint32_t w[4]; };

__device__ __forceinline__ uint32_t range_bits32(int lo, int hi, int base) {      // bits of [lo,hi] in [base, base+31]
  const int a = lo - base > 0 ? lo - base : 0;
  const int b = hi - base < 31 ? hi - base : 31;
  if (a > b) return 0u;
  return (0xffffffffu >> (31 - b)) & (0xffffffffu << a);
}
__device__ __forceinline__ void mask_or_range(Mask128& m, int lo, int hi, int t0) {
#pragma unroll
  for (int c = 0; c < 4; ++c) m.w[c] |= range_bits32(lo, hi, t0 + 32 * c);
}
__device__ __forceinline__ void mask_or_bit(Mask128& m, int pos, int t0) {
  const int d = pos - t0;
  if (d >= 0 && d < 128) m.w[d >> 5] |= 1u << (d & 31);
}

// bit b = may query i (absolute position) attend key k0 + b ;  keys >= n_k are never allowed
__device__ __forceinline__ Mask128 attn_row_bits(const AttnGeom& g, int i, int k0, const uint8_t* km, int width = 128) {
  Mask128 m = {{0u, 0u, 0u, 0u}};
  const int kend = (k0 + width - 1 < g.n_k - 1) ? k0 + width - 1 : g.n_k - 1;
  const int T = g.text_len, fm = g.fmap;
  switch (g.pattern) {
    case DB200_ATTN_FULL:
      mask_or_range(m, k0, g.causal ? (i < kend ? i : kend) : kend, k0);
      break;
    case DB200_ATTN_STATIC:
      for (int b = 0; k0 + b <= kend; ++b)
        if ((!g.causal || k0 + b <= i) && g.static_mask[(long long)i * g.static_ld + k0 + b]) m.w[b >> 5] |= 1u << (b & 31);
      break;
    default: {
      if (i < T) { mask_or_range(m, 0, i < kend ? i : kend, k0); break; }
      mask_or_range(m, 0, T - 1 < kend ? T - 1 : kend, k0);
      const int qi = i - T, qr = qi / fm, qc = qi - qr * fm;
      const int hi = i < kend ? i : kend;
      if (g.pattern == DB200_ATTN_AXIAL_ROW) {
        mask_or_range(m, T + qr * fm, hi, k0);
      } else if (g.pattern == DB200_ATTN_AXIAL_COL) {
        int j = T + qc;                                   // keys (r', qc), r' = 0..qr
        if (j < k0) j += ((k0 - j + fm - 1) / fm) * fm;
        for (; j <= hi; j += fm) mask_or_bit(m, j, k0);
      } else {                                            // conv_like: rows qr - a*dil, columns qc - b*dil
        const int span = (g.ksize - 1) * g.dil;
        for (int a = 0; a < g.ksize; ++a) {
          const int rr = qr - a * g.dil;
          if (rr < 0) break;
          const int rowbase = T + rr * fm;
          if (g.dil == 1) {
            mask_or_range(m, rowbase + (qc - span > 0 ? qc - span : 0), rowbase + qc < hi ? rowbase + qc : hi, k0);
          } else {
            for (int bb = 0; bb < g.ksize; ++bb) {
              const int cc = qc - bb * g.dil;
              if (cc < 0) break;
              if (rowbase + cc <= hi) mask_or_bit(m, rowbase + cc, k0);
            }
          }
        }
      }
    }
  }
  if (km) {
    for (int b = 0; k0 + b <= kend; ++b)
      if (!km[k0 + b]) m.w[b >> 5] &= ~(1u << (b & 31));
  }
  return m;
}

// bit b = may query q0 + b attend key j ;  queries >= n are never allowed (training: n_q == n_k == n, positions absolute)
__device__ __forceinline__ Mask128 attn_col_bits(const AttnGeom& g, int j, int q0, int n, int width = 128) {
  Mask128 m = {{0u, 0u, 0u, 0u}};
  const int qend = (q0 + width - 1 < n - 1) ? q0 + width - 1 : n - 1;
  const int T = g.text_len, fm = g.fmap;
  switch (g.pattern) {
    case DB200_ATTN_FULL:
      mask_or_range(m, g.causal ? (j > q0 ? j : q0) : q0, qend, q0);
      break;
    case DB200_ATTN_STATIC:
      for (int b = 0; q0 + b <= qend; ++b)
        if ((!g.causal || j <= q0 + b) && g.static_mask[(long long)(q0 + b) * g.static_ld + j]) m.w[b >> 5] |= 1u << (b & 31);
      break;
    default: {
      if (j < T) { mask_or_range(m, j > q0 ? j : q0, qend, q0); break; }   // text key: text queries >= j, every image query
      const int kj = j - T, kr = kj / fm, kc = kj - kr * fm;
      if (g.pattern == DB200_ATTN_AXIAL_ROW) {
        const int rowend = T + kr * fm + fm - 1;
        mask_or_range(m, j, rowend < qend ? rowend : qend, q0);
      } else if (g.pattern == DB200_ATTN_AXIAL_COL) {
        int i = j;
        if (i < q0) i += ((q0 - i + fm - 1) / fm) * fm;
        for (; i <= qend; i += fm) mask_or_bit(m, i, q0);
      } else {
        const int span = (g.ksize - 1) * g.dil;
        for (int a = 0; a < g.ksize; ++a) {
          const int rr = kr + a * g.dil;
          if (rr >= fm) break;
          const int rowbase = T + rr * fm;
          if (g.dil == 1) {
            const int chi = kc + span < fm - 1 ? kc + span : fm - 1;
            mask_or_range(m, rowbase + kc, rowbase + chi < qend ? rowbase + chi : qend, q0);
          } else {
            for (int bb = 0; bb < g.ksize; ++bb) {
              const int cc = kc + bb * g.dil;
              if (cc >= fm) break;
              if (rowbase + cc <= qend) mask_or_bit(m, rowbase + cc, q0);
            }
          }
        }
      }
    }
  }
  return m;
}

// inverse of the interleaved-pair rotary rotation (adjoint of epi_qkv_pair): given the gradient (g0,g1) of the
// rotated pair, returns the gradient of the unrotated pair.
__device__ __forceinline__ void rotary_adjoint(float c, float s, float& g0, float& g1) {
  const float a = g0 * c + g1 * s;
  const float b = g1 * c - g0 * s;
  g0 = a; g1 = b;
}

}  // namespace db200
