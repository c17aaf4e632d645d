// Attention sparsity patterns of the reference expressed as an allowed(query, key) predicate plus a
// conservative tile-skip test, shared by every attention kernel (SIMT fp32 and tensor-core bf16).
#pragma once
#include "common.cuh"

namespace db200 {

struct AttnGeom {
  int pattern, causal, text_len, fmap, ksize, dil;
  int n_q, n_k;                       // queries are the last n_q of the n_k key positions
  const uint8_t* static_mask; long long static_ld;
};

inline AttnGeom make_geom(const db200_attn_fwd_params& p) {
  AttnGeom g;
  g.pattern = p.pattern; g.causal = p.causal; g.text_len = p.text_len; g.fmap = p.fmap > 0 ? p.fmap : 1;
  g.ksize = p.kernel_size; g.dil = p.dilation > 0 ? p.dilation : 1;
  g.n_q = p.n_q; g.n_k = p.n_k; g.static_mask = p.static_mask; g.static_ld = p.static_ld;
  return g;
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

// inverse of the interleaved-pair rotary rotation (adjoint of epi_qkv_pair): given the gradient (g0,g1) of the
// rotated pair, returns the gradient of the unrotated pair.
__device__ __forceinline__ void rotary_adjoint(float c, float s, float& g0, float& g1) {
  const float a = g0 * c + g1 * s;
  const float b = g1 * c - g0 * s;
  g0 = a; g1 = b;
}

}  // namespace db200
