// Fused attention on the 5th-generation tensor cores (tcgen05.mma, accumulators and P in tensor memory), for every
// sparsity pattern of the reference.  bf16 operands, fp32 accumulation / softmax statistics.
//
// One CTA = 128 queries of one (batch, head); 64-key tiles stream through a 2-stage TMA ring.
//   warp 0     : TMA producer (Q once; K/V tiles)
//   warp 1     : MMA issuer   (converged warp + elect.sync, bursts software-pipelined: S of tile i+1 goes out with O += P V of tile i)
//                             S = Q K^T          (SS: both operands in 128B-swizzled smem, D in TMEM cols [0,64))
//                             O += P V           (TS: A = P read from TMEM, B = V tile read N-major from the same smem image)
//   warps 2..5 : softmax      thread = query row (tcgen05.ld 32x32b): row max / sum need no shuffles; P is written back to
//                             TMEM as packed bf16 (tcgen05.st) and O is rescaled in TMEM only when the running max moved by
//                             more than 2^8 (lazy rescale); final O / l and the log-sum-exp are written by the same threads.
// Three CTAs fit per SM (48 KB smem, 128 TMEM columns each), so one CTA's softmax overlaps the others' MMAs.
//
// Axial row / column attention runs the same kernels in "gathered" mode (attn_common.cuh: virtual axis-major order, tiles cut per
// segment): an image query tile is four whole lines of the 32 x 32 grid and needs the text key tiles plus its own 128 keys --
// 7 key tiles instead of the ~12 a causal sweep over the token sequence touches for the column pattern -- and for axis 1 the Q / K /
// V / dO tiles are fetched as strided 4-D / 5-D TMA boxes ({dh, 32 image rows at a pitch of 32 tokens, 4 columns, (b,h)}).
//
// Backward: delta = rowsum(dO*O), then a dK/dV kernel (CTA = 128 keys, loops over query tiles, works on the transposed
// score tile so that thread = key row) and a dQ kernel (CTA = 128 queries, loops over key tiles); both recompute P from
// the saved log-sum-exp, keep P / dS in TMEM as the A operand of the second-stage MMAs, and fold the rotary adjoint and
// the q scale into their final stores (which write the [rows, 3*h*64] gradient of the to_qkv output directly).  The backward
// kernels run eight softmax warps (two per TMEM lane quarter, 32 score columns each) on 128 x 64 tiles, two CTAs per SM.
#include <cstdlib>
#include <cstring>

#include "attn_common.cuh"
#include "tc_common.cuh"

namespace db200 {

using namespace tc;

namespace {

constexpr int TQ = 128, TK = 128, DH = 64;
constexpr int TILE_BYTES = 128 * DH * 2;     // one [128 x 64] bf16 tile = 16 KB
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_TAU = 8.0f / LOG2E;  // lazy rescale threshold (raw score units): exponentials stay <= 2^8
using bf16 = __nv_bfloat16;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// one full 32-byte sector per lane and store instruction (sm_100 256-bit global stores): the 16-byte stores of the output rows left
// half of every sector to a second instruction (see epilogue.cuh::st_global_v8)
__device__ __forceinline__ void st_global_32B(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z),
               "r"(b.w)
               : "memory");
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sel_bit(uint32_t bits, int i, float a, float b) { return (bits >> i) & 1u ? a : b; }

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}


// The list of tiles a CTA streams (those the pattern needs), built ONCE by one warp into shared memory before the role split:
// every role then walks the same compact list instead of re-evaluating the pattern predicates (integer divisions, segment
// arithmetic) in every thread for every tile -- ncu counted ~400 of the ~600 warp-instructions per tile iteration as such
// bookkeeping.  Entry: bits [0,16) first position of the tile, [16,24) number of valid positions - 1, bit 30 = image tile of the
// column gather (strided TMA box), bit 31 = every (query, key) pair of the tile is allowed (no mask needed).
constexpr int TL_CAP = 256;                   // entries (n <= 8192 with 32-wide tiles)
__device__ __forceinline__ int tl_pos(uint32_t e) { return static_cast<int>(e & 0xffffu); }
__device__ __forceinline__ int tl_width(uint32_t e) { return static_cast<int>((e >> 16) & 0xffu) + 1; }
__device__ __forceinline__ bool tl_img(uint32_t e) { return (e >> 30) & 1u; }
__device__ __forceinline__ bool tl_full(uint32_t e) { return (e >> 31) != 0u; }
// all 32 lanes of one warp; returns the number of entries (also stored in list[TL_CAP])
template <class Needed, class Full>
__device__ __forceinline__ int build_tile_list(uint32_t* list, int lane, const SegTiles& S, bool col, Needed needed, Full full) {
  int cnt = 0;
  const int nt = S.count();
  for (int base = 0; base < nt; base += 32) {
    const int t = base + lane;
    const bool ok = t < nt && needed(t);
    const uint32_t bal = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
      if (pos < TL_CAP) {
        const int o = S.origin(t), w = S.limit(t) - o;
        list[pos] = static_cast<uint32_t>(o) | (static_cast<uint32_t>(w - 1) << 16) | ((col && S.is_img(t)) ? (1u << 30) : 0u) |
                    (full(t) ? (1u << 31) : 0u);
      }
    }
    cnt += __popc(bal);
  }
  if (cnt > TL_CAP) cnt = TL_CAP;             // (launchers refuse shapes that could exceed the capacity)
  if (lane == 0) list[TL_CAP] = static_cast<uint32_t>(cnt);
  return cnt;
}

struct FwdArgs {
  bf16* out; float* lse; const uint8_t* key_mask; int heads, batch;
};

// Forward tiles are 128 queries x FK = 64 keys: S (fp32) needs 64 TMEM columns and P (bf16) is written over the first 32 of
// them, O needs 64 -> 128 columns and 48 KB of smem per CTA, so up to FOUR CTAs are resident per SM and their softmax phases
// overlap each other's MMAs (the per-tile MMA -> softmax -> MMA chain is latency-bound, not throughput-bound).
constexpr int FK = 64;
constexpr int FK_BYTES = FK * DH * 2;          // one [64 x 64] bf16 tile = 8 KB

// P = exp2(S*log2e - m*log2e) for one 32-column chunk held in registers, packed to bf16x2; returns the chunk's row sum
__device__ __forceinline__ float fwd_exp_pack(const uint32_t (&r)[32], uint32_t mb, bool skip, float mb2, uint32_t (&pk)[16]) {
  float2 rs2 = make_float2(0.f, 0.f), rs3 = make_float2(0.f, 0.f);     // two independent chains (a packed add has 4 cycles of latency)
  const float2 kLog2e = make_float2(LOG2E, LOG2E), nm = make_float2(-mb2, -mb2);
  if (skip) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
  } else if (mb == 0xffffffffu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 t = fma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), kLog2e, nm);
      const float2 p = make_float2(ex2(t.x), ex2(t.y));
      if (i & 1) rs3 = add2(rs3, p); else rs2 = add2(rs2, p);
      pk[i] = pack2(p.x, p.y);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 t = fma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), kLog2e, nm);
      const float2 p = make_float2(sel_bit(mb, 2 * i, ex2(t.x), 0.f), sel_bit(mb, 2 * i + 1, ex2(t.y), 0.f));
      if (i & 1) rs3 = add2(rs3, p); else rs2 = add2(rs2, p);
      pk[i] = pack2(p.x, p.y);
    }
  }
  rs2 = add2(rs2, rs3);
  return rs2.x + rs2.y;
}
template <bool P_TMEM>
__device__ __forceinline__ void fwd_store_p(const uint32_t (&pk)[16], int c, uint32_t tP_lane, uint8_t* sP, int row) {
  if constexpr (P_TMEM) {
    tmem_st16(tP_lane + c * 16, pk);
  } else {   // K-major [128 rows][64 keys] 128B-swizzled tile; this thread owns row `row`
    uint8_t* prow = sP + row * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int chunk = c * 4 + j;
      *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
  }
}

// smem map (after 1 KB alignment): Q | K0 | K1 | V0 | V1 | [P (SS mode only, 16 KB)] | barriers
template <bool P_TMEM>
struct FwdSmem {
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = TILE_BYTES;
  static constexpr int V_OFF = TILE_BYTES + 2 * FK_BYTES;
  static constexpr int P_OFF = TILE_BYTES + 4 * FK_BYTES;
  static constexpr int BAR_OFF = P_OFF + (P_TMEM ? 0 : TILE_BYTES);
  static constexpr int LIST_OFF = BAR_OFF + 128;
  static constexpr int TOTAL = LIST_OFF + (TL_CAP + 4) * 4 + 1024;
};

template <bool P_TMEM>
__global__ void __launch_bounds__(192, 3) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQg,
                                                          const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg,
                                                          FwdArgs P, AttnGeom g) {
  using L = FwdSmem<P_TMEM>;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem + L::Q_OFF);
  const uint32_t sK = smem_u32(smem + L::K_OFF), sV = smem_u32(smem + L::V_OFF), sP = smem_u32(smem + L::P_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  const uint32_t q_full = smem_u32(bars), kv_full = smem_u32(bars + 1), kv_empty = smem_u32(bars + 3);
  const uint32_t s_full = smem_u32(bars + 5), p_ready = smem_u32(bars + 6), o_done = smem_u32(bars + 7);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  constexpr uint32_t TMEM_COLS = 128;          // S [0,64) with P (bf16) aliased on [0,32) | O [64,128)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  // heaviest query tiles first: with a causal pattern the last tile of a head does the most key tiles; scheduling it first
  // keeps the final partial wave short
  const SegTiles SQ(g, TQ, g.n_q), SK(g, FK, g.n_k);
  const int qt = gridDim.x - 1 - blockIdx.x;
  const int q0 = SQ.origin(qt), q_lim = SQ.limit(qt);       // valid queries of this tile: [q0, q_lim)
  const int off = g.n_k - g.n_q;
  const int q_last = q_lim - 1;
  const int rows_q = g.gather ? g.n_alloc : g.n_q, rows_k = g.gather ? g.n_alloc : g.kv_rows;   // rows per (b,h) of q / k,v
  uint32_t* tlist = reinterpret_cast<uint32_t*>(smem + L::LIST_OFF);
  if (warp == 2) {                              // key tiles of this query tile, once per CTA (geometry only: before pdl_wait)
    const bool no_km = P.key_mask == nullptr;
    build_tile_list(tlist, lane, SK, g.col != 0,
                    [&](int kt) { return attn_tile_needed(g, q0 + off, q_last + off, SK.origin(kt), SK.limit(kt) - 1); },
                    [&](int kt) { return no_km && SK.limit(kt) - SK.origin(kt) == FK &&
                                         attn_tile_full(g, q0 + off, q_last + off, SK.origin(kt), SK.limit(kt) - 1); });
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    if (g.col) { tma_prefetch_desc(&tmQg); tma_prefetch_desc(&tmKg); tma_prefetch_desc(&tmVg); }
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full + 8 * s, 1); mbar_init(kv_empty + 8 * s, 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(o_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_its = static_cast<int>(tlist[TL_CAP]);
  pdl_wait();                                   // set-up above overlapped the previous kernel's tail
  const uint32_t tS = tmem, tO = tmem + 64, tP = tmem;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE_BYTES);
      if (g.col && SQ.is_img(qt)) tma_load_4d(sQ, &tmQg, q_full, 0, 0, (q0 - g.text_len) / g.fmap, bh);   // 4 image columns, all rows
      else tma_load_2d(sQ, &tmQ, q_full, 0, bh * rows_q + q0);
      for (int it = 0; it < n_its; ++it) {
        const int s = it & 1;
        const uint32_t te = tlist[it];
        const int k0 = tl_pos(te);
        mbar_wait(kv_empty + 8 * s, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full + 8 * s, 2 * FK_BYTES);
        if (tl_img(te)) {
          const int c0 = (k0 - g.text_len) / g.fmap;
          tma_load_4d(sK + s * FK_BYTES, &tmKg, kv_full + 8 * s, 0, 0, c0, bh);
          tma_load_4d(sV + s * FK_BYTES, &tmVg, kv_full + 8 * s, 0, 0, c0, bh);
        } else {
          tma_load_2d(sK + s * FK_BYTES, &tmK, kv_full + 8 * s, 0, bh * rows_k + k0);
          tma_load_2d(sV + s * FK_BYTES, &tmV, kv_full + 8 * s, 0, bh * rows_k + k0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // (whole warp converged; one elected lane issues each burst of tcgen05 instructions, see elect_one().  Software-pipelined:
    //  S = Q K^T of tile i+1 is issued in the same burst as O += P V of tile i.)
    {
      constexpr uint32_t IDESC_S = make_idesc_bf16(128, FK, false, false);    // S = Q K^T : both K-major
      constexpr uint32_t IDESC_O = make_idesc_bf16(128, 64, false, true);     // O = P V   : A K-major (TMEM / smem), B N-major
      mbar_wait(q_full, 0);
      const uint64_t dq = make_smem_desc(sQ, 16, 1024);
      const uint64_t dp = make_smem_desc(sP, 16, 1024);
      if (n_its > 0) {
        mbar_wait(kv_full, 0);
        tc_fence_after();
        const uint64_t dk = make_smem_desc(sK, 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tS, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0);
          umma_commit(s_full);
        }
        __syncwarp();
      }
      for (int it = 0; it < n_its; ++it) {
        const int s = it & 1, sn = s ^ 1;
        const bool more = it + 1 < n_its;
        if (more) { mbar_wait(kv_full + 8 * sn, ((it + 1) >> 1) & 1); }
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        // V tile image [64 keys][64 dh] read as the N-major B operand: K = keys (16 rows = 2048 B per step), N = dh
        const uint64_t dv = make_smem_desc(sV + s * FK_BYTES, FK_BYTES, 1024);
        // (S of the next tile overwrites the columns P of this tile is read from: safe, the tensor pipe executes this
        //  O += P V before the next S = Q K^T because both are issued in order by the same thread)
        const uint64_t dk = make_smem_desc(sK + sn * FK_BYTES, 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < FK / 16; ++k) {
            if constexpr (P_TMEM) umma_bf16_ts(tO, tP + 8 * k, dv + 128 * k, IDESC_O, (it | k) != 0);
            else umma_bf16(tO, dp + 2 * k, dv + 128 * k, IDESC_O, (it | k) != 0);
          }
          umma_commit(kv_empty + 8 * s);
          umma_commit(o_done);
          if (more) {
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16(tS, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0);
            umma_commit(s_full);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== softmax / epilogue =====================================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int qi = q0 + row;                       // query index within [0, n_q)
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * g.n_k : nullptr;
    float m_run = -1.0e30f, l_run = 0.f;
    for (int it = 0; it < n_its; ++it) {
      const uint32_t te = tlist[it];
      const int k0 = tl_pos(te), k1 = k0 + tl_width(te) - 1;
      const bool full = tl_full(te);
      Mask128 mk = {{0xffffffffu, 0xffffffffu, 0u, 0u}};
      if (!full) {
        if (qi < q_lim) mk = attn_row_bits(g, qi + off, k0, km, k1 - k0 + 1);
        else mk.w[0] = mk.w[1] = 0u;
      }
      mbar_wait(s_full, it & 1);                   // also implies the previous tile's O += P V has completed (in-order pipe)
      tc_fence_after();
      // ---- pass 1: row max (32 columns at a time keeps the register footprint small enough for 3-4 CTAs per SM) ----
      const bool skip0 = __all_sync(0xffffffffu, mk.w[0] == 0u), skip1 = __all_sync(0xffffffffu, mk.w[1] == 0u);
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t mb = c == 0 ? mk.w[0] : mk.w[1];
        if (c == 0 ? skip0 : skip1) continue;
        uint32_t r[32];
        tmem_ld32(tS + lane_off + c * 32, r);
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // four independent chains instead of one of 32
        if (mb == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], sel_bit(mb, i, __uint_as_float(r[i]), -INFINITY));
        }
        mx = fmaxf(fmaxf(mx, m4[0]), fmaxf(fmaxf(m4[1], m4[2]), m4[3]));
      }
      const float m_new = (mx > m_run + RESCALE_TAU) ? mx : m_run;
      const float corr = ex2((m_run - m_new) * LOG2E);
      const float mb2 = m_new * LOG2E;
      // ---- pass 2: P = exp2(S*log2e - m*log2e) -> bf16 over the S columns already consumed ----
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t mb = c == 0 ? mk.w[0] : mk.w[1];
        const bool skip = c == 0 ? skip0 : skip1;
        uint32_t r[32], pk[16];
        if (!skip) { tmem_ld32(tS + lane_off + c * 32, r); tmem_ld_wait(); }
        rs += fwd_exp_pack(r, mb, skip, mb2, pk);
        fwd_store_p<P_TMEM>(pk, c, tP + lane_off, smem + L::P_OFF, row);
      }
      l_run = l_run * corr + rs;
      m_run = m_new;
      // ---- lazy rescale of O (warp-uniform decision: tcgen05.ld/st are warp-collective) ----
      if (it > 0 && __any_sync(0xffffffffu, corr != 1.0f)) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(tO + lane_off + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
          tmem_st32(tO + lane_off + c * 32, r);
        }
      }
      if constexpr (P_TMEM) tmem_st_wait(); else { tmem_st_wait(); fence_proxy_async(); }
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // ---- epilogue: O / l -> out[b, qi, h*64 ..] (bf16), lse ----
    const int inner = P.heads * DH;
    const int it = n_its;
    if (it > 0) {
      mbar_wait(o_done, (it - 1) & 1);
      tc_fence_after();
    }
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      if (it > 0) { tmem_ld32(tO + lane_off + c * 32, r); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (qi < q_lim) {
        bf16* orow = P.out + ((long long)b * g.n_q + attn_nat(g, qi)) * inner + h * DH + c * 32;
#pragma unroll
        for (int j2 = 0; j2 < 2; ++j2) {
          uint4 u[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int j = 2 * j2 + t;
            u[t].x = pack2(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
            u[t].y = pack2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
            u[t].z = pack2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
            u[t].w = pack2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
          }
          st_global_32B(orow + 16 * j2, u[0], u[1]);
        }
      }
    }
    if (qi < q_lim) P.lse[(long long)bh * g.n_stat + attn_sidx(g, qi)] = l_run > 0.f ? m_run + logf(l_run) : 0.f;
    else if (g.gather) {
      // padding entries of the statistics array (text tail up to t_pad, the image token that does not exist in training) are
      // bulk-copied by the backward kernels together with real ones: keep them finite
      if (SQ.is_img(qt)) P.lse[(long long)bh * g.n_stat + g.t_pad + (qi - g.text_len)] = 0.f;
      else if (qi < g.t_pad) P.lse[(long long)bh * g.n_stat + qi] = 0.f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Forward, version 2: S double-buffered in tensor memory.
//
// In the kernel above S(i+1) = Q K(i+1)^T can only be issued after softmax(i) has written P(i) (it shares the burst with
// O += P(i) V(i)), so every iteration of a CTA is the serial chain  S-MMA -> commit -> softmax -> arrive -> PV-MMA: ncu's
// warp-state sampling shows the softmax warps asleep on s_full for most of their life (profiles/r02_attn_*).  Here two S
// buffers let the MMA warp run one tile AHEAD: S(i+2) is issued together with O += P(i) V(i), S(i+1) has long completed when
// softmax(i) ends, and the softmax warps go from tile to tile without waiting; the PV MMAs run in the shadow of the next
// softmax.  P(i) is still written over the first half of the S buffer it was computed from (the in-order tensor pipe executes
// O += P(i) V(i) before S(i+2) overwrites those columns).  K and V have separate 3-stage rings: a K stage is free as soon as
// its S MMA has executed, a V stage only after its PV MMA.  One tmem read per score (the row's 64 scores stay in registers
// between the max and the exponentials).  192 TMEM columns (256 allocated) and 64 KB of smem -> two CTAs per SM.
// ---------------------------------------------------------------------------------------------------------------
constexpr int F2_STAGES = 3;
template <int FKT>
struct Fwd2Smem {
  static constexpr int KV_BYTES = FKT * DH * 2;
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = TILE_BYTES;
  static constexpr int V_OFF = K_OFF + F2_STAGES * KV_BYTES;
  static constexpr int BAR_OFF = V_OFF + F2_STAGES * KV_BYTES;
  static constexpr int LIST_OFF = BAR_OFF + 256;
  static constexpr int TOTAL = LIST_OFF + (TL_CAP + 4) * 4 + 1024;
};

// FKT = keys per tile: 64 (192 TMEM columns, two CTAs per SM) or 32 (128 columns, four CTAs per SM = 16 softmax warps that
// never wait for an MMA; the per-tile hand-offs double but they are off the critical path)
template <int FKT>
__global__ void __launch_bounds__(192, FKT == 32 ? 4 : 2) attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQg,
                                                           const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg,
                                                           FwdArgs P, AttnGeom g) {
  using L = Fwd2Smem<FKT>;
  constexpr int FK_BYTES = L::KV_BYTES;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem + L::Q_OFF), sK = smem_u32(smem + L::K_OFF), sV = smem_u32(smem + L::V_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  // barriers: q_full | k_full[3] | k_empty[3] | v_full[3] | v_empty[3] | s_full[2] | p_ready[2] | o_done[2]
  // (o_done alternates between two barriers: the softmax warps run up to two tiles ahead of the O += P V MMAs, and a parity
  //  wait can only tell adjacent phases apart -- with one barrier per tile parity the waiter is never more than one phase off)
  const uint32_t q_full = smem_u32(bars), k_full = smem_u32(bars + 1), k_empty = smem_u32(bars + 4), v_full = smem_u32(bars + 7);
  const uint32_t v_empty = smem_u32(bars + 10), s_full = smem_u32(bars + 13), p_ready = smem_u32(bars + 15), o_done = smem_u32(bars + 17);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  constexpr uint32_t TMEM_COLS = FKT == 32 ? 128 : 256;   // S0 [0,FKT) | S1 [FKT,2 FKT) (P bf16 aliased on the first half of each) | O [2 FKT, 2 FKT + 64)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  const SegTiles SQ(g, TQ, g.n_q), SK(g, FKT, g.n_k);
  const int qt = gridDim.x - 1 - blockIdx.x;               // heaviest (causal) query tiles first
  const int q0 = SQ.origin(qt), q_lim = SQ.limit(qt);
  const int off = g.n_k - g.n_q;
  const int q_last = q_lim - 1;
  const int rows_q = g.gather ? g.n_alloc : g.n_q, rows_k = g.gather ? g.n_alloc : g.kv_rows;
  uint32_t* tlist = reinterpret_cast<uint32_t*>(smem + L::LIST_OFF);
  if (warp == 2) {                              // key tiles of this query tile, once per CTA (geometry only: before pdl_wait)
    const bool no_km = P.key_mask == nullptr;
    build_tile_list(tlist, lane, SK, g.col != 0,
                    [&](int kt) { return attn_tile_needed(g, q0 + off, q_last + off, SK.origin(kt), SK.limit(kt) - 1); },
                    [&](int kt) { return no_km && SK.limit(kt) - SK.origin(kt) == FKT &&
                                         attn_tile_full(g, q0 + off, q_last + off, SK.origin(kt), SK.limit(kt) - 1); });
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    if (g.col) { tma_prefetch_desc(&tmQg); tma_prefetch_desc(&tmKg); tma_prefetch_desc(&tmVg); }
    mbar_init(q_full, 1);
    for (int s = 0; s < F2_STAGES; ++s) {
      mbar_init(k_full + 8 * s, 1); mbar_init(k_empty + 8 * s, 1); mbar_init(v_full + 8 * s, 1); mbar_init(v_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(s_full + 8 * s, 1); mbar_init(p_ready + 8 * s, 128); }
    mbar_init(o_done, 1); mbar_init(o_done + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_its = static_cast<int>(tlist[TL_CAP]);
  pdl_wait();
  const uint32_t tO = tmem + 2 * FKT;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE_BYTES);
      if (g.col && SQ.is_img(qt)) tma_load_4d(sQ, &tmQg, q_full, 0, 0, (q0 - g.text_len) / g.fmap, bh);
      else tma_load_2d(sQ, &tmQ, q_full, 0, bh * rows_q + q0);
      for (int it = 0; it < n_its; ++it) {
        const int s = it % F2_STAGES;
        const uint32_t ph = ((it / F2_STAGES) & 1) ^ 1;
        const uint32_t te = tlist[it];
        const int k0 = tl_pos(te);
        const bool gat = tl_img(te);
        const int c0 = gat ? (k0 - g.text_len) / g.fmap : 0, r0 = gat ? (k0 - g.text_len) - c0 * g.fmap : 0;   // (r0 != 0 only if FKT < fmap)
        mbar_wait(k_empty + 8 * s, ph);
        mbar_expect_tx(k_full + 8 * s, FK_BYTES);
        if (gat) tma_load_4d(sK + s * FK_BYTES, &tmKg, k_full + 8 * s, 0, r0, c0, bh);
        else tma_load_2d(sK + s * FK_BYTES, &tmK, k_full + 8 * s, 0, bh * rows_k + k0);
        mbar_wait(v_empty + 8 * s, ph);
        mbar_expect_tx(v_full + 8 * s, FK_BYTES);
        if (gat) tma_load_4d(sV + s * FK_BYTES, &tmVg, v_full + 8 * s, 0, r0, c0, bh);
        else tma_load_2d(sV + s * FK_BYTES, &tmV, v_full + 8 * s, 0, bh * rows_k + k0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (converged warp, elected lane) =====================================
    constexpr uint32_t IDESC_S = make_idesc_bf16(128, FKT, false, false);
    constexpr uint32_t IDESC_O = make_idesc_bf16(128, 64, false, true);
    mbar_wait(q_full, 0);
    const uint64_t dq = make_smem_desc(sQ, 16, 1024);
    auto issue_s = [&](int j) {           // S(j) = Q K(j)^T into buffer j & 1; frees the K stage when it has executed
      const int s = j % F2_STAGES;
      mbar_wait(k_full + 8 * s, (j / F2_STAGES) & 1);
      tc_fence_after();
      const uint64_t dk = make_smem_desc(sK + s * FK_BYTES, 16, 1024);
      const uint32_t tS = tmem + FKT * (j & 1);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tS, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0);
        umma_commit(s_full + 8 * (j & 1));
        umma_commit(k_empty + 8 * s);
      }
      __syncwarp();
    };
    if (n_its > 0) issue_s(0);
    if (n_its > 1) issue_s(1);
    for (int i = 0; i < n_its; ++i) {
      const int s = i % F2_STAGES;
      mbar_wait(v_full + 8 * s, (i / F2_STAGES) & 1);
      mbar_wait(p_ready + 8 * (i & 1), (i >> 1) & 1);
      tc_fence_after();
      const uint64_t dv = make_smem_desc(sV + s * FK_BYTES, FK_BYTES, 1024);     // V tile read N-major: K = keys, N = dh
      const uint32_t tP = tmem + FKT * (i & 1);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FKT / 16; ++k) umma_bf16_ts(tO, tP + 8 * k, dv + 128 * k, IDESC_O, (i | k) != 0);
        umma_commit(v_empty + 8 * s);
        umma_commit(o_done + 8 * (i & 1));
      }
      __syncwarp();
      if (i + 2 < n_its) issue_s(i + 2);
    }
  } else {
    // ===================================== softmax / epilogue =====================================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * g.n_k : nullptr;
    float m_run = -1.0e30f, l_run = 0.f;
    for (int it = 0; it < n_its; ++it) {
      const uint32_t te = tlist[it];
      const int k0 = tl_pos(te), k1 = k0 + tl_width(te) - 1;
      const bool full = tl_full(te);
      Mask128 mk = {{0xffffffffu, FKT == 64 ? 0xffffffffu : 0u, 0u, 0u}};
      if (!full) {
        if (qi < q_lim) mk = attn_row_bits(g, qi + off, k0, km, k1 - k0 + 1);
        else mk.w[0] = mk.w[1] = 0u;
      }
      const uint32_t tS = tmem + FKT * (it & 1) + lane_off;
      mbar_wait(s_full + 8 * (it & 1), (it >> 1) & 1);
      tc_fence_after();
      const bool skip0 = __all_sync(0xffffffffu, mk.w[0] == 0u), skip1 = FKT == 32 || __all_sync(0xffffffffu, mk.w[1] == 0u);
      uint32_t r0[32], r1[FKT == 64 ? 32 : 1];
      if (!skip0) tmem_ld32(tS, r0);
      if constexpr (FKT == 64) { if (!skip1) tmem_ld32(tS + 32, r1); }
      tmem_ld_wait();
      float mx = -INFINITY;
      if (!skip0) {
        if (mk.w[0] == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r0[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, sel_bit(mk.w[0], i, __uint_as_float(r0[i]), -INFINITY));
        }
      }
      if constexpr (FKT == 64) if (!skip1) {
        if (mk.w[1] == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r1[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, sel_bit(mk.w[1], i, __uint_as_float(r1[i]), -INFINITY));
        }
      }
      const float m_new = (mx > m_run + RESCALE_TAU) ? mx : m_run;
      const float corr = ex2((m_run - m_new) * LOG2E);
      const float mb2 = m_new * LOG2E;
      uint32_t pk[16];
      float rs = fwd_exp_pack(r0, mk.w[0], skip0, mb2, pk);
      tmem_st16(tS, pk);
      if constexpr (FKT == 64) {
        rs += fwd_exp_pack(r1, mk.w[1], skip1, mb2, pk);
        tmem_st16(tS + 16, pk);
      }
      l_run = l_run * corr + rs;
      m_run = m_new;
      // ---- lazy rescale of O: needs every earlier O += P V to have executed (they run in the shadow of this softmax) ----
      if (it > 0 && __any_sync(0xffffffffu, corr != 1.0f)) {
        mbar_wait(o_done + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(tO + lane_off + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
          tmem_st32(tO + lane_off + c * 32, r);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready + 8 * (it & 1));
    }
    // ---- epilogue: O / l -> out[b, token, h*64 ..] (bf16), lse ----
    const int inner = P.heads * DH;
    const int it = n_its;
    if (it > 0) {
      mbar_wait(o_done + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
      tc_fence_after();
    }
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      if (it > 0) { tmem_ld32(tO + lane_off + c * 32, r); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (qi < q_lim) {
        bf16* orow = P.out + ((long long)b * g.n_q + attn_nat(g, qi)) * inner + h * DH + c * 32;
#pragma unroll
        for (int j2 = 0; j2 < 2; ++j2) {
          uint4 u[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int j = 2 * j2 + t;
            u[t].x = pack2(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
            u[t].y = pack2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
            u[t].z = pack2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
            u[t].w = pack2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
          }
          st_global_32B(orow + 16 * j2, u[0], u[1]);
        }
      }
    }
    if (qi < q_lim) P.lse[(long long)bh * g.n_stat + attn_sidx(g, qi)] = l_run > 0.f ? m_run + logf(l_run) : 0.f;
    else if (g.gather) {
      if (SQ.is_img(qt)) P.lse[(long long)bh * g.n_stat + g.t_pad + (qi - g.text_len)] = 0.f;
      else if (qi < g.t_pad) P.lse[(long long)bh * g.n_stat + qi] = 0.f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// =============================================== backward ====================================================
// delta[b,h,i] = sum_d O[b,i,h,d] * dO[b,i,h,d].  Rows of 64 bf16 (128 B) are contiguous in (b,i,h) order: 8 lanes per row,
// one 16-byte load of O and dO per lane, 4 rows per warp per step, 4 steps in flight.
// Gathered mode: delta is stored like lse ([b*h, n_stat], virtual order); the thread of token 0 of every (b,h) also zeroes the
// padding entries, and block 0 zeroes the row after the last token of dO that the strided 5-D box of the last batch reaches.
__global__ void __launch_bounds__(256) attn_delta_tc_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ delta,
                                                            int batch, int heads, int n, AttnGeom g, bf16* dO_pad_row) {
  pdl_launch();
  pdl_wait();
  if (dO_pad_row != nullptr && blockIdx.x == 0) {
    for (int j = threadIdx.x * 8; j < heads * DH; j += 256 * 8) *reinterpret_cast<uint4*>(dO_pad_row + j) = make_uint4(0, 0, 0, 0);
  }
  const long long total = (long long)batch * heads * n;
  const int part = threadIdx.x & 7;
  const long long stride = (long long)gridDim.x * 32 * 4;            // rows per grid step (32 rows per 256-thread block, x4 unroll)
  // (loop bound is warp-uniform so that the full-mask shuffles below are always executed by all 32 lanes)
  for (long long wb = (long long)blockIdx.x * 32 + (threadIdx.x >> 5) * 4; wb < total; wb += stride) {
    const long long base = wb + ((threadIdx.x & 31) >> 3);
    uint4 o[4], d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = base + (long long)u * gridDim.x * 32;
      o[u] = make_uint4(0, 0, 0, 0); d[u] = make_uint4(0, 0, 0, 0);
      if (r < total) {
        o[u] = *reinterpret_cast<const uint4*>(O + r * DH + part * 8);
        d[u] = *reinterpret_cast<const uint4*>(dO + r * DH + part * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = base + (long long)u * gridDim.x * 32;
      const __nv_bfloat162* po = reinterpret_cast<const __nv_bfloat162*>(&o[u]);
      const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(&d[u]);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = __bfloat1622float2(po[k]), b2 = __bfloat1622float2(pd[k]);
        acc = fmaf(a.x, b2.x, fmaf(a.y, b2.y, acc));
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (part == 0 && r < total) {
        const long long bi = r / heads;                              // r = (b*n + i)*heads + h
        const int h = static_cast<int>(r - bi * heads);
        const long long b = bi / n;
        const int i = static_cast<int>(bi - b * n);
        float* drow = delta + (b * heads + h) * (long long)g.n_stat;
        drow[attn_sidx(g, attn_virt(g, i))] = acc;
        if (g.gather && i == 0) {
          for (int s = g.text_len; s < g.t_pad; ++s) drow[s] = 0.f;
          if (n < g.text_len + g.fmap * g.fmap) drow[g.n_stat - 1] = 0.f;
        }
      }
    }
  }
}

// Timeline probe (debug only, DALLE_B200_ATTN_WAIT bit 2): CTA (1,0) of the dK/dV kernel records clock64() at the hand-off
// points of its first iterations; read back with dalle_b200_debug_attn_timeline().
__device__ long long g_attn_dbg[256];
#define ATTN_DBG(slot) do { if (dbg && it < 6) g_attn_dbg[(slot) * 6 + it] = clock64(); } while (0)

struct BwdArgs {
  const float* lse; const float* delta; const uint8_t* key_mask; const float* cos_t; const float* sin_t; float q_scale;
  bf16* dqkv; int heads, batch; int wait_mode;
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// 32 fp32 accumulator columns [32c, 32c+32) of this thread's row -> (rotary adjoint, optional scale) -> bf16 -> dst[32c ..]
__device__ __forceinline__ void store_grad_cols(uint32_t taddr, bf16* dst, const float* cos_row, const float* sin_row, float scale, bool valid,
                                                int c) {
  uint32_t r[32];
  tmem_ld32(taddr + c * 32, r);
  tmem_ld_wait();
  if (valid) {
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]) * scale;
      if (cos_row) {
        const int pi = (c * 32 + j * 8) >> 1;                      // pair index of the first pair of this granule
        const float4 cc = *reinterpret_cast<const float4*>(cos_row + pi);
        const float4 ss = *reinterpret_cast<const float4*>(sin_row + pi);
        rotary_adjoint(cc.x, ss.x, v[0], v[1]); rotary_adjoint(cc.y, ss.y, v[2], v[3]);
        rotary_adjoint(cc.z, ss.z, v[4], v[5]); rotary_adjoint(cc.w, ss.w, v[6], v[7]);
      }
      u[j].x = pack2(v[0], v[1]); u[j].y = pack2(v[2], v[3]); u[j].z = pack2(v[4], v[5]); u[j].w = pack2(v[6], v[7]);
    }
    st_global_32B(dst + c * 32, u[0], u[1]);          // (rows of dqkv start on 128-byte boundaries: heads * 64 bf16 per section)
    st_global_32B(dst + c * 32 + 16, u[2], u[3]);
  }
}
__device__ __forceinline__ void store_grad_row(uint32_t taddr, bf16* dst, const float* cos_row, const float* sin_row, float scale, bool valid) {
#pragma unroll 1
  for (int c = 0; c < 2; ++c) store_grad_cols(taddr, dst, cos_row, sin_row, scale, valid, c);
}

// softmax-backward of 16 columns of one row: p = exp(s - lse), ds = p * (dp - delta); packs 8 bf16 pairs each.
// kLseCol: lse/delta vary along the columns (dK/dV kernel, read from smem at ls_addr / dl_addr) instead of being per-row
// constants (dQ kernel; lse_r already carries the log2(e) factor).  Four packed fp32 instructions + two MUFU + two
// conversions per element pair.
template <bool kLseCol, bool kWantP, bool kNoExp = false>      // kNoExp: diagnosis only (DALLE_B200_ATTN_WAIT bit 3): a multiply instead of ex2
__device__ __forceinline__ void bwd_softmax16(const uint32_t* rs, const uint32_t* rd, uint32_t mb16, uint32_t ls_addr, uint32_t dl_addr,
                                              float lse_r, float delta_r, uint32_t* pk, uint32_t* dk_) {
  const float2 kLog2e = make_float2(LOG2E, LOG2E), kNegLog2e = make_float2(-LOG2E, -LOG2E), kNegOne = make_float2(-1.f, -1.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {                    // 4 columns per step
    float2 nl0, nl1, d0, d1;                       // -lse*log2e and delta of columns (4j, 4j+1), (4j+2, 4j+3)
    if constexpr (kLseCol) {
      const float4 a = lds128(ls_addr + 16 * j), b = lds128(dl_addr + 16 * j);
      nl0 = mul2(make_float2(a.x, a.y), kNegLog2e); nl1 = mul2(make_float2(a.z, a.w), kNegLog2e);
      d0 = make_float2(b.x, b.y); d1 = make_float2(b.z, b.w);
    } else {
      nl0 = nl1 = make_float2(-lse_r, -lse_r);
      d0 = d1 = make_float2(delta_r, delta_r);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * j + h;                     // pair index: columns 2i, 2i+1
      const float2 sv = make_float2(__uint_as_float(rs[2 * i]), __uint_as_float(rs[2 * i + 1]));
      const float2 dv = make_float2(__uint_as_float(rd[2 * i]), __uint_as_float(rd[2 * i + 1]));
      const float2 t = fma2(sv, kLog2e, h == 0 ? nl0 : nl1);
      float2 p = kNoExp ? make_float2(t.x * 1e-3f, t.y * 1e-3f) : make_float2(ex2(t.x), ex2(t.y));
      if (mb16 != 0xffffu) { p.x = sel_bit(mb16, 2 * i, p.x, 0.f); p.y = sel_bit(mb16, 2 * i + 1, p.y, 0.f); }
      const float2 e = fma2(h == 0 ? d0 : d1, kNegOne, dv);      // dp - delta
      const float2 ds = mul2(p, e);
      if constexpr (kWantP) pk[i] = pack2(p.x, p.y);
      dk_[i] = pack2(ds.x, ds.y);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward tiles are 128 rows x BW = 64 columns so that each kernel needs only 256 TMEM columns and ~64 KB of smem: TWO CTAs
// run per SM and one CTA's softmax-backward overlaps the other's MMAs.  Eight softmax warps per CTA: warps 2-5 own columns
// [0,32) of the score tile, warps 6-9 columns [32,64) (a warp may only touch the TMEM lane quarter warp%4).  P / dS (bf16)
// are written over the first half of the 32 columns each warp has itself read, i.e. packed columns [32c, 32c+16) for column
// half c, so no warp overwrites scores another warp has not read yet; the TS-mode MMAs take their A operand from there.
// ---------------------------------------------------------------------------------------------------------------
constexpr int BW = 64;
constexpr int HALF_TILE = BW * DH * 2;            // one [64 x 64] bf16 tile = 8 KB
constexpr int BWD_THREADS = 320;                  // TMA warp, MMA warp, 8 softmax warps
__device__ __forceinline__ uint32_t packed_kcol(int k) { return static_cast<uint32_t>(8 * k + (k >= 2 ? 16 : 0)); }   // A-operand column of k-step k

// dK, dV : CTA = 128 keys; transposed score tile (rows = keys, columns = 64 queries) so that thread = key row.
//   TMEM: S^T [0,64) (P^T bf16 aliased on [0,16)+[32,48)) | dP^T [64,128) (dS^T aliased on [64,80)+[96,112)) | dV [128,192) | dK [192,256)
struct DkvSmem {
  static constexpr int K_OFF = 0, V_OFF = TILE_BYTES, Q_OFF = 2 * TILE_BYTES, DO_OFF = 2 * TILE_BYTES + 2 * HALF_TILE;
  static constexpr int STAT_OFF = 2 * TILE_BYTES + 4 * HALF_TILE;   // lse [2][64], delta [2][64] (bulk-copied with the Q / dO tiles)
  static constexpr int BAR_OFF = STAT_OFF + 4 * BW * 4;
  static constexpr int LIST_OFF = BAR_OFF + 128;
  static constexpr int TOTAL = LIST_OFF + (TL_CAP + 4) * 4 + 1024;
};

__global__ void __launch_bounds__(BWD_THREADS, 2) attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                                                               const __grid_constant__ CUtensorMap tmQg, const __grid_constant__ CUtensorMap tmKg,
                                                               const __grid_constant__ CUtensorMap tmVg, const __grid_constant__ CUtensorMap tmdOg,
                                                               BwdArgs P, AttnGeom g) {
  using L = DkvSmem;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t sK = smem_u32(smem + L::K_OFF), sV = smem_u32(smem + L::V_OFF), sQ = smem_u32(smem + L::Q_OFF), sdO = smem_u32(smem + L::DO_OFF);
  float* s_lse = reinterpret_cast<float*>(smem + L::STAT_OFF);      // [2][64]
  float* s_delta = s_lse + 2 * BW;                                   // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  const uint32_t kv_full = smem_u32(bars), q_full = smem_u32(bars + 1), q_empty = smem_u32(bars + 3);
  const uint32_t st_full = smem_u32(bars + 5), ps_ready = smem_u32(bars + 6), acc_done = smem_u32(bars + 7);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  constexpr uint32_t TMEM_COLS = 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  const int n = g.n_k, inner = P.heads * DH;
  const SegTiles SK(g, TK, n), SQ(g, BW, n);
  const int ktile = blockIdx.x;
  const int k0 = SK.origin(ktile), k_lim = SK.limit(ktile), k1 = k_lim - 1;     // valid keys of this CTA: [k0, k_lim)
  const int rows = g.gather ? g.n_alloc : n;                                   // rows per (b,h) of q / k / v
  uint32_t* tlist = reinterpret_cast<uint32_t*>(smem + L::LIST_OFF);
  if (warp == 2) {                              // query tiles this key tile is visible to, once per CTA (before pdl_wait)
    const bool plain = P.key_mask == nullptr && k1 - k0 == TK - 1;
    build_tile_list(tlist, lane, SQ, g.col != 0,
                    [&](int qt) { return attn_tile_needed(g, SQ.origin(qt), SQ.limit(qt) - 1, k0, k1); },
                    [&](int qt) { return plain && SQ.limit(qt) - SQ.origin(qt) == BW &&
                                         attn_tile_full(g, SQ.origin(qt), SQ.limit(qt) - 1, k0, k1); });
  }
  // per-query statistics ride along with the Q / dO tiles as two 1-D bulk copies when the rows are 16-byte aligned
  const bool stats_tma = (g.n_stat & 3) == 0 && ((reinterpret_cast<uintptr_t>(P.lse) | reinterpret_cast<uintptr_t>(P.delta)) & 15) == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    if (g.col) { tma_prefetch_desc(&tmQg); tma_prefetch_desc(&tmKg); tma_prefetch_desc(&tmVg); tma_prefetch_desc(&tmdOg); }
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(q_full + 8 * s, 1); mbar_init(q_empty + 8 * s, 1); }
    mbar_init(st_full, 1); mbar_init(ps_ready, BWD_THREADS - 64); mbar_init(acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64) s_lse[threadIdx.x - 64] = 0.f;      // [2][64] lse + [2][64] delta: stale entries must stay finite
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_its = static_cast<int>(tlist[TL_CAP]);
  pdl_wait();                                   // set-up above overlapped the previous kernel's tail
  const uint32_t tSt = tmem, tdPt = tmem + 64, tdV = tmem + 128, tdK = tmem + 192, tPt = tmem, tdSt = tmem + 64;
  const bool dbg = (P.wait_mode & 4) && blockIdx.x == 1 && blockIdx.y == 0 && (threadIdx.x == 32 || threadIdx.x == 64 || threadIdx.x == 192);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      if (g.col && SK.is_img(ktile)) {
        const int c0 = (k0 - g.text_len) / g.fmap;
        tma_load_4d(sK, &tmKg, kv_full, 0, 0, c0, bh);
        tma_load_4d(sV, &tmVg, kv_full, 0, 0, c0, bh);
      } else {
        tma_load_2d(sK, &tmK, kv_full, 0, bh * rows + k0);
        tma_load_2d(sV, &tmV, kv_full, 0, bh * rows + k0);
      }
      for (int it = 0; it < n_its; ++it) {
        const int s = it & 1;
        const uint32_t te = tlist[it];
        const int qq0 = tl_pos(te);
        mbar_wait(q_empty + 8 * s, ((it >> 1) & 1) ^ 1);
        // (gathered mode: the statistics arrays are padded so that a full 64-entry run exists behind every tile origin)
        const uint32_t stat_bytes = stats_tma ? static_cast<uint32_t>(g.gather ? BW : min(BW, n - qq0)) * 4u : 0u;
        mbar_expect_tx(q_full + 8 * s, 2 * HALF_TILE + 2 * stat_bytes);
        if (tl_img(te)) {
          const int c0 = (qq0 - g.text_len) / g.fmap;
          tma_load_4d(sQ + s * HALF_TILE, &tmQg, q_full + 8 * s, 0, 0, c0, bh);
          tma_load_5d(sdO + s * HALF_TILE, &tmdOg, q_full + 8 * s, 0, 0, c0, h, b);
        } else {
          tma_load_2d(sQ + s * HALF_TILE, &tmQ, q_full + 8 * s, 0, bh * rows + qq0);
          tma_load_2d(sdO + s * HALF_TILE, &tmdO, q_full + 8 * s, h * DH, b * n + qq0);
        }
        if (stats_tma) {
          const long long so = (long long)bh * g.n_stat + attn_sidx(g, qq0);
          bulk_load_1d(smem_u32(s_lse + s * BW), P.lse + so, stat_bytes, q_full + 8 * s);
          bulk_load_1d(smem_u32(s_delta + s * BW), P.delta + so, stat_bytes, q_full + 8 * s);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // whole warp converged; an elected lane issues (see elect_one())
      // Software-pipelined issue: the S^T / dP^T MMAs of tile i+1 go out in the same burst as the dV / dK MMAs of tile i (the
      // in-order tensor pipe keeps the aliasing safe), so nothing but the softmax separates consecutive bursts.
      constexpr uint32_t IDESC_T = make_idesc_bf16(128, BW, false, false);    // S^T = K Q^T, dP^T = V dO^T
      constexpr uint32_t IDESC_G = make_idesc_bf16(128, 64, false, true);     // dV = P^T dO, dK = dS^T Q   (B N-major)
      mbar_wait(kv_full, 0);
      const uint64_t dk = make_smem_desc(sK, 16, 1024), dv = make_smem_desc(sV, 16, 1024);
      if (n_its > 0) {
        mbar_wait(q_full, 0);
        tc_fence_after();
        const uint64_t dq = make_smem_desc(sQ, 16, 1024), ddo = make_smem_desc(sdO, 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tSt, dk + 2 * k, dq + 2 * k, IDESC_T, k != 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tdPt, dv + 2 * k, ddo + 2 * k, IDESC_T, k != 0);
          umma_commit(st_full);
        }
        __syncwarp();
      }
      for (int it = 0; it < n_its; ++it) {
        const int s = it & 1, sn = s ^ 1;
        const bool more = it + 1 < n_its;
        if (more) { mbar_wait(q_full + 8 * sn, ((it + 1) >> 1) & 1); }        // next Q / dO tile (prefetched long ago)
        ATTN_DBG(1);
        mbar_wait_mode(ps_ready, it & 1, (P.wait_mode >> 1) & 1);
        tc_fence_after();
        ATTN_DBG(2);                                         // MMA: P^T / dS^T ready
        const uint64_t bq = make_smem_desc(sQ + s * HALF_TILE, HALF_TILE, 1024), bdo = make_smem_desc(sdO + s * HALF_TILE, HALF_TILE, 1024);
        const uint64_t dq = make_smem_desc(sQ + sn * HALF_TILE, 16, 1024), ddo = make_smem_desc(sdO + sn * HALF_TILE, 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BW / 16; ++k) umma_bf16_ts(tdV, tPt + packed_kcol(k), bdo + 128 * k, IDESC_G, (it | k) != 0);
#pragma unroll
          for (int k = 0; k < BW / 16; ++k) umma_bf16_ts(tdK, tdSt + packed_kcol(k), bq + 128 * k, IDESC_G, (it | k) != 0);
          umma_commit(q_empty + 8 * s);
          umma_commit(acc_done);
          if (more) {
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16(tSt, dk + 2 * k, dq + 2 * k, IDESC_T, k != 0);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16(tdPt, dv + 2 * k, ddo + 2 * k, IDESC_T, k != 0);
            umma_commit(st_full);
          }
        }
        __syncwarp();
        ATTN_DBG(3);                                         // MMA: burst issued
      }
    }
  } else {
    const int quarter = warp & 3, chunk = (warp - 2) >> 2;    // TMEM lane quarter; column half [32*chunk, 32*chunk+32)
    const int row = quarter * 32 + lane;
    const int kj = k0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * n : nullptr;
    const bool key_ok = kj < k_lim && (km == nullptr || km[kj] != 0);
    for (int it = 0; it < n_its; ++it) {
      const int s = it & 1;
      const uint32_t te = tlist[it];
      const int qq0 = tl_pos(te), qq1 = qq0 + tl_width(te) - 1;
      if (stats_tma) {
        mbar_wait(q_full + 8 * s, (it >> 1) & 1);
      } else {
        if (chunk == 0 && row < BW) {   // per-query statistics of this tile -> smem
          const int qi = qq0 + row;
          const long long so = (long long)bh * g.n_stat + attn_sidx(g, qi);
          s_lse[s * BW + row] = qi <= qq1 ? P.lse[so] : 0.f;
          s_delta[s * BW + row] = qi <= qq1 ? P.delta[so] : 0.f;
        }
        named_bar_sync(1, BWD_THREADS - 64);
      }
      const bool full = tl_full(te);
      uint32_t mb = 0xffffffffu;
      if (!full) {
        if (key_ok) { const Mask128 mk = attn_col_bits(g, kj, qq0, n, qq1 - qq0 + 1); mb = chunk == 0 ? mk.w[0] : mk.w[1]; }
        else mb = 0u;
      }
      ATTN_DBG(threadIdx.x == 64 ? 4 : 8);                   // softmax: masks ready, waiting for S^T / dP^T
      mbar_wait_mode(st_full, it & 1, P.wait_mode & 1);
      tc_fence_after();
      ATTN_DBG(threadIdx.x == 64 ? 5 : 9);                   // softmax: S^T / dP^T complete
      // (the MMAs of this tile were issued after the previous tile's dV/dK MMAs, which read P^T / dS^T from the columns
      //  written below: the tensor pipe executes in order, so st_full also means those reads are complete)
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const uint32_t mb16 = (mb >> (16 * sub)) & 0xffffu;
        const int col = chunk * 32 + sub * 16;
        uint32_t pk[8], dk_[8];
        if (__all_sync(0xffffffffu, mb16 == 0u)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { pk[i] = 0u; dk_[i] = 0u; }
        } else {
          uint32_t rs[16], rd[16];
          tmem_ld16(tSt + lane_off + col, rs);
          tmem_ld16(tdPt + lane_off + col, rd);
          tmem_ld_wait();
          if (P.wait_mode & 8) bwd_softmax16<true, true, true>(rs, rd, mb16, smem_u32(s_lse + s * BW + col), smem_u32(s_delta + s * BW + col), 0.f, 0.f, pk, dk_);
          else bwd_softmax16<true, true>(rs, rd, mb16, smem_u32(s_lse + s * BW + col), smem_u32(s_delta + s * BW + col), 0.f, 0.f, pk, dk_);
        }
        tmem_st8(tPt + lane_off + chunk * 32 + sub * 8, pk);
        tmem_st8(tdSt + lane_off + chunk * 32 + sub * 8, dk_);
      }
      ATTN_DBG(threadIdx.x == 64 ? 6 : 10);                  // softmax: both 16-column passes computed, stores issued
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ps_ready);
      ATTN_DBG(threadIdx.x == 64 ? 7 : 11);                  // softmax: arrived
    }
    // warps 2-5 write dK, warps 6-9 write dV
    const uint32_t tacc = chunk == 0 ? tdK : tdV;
    const int sel = chunk == 0 ? 1 : 2;
    const int it = n_its;
    if (it > 0) {
      mbar_wait(acc_done, (it - 1) & 1);
      tc_fence_after();
      const int pk = kj < k_lim ? attn_nat(g, kj) : 0;                 // token position of this key row
      bf16* rowp = P.dqkv + ((long long)b * n + pk) * (3 * inner) + h * DH;
      const float* cr = P.cos_t ? P.cos_t + (long long)pk * (DH / 2) : nullptr;
      const float* sr = P.sin_t ? P.sin_t + (long long)pk * (DH / 2) : nullptr;
      store_grad_row(tacc + lane_off, rowp + sel * inner, cr, sr, 1.0f, kj < k_lim);
    } else if (kj < k_lim) {
      bf16* rowp = P.dqkv + ((long long)b * n + attn_nat(g, kj)) * (3 * inner) + h * DH;
      for (int d = 0; d < DH; d += 8) *reinterpret_cast<uint4*>(rowp + sel * inner + d) = make_uint4(0, 0, 0, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dQ : CTA = 128 queries, thread = query row, 64-key tiles.
//   TMEM: S [0,64) | dP [64,128) (dS bf16 aliased on [64,80)+[96,112)) | dQ [128,192)
// ---------------------------------------------------------------------------------------------------------------
struct DqSmem {
  static constexpr int Q_OFF = 0, DO_OFF = TILE_BYTES, K_OFF = 2 * TILE_BYTES, V_OFF = 2 * TILE_BYTES + 2 * HALF_TILE;
  static constexpr int BAR_OFF = 2 * TILE_BYTES + 4 * HALF_TILE;
  static constexpr int LIST_OFF = BAR_OFF + 128;
  static constexpr int TOTAL = LIST_OFF + (TL_CAP + 4) * 4 + 1024;
};

__global__ void __launch_bounds__(BWD_THREADS, 2) attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                                                              const __grid_constant__ CUtensorMap tmQg, const __grid_constant__ CUtensorMap tmKg,
                                                              const __grid_constant__ CUtensorMap tmVg, const __grid_constant__ CUtensorMap tmdOg,
                                                              BwdArgs P, AttnGeom g) {
  using L = DqSmem;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem + L::Q_OFF), sdO = smem_u32(smem + L::DO_OFF), sK = smem_u32(smem + L::K_OFF), sV = smem_u32(smem + L::V_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  const uint32_t q_full = smem_u32(bars), kv_full = smem_u32(bars + 1), kv_empty = smem_u32(bars + 3);
  const uint32_t s_full = smem_u32(bars + 5), ds_ready = smem_u32(bars + 6), acc_done = smem_u32(bars + 7);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  constexpr uint32_t TMEM_COLS = 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / P.heads, h = bh - b * P.heads;
  const int n = g.n_k, inner = P.heads * DH;
  const SegTiles SQ(g, TQ, n), SK(g, BW, n);
  const int qtile = gridDim.x - 1 - blockIdx.x;                                // heaviest (causal) tiles first
  const int q0 = SQ.origin(qtile), q_lim = SQ.limit(qtile), q1 = q_lim - 1;     // valid queries of this CTA: [q0, q_lim)
  const int rows = g.gather ? g.n_alloc : n;
  uint32_t* tlist = reinterpret_cast<uint32_t*>(smem + L::LIST_OFF);
  if (warp == 2) {                              // key tiles of this query tile, once per CTA (before pdl_wait)
    const bool no_km = P.key_mask == nullptr;
    build_tile_list(tlist, lane, SK, g.col != 0,
                    [&](int kt) { return attn_tile_needed(g, q0, q1, SK.origin(kt), SK.limit(kt) - 1); },
                    [&](int kt) { return no_km && SK.limit(kt) - SK.origin(kt) == BW &&
                                         attn_tile_full(g, q0, q1, SK.origin(kt), SK.limit(kt) - 1); });
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    if (g.col) { tma_prefetch_desc(&tmQg); tma_prefetch_desc(&tmKg); tma_prefetch_desc(&tmVg); tma_prefetch_desc(&tmdOg); }
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full + 8 * s, 1); mbar_init(kv_empty + 8 * s, 1); }
    mbar_init(s_full, 1); mbar_init(ds_ready, BWD_THREADS - 64); mbar_init(acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_its = static_cast<int>(tlist[TL_CAP]);
  pdl_wait();                                   // set-up above overlapped the previous kernel's tail
  const uint32_t tS = tmem, tdP = tmem + 64, tdS = tmem + 64, tdQ = tmem + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      if (g.col && SQ.is_img(qtile)) {
        const int c0 = (q0 - g.text_len) / g.fmap;
        tma_load_4d(sQ, &tmQg, q_full, 0, 0, c0, bh);
        tma_load_5d(sdO, &tmdOg, q_full, 0, 0, c0, h, b);
      } else {
        tma_load_2d(sQ, &tmQ, q_full, 0, bh * rows + q0);
        tma_load_2d(sdO, &tmdO, q_full, h * DH, b * n + q0);
      }
      for (int it = 0; it < n_its; ++it) {
        const int s = it & 1;
        const uint32_t te = tlist[it];
        const int k0 = tl_pos(te);
        mbar_wait(kv_empty + 8 * s, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full + 8 * s, 2 * HALF_TILE);
        if (tl_img(te)) {
          const int c0 = (k0 - g.text_len) / g.fmap;
          tma_load_4d(sK + s * HALF_TILE, &tmKg, kv_full + 8 * s, 0, 0, c0, bh);
          tma_load_4d(sV + s * HALF_TILE, &tmVg, kv_full + 8 * s, 0, 0, c0, bh);
        } else {
          tma_load_2d(sK + s * HALF_TILE, &tmK, kv_full + 8 * s, 0, bh * rows + k0);
          tma_load_2d(sV + s * HALF_TILE, &tmV, kv_full + 8 * s, 0, bh * rows + k0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // whole warp converged; an elected lane issues (see elect_one()); software-pipelined like the dK/dV kernel
      constexpr uint32_t IDESC_S = make_idesc_bf16(128, BW, false, false);
      constexpr uint32_t IDESC_G = make_idesc_bf16(128, 64, false, true);
      mbar_wait(q_full, 0);
      const uint64_t dq = make_smem_desc(sQ, 16, 1024), ddo = make_smem_desc(sdO, 16, 1024);
      if (n_its > 0) {
        mbar_wait(kv_full, 0);
        tc_fence_after();
        const uint64_t dk = make_smem_desc(sK, 16, 1024), dv = make_smem_desc(sV, 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tS, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tdP, ddo + 2 * k, dv + 2 * k, IDESC_S, k != 0);
          umma_commit(s_full);
        }
        __syncwarp();
      }
      for (int it = 0; it < n_its; ++it) {
        const int s = it & 1, sn = s ^ 1;
        const bool more = it + 1 < n_its;
        if (more) { mbar_wait(kv_full + 8 * sn, ((it + 1) >> 1) & 1); }
        mbar_wait_mode(ds_ready, it & 1, (P.wait_mode >> 1) & 1);
        tc_fence_after();
        const uint64_t bk = make_smem_desc(sK + s * HALF_TILE, HALF_TILE, 1024);
        const uint64_t dk = make_smem_desc(sK + sn * HALF_TILE, 16, 1024), dv = make_smem_desc(sV + sn * HALF_TILE, 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BW / 16; ++k) umma_bf16_ts(tdQ, tdS + packed_kcol(k), bk + 128 * k, IDESC_G, (it | k) != 0);
          umma_commit(kv_empty + 8 * s);
          umma_commit(acc_done);
          if (more) {
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16(tS, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16(tdP, ddo + 2 * k, dv + 2 * k, IDESC_S, k != 0);
            umma_commit(s_full);
          }
        }
        __syncwarp();
      }
    }
  } else {
    const int quarter = warp & 3, chunk = (warp - 2) >> 2;    // TMEM lane quarter; column half [32*chunk, 32*chunk+32)
    const int row = quarter * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint8_t* km = P.key_mask ? P.key_mask + (long long)b * n : nullptr;
    const long long so = (long long)bh * g.n_stat + attn_sidx(g, qi);
    const float lse_r = qi < q_lim ? P.lse[so] * LOG2E : 0.f;
    const float delta_r = qi < q_lim ? P.delta[so] : 0.f;
    for (int it = 0; it < n_its; ++it) {
      const uint32_t te = tlist[it];
      const int k0 = tl_pos(te), k1 = k0 + tl_width(te) - 1;
      const bool full = tl_full(te);
      uint32_t mb = 0xffffffffu;
      if (!full) {
        if (qi < q_lim) { const Mask128 mk = attn_row_bits(g, qi, k0, km, k1 - k0 + 1); mb = chunk == 0 ? mk.w[0] : mk.w[1]; }
        else mb = 0u;
      }
      mbar_wait_mode(s_full, it & 1, P.wait_mode & 1);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const uint32_t mb16 = (mb >> (16 * sub)) & 0xffffu;
        const int col = chunk * 32 + sub * 16;
        uint32_t dk_[8];
        if (__all_sync(0xffffffffu, mb16 == 0u)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) dk_[i] = 0u;
        } else {
          uint32_t rs[16], rd[16];
          tmem_ld16(tS + lane_off + col, rs);
          tmem_ld16(tdP + lane_off + col, rd);
          tmem_ld_wait();
          bwd_softmax16<false, false>(rs, rd, mb16, 0u, 0u, lse_r, delta_r, nullptr, dk_);
        }
        tmem_st8(tdS + lane_off + chunk * 32 + sub * 8, dk_);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ds_ready);
    }
    const int it = n_its;
    if (it > 0) {
      mbar_wait(acc_done, (it - 1) & 1);
      tc_fence_after();
      const int qs = qi < q_lim ? attn_nat(g, qi) : 0;                     // token position of this query row
      bf16* rowp = P.dqkv + ((long long)b * n + qs) * (3 * inner) + h * DH;
      const float* cr = P.cos_t ? P.cos_t + (long long)qs * (DH / 2) : nullptr;
      const float* sr = P.sin_t ? P.sin_t + (long long)qs * (DH / 2) : nullptr;
      store_grad_cols(tdQ + lane_off, rowp, cr, sr, P.q_scale, qi < q_lim, chunk);    // q = rot(x) * scale (attention.py:69)
    } else if (qi < q_lim) {
      bf16* rowp = P.dqkv + ((long long)b * n + attn_nat(g, qi)) * (3 * inner) + h * DH + chunk * 32;
      for (int d = 0; d < 32; d += 8) *reinterpret_cast<uint4*>(rowp + d) = make_uint4(0, 0, 0, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int tc_mode_env() {   // DALLE_B200_ATTN_P=smem forces the SS variant (P staged through shared memory)
  const char* v = getenv("DALLE_B200_ATTN_P");
  return (v && !strcmp(v, "smem")) ? 1 : 0;
}

// Strided box over the image tokens of a [b*h, n_alloc, 64] tensor for axis 1: dims {dh, image row r, image column c, (b,h)},
// pitches {1, fm, 1, n_alloc} tokens -- the box {64, fm, W/fm, 1} lands in shared memory with the image ROW as the faster index,
// i.e. as W consecutive tokens of the column-major virtual order (attn_common.cuh).
int make_gather_map_qkv(CUtensorMap* map, const void* base, const AttnGeom& g, uint64_t bh, int W) {
  const uint64_t fm = g.fmap;
  const uint64_t dims[4] = {DH, fm, fm, bh};
  const uint64_t strides[4] = {1, fm * DH, DH, (uint64_t)g.n_alloc * DH};
  const uint32_t box[4] = {DH, (uint32_t)(W < (int)fm ? W : fm), (uint32_t)(W < (int)fm ? 1 : W / fm), 1};   // (part of) whole columns
  return make_tensor_map_bf16_nd(map, reinterpret_cast<const bf16*>(base) + (size_t)g.text_len * DH, 4, dims, strides, box);
}
// the same box over the [b, n, heads*64] gradient of the attention output: dims {dh, r, c, head, batch}
int make_gather_map_do(CUtensorMap* map, const void* base, const AttnGeom& g, int batch, int heads, int n, int W) {
  const uint64_t fm = g.fmap, inner = (uint64_t)heads * DH;
  const uint64_t dims[5] = {DH, fm, fm, (uint64_t)heads, (uint64_t)batch};
  const uint64_t strides[5] = {1, fm * inner, inner, DH, (uint64_t)n * inner};
  const uint32_t box[5] = {DH, (uint32_t)fm, (uint32_t)(W / fm), 1, 1};
  return make_tensor_map_bf16_nd(map, reinterpret_cast<const bf16*>(base) + (size_t)g.text_len * inner, 5, dims, strides, box);
}

template <bool P_TMEM>
int launch_fwd(const db200_attn_fwd_params& p, cudaStream_t st) {
  using L = FwdSmem<P_TMEM>;
  CUtensorMap tmQ, tmK, tmV;
  const AttnGeom g = make_geom(p);
  const uint64_t bh = (uint64_t)p.batch * p.heads;
  const uint64_t rows_q = g.gather ? g.n_alloc : p.n_q, rows_k = g.gather ? g.n_alloc : g.kv_rows;
  int rc = make_tensor_map_bf16(&tmQ, p.q, DH, bh * rows_q, DH, DH, TQ);
  if (rc) return rc;
  if ((rc = make_tensor_map_bf16(&tmK, p.k, DH, bh * rows_k, DH, DH, FK))) return rc;
  if ((rc = make_tensor_map_bf16(&tmV, p.v, DH, bh * rows_k, DH, DH, FK))) return rc;
  CUtensorMap tmQg = tmQ, tmKg = tmK, tmVg = tmV;          // placeholders unless the column gather is on
  if (g.col) {
    if ((rc = make_gather_map_qkv(&tmQg, p.q, g, bh, TQ))) return rc;
    if ((rc = make_gather_map_qkv(&tmKg, p.k, g, bh, FK))) return rc;
    if ((rc = make_gather_map_qkv(&tmVg, p.v, g, bh, FK))) return rc;
  }
  FwdArgs A{reinterpret_cast<bf16*>(p.out), p.lse, p.key_mask, p.heads, p.batch};
  dim3 grid(seg_tile_count(g, TQ, p.n_q), p.batch * p.heads);
  // DALLE_B200_ATTN_FWD = v1 (single S buffer, 64-key tiles, three CTAs per SM; default: measured fastest) | v2_64 | v2_32
  // (double-buffered S, profiles/r02_summary.md): A/B timing
  static const int variant = [] {
    const char* v = getenv("DALLE_B200_ATTN_FWD");
    return !v ? 0 : !strcmp(v, "v2_32") ? 32 : !strcmp(v, "v2_64") ? 64 : 0;
  }();
  if (P_TMEM && variant != 0) {
    static std::atomic<bool> attr2_done{false};
    if (!attr2_done.load(std::memory_order_acquire)) {
      DB200_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Smem<32>::TOTAL));
      DB200_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Smem<64>::TOTAL));
      attr2_done.store(true, std::memory_order_release);
    }
    if (variant == 32) {
      CUtensorMap k32, v32, kg32 = tmKg, vg32 = tmVg;
      if ((rc = make_tensor_map_bf16(&k32, p.k, DH, bh * rows_k, DH, DH, 32))) return rc;
      if ((rc = make_tensor_map_bf16(&v32, p.v, DH, bh * rows_k, DH, DH, 32))) return rc;
      if (g.col) {
        if ((rc = make_gather_map_qkv(&kg32, p.k, g, bh, 32))) return rc;
        if ((rc = make_gather_map_qkv(&vg32, p.v, g, bh, 32))) return rc;
      }
      DB200_CUDA_OK(launch_pdl(attn_fwd_tc2_kernel<32>, grid, dim3(192), Fwd2Smem<32>::TOTAL, st, tmQ, k32, v32, tmQg, kg32, vg32, A, g));
    } else {
      DB200_CUDA_OK(launch_pdl(attn_fwd_tc2_kernel<64>, grid, dim3(192), Fwd2Smem<64>::TOTAL, st, tmQ, tmK, tmV, tmQg, tmKg, tmVg, A, g));
    }
    DB200_LAUNCH_OK("attn_fwd_tc2_kernel");
    return DB200_OK;
  }
  auto kern = attn_fwd_tc_kernel<P_TMEM>;
  static std::atomic<bool> attr_done{false};   // idempotent set-up; atomic because forward and autograd threads both launch
  if (!attr_done.load(std::memory_order_acquire)) {
    DB200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done.store(true, std::memory_order_release);
  }
  DB200_CUDA_OK(launch_pdl(kern, grid, dim3(192), L::TOTAL, st, tmQ, tmK, tmV, tmQg, tmKg, tmVg, A, g));
  DB200_LAUNCH_OK("attn_fwd_tc_kernel");
  return DB200_OK;
}

}  // namespace

int attn_debug_timeline(long long* out, int count) {
  if (count > 256) count = 256;
  return cudaMemcpyFromSymbol(out, g_attn_dbg, sizeof(long long) * count) == cudaSuccess ? 0 : -1;
}

// Gathered axial mode (db200_attn_fwd_params::gather): what the caller must have laid out, see include/dalle_b200.h
bool attn_gather_ok(const db200_attn_fwd_params& p, const char** why) {
  const int fm = p.fmap, T = p.text_len;
  const char* w = nullptr;
  if (p.pattern != DB200_ATTN_AXIAL_ROW && p.pattern != DB200_ATTN_AXIAL_COL) w = "gather needs an axial pattern";
  else if (p.dtype != DB200_BF16 || p.dim_head != 64) w = "gather needs bf16, dim_head 64";
  else if (fm != 16 && fm != 32 && fm != 64) w = "gather needs fmap in {16, 32, 64}";
  else if (p.n_q != p.n_k) w = "gather needs n_q == n_k";
  else if (p.n_k != T + fm * fm - 1 && p.n_k != T + fm * fm) w = "gather needs the full image (n = text_len + fmap^2 [- 1])";
  else if (p.key_mask != nullptr) w = "gather does not take a key mask";
  if (why) *why = w;
  return w == nullptr;
}

bool attn_tc_supported(const db200_attn_fwd_params& p) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 && p.dtype == DB200_BF16 && p.dim_head == 64 && al16(p.q) && al16(p.k) && al16(p.v) && al16(p.out) &&
         p.n_k <= 32 * (TL_CAP - 8);      // capacity of the per-CTA tile list (32-wide key tiles)
}

int attn_fwd_tc_launch(const db200_attn_fwd_params& p, cudaStream_t st) {
  return tc_mode_env() ? launch_fwd<false>(p, st) : launch_fwd<true>(p, st);
}

int attn_bwd_tc_launch(const db200_attn_bwd_params& p, cudaStream_t st) {
  const db200_attn_fwd_params& f = p.f;
  const int n = f.n_k;
  if (!al16(p.d_out) || !al16(p.dqkv)) return set_error(DB200_ERR_BAD_ARG, "attn_bwd: d_out / dqkv must be 16-byte aligned");
  const AttnGeom g = make_geom(f);
  // 128-row boxes for the CTA's own rows, 64-row boxes for the streamed column tiles
  CUtensorMap tmQ128, tmK128, tmV128, tmdO128, tmQ64, tmK64, tmV64, tmdO64;
  const uint64_t bh = (uint64_t)f.batch * f.heads;
  const uint64_t inner = (uint64_t)f.heads * DH;
  const uint64_t rows = g.gather ? g.n_alloc : n;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmQ128, f.q, DH, bh * rows, DH, DH, 128))) return rc;
  if ((rc = make_tensor_map_bf16(&tmK128, f.k, DH, bh * rows, DH, DH, 128))) return rc;
  if ((rc = make_tensor_map_bf16(&tmV128, f.v, DH, bh * rows, DH, DH, 128))) return rc;
  if ((rc = make_tensor_map_bf16(&tmdO128, p.d_out, inner, (uint64_t)f.batch * n, inner, DH, 128))) return rc;
  if ((rc = make_tensor_map_bf16(&tmQ64, f.q, DH, bh * rows, DH, DH, BW))) return rc;
  if ((rc = make_tensor_map_bf16(&tmK64, f.k, DH, bh * rows, DH, DH, BW))) return rc;
  if ((rc = make_tensor_map_bf16(&tmV64, f.v, DH, bh * rows, DH, DH, BW))) return rc;
  if ((rc = make_tensor_map_bf16(&tmdO64, p.d_out, inner, (uint64_t)f.batch * n, inner, DH, BW))) return rc;
  CUtensorMap gQ128 = tmQ128, gK128 = tmK128, gV128 = tmV128, gdO128 = tmdO128, gQ64 = tmQ64, gK64 = tmK64, gV64 = tmV64, gdO64 = tmdO64;
  bf16* dO_pad_row = nullptr;
  if (g.col) {
    if ((rc = make_gather_map_qkv(&gQ128, f.q, g, bh, 128))) return rc;
    if ((rc = make_gather_map_qkv(&gK128, f.k, g, bh, 128))) return rc;
    if ((rc = make_gather_map_qkv(&gV128, f.v, g, bh, 128))) return rc;
    if ((rc = make_gather_map_do(&gdO128, p.d_out, g, f.batch, f.heads, n, 128))) return rc;
    if ((rc = make_gather_map_qkv(&gQ64, f.q, g, bh, BW))) return rc;
    if ((rc = make_gather_map_qkv(&gK64, f.k, g, bh, BW))) return rc;
    if ((rc = make_gather_map_qkv(&gV64, f.v, g, bh, BW))) return rc;
    if ((rc = make_gather_map_do(&gdO64, p.d_out, g, f.batch, f.heads, n, BW))) return rc;
    // the box of the last image column of the last batch reaches one token past the end of d_out: the caller allocates that
    // row (include/dalle_b200.h), the delta kernel zeroes it
    if (n < g.text_len + g.fmap * g.fmap)
      dO_pad_row = reinterpret_cast<bf16*>(const_cast<void*>(p.d_out)) + (size_t)f.batch * n * inner;
  }
  static std::atomic<bool> attr_done{false};   // idempotent set-up; atomic because forward and autograd threads both launch
  if (!attr_done.load(std::memory_order_acquire)) {
    DB200_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::TOTAL));
    DB200_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::TOTAL));
    attr_done.store(true, std::memory_order_release);
  }
  const int total_rows = f.batch * f.heads * n;
  {
    int blocks = ceil_div(total_rows, 32 * 4);
    const int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    DB200_CUDA_OK(launch_pdl(attn_delta_tc_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const bf16*>(f.out),
                             reinterpret_cast<const bf16*>(p.d_out), p.delta, f.batch, f.heads, n, g, dO_pad_row));
  }
  DB200_LAUNCH_OK("attn_delta_tc_kernel");
  static const int wait_mode = [] { const char* v = getenv("DALLE_B200_ATTN_WAIT"); return v ? atoi(v) : 0; }();
  BwdArgs A{f.lse, p.delta, f.key_mask, p.cos_t, p.sin_t, p.q_scale, reinterpret_cast<bf16*>(p.dqkv), f.heads, f.batch, wait_mode};
  dim3 grid(seg_tile_count(g, TQ, n), f.batch * f.heads);
  DB200_CUDA_OK(launch_pdl(attn_bwd_dkv_tc_kernel, grid, dim3(BWD_THREADS), DkvSmem::TOTAL, st, tmQ64, tmK128, tmV128, tmdO64, gQ64, gK128, gV128,
                           gdO64, A, g));
  DB200_LAUNCH_OK("attn_bwd_dkv_tc_kernel");
  DB200_CUDA_OK(launch_pdl(attn_bwd_dq_tc_kernel, grid, dim3(BWD_THREADS), DqSmem::TOTAL, st, tmQ128, tmK64, tmV64, tmdO128, gQ128, gK64, gV64,
                           gdO128, A, g));
  DB200_LAUNCH_OK("attn_bwd_dq_tc_kernel");
  return DB200_OK;
}

}  // namespace db200
