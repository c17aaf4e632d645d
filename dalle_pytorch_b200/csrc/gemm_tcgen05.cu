// bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA
// into 128B-swizzled shared memory, warp-specialised producer / MMA-issuer / epilogue roles, persistent CTAs
// (one per SM) with a double-buffered TMEM accumulator so the epilogue of tile i overlaps the MMAs of tile i+1.
//
//   warp 0      : TMA producer + dynamic tile scheduler (whole warp converged, elect.sync picks the issuing lane per k-block;
//                                 cp.async.bulk.tensor -> smem ring, mbarrier complete_tx; tile ids drawn with atomicAdd and
//                                 published to the other roles through a small smem ring)
//   warp 1      : MMA issuer     (converged warp + elect.sync; tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16;
//                                 tcgen05.commit releases smem stages / publishes the accumulator)
//   warps 2..9  : epilogue       (tcgen05.ld 32x32b gives thread = accumulator row; each 32x32 chunk is transposed through a
//                                 private 4 KB smem buffer so that LANES RUN ALONG N: every global load/store of the fused
//                                 epilogues (epilogue.cuh) is then a contiguous 64-128 B row segment per half-warp)
//
// All three GEMM shapes of training map onto the same kernel through the operand "major" flags:
//   forward : A K-major (activations [M,K]),  B K-major (weight [N,K])
//   dgrad   : A K-major (dY [M,K=out]),       B MN-major (weight [out,in] read as B(n=in, k=out))
//   wgrad   : A MN-major (dY [rows, out]),    B MN-major (X [rows, in])     -> C = dW [out, in] fp32
// (UMMA shared-memory descriptors support both majors for bf16; MN-major tiles are loaded as 64x64 boxes.)
#include <cuda.h>
#include <atomic>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "epilogue.cuh"
#include "tc_common.cuh"

namespace db200 {

using namespace tc;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;          // 2 + 8 warps
constexpr int NUM_EPI_WARPS = 8;
constexpr int SCHED_DEPTH = 4;                 // tile-id ring between the producer and the MMA / epilogue roles
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;   // per epilogue warp
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int BOX_MN_BYTES = 64 * BLOCK_K * 2;         // one 64(mn) x 64(k) MN-major box = 8 KB
constexpr int SMEM_BUDGET = 193 * 1024;   // operand ring; + 32 KB epilogue staging + barriers stays under 227 KB

// 32x32 fp32 chunk: in[] = 32 consecutive columns of this lane's row  ->  out[2*it], out[2*it+1] = columns
// 2*(lane&15), +1 of row 2*it + (lane>>4).  Private per-warp buffer, float4 slots XOR-swizzled by the row.
__device__ __forceinline__ void transpose_chunk(const uint32_t* in, float* out, float* stage, int lane) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(in[4 * j], in[4 * j + 1], in[4 * j + 2], in[4 * j + 3]);
  __syncwarp();
  const int j = (lane & 15) >> 1, sub = (lane & 1) << 1;
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int rr = 2 * it + (lane >> 4);
    const float2 v = *reinterpret_cast<const float2*>(stage + rr * 32 + ((j ^ (rr & 7)) << 2) + sub);
    out[2 * it] = v.x;
    out[2 * it + 1] = v.y;
  }
}

template <int BLOCK_N>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int BAR_BYTES = (2 * STAGES + 4 + 2 * SCHED_DEPTH) * 8 + 16 + SCHED_DEPTH * 4 + 12;
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = EPI_OFFSET + NUM_EPI_WARPS * EPI_STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + BAR_BYTES + 1024;   // +1024: manual 1 KB alignment slack
};

// Dynamic tile scheduler: every tile is drawn with atomicAdd(counter).  A CTA that is not resident when the kernel starts
// (an SM is held by a concurrent kernel, e.g. the NCCL all-reduce of the previous gradient bucket, whose shared memory
// keeps a 200 KB GEMM CTA out) simply takes no tiles, instead of leaving its static share for the very end (measured:
// the dgrad GEMM overlapped by the all-reduce went 0.14 -> 0.24 ms with the static schedule).  Exactly
// num_tiles + gridDim.x draws happen per launch (one failing draw per CTA), so the thread that draws the last value
// resets the counter for the next launch using the slot; the host rotates over SCHED_SLOTS counters so launches that run
// concurrently on different streams never share one.
constexpr int SCHED_SLOTS = 256;
__device__ unsigned int g_sched_counter[SCHED_SLOTS];
static std::atomic<unsigned int> g_launch_seq{0};

template <int BLOCK_N, int EPI, bool A_MN, bool B_MN, bool EPI_COLS>
// 320 threads are allocated as 12 warps of registers (4-warp granularity) -> 168 registers per thread at most
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                    int M, int N, int K, int k_splits, int sched_slot, EpiArgs e) {
  using L = SmemLayout<BLOCK_N>;
  constexpr int STAGES = L::STAGES;
  constexpr int B_STAGE_BYTES = L::B_STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;                 // double-buffered fp32 accumulator (power of two >= 32)
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");
  static_assert(EPI != DB200_EPI_GEGLU || (BLOCK_N == 256 && !B_MN), "GEGLU pairs the two 128-column halves of a 256 tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  const uint32_t full_bar = smem_u32(bars);                    // [STAGES]
  const uint32_t empty_bar = smem_u32(bars + STAGES);          // [STAGES]
  const uint32_t tfull_bar = smem_u32(bars + 2 * STAGES);      // [2]
  const uint32_t tempty_bar = smem_u32(bars + 2 * STAGES + 2); // [2]
  const uint32_t sfull_bar = smem_u32(bars + 2 * STAGES + 4);                 // [SCHED_DEPTH] tile id published
  const uint32_t sempty_bar = smem_u32(bars + 2 * STAGES + 4 + SCHED_DEPTH);  // [SCHED_DEPTH] tile id consumed by MMA + epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + 2 * SCHED_DEPTH);
  volatile int* sched_tile = reinterpret_cast<volatile int*>(tmem_slot + 4);  // [SCHED_DEPTH]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_launch();
  unsigned int first_draw = 0;
  const bool dyn = sched_slot >= 0;                            // sched_slot < 0: static round-robin schedule (A/B switch)
  if (threadIdx.x == 0) first_draw = dyn ? atomicAdd(&g_sched_counter[sched_slot], 1u) : blockIdx.x;   // latency hides behind the setup below
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, NUM_EPI_WARPS * 32); }
    for (int s = 0; s < SCHED_DEPTH; ++s) { mbar_init(sfull_bar + 8 * s, 1); mbar_init(sempty_bar + 8 * s, 1 + NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                                 // everything above overlapped the previous kernel's tail

  const int num_m = (M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_mn = num_m * num_n;
  const int num_tiles = num_mn * k_splits;                    // split-K: tile = (split, n, m); partial sums are reduced with
  const int total_kb = (K + BLOCK_K - 1) / BLOCK_K;           // fp32 red.global.add in the STORE epilogue (C pre-zeroed)
  const int kb_per = (total_kb + k_splits - 1) / k_splits;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // Whole warp converged; the TMA instructions of a k-block are issued by one elected lane (operands stay warp-uniform, no
    // R2UR / ELECT waterfall per cp.async.bulk.tensor -- see elect_one()).  Lane 0 owns the scheduler traffic.
    {
      int s = 0; uint32_t ph = 0;
      int ss = 0; uint32_t sph = 0;
      unsigned int* counter = &g_sched_counter[dyn ? sched_slot : 0];
      const unsigned int last_draw = dyn ? static_cast<unsigned int>(num_tiles) + gridDim.x - 1u : 0xffffffffu;
      if (lane == 0 && first_draw == last_draw) atomicExch(counter, 0u);
      int tile = static_cast<int>(__shfl_sync(0xffffffffu, first_draw, 0));
      while (true) {
        mbar_wait(sempty_bar + 8 * ss, sph ^ 1);             // publish the tile id (or the end marker) to the other roles
        if (lane == 0) {
          sched_tile[ss] = tile;
          mbar_arrive(sfull_bar + 8 * ss);
        }
        __syncwarp();
        if (++ss == SCHED_DEPTH) { ss = 0; sph ^= 1; }
        if (tile >= num_tiles) break;
        unsigned int drawn = 0;
        if (lane == 0) drawn = dyn ? atomicAdd(counter, 1u) : static_cast<unsigned int>(tile) + gridDim.x;   // next tile; latency hides behind the loads
        const int mn = tile % num_mn, split = tile / num_mn;
        const int m0 = (mn / num_n) * BLOCK_M;          // n-fastest: CTAs running together share the A tile through L2
        const int n0 = (mn % num_n) * BLOCK_N;
        const int kb0 = split * kb_per, kb1 = min(total_kb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          const uint32_t fb = full_bar + 8 * s;
          const uint32_t sa = smem_u32(smem_a + s * A_STAGE_BYTES);
          const uint32_t sb = smem_u32(smem_b + s * B_STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          if (elect_one()) {
            mbar_expect_tx(fb, L::STAGE_BYTES);
            if constexpr (!A_MN) {
              tma_load_2d(sa, &tmA, fb, k0, m0);                                   // box {64 k, 128 m}
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d(sa + c * BOX_MN_BYTES, &tmA, fb, m0 + c * 64, k0);   // box {64 m, 64 k}
            }
            if constexpr (!B_MN) {
#pragma unroll
              for (int c = 0; c < BLOCK_N / 128; ++c) {
                int row = n0 + c * 128;
                if constexpr (EPI == DB200_EPI_GEGLU) row = (c == 0) ? (n0 >> 1) : e.hidden + (n0 >> 1);
                tma_load_2d(sb + c * (128 * BLOCK_K * 2), &tmB, fb, k0, row);      // box {64 k, 128 n}
              }
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_N / 64; ++c) tma_load_2d(sb + c * BOX_MN_BYTES, &tmB, fb, n0 + c * 64, k0);   // box {64 n, 64 k}
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        drawn = __shfl_sync(0xffffffffu, drawn, 0);
        if (lane == 0 && drawn == last_draw) atomicExch(counter, 0u);    // last draw of this launch: leave the slot clean
        tile = static_cast<int>(drawn);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // Whole warp converged, one elected lane issues each k-block (see elect_one() in tc_common.cuh): descriptors stay in
    // uniform registers instead of paying an R2UR + ELECT waterfall per tcgen05.mma.
    {
      // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, a_major bit15, b_major bit16,
      // N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u) |
                             (static_cast<uint32_t>(BLOCK_N >> 3) << 17) | (static_cast<uint32_t>(BLOCK_M >> 4) << 24);
      // K-major  : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (1).      k-step = 32 B
      // MN-major : 64-wide mn chunks 8 KB apart (LBO), 8-k-row groups 1024 B apart (SBO). k-step = 16 rows * 128 B
      constexpr uint32_t A_LBO = A_MN ? BOX_MN_BYTES : 16, B_LBO = B_MN ? BOX_MN_BYTES : 16;
      constexpr uint32_t A_KSTEP = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      constexpr uint32_t B_KSTEP = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      int ss = 0; uint32_t sph = 0;
      while (true) {
        mbar_wait(sfull_bar + 8 * ss, sph);
        const int tile = sched_tile[ss];
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty_bar + 8 * ss);
        if (++ss == SCHED_DEPTH) { ss = 0; sph ^= 1; }
        if (tile >= num_tiles) break;
        mbar_wait(tempty_bar + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + as * BLOCK_N;
        const int split = tile / num_mn;
        const int kb0 = split * kb_per, kb1 = min(total_kb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES), A_LBO, 1024);
          const uint64_t db = make_smem_desc(smem_u32(smem_b + s * B_STAGE_BYTES), B_LBO, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(tmem_c, da + static_cast<uint64_t>(k * A_KSTEP), db + static_cast<uint64_t>(k * B_KSTEP), idesc, (kb > kb0) || (k != 0));
            umma_commit(empty_bar + 8 * s);                    // smem stage free once these MMAs retire
            if (kb == kb1 - 1) umma_commit(tfull_bar + 8 * as);   // accumulator complete
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    const int ew = warp - 2;
    const int quarter = warp & 3;                            // TMEM lane quarter this warp may access
    const int half = ew >> 2;                                // which half of the tile's columns
    float* stage = reinterpret_cast<float*>(smem + L::EPI_OFFSET + ew * EPI_STAGE_BYTES);
    const int lrow = lane >> 4, lcol = (lane & 15) << 1;
    int as = 0; uint32_t aph = 0;
    int ss = 0; uint32_t sph = 0;
    while (true) {
      mbar_wait(sfull_bar + 8 * ss, sph);
      const int tile = sched_tile[ss];
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty_bar + 8 * ss);
      if (++ss == SCHED_DEPTH) { ss = 0; sph ^= 1; }
      if (tile >= num_tiles) break;
      const int mn = tile % num_mn;
      const int m0 = (mn / num_n) * BLOCK_M;          // n-fastest: CTAs running together share the A tile through L2
      const int n0 = (mn % num_n) * BLOCK_N;
      const int mbase = m0 + quarter * 32 + lrow;            // + 2*it
      mbar_wait(tfull_bar + 8 * as, aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BLOCK_N;
      if (e.atomic_c == 4) {
        // (diagnosis: no epilogue work at all)
      } else if (EPI == DB200_EPI_STORE && !EPI_COLS && e.tma_store) {
        // ---- bf16 STORE through TMA: thread = accumulator row packs 32 columns (64 B) per chunk into this warp's 4 KB staging
        // buffer in the 128B-swizzled layout; two chunks make one [32 rows x 64 columns] box = one cp.async.bulk.tensor store (full
        // 128-byte lines, issued by one lane, clipped at the M / N edges by the hardware).  The 16-byte stores of the register
        // path touched 32 different lines per warp instruction and were 27 % of the time of the K = 1024 GEMMs (tools/gemm_gap.py).
        constexpr int NCH = (BLOCK_N / 2) / 32;
        const uint32_t stage_u32 = smem_u32(stage);
        uint32_t r[2][32];
        tmem_ld32(taddr + half * (BLOCK_N / 2), r[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tmem_ld_wait();
          if (ch + 1 < NCH) tmem_ld32(taddr + half * (BLOCK_N / 2) + (ch + 1) * 32, r[(ch + 1) & 1]);
          const int n = n0 + half * (BLOCK_N / 2) + ch * 32;
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = __uint_as_float(r[ch & 1][2 * i]), b = __uint_as_float(r[ch & 1][2 * i + 1]);
            if (e.bias && n + 2 * i < N) { const float2 bb = __ldg(reinterpret_cast<const float2*>(e.bias + n + 2 * i)); a += bb.x; b += bb.y; }
            const __nv_bfloat162 p2 = __floats2bfloat162_rn(a, b);
            w[i] = *reinterpret_cast<const uint32_t*>(&p2);
          }
          if ((ch & 1) == 0) {                                 // the previous box must have been read out of the staging buffer
            if (lane == 0) bulk_wait_group_read0();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int unit = (ch & 1) * 4 + j;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_u32 + lane * 128 + ((unit ^ (lane & 7)) << 4)), "r"(w[4 * j]),
                         "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                         : "memory");
          }
          if (ch & 1) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, stage_u32, n0 + half * (BLOCK_N / 2) + (ch >> 1) * 64, m0 + quarter * 32);
              bulk_commit_group();
            }
          }
        }
      } else
      if constexpr (!EPI_COLS) {
        // ---- row mode: thread = accumulator row, 8-column (16-byte) granules straight from registers ----
        const int m = m0 + quarter * 32 + lane;
        if constexpr (EPI == DB200_EPI_GEGLU) {
          uint32_t ra[2][32], rg[2][32];
          tmem_ld32(taddr + half * 64, ra[0]);
          tmem_ld32(taddr + 128 + half * 64, rg[0]);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            tmem_ld_wait();
            if (ch + 1 < 2) {
              tmem_ld32(taddr + half * 64 + 32, ra[1]);
              tmem_ld32(taddr + 128 + half * 64 + 32, rg[1]);
            }
            const int j0 = (n0 >> 1) + half * 64 + ch * 32;
            if (m < M) {
              if ((e.hidden & 15) == 0) {                       // full-sector stores of u (a | g) and h
#pragma unroll
                for (int o2 = 0; o2 < 2; ++o2)
                  if (j0 + o2 * 16 < e.hidden) epi_geglu16<__nv_bfloat16>(e, m, j0 + o2 * 16, &ra[ch][o2 * 16], &rg[ch][o2 * 16]);
              } else {
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                  const int j = j0 + o * 8;
                  if (j < e.hidden) {
                    float a[8], g[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(ra[ch][o * 8 + i]); g[i] = __uint_as_float(rg[ch][o * 8 + i]); }
                    epi_geglu_vec8<__nv_bfloat16>(e, m, j, a, g);
                  }
                }
              }
            }
          }
        } else {
          constexpr int NCH = (BLOCK_N / 2) / 32;
          uint32_t r[2][32];
          tmem_ld32(taddr + half * (BLOCK_N / 2), r[0]);
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            tmem_ld_wait();
            if (ch + 1 < NCH) tmem_ld32(taddr + half * (BLOCK_N / 2) + (ch + 1) * 32, r[(ch + 1) & 1]);
            if (m < M) {
#pragma unroll
              for (int o2 = 0; o2 < 2; ++o2) {
                const int nn = n0 + half * (BLOCK_N / 2) + ch * 32 + o2 * 16;
                bool done = false;
                if constexpr (EPI == DB200_EPI_STORE) {          // one 32-byte sector per lane (two for an fp32 result)
                  if (nn + 16 <= N) done = epi_store16<__nv_bfloat16>(e, m, nn, &r[ch & 1][o2 * 16]);
                }
                if constexpr (EPI == DB200_EPI_RESID) {          // full-sector residual loads / out, y stores (N % 8 == 0 checked by the launcher)
                  if (nn + 16 <= N) { epi_resid16<__nv_bfloat16>(e, m, nn, &r[ch & 1][o2 * 16]); done = true; }
                }
                if (!done) {
#pragma unroll
                  for (int o = 2 * o2; o < 2 * o2 + 2; ++o) {
                    const int n = n0 + half * (BLOCK_N / 2) + ch * 32 + o * 8;
                    if (n < N) {
                      float v[8];
#pragma unroll
                      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[ch & 1][o * 8 + i]);
                      epi_vec8<EPI, __nv_bfloat16>(e, m, n, v);
                    }
                  }
                }
              }
            }
          }
        }
      } else
      if constexpr (EPI == DB200_EPI_GEGLU) {
        // a = columns [0,128), g = columns [128,256) of the accumulator; this warp pairs a/g of hidden [half*64, +64)
        uint32_t ra[32], rg[32];
        tmem_ld32(taddr + half * 64, ra);
        tmem_ld32(taddr + 128 + half * 64, rg);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          tmem_ld_wait();
          float ta[32], tg[32];
          transpose_chunk(ra, ta, stage, lane);
          transpose_chunk(rg, tg, stage, lane);
          if (ch + 1 < 2) {                                  // the raw registers are free again: fetch the next chunk now
            tmem_ld32(taddr + half * 64 + 32, ra);
            tmem_ld32(taddr + 128 + half * 64 + 32, rg);
          }
          const int j = (n0 >> 1) + half * 64 + ch * 32 + lcol;
          if (j < e.hidden) epi_chunk_cols_geglu<__nv_bfloat16>(e, mbase, j, ta, tg, M);
        }
      } else {
        constexpr int NCH = (BLOCK_N / 2) / 32;
        uint32_t r[32];
        tmem_ld32(taddr + half * (BLOCK_N / 2), r);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tmem_ld_wait();
          float t[32];
          transpose_chunk(r, t, stage, lane);
          if (ch + 1 < NCH) tmem_ld32(taddr + half * (BLOCK_N / 2) + (ch + 1) * 32, r);
          const int n = n0 + half * (BLOCK_N / 2) + ch * 32 + lcol;
          if (n < N) epi_chunk_cols<EPI, __nv_bfloat16>(e, mbase, n, t, M);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + 8 * as);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (EPI == DB200_EPI_STORE && !EPI_COLS && e.tma_store && lane == 0) bulk_wait_group0();   // outstanding tensor stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

inline bool al16(const void* p);
inline bool tma_store_env() {      // DALLE_B200_GEMM_TMA_STORE=0: register-path stores (A/B timing)
  static const bool on = [] { const char* v = getenv("DALLE_B200_GEMM_TMA_STORE"); return !(v && !strcmp(v, "0")); }();
  return on;
}

template <int BLOCK_N, int EPI, bool A_MN, bool B_MN, bool EPI_COLS>
int launch_cfg_mode(const db200_gemm_params& p, cudaStream_t st) {
  using L = SmemLayout<BLOCK_N>;
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_tensor_map_bf16(&tmA, p.A, p.K, p.M, p.lda, BLOCK_K, BLOCK_M);
  else rc = make_tensor_map_bf16(&tmA, p.A, p.M, p.K, p.lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!B_MN) rc = make_tensor_map_bf16(&tmB, p.B, p.K, p.N, p.ldb, BLOCK_K, 128);
  else rc = make_tensor_map_bf16(&tmB, p.B, p.N, p.K, p.ldb, 64, BLOCK_K);
  if (rc) return rc;
  CUtensorMap tmC = tmA;                                      // placeholder unless the TMA-store epilogue is used
  EpiArgs e = make_epi_args(p);
  if (EPI == DB200_EPI_STORE && !EPI_COLS && e.atomic_c == 0 && p.c_dtype == DB200_BF16 && !p.split_k_ok && al16(p.C) && (p.ldc % 8) == 0 &&
      tma_store_env()) {
    if ((rc = make_tensor_map_bf16(&tmC, p.C, p.N, p.M, p.ldc, 64, 32))) return rc;
    e.tma_store = 1;
  }
  auto kern = gemm_tcgen05_kernel<BLOCK_N, EPI, A_MN, B_MN, EPI_COLS>;
  static std::atomic<bool> attr_done{false};   // idempotent set-up; atomic because forward and autograd threads both launch
  if (!attr_done.load(std::memory_order_acquire)) {
    DB200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done.store(true, std::memory_order_release);
  }
  const int num_mn = ceil_div(p.M, BLOCK_M) * ceil_div(p.N, BLOCK_N);
  // split-K only where the reduction is free to express: fp32 STORE without bias into a buffer the caller has zeroed
  int k_splits = 1;
  if (EPI == DB200_EPI_STORE && p.split_k_ok && p.c_dtype == DB200_F32 && p.bias == nullptr) {
    const int total_kb = ceil_div(p.K, BLOCK_K);
    k_splits = sm_count() / num_mn;
    if (k_splits > 8) k_splits = 8;
    if (k_splits > total_kb / 8) k_splits = total_kb / 8;       // keep >= 8 k-blocks per split
    if (k_splits < 1) k_splits = 1;
  }
  const int num_tiles = num_mn * k_splits;
  const int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  if (e.atomic_c < 2) e.atomic_c = k_splits > 1;      // (multimem reductions are additive already: split-K needs nothing more)
  static const bool static_sched = [] { const char* v = std::getenv("DALLE_B200_SCHED"); return v && !std::strcmp(v, "static"); }();
  const int sched_slot = static_sched ? -1 : static_cast<int>(g_launch_seq.fetch_add(1, std::memory_order_relaxed) % SCHED_SLOTS);
  DB200_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), L::TOTAL, st, tmA, tmB, tmC, p.M, p.N, p.K, k_splits, sched_slot, e));
  DB200_LAUNCH_OK("gemm_tcgen05_kernel");
  return DB200_OK;
}

// Epilogue mode: "rows" (thread = accumulator row, 16-byte granules) or "cols" (32x32 chunks transposed through smem so
// lanes run along N).  Measured on B200 (profiles/): cols wins only for the GEGLU forward epilogue; DALLE_B200_EPI=rows|cols
// forces one mode for A/B timing.
inline int epi_mode_env() {
  static const int mode = [] {
    const char* v = getenv("DALLE_B200_EPI");
    return (v && !strcmp(v, "rows")) ? 1 : (v && !strcmp(v, "cols")) ? 2 : 0;
  }();
  return mode;
}

template <int BLOCK_N, int EPI, bool A_MN, bool B_MN>
int launch_cfg(const db200_gemm_params& p, cudaStream_t st) {
  const int env = epi_mode_env();
  // (multicast reduction: lanes must run along N so that one warp instruction is one contiguous 256-byte run -- NVLink packets of
  //  16 scattered bytes per lane, the "rows" layout, were measured 10 ms per C2 step slower)
  // (r02, tools/gemm_gap.py: with one full 32-byte sector per lane and instruction the row layout beats the transposed one for
  //  GEGLU-forward, 303 vs 323 us; LayerScale + residual stays faster transposed, 77 vs 83 us)
  const bool cols = env == 2 || (env == 0 && EPI == DB200_EPI_RESID) ||
                    (EPI == DB200_EPI_STORE && p.C_multicast != nullptr);   // measured, see profiles/
  if (cols) return launch_cfg_mode<BLOCK_N, EPI, A_MN, B_MN, true>(p, st);
  return launch_cfg_mode<BLOCK_N, EPI, A_MN, B_MN, false>(p, st);
}

template <int EPI, bool A_MN, bool B_MN>
int launch_bn(const db200_gemm_params& p, cudaStream_t st) {
  if constexpr (EPI == DB200_EPI_GEGLU) {
    return launch_cfg<256, EPI, A_MN, B_MN>(p, st);
  } else {
    // 128x256 tiles halve the smem traffic per MMA; fall back to 128x128 when N is small or when the wider tile would
    // leave most SMs idle
    const long long tiles256 = (long long)ceil_div(p.M, BLOCK_M) * ceil_div(p.N, 256);
    if (p.N >= 256 && (p.N % 256 == 0 || p.N > 1024) && tiles256 >= sm_count() / 2) return launch_cfg<256, EPI, A_MN, B_MN>(p, st);
    return launch_cfg<128, EPI, A_MN, B_MN>(p, st);
  }
}

bool device_is_sm100() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  if (cached[dev] == 0) {
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cached[dev] = (major == 10) ? 1 : -1;
  }
  return cached[dev] == 1;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

namespace {
// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, outer stride `ld` elements
int make_tensor_map_bf16_impl(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(DB200_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(DB200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu)", (int)r,
                                          (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return DB200_OK;
}

// rank-N (2..5) bf16 tensor map: dims[0] is the contiguous dimension; strides_elems[i] (i >= 1) is the pitch of dimension i in
// elements (need not be monotonic: the axial-column gather walks the token grid with the ROW of the image as the faster dim)
int make_tensor_map_bf16_nd_impl(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                                 const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(DB200_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  if (rank < 2 || rank > 5) return set_error(DB200_ERR_BAD_ARG, "tensor map rank %d", rank);
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
  for (int i = 1; i < rank; ++i) gstride[i - 1] = strides_elems[i] * 2;
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstride, bx, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(DB200_ERR_CUDA, "cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank, (int)r);
  return DB200_OK;
}

}  // namespace

namespace tc {
int make_tensor_map_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                         uint32_t box_outer) {
  return make_tensor_map_bf16_impl(map, base, inner, outer, ld, box_inner, box_outer);
}
int make_tensor_map_bf16_nd(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                            const uint32_t* box) {
  return make_tensor_map_bf16_nd_impl(map, base, rank, dims, strides_elems, box);
}
}  // namespace tc

bool gemm_tcgen05_supported(const db200_gemm_params& p, const char** why) {
  const char* w = nullptr;
  if (p.dtype != DB200_BF16) w = "operands are not bf16";
  else if (!device_is_sm100()) w = "device is not sm_100";
  else if (!al16(p.A) || !al16(p.B)) w = "A/B not 16-byte aligned";
  else if (p.lda % 8 || p.ldb % 8) w = "leading dimensions must be multiples of 8 elements";
  else if (p.K % 8 || p.N % 8) w = "K and N must be multiples of 8";
  else if (p.a_mn_major && p.M % 8) w = "M-major A needs M % 8 == 0";
  else if (p.a_mn_major && !p.b_mn_major && p.epilogue != DB200_EPI_STORE) w = "M-major A x K-major B only with the STORE epilogue";
  else {
    switch (p.epilogue) {
      case DB200_EPI_STORE:
        if (!al16(p.C) || p.ldc % 8 || (p.bias && !al16(p.bias))) w = "C/bias alignment";
        break;
      case DB200_EPI_QKV:
        if (p.a_mn_major || p.b_mn_major) w = "QKV epilogue is forward-only (K-major operands)";
        else if (p.dim_head % 8 || !al16(p.q) || !al16(p.k) || !al16(p.v) || (p.cos_t && (!al16(p.cos_t) || !al16(p.sin_t) || p.dim_head % 16)))
          w = "q/k/v/rotary alignment";
        break;
      case DB200_EPI_RESID:
        if (p.a_mn_major || p.b_mn_major) w = "RESID epilogue is forward-only";
        else if (!al16(p.out) || (p.resid && !al16(p.resid)) || (p.scale && !al16(p.scale)) || (p.bias && !al16(p.bias)) || (p.y_out && !al16(p.y_out)))
          w = "RESID tensor alignment";
        break;
      case DB200_EPI_GEGLU:
        if (p.a_mn_major || p.b_mn_major) w = "GEGLU epilogue is forward-only";
        else if (p.hidden % 128) w = "GEGLU needs hidden % 128 == 0";
        else if (!al16(p.h_out) || (p.u_out && !al16(p.u_out)) || (p.bias && !al16(p.bias))) w = "GEGLU tensor alignment";
        break;
      case DB200_EPI_GEGLU_BWD:
        if (p.a_mn_major) w = "GEGLU_BWD expects K-major A";
        else if (!al16(p.u_in) || !al16(p.du_out) || p.hidden % 8) w = "GEGLU_BWD tensor alignment";
        break;
      default: w = "unknown epilogue";
    }
  }
  if (why) *why = w ? w : "";
  return w == nullptr;
}

int gemm_tcgen05_launch(const db200_gemm_params& p, cudaStream_t st) {
  const bool a = p.a_mn_major != 0, b = p.b_mn_major != 0;
  switch (p.epilogue) {
    case DB200_EPI_STORE:
      if (!a && !b) return launch_bn<DB200_EPI_STORE, false, false>(p, st);
      if (!a && b) return launch_bn<DB200_EPI_STORE, false, true>(p, st);
      if (a && b) return launch_bn<DB200_EPI_STORE, true, true>(p, st);
      return launch_bn<DB200_EPI_STORE, true, false>(p, st);
    case DB200_EPI_QKV: return launch_bn<DB200_EPI_QKV, false, false>(p, st);
    case DB200_EPI_RESID: return launch_bn<DB200_EPI_RESID, false, false>(p, st);
    case DB200_EPI_GEGLU: return launch_bn<DB200_EPI_GEGLU, false, false>(p, st);
    case DB200_EPI_GEGLU_BWD:
      if (!b) return launch_bn<DB200_EPI_GEGLU_BWD, false, false>(p, st);
      return launch_bn<DB200_EPI_GEGLU_BWD, false, true>(p, st);
    default: break;
  }
  return set_error(DB200_ERR_BAD_ARG, "gemm_tcgen05: unsupported epilogue/major combination");
}

}  // namespace db200
