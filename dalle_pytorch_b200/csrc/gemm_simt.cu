// CUDA-core GEMM with the fused epilogues of epilogue.cuh.
//
// Role: (1) the fp32 PARITY path — FFMA accumulation in fp32 so that logits meet rtol 1e-3 / atol 1e-5
// against the reference's CPU fp32 arithmetic (bf16 and plain TF32 cannot); (2) the
// always-available fallback / cross-check for the tcgen05 kernel (gemm_tcgen05.cu) in bf16 mode.
// 64x64x16 tiles, 256 threads, 4x4 register blocking; every epilogue sees adjacent column pairs.
#include "common.cuh"
#include "epilogue.cuh"

namespace db200 {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int LDS = BM + 4;

template <typename T>
struct TileLoader {
  // Loads a (64 x 16) operand tile into smem as S[k][mn] (fp32).
  // mn_major == 0: src(mn, k) = src[mn*ld + k] ;  == 1: src(mn, k) = src[k*ld + mn]
  // `split`: GEGLU column remap — tile-local mn < 32 -> mn0/2 + mn ; mn >= 32 -> hidden + mn0/2 + (mn-32)
  __device__ static void load(float (*S)[LDS], const T* __restrict__ src, long long ld, int mn_major,
                              int mn0, int k0, int MN, int K, int split_hidden) {
    const int t = threadIdx.x;
    if (!mn_major) {
      // thread -> (mn = t/4, k = (t%4)*4 .. +3)
      const int mn_l = t >> 2, kq = (t & 3) << 2;
      int mn = mn0 + mn_l;
      if (split_hidden) mn = (mn_l < 32) ? (mn0 >> 1) + mn_l : split_hidden + (mn0 >> 1) + (mn_l - 32);
      const bool row_ok = mn < MN;
      const T* p = src + (long long)mn * ld + k0 + kq;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        if (row_ok && (k0 + kq + j) < K) v = to_f32(p[j]);
        S[kq + j][mn_l] = v;
      }
    } else {
      // thread -> (k = t/16, mn = (t%16)*4 .. +3)
      const int k_l = t >> 4, mq = (t & 15) << 2;
      const bool k_ok = (k0 + k_l) < K;
      const T* p = src + (long long)(k0 + k_l) * ld;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int mn_l = mq + j;
        int mn = mn0 + mn_l;
        if (split_hidden) mn = (mn_l < 32) ? (mn0 >> 1) + mn_l : split_hidden + (mn0 >> 1) + (mn_l - 32);
        float v = 0.f;
        if (k_ok && mn < MN) v = to_f32(p[mn]);
        S[k_l][mn_l] = v;
      }
    }
  }
};

template <typename T, int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, long long lda, int a_mn,
                                                        const T* __restrict__ B, long long ldb, int b_mn,
                                                        int M, int N, int K, EpiArgs e) {
  __shared__ float As[2][BK][LDS];
  __shared__ float Bs[2][BK][LDS];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int split_hidden = (EPI == DB200_EPI_GEGLU) ? e.hidden : 0;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  TileLoader<T>::load(As[0], A, lda, a_mn, m0, 0, M, K, 0);
  TileLoader<T>::load(Bs[0], B, ldb, b_mn, n0, 0, N, K, split_hidden);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      TileLoader<T>::load(As[cur ^ 1], A, lda, a_mn, m0, (kt + 1) * BK, M, K, 0);
      TileLoader<T>::load(Bs[cur ^ 1], B, ldb, b_mn, n0, (kt + 1) * BK, N, K, split_hidden);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float2 b0 = *reinterpret_cast<const float2*>(&Bs[cur][k][tx * 2]);
      const float2 b1 = *reinterpret_cast<const float2*>(&Bs[cur][k][32 + tx * 2]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    if constexpr (EPI == DB200_EPI_GEGLU) {
      const int j = (n0 >> 1) + tx * 2;           // hidden index of the pair
      if (j < e.hidden) epi_geglu_pair<T>(e, m, j, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
      const int na = n0 + tx * 2, nb = n0 + 32 + tx * 2;
      if (na < N) epi_pair<EPI, T>(e, m, na, acc[i][0], acc[i][1]);
      if (nb < N) epi_pair<EPI, T>(e, m, nb, acc[i][2], acc[i][3]);
    }
  }
}

template <typename T>
int launch_simt_t(const db200_gemm_params& p, cudaStream_t st) {
  const EpiArgs e = make_epi_args(p);
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
  dim3 block(256);
  const T* A = reinterpret_cast<const T*>(p.A);
  const T* B = reinterpret_cast<const T*>(p.B);
#define DB200_LAUNCH_SIMT(EPI)                                                                             \
  gemm_simt_kernel<T, EPI><<<grid, block, 0, st>>>(A, p.lda, p.a_mn_major, B, p.ldb, p.b_mn_major, p.M, p.N, \
                                                    p.K, e)
  switch (p.epilogue) {
    case DB200_EPI_STORE: DB200_LAUNCH_SIMT(DB200_EPI_STORE); break;
    case DB200_EPI_QKV: DB200_LAUNCH_SIMT(DB200_EPI_QKV); break;
    case DB200_EPI_RESID: DB200_LAUNCH_SIMT(DB200_EPI_RESID); break;
    case DB200_EPI_GEGLU: DB200_LAUNCH_SIMT(DB200_EPI_GEGLU); break;
    case DB200_EPI_GEGLU_BWD: DB200_LAUNCH_SIMT(DB200_EPI_GEGLU_BWD); break;
    default: return set_error(DB200_ERR_BAD_ARG, "gemm: unknown epilogue %d", p.epilogue);
  }
#undef DB200_LAUNCH_SIMT
  DB200_LAUNCH_OK("gemm_simt_kernel");
  return DB200_OK;
}

}  // namespace

int gemm_simt_launch(const db200_gemm_params& p, cudaStream_t st) {
  if (p.dtype == DB200_F32) return launch_simt_t<float>(p, st);
  return launch_simt_t<__nv_bfloat16>(p, st);
}

}  // namespace db200
