// gemm_simt.cuh -- fp32 CUDA-core GEMM with arbitrary strides and the shared epilogues.
//
// Two jobs: (1) the whole GEMM path of the fp32 validation mode (`precision=fp32`, the mode
// the rel-1e-5 parity tests run in; bf16 tensor-core math cannot reach that), and (2) in bf16
// mode, the handful of contractions that are too thin for a tensor-core tile (K or N = number
// of mixture components: prior_gmm, the y-columns of encoder_gmm layer 0, encoder_y's last
// layer and their gradients; < 0.5 % of the step's FLOPs).
//
//   C[m,n] = sum_k A(m,k) * B(k,n),  A(m,k) = A[m*sAm + k*sAk],  B(k,n) = B[k*sBk + n*sBn]
// Transposes are expressed through the strides.  gridDim.z splits K (epilogue must be atomic).
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace gmvae {

constexpr int SIMT_BM = 64, SIMT_BN = 64, SIMT_BK = 16, SIMT_THREADS = 256;

template <typename TA, typename TB, class Epi>
__global__ void __launch_bounds__(SIMT_THREADS)
gemm_simt_kernel(const TA* __restrict__ A, int64_t sAm, int64_t sAk, const TB* __restrict__ B, int64_t sBk,
                 int64_t sBn, int M, int N, int K, int k_per_split, Epi epi_in) {
  __shared__ float As[SIMT_BK][SIMT_BM + 4];
  __shared__ float Bs[SIMT_BK][SIMT_BN + 4];
  griddep_wait();
  griddep_launch();
  Epi epi = epi_in;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SIMT_BM, n0 = blockIdx.x * SIMT_BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += SIMT_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * SIMT_THREADS;
      int mm, kk;
      if (sAk == 1) { kk = idx % SIMT_BK; mm = idx / SIMT_BK; } else { mm = idx % SIMT_BM; kk = idx / SIMT_BM; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? to_f32<TA>(A[(int64_t)gm * sAm + (int64_t)gk * sAk]) : 0.f;
      int nn, kb;
      if (sBn == 1) { nn = idx % SIMT_BN; kb = idx / SIMT_BN; } else { kb = idx % SIMT_BK; nn = idx / SIMT_BK; }
      int gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < N && gkb < kend) ? to_f32<TB>(B[(int64_t)gkb * sBk + (int64_t)gn * sBn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SIMT_BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (kbeg < kend) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = m0 + ty * 4 + i;
      if (m < M && n < N) {
        const int nv = min(4, N - n);
        const EpiCtx ctx{nullptr, 0, nullptr, nullptr};
        auto pre = epi.template prefetch<4>(m, n, nv, true, ctx);
        epi.template row<4>(m, n, acc[i], nv, true, pre, ctx);
      }
    }
  }
  epi.finish_warp();
}

template <typename TA, typename TB, class Epi>
inline cudaError_t launch_gemm_simt(const TA* A, int64_t sAm, int64_t sAk, const TB* B, int64_t sBk, int64_t sBn,
                                    int M, int N, int K, int split_k, const Epi& epi, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  if (split_k < 1) split_k = 1;
  int k_per = (K + split_k - 1) / split_k;
  k_per = (k_per + SIMT_BK - 1) / SIMT_BK * SIMT_BK;
  split_k = (K + k_per - 1) / k_per;
  dim3 grid((N + SIMT_BN - 1) / SIMT_BN, (M + SIMT_BM - 1) / SIMT_BM, split_k);
  return launch_k(gemm_simt_kernel<TA, TB, Epi>, grid, dim3(SIMT_THREADS), 0, st, true, A, sAm, sAk, B, sBk, sBn, M, N, K, k_per, epi);
}

}  // namespace gmvae
