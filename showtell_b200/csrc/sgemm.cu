// fp32 GEMM on CUDA cores (FFMA) -- the arithmetic of fp32 mode, where parity with the fp32
// reference (1e-4 relative) rules out bf16/tf32 tensor-core inputs.  bf16 mode uses gemm_tc.cu.
//
//   C[M,N] = alpha * op(A) op(B) + beta * C + bias[N]
//
// 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread (two 4-wide strips per
// dimension so shared-memory reads are conflict-free float4), register-staged double buffering
// of the global loads.  Arbitrary M, N, K and leading dimensions; transposed operands are read
// with the thread mapping that keeps the contiguous dimension on consecutive threads.
#include "common.cuh"

namespace st {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NT = 256;

// Loads one BMxBK (or BNxBK) operand tile into registers: 8 elements per thread.
// KCONTIG: memory is (rows, K) with K contiguous   -> thread = (k = t%16, r = t/16 + 16 i)
// else:    memory is (K, rows) with rows contiguous -> thread = (r = t%128, k = t/128 + 2 i)
template <bool KCONTIG>
__device__ __forceinline__ void load_tile(float (&reg)[8], const float* __restrict__ P, int ld,
                                          int row0, int k0, int nrows, int K, int tid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r, k;
    if (KCONTIG) {
      k = tid & 15;
      r = (tid >> 4) + 16 * i;
    } else {
      r = tid & 127;
      k = (tid >> 7) + 2 * i;
    }
    int gr = row0 + r, gk = k0 + k;
    float v = 0.f;
    if (gr < nrows && gk < K) v = KCONTIG ? P[(size_t)gr * ld + gk] : P[(size_t)gk * ld + gr];
    reg[i] = v;
  }
}

template <bool KCONTIG>
__device__ __forceinline__ void store_tile(const float (&reg)[8], float (*S)[BM + PAD], int tid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r, k;
    if (KCONTIG) {
      k = tid & 15;
      r = (tid >> 4) + 16 * i;
    } else {
      r = tid & 127;
      k = (tid >> 7) + 2 * i;
    }
    S[k][r] = reg[i];
  }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(NT) sgemm_kernel(int M, int N, int K, float alpha,
                                                   const float* __restrict__ A, int lda,
                                                   const float* __restrict__ B, int ldb, float beta,
                                                   float* __restrict__ C, int ldc,
                                                   const float* __restrict__ bias) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  // A as stored: TA=0 -> (M,K) k-contiguous; TA=1 -> (K,M) m-contiguous
  // B as stored: TB=1 -> (N,K) k-contiguous; TB=0 -> (K,N) n-contiguous
  load_tile<!TA>(ra, A, lda, m0, 0, M, K, tid);
  load_tile<TB>(rb, B, ldb, n0, 0, N, K, tid);
  store_tile<!TA>(ra, As[0], tid);
  store_tile<TB>(rb, Bs[0], tid);
  __syncthreads();

  const int nk = (K + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_tile<!TA>(ra, A, lda, m0, (kt + 1) * BK, M, K, tid);
      load_tile<TB>(rb, B, ldb, n0, (kt + 1) * BK, N, K, tid);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile<!TA>(ra, As[cur ^ 1], tid);
      store_tile<TB>(rb, Bs[cur ^ 1], tid);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[n];
      if (beta != 0.f) v += beta * C[(size_t)m * ldc + n];
      C[(size_t)m * ldc + n] = v;
    }
  }
}

// Small-problem variant (per-step products of the attention / decode loops, M ~ batch): 32x32x32
// tiles so that a 128 x 512 output still spreads over 64 CTAs; 2x2 outputs per thread.
template <bool TA, bool TB>
__global__ void __launch_bounds__(NT) sgemm_small_kernel(int M, int N, int K, float alpha,
                                                         const float* __restrict__ A, int lda,
                                                         const float* __restrict__ B, int ldb, float beta,
                                                         float* __restrict__ C, int ldc,
                                                         const float* __restrict__ bias) {
  __shared__ float As[32][33], Bs[32][33];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      {  // A element (m, k): TA=0 memory (M,K) k-contiguous; TA=1 memory (K,M) m-contiguous
        const int k = TA ? (tid >> 5) + 8 * i : (tid & 31), m = TA ? (tid & 31) : (tid >> 5) + 8 * i;
        const int gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < M && gk < K) ? (TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk]) : 0.f;
      }
      {  // B element (n, k): TB=1 memory (N,K) k-contiguous; TB=0 memory (K,N) n-contiguous
        const int k = TB ? (tid & 31) : (tid >> 5) + 8 * i, n = TB ? (tid >> 5) + 8 * i : (tid & 31);
        const int gn = n0 + n, gk = k0 + k;
        Bs[k][n] = (gn < N && gk < K) ? (TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1], b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = m0 + ty * 2 + i, n = n0 + tx * 2 + j;
      if (m < M && n < N) {
        float v = alpha * acc[i][j];
        if (bias) v += bias[n];
        if (beta != 0.f) v += beta * C[(size_t)m * ldc + n];
        C[(size_t)m * ldc + n] = v;
      }
    }
}

}  // namespace
}  // namespace st

extern "C" int st_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A,
                        int lda, const float* B, int ldb, float beta, float* C, int ldc,
                        const float* bias, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(M >= 0 && N >= 0 && K >= 0, ST_ERR_BAD_SHAPE, "st_sgemm: negative dimension");
  if (M == 0 || N == 0) return ST_OK;
  ST_REQUIRE(A && B && C, ST_ERR_NULL, "st_sgemm: NULL operand");
  ST_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, ST_ERR_BAD_SHAPE,
             "st_sgemm: leading dimension too small (lda=%d ldb=%d ldc=%d)", lda, ldb, ldc);
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  ST_REQUIRE(grid.y <= 65535, ST_ERR_BAD_SHAPE, "st_sgemm: M=%d too large", M);
  cudaStream_t s = as_stream(stream);
  if (grid.x * grid.y < 48) {  // too few 128x128 tiles to fill the GPU: small-tile kernel
    dim3 g2((N + 31) / 32, (M + 31) / 32);
    if (!transA && !transB)
      sgemm_small_kernel<false, false><<<g2, NT, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
    else if (!transA && transB)
      sgemm_small_kernel<false, true><<<g2, NT, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
    else if (transA && !transB)
      sgemm_small_kernel<true, false><<<g2, NT, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
    else
      sgemm_small_kernel<true, true><<<g2, NT, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
    ST_LAUNCH_TRY("sgemm_small_kernel");
    return ST_OK;
  }
  if (!transA && !transB)
    sgemm_kernel<false, false><<<grid, NT, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  else if (!transA && transB)
    sgemm_kernel<false, true><<<grid, NT, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  else if (transA && !transB)
    sgemm_kernel<true, false><<<grid, NT, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  else
    sgemm_kernel<true, true><<<grid, NT, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
  ST_LAUNCH_TRY("sgemm_kernel");
  return ST_OK;
}
