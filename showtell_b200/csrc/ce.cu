// Row-wise kernels over materialised logits (fp32 mode and the drop-in forward()):
//   nn.CrossEntropyLoss()          main.py:94,149      -> st_ce_fwd_bwd
//   result_state.max(1)[1]         rnn.py:51           -> st_argmax_rows
//   result_state.topk(k, dim=1)    rnn.py:63,90-91     -> st_topk_rows
// One CTA per row; the row is streamed from HBM once per pass (HBM-bound).
#include <cfloat>

#include "common.cuh"

namespace st {
namespace {

constexpr int NT = 256;

// Block-wide reductions: warp shuffle, then every thread folds the NT/32 per-warp partials.
__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) r += red[i];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(NT) ce_kernel(const float* logits, int ld,
                                                const int64_t* __restrict__ target, int V,
                                                float* __restrict__ loss_sum, float* __restrict__ lse_out,
                                                float* dlogits, float grad_scale) {
  __shared__ float red[NT / 32];
  const int n = blockIdx.x;
  const float* row = logits + (size_t)n * ld;
  float m = -FLT_MAX;
  for (int v = threadIdx.x; v < V; v += NT) m = fmaxf(m, row[v]);
  m = block_reduce_max(m, red);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += NT) s += expf(row[v] - m);
  s = block_reduce_sum(s, red);
  const float lse = m + logf(s);
  const int64_t tgt = target[n];
  if (threadIdx.x == 0) {
    if (lse_out) lse_out[n] = lse;
    atomicAdd(loss_sum, lse - row[tgt]);
  }
  if (dlogits) {
    __syncthreads();  // row[tgt] read above before a possible in-place overwrite
    float* drow = dlogits + (size_t)n * ld;
    for (int v = threadIdx.x; v < V; v += NT) {
      float p = expf(row[v] - lse);
      drow[v] = (p - (v == tgt ? 1.f : 0.f)) * grad_scale;
    }
  }
}

// (value desc, index asc) ordering packed so that a single max picks the winner.
struct VI {
  float v;
  int i;
};
__device__ __forceinline__ VI vi_better(VI a, VI b) {
  return (a.v > b.v || (a.v == b.v && a.i < b.i)) ? a : b;
}
__device__ __forceinline__ VI block_reduce_vi(VI x, VI* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    VI y{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o)};
    x = vi_better(x, y);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  VI r = red[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) r = vi_better(r, red[i]);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(NT) argmax_kernel(const float* __restrict__ X, int ld, int cols,
                                                    int64_t* __restrict__ idx, int idx_stride) {
  __shared__ VI red[NT / 32];
  const float* row = X + (size_t)blockIdx.x * ld;
  VI best{-FLT_MAX, 0x7fffffff};
  for (int c = threadIdx.x; c < cols; c += NT) best = vi_better(best, VI{row[c], c});
  best = block_reduce_vi(best, red);
  if (threadIdx.x == 0) idx[(size_t)blockIdx.x * idx_stride] = best.i;
}

// K rounds of block arg-max with exclusion of earlier winners (K <= 32, cols ~ 1e4: the row stays
// in L1/L2 after the first pass).
__global__ void __launch_bounds__(NT) topk_kernel(const float* __restrict__ X, int ld, int cols, int K,
                                                  float* __restrict__ val, int32_t* __restrict__ idx,
                                                  int out_stride) {
  __shared__ VI red[NT / 32];
  __shared__ int taken[32];
  const float* row = X + (size_t)blockIdx.x * ld;
  for (int k = 0; k < K; ++k) {
    VI best{-FLT_MAX, 0x7fffffff};
    for (int c = threadIdx.x; c < cols; c += NT) {
      bool skip = false;
      for (int j = 0; j < k; ++j) skip |= (taken[j] == c);
      if (!skip) best = vi_better(best, VI{row[c], c});
    }
    best = block_reduce_vi(best, red);
    if (threadIdx.x == 0) {
      taken[k] = best.i;
      val[(size_t)blockIdx.x * out_stride + k] = best.v;
      idx[(size_t)blockIdx.x * out_stride + k] = best.i;
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace st

extern "C" {

int st_ce_fwd_bwd(const float* logits, int ld, const int64_t* target, int N, int V, float* loss_sum,
                  float* lse, float* dlogits, float grad_scale, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(logits && target && loss_sum, ST_ERR_NULL, "st_ce_fwd_bwd: NULL pointer");
  ST_REQUIRE(N >= 1 && V >= 1 && ld >= V, ST_ERR_BAD_SHAPE, "st_ce_fwd_bwd: N=%d V=%d ld=%d", N, V, ld);
  cudaStream_t s = as_stream(stream);
  ST_CUDA_TRY(cudaMemsetAsync(loss_sum, 0, sizeof(float), s));
  ce_kernel<<<N, NT, 0, s>>>(logits, ld, target, V, loss_sum, lse, dlogits, grad_scale);
  ST_LAUNCH_TRY("ce_kernel");
  return ST_OK;
}

int st_argmax_rows(const float* X, int ld, int rows, int cols, int64_t* idx, int idx_stride,
                   st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(X && idx, ST_ERR_NULL, "st_argmax_rows: NULL pointer");
  ST_REQUIRE(rows >= 1 && cols >= 1 && ld >= cols && idx_stride >= 1, ST_ERR_BAD_SHAPE,
             "st_argmax_rows: bad shape");
  argmax_kernel<<<rows, NT, 0, as_stream(stream)>>>(X, ld, cols, idx, idx_stride);
  ST_LAUNCH_TRY("argmax_kernel");
  return ST_OK;
}

int st_topk_rows(const float* X, int ld, int rows, int cols, int K, float* val, int32_t* idx,
                 int out_stride, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(X && val && idx, ST_ERR_NULL, "st_topk_rows: NULL pointer");
  ST_REQUIRE(rows >= 1 && cols >= 1 && ld >= cols, ST_ERR_BAD_SHAPE, "st_topk_rows: bad shape");
  ST_REQUIRE(K >= 1 && K <= 32 && K <= cols && out_stride >= K, ST_ERR_BAD_SHAPE,
             "st_topk_rows: K=%d outside [1, min(32, cols)] or out_stride=%d < K", K, out_stride);
  topk_kernel<<<rows, NT, 0, as_stream(stream)>>>(X, ld, cols, K, val, idx, out_stride);
  ST_LAUNCH_TRY("topk_kernel");
  return ST_OK;
}

}  // extern "C"
