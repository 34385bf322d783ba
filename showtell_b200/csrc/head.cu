// Encoder head of the base models (cnn.py:37-38,49): BatchNorm1d(embed_dim, momentum=0.01) applied to
// Linear(2048, embed_dim)(pooled ResNet features) -- the step that produces the decoder's `cnn_feature`, and the
// first consumer of its gradient (main.py:96 trains exactly these two layers).  The Linear product runs on the
// library's GEMMs; these kernels are the batch normalisation over the B rows of the (B, E) activations.
// One CTA per 32 columns: 8 warps stride the rows, a warp reads 32 consecutive floats of a row (coalesced),
// column statistics meet in shared memory.  Bytes: forward reads Y twice (L2-resident) and writes out once.
#include "common.cuh"

namespace st {
namespace {

constexpr int BN_COLS = 32, BN_WARPS = 8;

// sum over the rows of f(row) for this thread's column, combined over the CTA's warps -> every thread of the
// column gets the total
template <typename F>
__device__ __forceinline__ float col_reduce(F f, int rows, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s = 0.f;
  for (int r = warp; r < rows; r += BN_WARPS) s += f(r);
  __syncthreads();                                 // red may still be read from the previous reduction
  red[warp * BN_COLS + lane] = s;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < BN_WARPS; ++w) t += red[w * BN_COLS + lane];
  return t;
}

// training: batch statistics (biased variance for the normalisation, unbiased for running_var), saves
// mean / invstd for the backward pass.  eval (use_running != 0): running statistics, nothing updated.
__global__ void __launch_bounds__(BN_COLS * BN_WARPS)
bn1d_fwd_kernel(const float* __restrict__ Y, int ldy, int B, int E, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, float momentum, int use_running,
                float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                float* __restrict__ save_invstd, float* __restrict__ out, int ldo) {
  __shared__ float red[BN_WARPS * BN_COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = blockIdx.x * BN_COLS + lane;
  const bool ok = c < E;
  float mean, invstd;
  if (use_running) {
    mean = ok ? running_mean[c] : 0.f;
    invstd = ok ? rsqrtf(running_var[c] + eps) : 0.f;
  } else {
    mean = col_reduce([&](int r) { return ok ? Y[(size_t)r * ldy + c] : 0.f; }, B, red) / (float)B;
    const float ss = col_reduce([&](int r) { const float d = ok ? Y[(size_t)r * ldy + c] - mean : 0.f; return d * d; }, B, red);
    const float var = ss / (float)B;
    invstd = rsqrtf(var + eps);
    if (ok && warp == 0) {
      if (save_mean) { save_mean[c] = mean; save_invstd[c] = invstd; }
      if (running_mean) {
        running_mean[c] = fmaf(momentum, mean - running_mean[c], running_mean[c]);
        const float unbiased = B > 1 ? ss / (float)(B - 1) : var;
        running_var[c] = fmaf(momentum, unbiased - running_var[c], running_var[c]);
      }
    }
  }
  if (!ok) return;
  const float g = gamma[c] * invstd, b = beta[c];
  for (int r = warp; r < B; r += BN_WARPS) out[(size_t)r * ldo + c] = fmaf(Y[(size_t)r * ldy + c] - mean, g, b);
}

// dgamma = sum dOut * xhat, dbeta = sum dOut, dY = gamma * invstd * (dOut - dbeta / B - xhat * dgamma / B)
__global__ void __launch_bounds__(BN_COLS * BN_WARPS)
bn1d_bwd_kernel(const float* __restrict__ Y, int ldy, const float* __restrict__ dOut, int ldd, int B, int E,
                const float* __restrict__ gamma, const float* __restrict__ save_mean,
                const float* __restrict__ save_invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                float* __restrict__ dY, int ldg) {
  __shared__ float red[BN_WARPS * BN_COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = blockIdx.x * BN_COLS + lane;
  const bool ok = c < E;
  const float mean = ok ? save_mean[c] : 0.f, invstd = ok ? save_invstd[c] : 0.f;
  const float db = col_reduce([&](int r) { return ok ? dOut[(size_t)r * ldd + c] : 0.f; }, B, red);
  const float dg = col_reduce([&](int r) {
    return ok ? dOut[(size_t)r * ldd + c] * (Y[(size_t)r * ldy + c] - mean) * invstd : 0.f; }, B, red);
  if (!ok) return;
  if (warp == 0) { dgamma[c] = dg; dbeta[c] = db; }
  const float k = gamma[c] * invstd, ib = 1.f / (float)B;
  for (int r = warp; r < B; r += BN_WARPS) {
    const float xh = (Y[(size_t)r * ldy + c] - mean) * invstd;
    dY[(size_t)r * ldg + c] = k * (dOut[(size_t)r * ldd + c] - db * ib - xh * dg * ib);
  }
}

}  // namespace
}  // namespace st

extern "C" {

int st_bn1d_fwd(const float* Y, int ldy, int B, int E, const float* gamma, const float* beta, float eps, float momentum,
                int use_running_stats, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                float* out, int ldo, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(Y && gamma && beta && out, ST_ERR_NULL, "st_bn1d_fwd: NULL pointer");
  ST_REQUIRE(B >= 1 && E >= 1 && ldy >= E && ldo >= E, ST_ERR_BAD_SHAPE, "st_bn1d_fwd: B=%d E=%d ldy=%d ldo=%d", B, E, ldy, ldo);
  ST_REQUIRE(!use_running_stats || (running_mean && running_var), ST_ERR_NULL, "st_bn1d_fwd: eval mode needs running statistics");
  ST_REQUIRE((running_mean == nullptr) == (running_var == nullptr) && (save_mean == nullptr) == (save_invstd == nullptr),
             ST_ERR_NULL, "st_bn1d_fwd: statistics buffers come in pairs");
  // nn.BatchNorm1d raises for a single row in training mode ("Expected more than 1 value per channel")
  ST_REQUIRE(use_running_stats || B > 1, ST_ERR_BAD_SHAPE, "st_bn1d_fwd: training mode needs more than 1 row");
  bn1d_fwd_kernel<<<(E + BN_COLS - 1) / BN_COLS, BN_COLS * BN_WARPS, 0, as_stream(stream)>>>(
      Y, ldy, B, E, gamma, beta, eps, momentum, use_running_stats, running_mean, running_var, save_mean, save_invstd, out, ldo);
  ST_LAUNCH_TRY("bn1d_fwd_kernel");
  return ST_OK;
}

int st_bn1d_bwd(const float* Y, int ldy, const float* dOut, int ldd, int B, int E, const float* gamma,
                const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* dY, int ldg,
                st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(Y && dOut && gamma && save_mean && save_invstd && dgamma && dbeta && dY, ST_ERR_NULL, "st_bn1d_bwd: NULL pointer");
  ST_REQUIRE(B >= 1 && E >= 1 && ldy >= E && ldd >= E && ldg >= E, ST_ERR_BAD_SHAPE, "st_bn1d_bwd: B=%d E=%d", B, E);
  bn1d_bwd_kernel<<<(E + BN_COLS - 1) / BN_COLS, BN_COLS * BN_WARPS, 0, as_stream(stream)>>>(
      Y, ldy, dOut, ldd, B, E, gamma, save_mean, save_invstd, dgamma, dbeta, dY, ldg);
  ST_LAUNCH_TRY("bn1d_bwd_kernel");
  return ST_OK;
}

}  // extern "C"
