// Soft-attention kernels of RNN_Attn (Attention/rnn_attn.py:8-31,60-76; rnn_attn_LSTM.py).
//
// The reference recomputes encoder_att(f) (a 52.6 GFLOP GEMM at B=128, P=196) inside every time
// step.  Here everything that does not depend on the recurrent state is hoisted out of the loop:
//   F    (B,P,C)  = f^T                      relayout of the channels-first grid (st_attn_relayout)
//   att1 (B,P,A)  = F W_enc^T + b_enc        one GEMM                                (rnn_attn.py:23)
//   Fe   (B,P,E)  = F W_embed^T              one GEMM: because sum_p alpha_p = 1,
//                    embed(sum_p alpha_p F_p) = sum_p alpha_p Fe_p + b_embed, so the context sum
//                    runs directly in the 512-wide embedding space (2.5x fewer bytes per step
//                    than in the 2048-wide feature space)                        (rnn_attn.py:29,70)
// Per step ONE fused kernel (st_attn_step_fwd), one CTA per batch row, HBM/L2-bound:
//   e_p = w_f . act(att1[b,p,:] + att2[b,:]) + b_f ; alpha = softmax_P(e) ; ctx_e = sum_p alpha_p Fe_p + b_embed
// with att2 = decoder_att(h) from a small GEMM.  It also writes alpha into alphas[b,t,:] and
// accumulates S[b,p] = sum_t alpha (the doubly-stochastic penalty, main_attn.py:131).
// Backward per step (st_attn_step_bwd): d alpha_p = <d ctx_e, Fe_p> + d pen_p, softmax backward,
// d att2 = sum_p de_p w_f act'(s_p).  The parts of the backward that do not feed the recurrence are
// hoisted again into single passes after the loop: d att1 and d w_f (st_attn_hoist_bwd) and the
// feature-space contexts needed for d W_embed (st_attn_ctx_all).
//
// act = LeakyReLU(0.2) as in the reference (rnn_attn.py:18); tanh is offered as an option.
#include <cuda_bf16.h>

#include <cfloat>

#include "attn_stream.cuh"
#include "common.cuh"

namespace st {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <int ACT>
__device__ __forceinline__ float act_f(float s) { return ACT == 0 ? (s > 0.f ? s : 0.2f * s) : tanhf(s); }
template <int ACT>
__device__ __forceinline__ float act_d(float s) {
  if (ACT == 0) return s > 0.f ? 1.f : 0.2f;
  const float t = tanhf(s);
  return 1.f - t * t;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) r += red[i];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}

// f (B,C,P) fp32 channels-first (cnn_attn.py:49) -> F (B,P,C) [T], FT (C, ldft) [T] with column b*P+p,
// mean_f (B,C) fp32 (rnn_attn.py:62 `cnn_feature.mean(dim=2)`).  One CTA per (32-channel slab, image).
template <typename T>
__global__ void relayout_kernel(const float* __restrict__ f, int C, int P, T* __restrict__ F, T* __restrict__ FT,
                                int ldft, float* __restrict__ mean_f) {
  __shared__ float tile[32][33];
  const int b = blockIdx.y, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32
  float msum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int p0 = 0; p0 < P; p0 += 32) {
    for (int i = ty; i < 32; i += 8) {  // read (c, p): p contiguous
      const int c = c0 + i, p = p0 + tx;
      float v = (c < C && p < P) ? f[((size_t)b * C + c) * P + p] : 0.f;
      tile[i][tx] = v;
      msum[i >> 3] += v;
      if (FT && c < C && p < P) stf(FT + (size_t)c * ldft + (size_t)b * P + p, v);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {  // write (p, c): c contiguous
      const int p = p0 + i, c = c0 + tx;
      if (p < P && c < C) stf(F + ((size_t)b * P + p) * C + c, tile[tx][i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float s = warp_sum(msum[k]);
    const int c = c0 + ty + 8 * k;
    if (tx == 0 && c < C) mean_f[(size_t)b * C + c] = s / (float)P;
  }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(NT)
attn_step_fwd_kernel(int P, int A, int E, const T* __restrict__ att1, const T* __restrict__ Fe,
                     const float* __restrict__ att2, const float* __restrict__ wf, const float* __restrict__ bfp,
                     const float* __restrict__ b_embed, float* __restrict__ alphas, int alpha_stride,
                     float* __restrict__ S, float* __restrict__ ctx_out, int ld_ctx,
                     __nv_bfloat16* __restrict__ ctx_bf16, int ld_ctx_bf16) {
  extern __shared__ float sm[];
  float* s_att2 = sm;            // [A]
  float* s_wf = sm + A;          // [A]
  float* s_e = sm + 2 * A;       // [P]
  __shared__ float red[NT / 32];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float bf = bfp[0];
  for (int a = tid; a < A; a += NT) {
    s_att2[a] = att2[(size_t)b * A + a];
    s_wf[a] = wf[a];
  }
  __syncthreads();
  // scores: one warp per grid location, lanes strided over the attention dimension
  const T* a1 = att1 + (size_t)b * P * A;
  for (int p = warp; p < P; p += NT / 32) {
    const T* row = a1 + (size_t)p * A;
    float acc = 0.f;
    for (int a = lane; a < A; a += 32) acc = fmaf(s_wf[a], act_f<ACT>(ldf(row + a) + s_att2[a]), acc);
    acc = warp_sum(acc);
    if (lane == 0) s_e[p] = acc + bf;
  }
  __syncthreads();
  // softmax over the P locations (nn.Softmax(dim=1), rnn_attn.py:26)
  float m = -FLT_MAX;
  for (int p = tid; p < P; p += NT) m = fmaxf(m, s_e[p]);
  m = block_max(m, red);
  float s = 0.f;
  for (int p = tid; p < P; p += NT) {
    const float w = expf(s_e[p] - m);
    s_e[p] = w;
    s += w;
  }
  s = block_sum(s, red);
  const float inv = 1.f / s;
  for (int p = tid; p < P; p += NT) {
    const float al = s_e[p] * inv;
    s_e[p] = al;
    alphas[(size_t)b * alpha_stride + p] = al;
    S[(size_t)b * P + p] += al;
  }
  __syncthreads();
  // context in embedding space: ctx_e = sum_p alpha_p Fe[b,p,:] + b_embed
  const T* fe = Fe + (size_t)b * P * E;
  for (int e = tid; e < E; e += NT) {
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc = fmaf(s_e[p], ldf(fe + (size_t)p * E + e), acc);
    ctx_out[(size_t)b * ld_ctx + e] = acc + b_embed[e];
    if (ctx_bf16) ctx_bf16[(size_t)b * ld_ctx_bf16 + e] = __float2bfloat16(acc + b_embed[e]);
  }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(NT)
attn_step_bwd_kernel(int P, int A, int E, const T* __restrict__ att1, const T* __restrict__ Fe,
                     const float* __restrict__ att2, const float* __restrict__ wf,
                     const float* __restrict__ alphas, int alpha_stride, const float* __restrict__ dal,
                     int dal_stride, const float* __restrict__ dctx, int ld_dctx, float* __restrict__ de_out,
                     float* __restrict__ datt2, __nv_bfloat16* __restrict__ datt2_bf16) {
  extern __shared__ float sm[];
  float* s_att2 = sm;             // [A]
  float* s_wf = sm + A;           // [A]
  float* s_dctx = sm + 2 * A;     // [E]
  float* s_de = sm + 2 * A + E;   // [P]
  __shared__ float red[NT / 32];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int a = tid; a < A; a += NT) {
    s_att2[a] = att2[(size_t)b * A + a];
    s_wf[a] = wf[a];
  }
  for (int e = tid; e < E; e += NT) s_dctx[e] = dctx[(size_t)b * ld_dctx + e];
  __syncthreads();
  // d alpha_p = <d ctx_e, Fe_p> + d pen_p
  const T* fe = Fe + (size_t)b * P * E;
  for (int p = warp; p < P; p += NT / 32) {
    float acc = 0.f;
    for (int e = lane; e < E; e += 32) acc = fmaf(s_dctx[e], ldf(fe + (size_t)p * E + e), acc);
    acc = warp_sum(acc);
    if (lane == 0) s_de[p] = acc + (dal ? dal[(size_t)b * dal_stride + p] : 0.f);
  }
  __syncthreads();
  // softmax backward: de = alpha * (d alpha - sum_p alpha_p d alpha_p)
  float dot = 0.f;
  for (int p = tid; p < P; p += NT) dot += alphas[(size_t)b * alpha_stride + p] * s_de[p];
  dot = block_sum(dot, red);
  for (int p = tid; p < P; p += NT) {
    const float de = alphas[(size_t)b * alpha_stride + p] * (s_de[p] - dot);
    s_de[p] = de;
    de_out[(size_t)b * P + p] = de;
  }
  __syncthreads();
  // d att2[a] = sum_p de_p w_f[a] act'(att1[b,p,a] + att2[b,a])
  const T* a1 = att1 + (size_t)b * P * A;
  for (int a = tid; a < A; a += NT) {
    float acc = 0.f;
    const float a2 = s_att2[a];
    for (int p = 0; p < P; ++p) acc = fmaf(s_de[p], act_d<ACT>(ldf(a1 + (size_t)p * A + a) + a2), acc);
    datt2[(size_t)b * A + a] = acc * s_wf[a];
    if (datt2_bf16) datt2_bf16[(size_t)b * A + a] = __float2bfloat16(acc * s_wf[a]);
  }
}

// After the time loop: d att1[b,p,a] = w_f[a] sum_t de[t,b,p] act'(att1[b,p,a] + att2[t,b,a]) and
// d w_f[a] += sum_{t,b,p} de[t,b,p] act(att1[b,p,a] + att2[t,b,a]).  de (N,P), att2 (N,A) packed.
// One CTA per (image, 8 grid locations); threads over A.
template <typename T, typename TO, int ACT>
__global__ void __launch_bounds__(NT)
attn_hoist_bwd_kernel(const __grid_constant__ StepTable tab, int P, int A, const T* __restrict__ att1,
                      const float* __restrict__ att2, const float* __restrict__ de, const float* __restrict__ wf,
                      TO* __restrict__ datt1, TO* __restrict__ datt1T, int ldt, float* __restrict__ dwf) {
  const int b = blockIdx.y, p0 = blockIdx.x * 8;
  int len = 0;
  while (len < tab.nsteps && tab.bs[len] > b) ++len;  // steps in which row b is live
  for (int a = threadIdx.x; a < A; a += NT) {
    float w = wf[a], dw = 0.f;
    for (int pi = 0; pi < 8; ++pi) {
      const int p = p0 + pi;
      if (p >= P) break;
      const float s1 = ldf(att1 + ((size_t)b * P + p) * A + a);
      float acc = 0.f;
      for (int t = 0; t < len; ++t) {
        const size_t n = (size_t)tab.off[t] + b;
        const float d = de[n * P + p], s = s1 + att2[n * A + a];
        acc = fmaf(d, act_d<ACT>(s), acc);
        dw = fmaf(d, act_f<ACT>(s), dw);
      }
      const float v = acc * w;
      stf(datt1 + ((size_t)b * P + p) * A + a, v);
      if (datt1T) stf(datt1T + (size_t)a * ldt + (size_t)b * P + p, v);
    }
    atomicAdd(dwf + a, dw);
  }
}

// ctx[n=(t,b), c] = sum_p alpha[b,t,p] F[b,p,c]  (feature-space context, needed only for d W_embed).
// One CTA per (image, 256-channel slab); alpha rows of the image in shared memory; 8 steps at a time.
template <typename T, typename TO>
__global__ void __launch_bounds__(NT)
attn_ctx_all_kernel(const __grid_constant__ StepTable tab, int P, int C, int Tcap, const T* __restrict__ F,
                    const float* __restrict__ alphas, TO* __restrict__ ctx, TO* __restrict__ ctxT, int ldt) {
  extern __shared__ float s_al[];  // [8][P]
  const int b = blockIdx.y, c = blockIdx.x * NT + threadIdx.x;
  int len = 0;
  while (len < tab.nsteps && tab.bs[len] > b) ++len;
  for (int t0 = 0; t0 < len; t0 += 8) {
    const int nt = min(8, len - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < nt * P; i += NT)
      s_al[i] = alphas[((size_t)b * Tcap + t0 + i / P) * P + (i % P)];
    __syncthreads();
    if (c < C) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int p = 0; p < P; ++p) {
        const float fv = ldf(F + ((size_t)b * P + p) * C + c);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < nt) acc[k] = fmaf(s_al[k * P + p], fv, acc[k]);
      }
      for (int k = 0; k < nt; ++k) {
        const size_t n = (size_t)tab.off[t0 + k] + b;
        if (ctx) stf(ctx + n * C + c, acc[k]);
        if (ctxT) stf(ctxT + (size_t)c * ldt + n, acc[k]);
      }
    }
  }
}

// pen_sum = sum_{b,p} (1 - S)^2 ; Gpen = -2 * coef * (1 - S)   (coef = alpha_c / (B_global * P))
__global__ void attn_penalty_kernel(int n, const float* __restrict__ S, float coef, float* __restrict__ pen_sum,
                                    float* __restrict__ Gpen) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.f;
  if (i < n) {
    const float d = 1.f - S[i];
    v = d * d;
    Gpen[i] = -2.f * coef * d;
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) atomicAdd(pen_sum, v);
}

// dst[r, :] += src[r, :] for r < rows (adds the attention query gradient into the carried dh rows)
__global__ void add_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int cols) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)rows * cols) dst[i] += src[i];
}

}  // namespace
}  // namespace st

extern "C" {

int st_attn_relayout(const float* f, int B, int C, int P, void* F, void* FT, int ldft, int out_bf16,
                     float* mean_f, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(f && F && mean_f, ST_ERR_NULL, "st_attn_relayout: NULL pointer");
  ST_REQUIRE(B >= 1 && C >= 1 && P >= 1 && (!FT || ldft >= B * P), ST_ERR_BAD_SHAPE,
             "st_attn_relayout: B=%d C=%d P=%d ldft=%d", B, C, P, ldft);
  dim3 grid((C + 31) / 32, B);
  ST_REQUIRE(grid.y <= 65535, ST_ERR_BAD_SHAPE, "st_attn_relayout: batch too large");
  cudaStream_t s = as_stream(stream);
  if (out_bf16)
    relayout_kernel<__nv_bfloat16><<<grid, NT, 0, s>>>(f, C, P, (__nv_bfloat16*)F, (__nv_bfloat16*)FT, ldft, mean_f);
  else
    relayout_kernel<float><<<grid, NT, 0, s>>>(f, C, P, (float*)F, (float*)FT, ldft, mean_f);
  ST_LAUNCH_TRY("relayout_kernel");
  return ST_OK;
}

int st_attn_step_fwd(int rows, int P, int A, int E, const void* att1, const void* Fe, int in_bf16,
                     const float* att2, const float* wf, const float* bf, const float* b_embed, float* alphas,
                     int alpha_stride, float* S, float* ctx_out, int ld_ctx, void* ctx_out_bf16, int ld_ctx_bf16,
                     int act, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(att1 && Fe && att2 && wf && bf && b_embed && alphas && S && ctx_out, ST_ERR_NULL,
             "st_attn_step_fwd: NULL pointer");
  ST_REQUIRE(rows >= 1 && P >= 1 && A >= 1 && E >= 1 && alpha_stride >= P && ld_ctx >= E &&
                 (!ctx_out_bf16 || ld_ctx_bf16 >= E),
             ST_ERR_BAD_SHAPE, "st_attn_step_fwd: rows=%d P=%d A=%d E=%d", rows, P, A, E);
  ST_REQUIRE(act == 0 || act == 1, ST_ERR_UNSUPPORTED, "st_attn_step_fwd: act=%d", act);
  {
    StreamParams sp{};
    sp.P = P; sp.A = A; sp.E = E; sp.att1 = att1; sp.Fe = Fe; sp.att2 = att2; sp.wf = wf;
    sp.bfp = bf; sp.b_embed = b_embed; sp.alphas_w = alphas; sp.S = S; sp.ctx_out = ctx_out;
    sp.ctx_bf16 = reinterpret_cast<__nv_bfloat16*>(ctx_out_bf16);
    sp.alpha_stride = alpha_stride; sp.ld_ctx = ld_ctx; sp.ld_ctx_bf16 = ld_ctx_bf16;
    const int r = attn_stream_try(false, rows, in_bf16, act, sp, as_stream(stream));
    if (r != 0) return r < 0 ? r : ST_OK;
  }
  const size_t smem = sizeof(float) * (2 * (size_t)A + P);
  ST_REQUIRE(smem <= 48 * 1024, ST_ERR_BAD_SHAPE, "st_attn_step_fwd: A=%d P=%d too large", A, P);
  cudaStream_t s = as_stream(stream);
#define ST_LAUNCH_FWD(T, ACT)                                                                         \
  attn_step_fwd_kernel<T, ACT><<<rows, NT, smem, s>>>(P, A, E, (const T*)att1, (const T*)Fe, att2, wf, bf, \
                                                      b_embed, alphas, alpha_stride, S, ctx_out, ld_ctx, \
                                                      (__nv_bfloat16*)ctx_out_bf16, ld_ctx_bf16)
  if (in_bf16) { if (act == 0) ST_LAUNCH_FWD(__nv_bfloat16, 0); else ST_LAUNCH_FWD(__nv_bfloat16, 1); }
  else         { if (act == 0) ST_LAUNCH_FWD(float, 0); else ST_LAUNCH_FWD(float, 1); }
#undef ST_LAUNCH_FWD
  ST_LAUNCH_TRY("attn_step_fwd_kernel");
  return ST_OK;
}

int st_attn_step_bwd(int rows, int P, int A, int E, const void* att1, const void* Fe, int in_bf16,
                     const float* att2, const float* wf, const float* alphas, int alpha_stride,
                     const float* dalpha, int dalpha_stride, const float* dctx, int ld_dctx, float* de_out,
                     float* datt2, void* datt2_bf16, int act, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(att1 && Fe && att2 && wf && alphas && dctx && de_out && datt2, ST_ERR_NULL,
             "st_attn_step_bwd: NULL pointer");
  ST_REQUIRE(rows >= 1 && P >= 1 && A >= 1 && E >= 1 && alpha_stride >= P && ld_dctx >= E, ST_ERR_BAD_SHAPE,
             "st_attn_step_bwd: rows=%d P=%d A=%d E=%d", rows, P, A, E);
  ST_REQUIRE(act == 0 || act == 1, ST_ERR_UNSUPPORTED, "st_attn_step_bwd: act=%d", act);
  {
    StreamParams sp{};
    sp.P = P; sp.A = A; sp.E = E; sp.att1 = att1; sp.Fe = Fe; sp.att2 = att2; sp.wf = wf;
    sp.alphas_r = alphas; sp.alpha_stride = alpha_stride; sp.dal = dalpha; sp.dal_stride = dalpha_stride;
    sp.dctx = dctx; sp.ld_dctx = ld_dctx; sp.de_out = de_out; sp.datt2 = datt2;
    sp.datt2_bf16 = reinterpret_cast<__nv_bfloat16*>(datt2_bf16);
    const int r = attn_stream_try(true, rows, in_bf16, act, sp, as_stream(stream));
    if (r != 0) return r < 0 ? r : ST_OK;
  }
  const size_t smem = sizeof(float) * (2 * (size_t)A + E + P);
  ST_REQUIRE(smem <= 48 * 1024, ST_ERR_BAD_SHAPE, "st_attn_step_bwd: A=%d E=%d P=%d too large", A, E, P);
  cudaStream_t s = as_stream(stream);
#define ST_LAUNCH_BWD(T, ACT)                                                                         \
  attn_step_bwd_kernel<T, ACT><<<rows, NT, smem, s>>>(P, A, E, (const T*)att1, (const T*)Fe, att2, wf, alphas, \
                                                      alpha_stride, dalpha, dalpha_stride, dctx, ld_dctx, de_out, datt2, \
                                                      (__nv_bfloat16*)datt2_bf16)
  if (in_bf16) { if (act == 0) ST_LAUNCH_BWD(__nv_bfloat16, 0); else ST_LAUNCH_BWD(__nv_bfloat16, 1); }
  else         { if (act == 0) ST_LAUNCH_BWD(float, 0); else ST_LAUNCH_BWD(float, 1); }
#undef ST_LAUNCH_BWD
  ST_LAUNCH_TRY("attn_step_bwd_kernel");
  return ST_OK;
}

int st_attn_hoist_bwd(int nsteps, const int* batch_sizes_host, int P, int A, const void* att1, int in_bf16,
                      const float* att2, const float* de, const float* wf, void* datt1, void* datt1T, int ldt,
                      int out_bf16, float* dwf, int act, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(att1 && att2 && de && wf && datt1 && dwf, ST_ERR_NULL, "st_attn_hoist_bwd: NULL pointer");
  ST_REQUIRE(act == 0 || act == 1, ST_ERR_UNSUPPORTED, "st_attn_hoist_bwd: act=%d", act);
  ST_REQUIRE(in_bf16 == out_bf16, ST_ERR_UNSUPPORTED, "st_attn_hoist_bwd: mixed storage types");
  const int B = tab.bs[0];
  ST_REQUIRE(!datt1T || ldt >= B * P, ST_ERR_BAD_SHAPE, "st_attn_hoist_bwd: ldt=%d", ldt);
  dim3 grid((P + 7) / 8, B);
  cudaStream_t s = as_stream(stream);
  ST_CUDA_TRY(cudaMemsetAsync(dwf, 0, sizeof(float) * A, s));
#define ST_LAUNCH_H(T, ACT)                                                                               \
  attn_hoist_bwd_kernel<T, T, ACT><<<grid, NT, 0, s>>>(tab, P, A, (const T*)att1, att2, de, wf, (T*)datt1, \
                                                       (T*)datt1T, ldt, dwf)
  if (in_bf16) { if (act == 0) ST_LAUNCH_H(__nv_bfloat16, 0); else ST_LAUNCH_H(__nv_bfloat16, 1); }
  else         { if (act == 0) ST_LAUNCH_H(float, 0); else ST_LAUNCH_H(float, 1); }
#undef ST_LAUNCH_H
  ST_LAUNCH_TRY("attn_hoist_bwd_kernel");
  return ST_OK;
}

int st_attn_ctx_all(int nsteps, const int* batch_sizes_host, int P, int C, int T_cap, const void* F, int in_bf16,
                    const float* alphas, void* ctx, void* ctxT, int ldt, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(F && alphas && (ctx || ctxT), ST_ERR_NULL, "st_attn_ctx_all: NULL pointer");
  ST_REQUIRE(T_cap >= nsteps && (!ctxT || ldt >= tab.off[nsteps]), ST_ERR_BAD_SHAPE,
             "st_attn_ctx_all: T_cap=%d ldt=%d", T_cap, ldt);
  const size_t smem = sizeof(float) * 8 * (size_t)P;
  ST_REQUIRE(smem <= 48 * 1024, ST_ERR_BAD_SHAPE, "st_attn_ctx_all: P=%d too large", P);
  dim3 grid((C + NT - 1) / NT, tab.bs[0]);
  cudaStream_t s = as_stream(stream);
  if (in_bf16)
    attn_ctx_all_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, NT, smem, s>>>(
        tab, P, C, T_cap, (const __nv_bfloat16*)F, alphas, (__nv_bfloat16*)ctx, (__nv_bfloat16*)ctxT, ldt);
  else
    attn_ctx_all_kernel<float, float><<<grid, NT, smem, s>>>(tab, P, C, T_cap, (const float*)F, alphas,
                                                             (float*)ctx, (float*)ctxT, ldt);
  ST_LAUNCH_TRY("attn_ctx_all_kernel");
  return ST_OK;
}

int st_attn_penalty(int n, const float* S, float coef, float* pen_sum, float* Gpen, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(S && pen_sum && Gpen, ST_ERR_NULL, "st_attn_penalty: NULL pointer");
  ST_REQUIRE(n >= 1, ST_ERR_BAD_SHAPE, "st_attn_penalty: n=%d", n);
  cudaStream_t s = as_stream(stream);
  ST_CUDA_TRY(cudaMemsetAsync(pen_sum, 0, sizeof(float), s));
  attn_penalty_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, S, coef, pen_sum, Gpen);
  ST_LAUNCH_TRY("attn_penalty_kernel");
  return ST_OK;
}

int st_add_rows(float* dst, const float* src, int rows, int cols, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(dst && src, ST_ERR_NULL, "st_add_rows: NULL pointer");
  if (rows <= 0 || cols <= 0) return ST_OK;
  const size_t n = (size_t)rows * cols;
  add_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dst, src, rows, cols);
  ST_LAUNCH_TRY("add_rows_kernel");
  return ST_OK;
}

}  // extern "C"
