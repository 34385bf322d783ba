// Optimizer step of the training loop (main.py:97-100,152; main_attn.py:91-94,134): torch.optim.SGD(lr,
// momentum) and torch.optim.Adam(lr) over every trainable tensor in ONE launch -- the step right after the
// path's backward.  HBM-bound: SGD reads p, g, buf and writes p, buf (20 B per parameter),
// Adam reads p, g, m, v and writes p, m, v (28 B per parameter).  The tensors form one index space of
// 4096-element chunks (as st_scale_multi), 128-bit accesses, grid = a multiple of the SM count.
//
// Arithmetic follows torch's single-tensor reference implementations (torch/optim/sgd.py, adam.py):
//   SGD   buf = g (first step) | momentum * buf + g;  p -= lr * buf          (dampening 0, no nesterov)
//   Adam  m += (g - m) * (1 - b1);  v = b2 * v + (1 - b2) * g * g;
//         p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// `grad_scale` (device scalar, may be NULL = 1) multiplies every gradient first: autograd's grad_output of
// forward_loss, so the chain-rule pass over the gradients folds into the optimizer's read of them.
// `shadow_bf16[i]` (table and entries may be NULL): a contiguous bf16 copy of parameter i, rewritten from the
// updated fp32 value in the same pass -- the operand the tensor-core kernels of the next step read, so no
// separate fp32 -> bf16 cast launches exist in the training loop (2 more bytes written per parameter).
#include <cuda_bf16.h>

#include "common.cuh"

namespace st {
namespace {

constexpr int OPT_CHUNK = 4096;
struct OptTable {
  int n;
  float* p[ST_OPT_MAX];
  const float* g[ST_OPT_MAX];
  float* m[ST_OPT_MAX];
  float* v[ST_OPT_MAX];
  __nv_bfloat16* s[ST_OPT_MAX];
  long long count[ST_OPT_MAX];
  int first[ST_OPT_MAX + 1];
};
struct OptHyper {
  float lr, momentum, b1, b2, omb1, omb2, eps, bc1, sqrt_bc2;   // omb = 1 - beta (rounded from double, as torch passes it); bc1 = 1 - b1^t, sqrt_bc2 = sqrt(1 - b2^t)
  int first_step;
};

template <bool ADAM>
__device__ __forceinline__ void update(float& p, float g, float& m, float& v, const OptHyper& h) {
  if (ADAM) {
    m = fmaf(g - m, h.omb1, m);                           // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(h.omb2 * g, g, h.b2 * v);                    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / h.sqrt_bc2 + h.eps;
    p -= (h.lr / h.bc1) * (m / denom);
  } else {
    if (h.momentum != 0.f) {
      m = h.first_step ? g : fmaf(h.momentum, m, g);
      g = m;
    }
    p -= h.lr * g;
  }
}

template <bool ADAM>
__global__ void __launch_bounds__(256) optim_multi_kernel(const OptTable tab, const OptHyper h, const float* __restrict__ gscale) {
  const float a = gscale ? __ldg(gscale) : 1.f;
  const bool has_m = ADAM || h.momentum != 0.f;
  for (int c = blockIdx.x; c < tab.first[tab.n]; c += gridDim.x) {
    int i = 0;
    while (c >= tab.first[i + 1]) ++i;
    const long long base = (long long)(c - tab.first[i]) * OPT_CHUNK;
    const int n = (int)min((long long)OPT_CHUNK, tab.count[i] - base);
    float* __restrict__ P = tab.p[i] + base;
    const float* __restrict__ G = tab.g[i] + base;
    float* __restrict__ M = has_m ? tab.m[i] + base : nullptr;
    float* __restrict__ V = ADAM ? tab.v[i] + base : nullptr;
    __nv_bfloat16* __restrict__ S = tab.s[i] ? tab.s[i] + base : nullptr;
    const uintptr_t al = (reinterpret_cast<uintptr_t>(S) << 1) | reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(M) |
                         reinterpret_cast<uintptr_t>(V);
    int done = 0;
    if ((al & 15) == 0) {
      const int n4 = n >> 2;
      for (int j = threadIdx.x; j < n4; j += blockDim.x) {
        float4 p4 = reinterpret_cast<float4*>(P)[j];
        const float4 g4 = __ldcs(reinterpret_cast<const float4*>(G) + j);
        float4 m4 = has_m ? reinterpret_cast<float4*>(M)[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v4 = ADAM ? reinterpret_cast<float4*>(V)[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        update<ADAM>(p4.x, g4.x * a, m4.x, v4.x, h);
        update<ADAM>(p4.y, g4.y * a, m4.y, v4.y, h);
        update<ADAM>(p4.z, g4.z * a, m4.z, v4.z, h);
        update<ADAM>(p4.w, g4.w * a, m4.w, v4.w, h);
        reinterpret_cast<float4*>(P)[j] = p4;
        if (S) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
          reinterpret_cast<uint2*>(S)[j] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
        if (has_m) reinterpret_cast<float4*>(M)[j] = m4;
        if (ADAM) reinterpret_cast<float4*>(V)[j] = v4;
      }
      done = n4 << 2;
    }
    for (int j = done + threadIdx.x; j < n; j += blockDim.x) {
      float pj = P[j], mj = has_m ? M[j] : 0.f, vj = ADAM ? V[j] : 0.f;
      update<ADAM>(pj, G[j] * a, mj, vj, h);
      P[j] = pj;
      if (S) S[j] = __float2bfloat16(pj);
      if (has_m) M[j] = mj;
      if (ADAM) V[j] = vj;
    }
  }
}

int build_table(OptTable& tab, int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                void* const* shadow, const int64_t* count, bool need_m, bool need_v, long long* chunks_out) {
  ST_REQUIRE(n >= 0 && n <= ST_OPT_MAX, ST_ERR_BAD_SHAPE, "optimizer step: n=%d tensors (max %d per call)", n, (int)ST_OPT_MAX);
  ST_REQUIRE(n == 0 || (p && g && count && (!need_m || m) && (!need_v || v)), ST_ERR_NULL, "optimizer step: NULL pointer table");
  tab.n = n;
  long long chunks = 0;
  for (int i = 0; i < n; ++i) {
    ST_REQUIRE(count[i] >= 0, ST_ERR_BAD_SHAPE, "optimizer step: tensor %d has count %lld", i, (long long)count[i]);
    ST_REQUIRE(count[i] == 0 || (p[i] && g[i] && (!need_m || m[i]) && (!need_v || v[i])), ST_ERR_NULL,
               "optimizer step: tensor %d has a NULL pointer", i);
    tab.p[i] = p[i]; tab.g[i] = g[i]; tab.m[i] = need_m ? m[i] : nullptr; tab.v[i] = need_v ? v[i] : nullptr;
    tab.s[i] = shadow ? reinterpret_cast<__nv_bfloat16*>(shadow[i]) : nullptr;
    tab.count[i] = count[i];
    tab.first[i] = (int)chunks;
    chunks += (count[i] + OPT_CHUNK - 1) / OPT_CHUNK;
    ST_REQUIRE(chunks < (1LL << 30), ST_ERR_BAD_SHAPE, "optimizer step: too many elements");
  }
  tab.first[n] = (int)chunks;
  *chunks_out = chunks;
  return ST_OK;
}

}  // namespace
}  // namespace st

extern "C" {

int st_sgd_step(int n, float* const* param, const float* const* grad, float* const* momentum_buf, void* const* shadow_bf16,
                const int64_t* count, float lr, float momentum, int first_step, const float* grad_scale, st_stream_t stream) {
  using namespace st;
  OptTable tab;
  long long chunks = 0;
  ST_TRY(build_table(tab, n, param, grad, momentum_buf, nullptr, shadow_bf16, count, momentum != 0.f, false, &chunks));
  if (chunks == 0) return ST_OK;
  OptHyper h{lr, momentum, 0.f, 0.f, 1.f, 1.f, 0.f, 1.f, 1.f, first_step};
  int sms = 0;
  ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
  const long long grid = chunks < 8LL * sms ? chunks : 8LL * sms;
  optim_multi_kernel<false><<<(unsigned)grid, 256, 0, as_stream(stream)>>>(tab, h, grad_scale);
  ST_LAUNCH_TRY("optim_multi_kernel<sgd>");
  return ST_OK;
}

int st_adam_step(int n, float* const* param, const float* const* grad, float* const* exp_avg, float* const* exp_avg_sq,
                 void* const* shadow_bf16, const int64_t* count, float lr, double beta1_d, double beta2_d, float eps, int64_t step,
                 const float* grad_scale, st_stream_t stream) {
  using namespace st;
  const float beta1 = (float)beta1_d, beta2 = (float)beta2_d;   // betas come as doubles: 1 - beta is rounded once, as torch does
  ST_REQUIRE(step >= 1, ST_ERR_BAD_SHAPE, "st_adam_step: step=%lld must be >= 1", (long long)step);
  ST_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, ST_ERR_BAD_SHAPE,
             "st_adam_step: betas (%g, %g) eps %g", beta1, beta2, eps);
  OptTable tab;
  long long chunks = 0;
  ST_TRY(build_table(tab, n, param, grad, exp_avg, exp_avg_sq, shadow_bf16, count, true, true, &chunks));
  if (chunks == 0) return ST_OK;
  const double bc1 = 1.0 - pow(beta1_d, (double)step), bc2 = 1.0 - pow(beta2_d, (double)step);
  OptHyper h{lr, 0.f, beta1, beta2, (float)(1.0 - beta1_d), (float)(1.0 - beta2_d), eps, (float)bc1, (float)sqrt(bc2), 0};
  int sms = 0;
  ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
  const long long grid = chunks < 8LL * sms ? chunks : 8LL * sms;
  optim_multi_kernel<true><<<(unsigned)grid, 256, 0, as_stream(stream)>>>(tab, h, grad_scale);
  ST_LAUNCH_TRY("optim_multi_kernel<adam>");
  return ST_OK;
}

}  // extern "C"
