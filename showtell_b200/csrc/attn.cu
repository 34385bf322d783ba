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
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace st {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
// two adjacent elements; p must be aligned to 2 elements
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st2(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

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
// mean_f (B,C) fp32 (rnn_attn.py:62 `cnn_feature.mean(dim=2)`).  One CTA per (64-channel slab, image):
// the slab (64 x P floats, contiguous in f) is read once into shared memory with coalesced loads, then
// written out twice: F rows get 64 channels = 128 B (bf16) per location, FT rows get P contiguous
// locations per channel.  HBM-bound: 4 + 2*sizeof(T) bytes per element.
constexpr int RL_CH = 64;
// TI = float (the reference's feature dtype, cnn_attn.py:49) or __nv_bfloat16 (a trunk run under autocast: half the
// bytes over PCIe / HBM for the largest input of the step).
template <typename T, typename TI>
__global__ void __launch_bounds__(NT)
relayout_kernel(const TI* __restrict__ f, int C, int P, T* __restrict__ F, T* __restrict__ FT, int ldft,
                float* __restrict__ mean_f) {
  extern __shared__ float tile[];                 // [RL_CH][PS], PS odd: conflict-free column reads
  const int PS = P | 1;
  const int b = blockIdx.y, c0 = blockIdx.x * RL_CH;
  const int nch = min(RL_CH, C - c0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TI* src = f + ((size_t)b * C + c0) * P;
  if (sizeof(TI) == 2) {
    if ((P & 3) == 0 && ((nch * P) & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      // 8 bf16 per 128-bit load; P % 4 == 0, so each half of a vector stays inside one channel row
      const int nv = nch * P / 8;
      for (int i = tid; i < nv; i += NT) {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(src) + (size_t)i * 16);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          const int e = i * 8 + hq * 4, row = e / P, col = e - row * P;
          float* d = tile + row * PS + col;
          d[0] = __uint_as_float(w[2 * hq] << 16); d[1] = __uint_as_float(w[2 * hq] & 0xffff0000u);
          d[2] = __uint_as_float(w[2 * hq + 1] << 16); d[3] = __uint_as_float(w[2 * hq + 1] & 0xffff0000u);
        }
      }
    } else {
      for (int i = tid; i < nch * P; i += NT) tile[(i / P) * PS + (i % P)] = (float)src[i];
    }
  } else if ((P & 3) == 0) {   // the slab is contiguous and 16-byte aligned: 128-bit loads, running (row, col)
    const int nv = nch * P / 4;
    int col = tid * 4, row = 0;
    while (col >= P) { col -= P; ++row; }
    for (int i = tid; i < nv; i += NT) {
      const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (size_t)i * 4);
      float* d = tile + row * PS + col;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      col += NT * 4;
      while (col >= P) { col -= P; ++row; }
    }
  } else {
    for (int i = tid; i < nch * P; i += NT) tile[(i / P) * PS + (i % P)] = (float)src[i];
  }
  __syncthreads();
  // F[(b*P + p), c0 + c]: warp per location, a lane writes two adjacent channels (128 B per warp in bf16)
  for (int p = warp; p < P; p += NT / 32) {
    T* dst = F + ((size_t)b * P + p) * C + c0;
    if (nch == RL_CH && (C & 1) == 0) st2(dst + 2 * lane, tile[(2 * lane) * PS + p], tile[(2 * lane + 1) * PS + p]);
    else
      for (int c = lane; c < nch; c += 32) stf(dst + c, tile[c * PS + p]);
  }
  // FT[c0 + c, b*P + p] and the channel means: warp per channel, a lane owns two adjacent locations
  const bool pair_ok = FT && ((ldft & 1) == 0) && (((size_t)b * P) & 1) == 0;
  for (int c = warp; c < nch; c += NT / 32) {
    float s = 0.f;
    T* dst = FT ? FT + (size_t)(c0 + c) * ldft + (size_t)b * P : nullptr;
    for (int p = 2 * lane; p < P; p += 64) {
      const float v0 = tile[c * PS + p], v1 = (p + 1 < P) ? tile[c * PS + p + 1] : 0.f;
      s += v0 + v1;
      if (FT) {
        if (pair_ok && p + 1 < P) st2(dst + p, v0, v1);
        else {
          stf(dst + p, v0);
          if (p + 1 < P) stf(dst + p + 1, v1);
        }
      }
    }
    s = warp_sum(s);
    if (lane == 0) mean_f[(size_t)b * C + c0 + c] = s / (float)P;
  }
}

// The common case of relayout_kernel -- P % 4 == 0, C % 64 == 0, bf16 F, no transposed copy -- with an order of
// magnitude fewer instructions (the general kernel issued 51 per element and was issue-bound at 2.4 TB/s):
//   load:  warp per CHANNEL PAIR; a lane reads 4 consecutive locations of both channels (128-bit / 64-bit loads, rows
//          are 16 / 8-byte aligned) and stores them as 4 (even, odd) pairs, 64-bit, conflict-free; the channel means
//          come out of the same registers (warp sum);
//   store: warp per LOCATION; a lane reads its channel pair (64-bit, row stride odd in pairs: conflict-free), converts
//          and stores 4 bytes -- 128 contiguous bytes per warp instruction.
template <typename TI>
__global__ void __launch_bounds__(NT)
relayout_fast_kernel(const TI* __restrict__ f, int C, int P, __nv_bfloat16* __restrict__ F, float* __restrict__ mean_f) {
  extern __shared__ float2 tile2[];               // [32 channel pairs][PS2], PS2 = P | 1
  const int PS2 = P | 1;
  const int b = blockIdx.y, c0 = blockIdx.x * RL_CH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TI* src = f + ((size_t)b * C + c0) * P;
  const int nq = P >> 2;                          // 4-location groups per channel row
  const float inv_p = 1.f / (float)P;
  for (int cp = warp; cp < RL_CH / 2; cp += NT / 32) {
    const TI* r0 = src + (size_t)(2 * cp) * P;
    const TI* r1 = r0 + P;
    float s0 = 0.f, s1 = 0.f;
    for (int q = lane; q < nq; q += 32) {
      float a[4], c[4];
      if (sizeof(TI) == 4) {
        const float4 va = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(r0) + 4 * q);
        const float4 vc = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(r1) + 4 * q);
        a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w; c[0] = vc.x; c[1] = vc.y; c[2] = vc.z; c[3] = vc.w;
      } else {
        const uint2 ua = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(r0) + 4 * q);
        const uint2 uc = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(r1) + 4 * q);
        a[0] = __uint_as_float(ua.x << 16); a[1] = __uint_as_float(ua.x & 0xffff0000u);
        a[2] = __uint_as_float(ua.y << 16); a[3] = __uint_as_float(ua.y & 0xffff0000u);
        c[0] = __uint_as_float(uc.x << 16); c[1] = __uint_as_float(uc.x & 0xffff0000u);
        c[2] = __uint_as_float(uc.y << 16); c[3] = __uint_as_float(uc.y & 0xffff0000u);
      }
      float2* d = tile2 + (size_t)cp * PS2 + 4 * q;
#pragma unroll
      for (int i = 0; i < 4; ++i) { d[i] = make_float2(a[i], c[i]); s0 += a[i]; s1 += c[i]; }
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) *reinterpret_cast<float2*>(mean_f + (size_t)b * C + c0 + 2 * cp) = make_float2(s0 * inv_p, s1 * inv_p);
  }
  __syncthreads();
  __nv_bfloat16* dst = F + ((size_t)b * P) * C + c0 + 2 * lane;
  for (int p = warp; p < P; p += NT / 32) {
    const float2 v = tile2[(size_t)lane * PS2 + p];
    *reinterpret_cast<__nv_bfloat162*>(dst + (size_t)p * C) = __floats2bfloat162_rn(v.x, v.y);
  }
}

// The common case again (bf16 F, no transposed copy, C % 64 == 0), bound by HBM instead of by load latency: the 64
// channel rows of a block are CONTIGUOUS in the channels-first grid, so four bulk copies (cp.async.bulk, 16 rows
// each) bring the whole 50 KB tile into shared memory with every byte in flight at once -- relayout_fast_kernel
// walked it with two 16-byte loads per thread at a time (8 dependent round trips per warp, 1 TB/s).  The tile keeps
// the global layout [channel][location]; a thread then owns one location and 16 channels: 16 conflict-free
// reads (consecutive lanes = consecutive locations), one 32-byte store into the channels-last row.
// Persistent blocks (two per SM) walk the tiles with a two-stage ring: the copy of the next tile is issued before the
// current one is processed, so reads stay in flight through the store phase (one tile per block, 4 blocks per SM,
// ran all blocks in lock step -- load, then store -- at 4.0 TB/s).  Consecutive tiles are consecutive in the grid.
constexpr int RL_NT = 512;
template <typename TI>
__global__ void __launch_bounds__(RL_NT)
relayout_bulk_kernel(const TI* __restrict__ f, int C, int P, int ntiles, __nv_bfloat16* __restrict__ F,
                     float* __restrict__ mean_f) {
  extern __shared__ __align__(128) uint8_t rl_raw[];
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tile_bytes = (uint32_t)RL_CH * (uint32_t)P * (uint32_t)sizeof(TI);
  const uint32_t stage_bytes = (tile_bytes + 127u) & ~127u;
  const int tiles_per_b = C / RL_CH;
  auto issue = [&](int t, int stage) {             // thread 0: four bulk copies of 16 channel rows each
    const uint32_t part = tile_bytes / 4;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(f) + (size_t)t * tile_bytes;   // tile t = (b, c0 / RL_CH)
    mbar_expect_tx(&bar[stage], tile_bytes);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(rl_raw) + stage * stage_bytes + i * part),
                   "l"(src + (size_t)i * part), "r"(part), "r"(smem_u32(&bar[stage]))
                   : "memory");
  };
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue(blockIdx.x, 0);
  }
  __syncthreads();
  const float inv_p = 1.f / (float)P;
  uint32_t ph[2] = {0u, 0u};
  int stage = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, stage ^= 1) {
    if (tid == 0 && t + (int)gridDim.x < ntiles) issue(t + gridDim.x, stage ^ 1);   // freed by the barrier below
    mbar_wait(&bar[stage], ph[stage]);
    ph[stage] ^= 1u;
    const TI* tile = reinterpret_cast<const TI*>(rl_raw + stage * stage_bytes);
    const int b = t / tiles_per_b, c0 = (t - b * tiles_per_b) * RL_CH;
    {                                              // channel means: a warp sums its RPW rows side by side
      constexpr int RPW = RL_CH / (RL_NT / 32);
      const TI* rows = tile + (warp * RPW) * P;
      float sm[RPW];
#pragma unroll
      for (int k = 0; k < RPW; ++k) sm[k] = 0.f;
      for (int p = lane; p < P; p += 32) {
#pragma unroll
        for (int k = 0; k < RPW; ++k) sm[k] += ldf(rows + k * P + p);
      }
      float mine = 0.f;
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        const float v = warp_sum(sm[k]);
        if (lane == k) mine = v;
      }
      if (lane < RPW) mean_f[(size_t)b * C + c0 + warp * RPW + lane] = mine * inv_p;
    }
    __nv_bfloat16* dst = F + ((size_t)b * P) * C + c0;
    for (int it = tid; it < P * (RL_CH / 16); it += RL_NT) {
      const int g = it / P, p = it - g * P;
      const TI* tp = tile + (16 * g) * P + p;
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(ldf(tp + (2 * i) * P), ldf(tp + (2 * i + 1) * P));
        o[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
      // one full 32-byte sector per lane (256-bit store): 16-byte stores left every sector half written per request
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + (size_t)p * C + 16 * g),
                   "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                   : "memory");
    }
    __syncthreads();                               // every read of this stage is done before it is filled again
  }
}

// Channels-last grids (B, P, C): F = the grid itself; only the channel means over the P locations are needed
// (rnn_attn.py:62 `cnn_feature.mean(dim=2)` of the channels-first tensor).  Thread per channel, coalesced rows.
template <typename T>
__global__ void __launch_bounds__(NT) grid_mean_bpc_kernel(const T* __restrict__ F, int P, int C, float* __restrict__ mean_f) {
  const int b = blockIdx.y, c = blockIdx.x * NT + threadIdx.x;
  if (c >= C) return;
  const T* src = F + (size_t)b * P * C + c;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int p = 0;
  for (; p + 3 < P; p += 4) {
    s0 += ldf(src + (size_t)p * C); s1 += ldf(src + (size_t)(p + 1) * C);
    s2 += ldf(src + (size_t)(p + 2) * C); s3 += ldf(src + (size_t)(p + 3) * C);
  }
  for (; p < P; ++p) s0 += ldf(src + (size_t)p * C);
  mean_f[(size_t)b * C + c] = ((s0 + s1) + (s2 + s3)) / (float)P;
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
                     float* __restrict__ datt2, __nv_bfloat16* __restrict__ datt2_bf16, float* __restrict__ gt_out) {
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
    if (gt_out) gt_out[(size_t)b * A + a] = acc;
    datt2[(size_t)b * A + a] = acc * s_wf[a];
    if (datt2_bf16) datt2_bf16[(size_t)b * A + a] = __float2bfloat16(acc * s_wf[a]);
  }
}

// After the time loop: d att1[b,p,a] = w_f[a] sum_t de[t,b,p] act'(att1[b,p,a] + att2[t,b,a]) and
// d w_f[a] += sum_{t,b,p} de[t,b,p] act(att1[b,p,a] + att2[t,b,a]).  de (N,P), att2 (N,A) packed.
// ALU-bound (B*P*A*T tuples): one CTA per (image, HB_P grid locations); a thread owns one attention
// unit a per pass and keeps att2[t,b,a] for up to HB_T steps in registers; de of the CTA's locations
// sits in shared memory as [p][t] so four steps come with one broadcast 128-bit read.  The
// transposed copy (the K-major operand of dW_enc = datt1^T F) is staged through shared memory so
// each attention unit's row gets HB_P contiguous locations.
constexpr int HB_P = 28, HB_T = 20;
template <typename T, typename TO, int ACT>
__global__ void __launch_bounds__(NT)
attn_hoist_bwd_kernel(const __grid_constant__ StepTable tab, int P, int A, const T* __restrict__ att1,
                      const float* __restrict__ att2, const float* __restrict__ de, const float* __restrict__ wf,
                      TO* __restrict__ datt1, TO* __restrict__ datt1T, int ldt, float* __restrict__ dwf) {
  extern __shared__ float hsm[];
  float* s_de = hsm;                                   // [HB_P][HB_T]
  TO* s_out = reinterpret_cast<TO*>(hsm + HB_P * HB_T); // [A][HB_P + 1] (only if datt1T)
  const int b = blockIdx.y, p0 = blockIdx.x * HB_P, np = min(HB_P, P - p0);
  int len = 0;
  while (len < tab.nsteps && tab.bs[len] > b) ++len;   // steps in which row b is live
  const int OS = HB_P + 1;
  for (int a0 = 0; a0 < A; a0 += NT) {
    const int a = a0 + threadIdx.x;
    const bool a_ok = a < A;
    const float w = a_ok ? wf[a] : 0.f;
    float dw = 0.f;
    float acc[HB_P];
#pragma unroll
    for (int i = 0; i < HB_P; ++i) acc[i] = 0.f;
    for (int t0 = 0; t0 < len; t0 += HB_T) {
      const int nt = min(HB_T, len - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < HB_P * HB_T; i += NT) {
        const int pi = i / HB_T, ti = i % HB_T;
        s_de[i] = (pi < np && ti < nt) ? de[((size_t)tab.off[t0 + ti] + b) * P + p0 + pi] : 0.f;
      }
      float a2[HB_T];
#pragma unroll
      for (int ti = 0; ti < HB_T; ++ti)
        a2[ti] = (a_ok && ti < nt) ? att2[((size_t)tab.off[t0 + ti] + b) * A + a] : 0.f;
      __syncthreads();
      if (a_ok) {
#pragma unroll
        for (int pi = 0; pi < HB_P; ++pi) {
          if (pi < np) {
            const float s1 = ldf(att1 + ((size_t)b * P + p0 + pi) * A + a);
#pragma unroll
            for (int ti = 0; ti < HB_T; ti += 4) {
              const float4 d4 = *reinterpret_cast<const float4*>(s_de + pi * HB_T + ti);
              const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float sv = s1 + a2[ti + k];
                if (ACT == 0) {   // LeakyReLU: act(s) = s * act'(s)
                  const float g = sv > 0.f ? dd[k] : 0.2f * dd[k];
                  acc[pi] += g;
                  dw = fmaf(sv, g, dw);
                } else {
                  acc[pi] = fmaf(dd[k], act_d<ACT>(sv), acc[pi]);
                  dw = fmaf(dd[k], act_f<ACT>(sv), dw);
                }
              }
            }
          }
        }
      }
    }
    if (a_ok) {
#pragma unroll
      for (int pi = 0; pi < HB_P; ++pi) {
        if (pi < np) {
          const float v = acc[pi] * w;
          stf(datt1 + ((size_t)b * P + p0 + pi) * A + a, v);
          if (datt1T) stf(s_out + (size_t)a * OS + pi, v);
        }
      }
      atomicAdd(dwf + a, dw);
    }
  }
  if (datt1T) {
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int a = warp; a < A; a += NT / 32) {
      TO* dst = datt1T + (size_t)a * ldt + (size_t)b * P + p0;
      for (int pi = lane; pi < np; pi += 32) dst[pi] = s_out[(size_t)a * OS + pi];
    }
  }
}

// ctx[n=(t,b), c] = sum_p alpha[b,t,p] F[b,p,c]  (feature-space context, needed only for d W_embed).
// Bound by one read of F: one CTA per (image, 4*NT-channel slab); a thread owns 4 adjacent channels
// (one 8- or 16-byte load per location) and accumulates CX_T steps at a time, the alpha rows of those
// steps sitting in shared memory as [p][t] (four steps per broadcast 128-bit read).
constexpr int CX_T = 20;
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
}
// dW_embed without rebuilding the feature-space contexts: with ctx[(t,b),:] = sum_p alpha[b,t,p] F[b,p,:],
//   dW_embed[e,c] = sum_{t,b} dctx_e[(t,b),e] ctx[(t,b),c] = sum_{(b,p)} Q[(b,p),e] F[(b,p),c],
//   Q[(b,p),e] = sum_t alpha[b,t,p] dctx_e[(t,b),e]      (t over the steps in which row b is live)
// so one pass forms Q (B*P, E) as bf16 -- reading alphas and the (N,E) gradients only -- and ONE tensor-core GEMM
// Q^T . F over the B*P locations (both operands in place, MN-major) replaces the pass that re-reads the whole grid
// F (B*P*C) once per live step chunk (attn_ctx_all: 148 us at config 3) plus a small GEMM.
// CTA = one batch row x QP locations; thread = 2 adjacent embedding columns (blockDim = ceil(E/2) rounded to warps).
constexpr int QP = 28, QT = 20;
__global__ void __launch_bounds__(NT)
attn_embed_q_kernel(const __grid_constant__ StepTable tab, int P, int E, int Tcap, const float* __restrict__ alphas,
                    const float* __restrict__ dctx, int ldd, __nv_bfloat16* __restrict__ Q) {
  __shared__ __align__(16) float s_al[QT][QP];                 // rows of 112 B: four locations per 128-bit broadcast read
  const int b = blockIdx.y, p0 = blockIdx.x * QP, np = min(QP, P - p0);
  int len = 0;
  while (len < tab.nsteps && tab.bs[len] > b) ++len;
  for (int eb = 0; eb < E; eb += 2 * NT) {                    // one pass when E <= 512; uniform trip count (barriers inside)
    const int e0 = eb + 2 * threadIdx.x;
    const bool on = e0 < E, two = e0 + 1 < E;
    const bool pairld = two && (ldd & 1) == 0 && (reinterpret_cast<uintptr_t>(dctx) & 7) == 0;
    float acc0[QP], acc1[QP];
#pragma unroll
    for (int i = 0; i < QP; ++i) acc0[i] = acc1[i] = 0.f;
    auto load_d = [&](int t, float& d0, float& d1) {
      const float* d = dctx + ((size_t)tab.off[t] + b) * ldd + e0;
      if (pairld) { const float2 v = *reinterpret_cast<const float2*>(d); d0 = v.x; d1 = v.y; }
      else { d0 = on ? d[0] : 0.f; d1 = two ? d[1] : 0.f; }
    };
    for (int t0 = 0; t0 < len; t0 += QT) {
      const int nt = min(QT, len - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < QT * QP; i += NT) {
        const int ti = i / QP, pi = i % QP;
        s_al[ti][pi] = (ti < nt && pi < np) ? alphas[((size_t)b * Tcap + t0 + ti) * P + p0 + pi] : 0.f;
      }
      float n0 = 0.f, n1 = 0.f;
      load_d(t0, n0, n1);
      __syncthreads();
      for (int ti = 0; ti < nt; ++ti) {
        const float d0 = n0, d1 = n1;
        if (ti + 1 < nt) load_d(t0 + ti + 1, n0, n1);          // next step's gradient row while this one is used
#pragma unroll
        for (int pi = 0; pi < QP; pi += 4) {
          const float4 a = *reinterpret_cast<const float4*>(&s_al[ti][pi]);
          acc0[pi] = fmaf(a.x, d0, acc0[pi]);         acc1[pi] = fmaf(a.x, d1, acc1[pi]);
          acc0[pi + 1] = fmaf(a.y, d0, acc0[pi + 1]); acc1[pi + 1] = fmaf(a.y, d1, acc1[pi + 1]);
          acc0[pi + 2] = fmaf(a.z, d0, acc0[pi + 2]); acc1[pi + 2] = fmaf(a.z, d1, acc1[pi + 2]);
          acc0[pi + 3] = fmaf(a.w, d0, acc0[pi + 3]); acc1[pi + 3] = fmaf(a.w, d1, acc1[pi + 3]);
        }
      }
    }
#pragma unroll
    for (int pi = 0; pi < QP; ++pi) {
      if (pi < np && on) {
        __nv_bfloat16* q = Q + ((size_t)b * P + p0 + pi) * E + e0;
        if (two && (E & 1) == 0) *reinterpret_cast<__nv_bfloat162*>(q) = __floats2bfloat162_rn(acc0[pi], acc1[pi]);
        else {
          q[0] = __float2bfloat16(acc0[pi]);
          if (two) q[1] = __float2bfloat16(acc1[pi]);
        }
      }
    }
  }
}

// LeakyReLU(0.2) form of attn_hoist_bwd_kernel (rnn_attn.py:26), without the transposed copy.  act'(s) is 1 or 0.2, so
//   g[t,p] = de[t,p] act'(s) = 0.2 de[t,p] + 0.8 de[t,p] [s > 0],    s = att1[b,p,a] + att2[t,b,a]  (> 0  <=>  att1 > -att2)
// and a tuple costs one compare and two predicated adds -- into a per-location and a per-step accumulator:
//   datt1[b,p,a] = w_f[a] (0.2 D_p + 0.8 A_p),        A_p = sum_t de[t,p][s > 0],   D_p = sum_t de[t,p]
//   dw_f[a]     += sum_p att1[b,p,a] (0.2 D_p + 0.8 A_p) + sum_t att2[t,b,a] (0.2 D_t + 0.8 A_t)      (act(s) = s act'(s))
// with D_p / D_t (independent of a) summed once per CTA.  63 M -> ~25 M warp instructions at config 3.
// out[c] += sum_r X[r,c] Y[r,c]  (the step part of dw_f: X = att2 (N,A), Y = gt (N,A); a few MB)
constexpr int CSP_ROWS = 16;
__global__ void __launch_bounds__(256) colsum_prod_kernel(float* __restrict__ out, const float* __restrict__ X,
                                                          const float* __restrict__ Y, int rows, int cols) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * CSP_ROWS, r1 = min(rows, r0 + CSP_ROWS);
  float s0 = 0.f, s1 = 0.f;
  if (r1 - r0 == CSP_ROWS) {                      // all 2 * CSP_ROWS loads in flight at once
    float x[CSP_ROWS], y[CSP_ROWS];
#pragma unroll
    for (int i = 0; i < CSP_ROWS; ++i) {
      x[i] = X[(size_t)(r0 + i) * cols + c];
      y[i] = Y[(size_t)(r0 + i) * cols + c];
    }
#pragma unroll
    for (int i = 0; i < CSP_ROWS; i += 2) {
      s0 = fmaf(x[i], y[i], s0);
      s1 = fmaf(x[i + 1], y[i + 1], s1);
    }
  } else {
    for (int r = r0; r < r1; ++r) s0 = fmaf(X[(size_t)r * cols + c], Y[(size_t)r * cols + c], s0);
  }
  atomicAdd(out + c, s0 + s1);
}

// FULL: P is a multiple of HB_P, A a multiple of NT and every offset fits 31 bits (the launcher checks): no tail
// predicates, 32-bit index arithmetic -- a third fewer instructions around the tuple loop.
// GT: the step part of dw_f, sum_{t,p} g[t,p] att2[t], is formed by the launcher as sum_rows att2 * gt from the tensor
// the per-step attention backward kernels already computed (gt[(t,b),a] = sum_p g[t,p]): a third of the tuple work less.
template <typename T, typename TO, bool FULL, bool GT>
__global__ void __launch_bounds__(NT, 3)
attn_hoist_bwd_lrelu_kernel(const __grid_constant__ StepTable tab, int P, int A, const T* __restrict__ att1,
                            const float* __restrict__ att2, const float* __restrict__ de, const float* __restrict__ wf,
                            TO* __restrict__ datt1, float* __restrict__ dwf) {
  __shared__ __align__(16) float s_de[HB_P][HB_T];       // [p][t] of the chunk, zero padded
  __shared__ float s_dp[HB_P], s_dpt[HB_P], s_dt[HB_T];  // chunk sums over t, their total over the chunks, chunk sums over p
  __shared__ int s_row[HB_T];                            // packed row of (t0 + ti, b)
  using IX = typename std::conditional<FULL, int, size_t>::type;   // offsets
  const int b = blockIdx.y, p0 = blockIdx.x * HB_P, np = FULL ? HB_P : min(HB_P, P - p0), tid = threadIdx.x;
  const int a = blockIdx.z * NT + tid;                   // grid.z = blocks of NT attention units
  const bool a_ok = FULL || a < A;
  int len = 0;
  while (len < tab.nsteps && tab.bs[len] > b) ++len;     // steps in which row b is live
  constexpr int GP = 7;                                  // locations per group: their att1 values are loaded together,
  static_assert(HB_P % GP == 0, "location groups");      // one group ahead of the arithmetic
  const T* a1 = att1 + ((IX)b * P + p0) * A + a;
  float ap[HB_P], dw = 0.f;
#pragma unroll
  for (int i = 0; i < HB_P; ++i) ap[i] = 0.f;
  if (tid < HB_P) s_dpt[tid] = 0.f;
  for (int t0 = 0; t0 < len; t0 += HB_T) {
    const int nt = min(HB_T, len - t0);
    __syncthreads();
    if (tid < HB_T) s_row[tid] = tid < nt ? tab.off[t0 + tid] + b : -1;
    __syncthreads();
    for (int i = tid; i < HB_P * HB_T; i += NT) {
      const int pi = i / HB_T, ti = i - pi * HB_T, r = s_row[ti];
      s_de[pi][ti] = (pi < np && r >= 0) ? de[(IX)r * P + p0 + pi] : 0.f;
    }
    float na2[HB_T];                                      // -att2[t,b,a]
#pragma unroll
    for (int ti = 0; ti < HB_T; ++ti) {
      const int r = s_row[ti];
      na2[ti] = (a_ok && r >= 0) ? -att2[(IX)r * A + a] : 0.f;
    }
    float s1n[GP];
#pragma unroll
    for (int j = 0; j < GP; ++j) s1n[j] = (a_ok && j < np) ? ldf(a1 + (IX)j * A) : 0.f;
    __syncthreads();
    if (tid < HB_P) {
      float v = 0.f;
#pragma unroll
      for (int ti = 0; ti < HB_T; ++ti) v += s_de[tid][ti];
      s_dp[tid] = v;
      s_dpt[tid] += v;
    } else if (tid >= 32 && tid < 32 + HB_T) {
      float v = 0.f;
      for (int pi = 0; pi < HB_P; ++pi) v += s_de[pi][tid - 32];
      s_dt[tid - 32] = v;
    }
    __syncthreads();
    // sum_{t,p} [s > 0] de a2[t]  (the step part of dw_f) as four independent predicated-FMA chains
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int g = 0; g < HB_P / GP; ++g) {
      float s1c[GP];
#pragma unroll
      for (int j = 0; j < GP; ++j) s1c[j] = s1n[j];
      if (g + 1 < HB_P / GP) {
#pragma unroll
        for (int j = 0; j < GP; ++j) {
          const int pn = (g + 1) * GP + j;
          s1n[j] = (a_ok && pn < np) ? ldf(a1 + (IX)pn * A) : 0.f;
        }
      }
      // two locations per pass of the step loop: their compare -> predicated-add chains are independent, which
      // gives the scheduler something to issue while a predicate is in flight
#pragma unroll
      for (int j = 0; j < GP; j += 2) {
        const int pa = g * GP + j, pb = pa + 1;
        const bool has_b = (j + 1 < GP) && pb < np;
        if (pa < np) {
          const float sa = s1c[j], sb = has_b ? s1c[(j + 1 < GP) ? j + 1 : j] : -FLT_MAX;   // -FLT_MAX: never above a threshold
          const int pbs = has_b ? pb : pa;
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
          for (int ti = 0; ti < HB_T; ti += 4) {
            const float4 d = *reinterpret_cast<const float4*>(&s_de[pa][ti]);
            const float4 e = *reinterpret_cast<const float4*>(&s_de[pbs][ti]);
            if (sa > na2[ti]) { a0 += d.x; if (!GT) q0 = fmaf(d.x, na2[ti], q0); }
            if (sb > na2[ti]) { b0 += e.x; if (!GT) q2 = fmaf(e.x, na2[ti], q2); }
            if (sa > na2[ti + 1]) { a1 += d.y; if (!GT) q1 = fmaf(d.y, na2[ti + 1], q1); }
            if (sb > na2[ti + 1]) { b1 += e.y; if (!GT) q3 = fmaf(e.y, na2[ti + 1], q3); }
            if (sa > na2[ti + 2]) { a0 += d.z; if (!GT) q0 = fmaf(d.z, na2[ti + 2], q0); }
            if (sb > na2[ti + 2]) { b0 += e.z; if (!GT) q2 = fmaf(e.z, na2[ti + 2], q2); }
            if (sa > na2[ti + 3]) { a1 += d.w; if (!GT) q1 = fmaf(d.w, na2[ti + 3], q1); }
            if (sb > na2[ti + 3]) { b1 += e.w; if (!GT) q3 = fmaf(e.w, na2[ti + 3], q3); }
          }
          const float apa = a0 + a1;
          ap[pa] += apa;
          dw = fmaf(sa, fmaf(0.8f, apa, 0.2f * s_dp[pa]), dw);
          if (has_b) {
            const float apb = b0 + b1;
            ap[pb] += apb;
            dw = fmaf(sb, fmaf(0.8f, apb, 0.2f * s_dp[pb]), dw);
          }
        }
      }
    }
    if (a_ok && !GT) {
      dw = fmaf(-0.8f, (q0 + q1) + (q2 + q3), dw);        // q sums de * (-att2)
#pragma unroll
      for (int ti = 0; ti < HB_T; ++ti) dw = fmaf(-0.2f * na2[ti], s_dt[ti], dw);
    }
  }
  __syncthreads();
  if (a_ok) {
    const float w = wf[a];
#pragma unroll
    for (int pi = 0; pi < HB_P; ++pi)
      if (pi < np) stf(datt1 + ((IX)b * P + p0 + pi) * A + a, w * fmaf(0.8f, ap[pi], 0.2f * s_dpt[pi]));
    atomicAdd(dwf + a, dw);
  }
}

template <typename T, typename TO>
__global__ void __launch_bounds__(NT)
attn_ctx_all_kernel(const __grid_constant__ StepTable tab, int P, int C, int Tcap, const T* __restrict__ F,
                    const float* __restrict__ alphas, TO* __restrict__ ctx, TO* __restrict__ ctxT, int ldt) {
  extern __shared__ float s_al[];  // [P][CX_T]
  const int b = blockIdx.y, c = (blockIdx.x * NT + threadIdx.x) * 4;
  const bool vec_ok = (C & 3) == 0 && c + 4 <= C;
  int len = 0;
  while (len < tab.nsteps && tab.bs[len] > b) ++len;
  for (int t0 = 0; t0 < len; t0 += CX_T) {
    const int nt = min(CX_T, len - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < P * CX_T; i += NT) {
      const int pi = i / CX_T, ti = i % CX_T;
      s_al[i] = ti < nt ? alphas[((size_t)b * Tcap + t0 + ti) * P + pi] : 0.f;
    }
    __syncthreads();
    if (c < C) {
      float acc[CX_T][4];
#pragma unroll
      for (int k = 0; k < CX_T; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
      for (int p0 = 0; p0 < P; p0 += 4) {       // four locations' loads in flight per thread
        float fv[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int p = min(p0 + u, P - 1);
          const T* src = F + ((size_t)b * P + p) * C + c;
          if (vec_ok) ld4(src, fv[u]);
          else
            for (int j = 0; j < 4; ++j) fv[u][j] = (c + j < C) ? ldf(src + j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (p0 + u < P) {
#pragma unroll
            for (int k = 0; k < CX_T; k += 4) {
              const float4 a4 = *reinterpret_cast<const float4*>(s_al + (p0 + u) * CX_T + k);
              const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
              for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[k + q][j] = fmaf(aa[q], fv[u][j], acc[k + q][j]);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < CX_T; ++k) {
        if (k < nt) {
          const size_t n = (size_t)tab.off[t0 + k] + b;
          for (int j = 0; j < 4; ++j) {
            if (c + j < C) {
              if (ctx) stf(ctx + n * C + c + j, acc[k][j]);
              if (ctxT) stf(ctxT + (size_t)(c + j) * ldt + n, acc[k][j]);
            }
          }
        }
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

static int g_relayout_legacy = 0;
int st_debug_relayout_legacy(int on) { g_relayout_legacy = on ? 1 : 0; return ST_OK; }

static int relayout_any(const void* f, int f_bf16, int B, int C, int P, void* F, void* FT, int ldft, int out_bf16,
                        float* mean_f, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(f && F && mean_f, ST_ERR_NULL, "st_attn_relayout: NULL pointer");
  ST_REQUIRE(B >= 1 && C >= 1 && P >= 1 && (!FT || ldft >= B * P), ST_ERR_BAD_SHAPE,
             "st_attn_relayout: B=%d C=%d P=%d ldft=%d", B, C, P, ldft);
  dim3 grid((C + RL_CH - 1) / RL_CH, B);
  ST_REQUIRE(grid.y <= 65535, ST_ERR_BAD_SHAPE, "st_attn_relayout: batch too large");
  const size_t smem = sizeof(float) * RL_CH * (size_t)(P | 1);
  ST_REQUIRE(smem <= 200 * 1024, ST_ERR_BAD_SHAPE, "st_attn_relayout: P=%d too large", P);
  cudaStream_t s = as_stream(stream);
  const size_t tile_bytes = (size_t)RL_CH * P * (f_bf16 ? 2 : 4);
  const size_t ring_bytes = 2 * ((tile_bytes + 127) & ~(size_t)127);
  if (out_bf16 && !FT && C % RL_CH == 0 && ring_bytes <= 200 * 1024 && (reinterpret_cast<uintptr_t>(f) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(F) & 31) == 0 && !g_relayout_legacy) {
    int dev = 0, sms = 0;
    ST_CUDA_TRY(cudaGetDevice(&dev));
    ST_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long ntiles = (long long)B * (C / RL_CH);
    ST_REQUIRE(ntiles < (1ll << 30), ST_ERR_BAD_SHAPE, "st_attn_relayout: grid too large");
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(208 * 1024) / (ring_bytes + 1024)));
    const int nblk = (int)std::min<long long>(ntiles, (long long)per_sm * sms);
    if (f_bf16) {
      ST_CUDA_TRY(cudaFuncSetAttribute(relayout_bulk_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
      relayout_bulk_kernel<__nv_bfloat16><<<nblk, RL_NT, ring_bytes, s>>>((const __nv_bfloat16*)f, C, P, (int)ntiles, (__nv_bfloat16*)F, mean_f);
    } else {
      ST_CUDA_TRY(cudaFuncSetAttribute(relayout_bulk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
      relayout_bulk_kernel<float><<<nblk, RL_NT, ring_bytes, s>>>((const float*)f, C, P, (int)ntiles, (__nv_bfloat16*)F, mean_f);
    }
    ST_LAUNCH_TRY("relayout_bulk_kernel");
    return ST_OK;
  }
  if (out_bf16 && !FT && (P & 3) == 0 && C % RL_CH == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(F) & 3) == 0 && (reinterpret_cast<uintptr_t>(mean_f) & 7) == 0) {
    const size_t sm2 = sizeof(float2) * (RL_CH / 2) * (size_t)(P | 1);
    if (f_bf16) {
      ST_CUDA_TRY(cudaFuncSetAttribute(relayout_fast_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      relayout_fast_kernel<__nv_bfloat16><<<grid, NT, sm2, s>>>((const __nv_bfloat16*)f, C, P, (__nv_bfloat16*)F, mean_f);
    } else {
      ST_CUDA_TRY(cudaFuncSetAttribute(relayout_fast_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      relayout_fast_kernel<float><<<grid, NT, sm2, s>>>((const float*)f, C, P, (__nv_bfloat16*)F, mean_f);
    }
    ST_LAUNCH_TRY("relayout_fast_kernel");
    return ST_OK;
  }
#define ST_RELAYOUT(T, TI)                                                                               \
  do {                                                                                                   \
    auto kern = relayout_kernel<T, TI>;                                                                  \
    ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<grid, NT, smem, s>>>((const TI*)f, C, P, (T*)F, (T*)FT, ldft, mean_f);                        \
  } while (0)
  if (out_bf16) { if (f_bf16) ST_RELAYOUT(__nv_bfloat16, __nv_bfloat16); else ST_RELAYOUT(__nv_bfloat16, float); }
  else          { if (f_bf16) ST_RELAYOUT(float, __nv_bfloat16); else ST_RELAYOUT(float, float); }
#undef ST_RELAYOUT
  ST_LAUNCH_TRY("relayout_kernel");
  return ST_OK;
}

int st_grid_mean_bpc(const void* F, int f_bf16, int B, int P, int C, float* mean_f, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(F && mean_f, ST_ERR_NULL, "st_grid_mean_bpc: NULL pointer");
  ST_REQUIRE(B >= 1 && B <= 65535 && P >= 1 && C >= 1, ST_ERR_BAD_SHAPE, "st_grid_mean_bpc: B=%d P=%d C=%d", B, P, C);
  dim3 grid((C + NT - 1) / NT, B);
  if (f_bf16) grid_mean_bpc_kernel<__nv_bfloat16><<<grid, NT, 0, as_stream(stream)>>>((const __nv_bfloat16*)F, P, C, mean_f);
  else grid_mean_bpc_kernel<float><<<grid, NT, 0, as_stream(stream)>>>((const float*)F, P, C, mean_f);
  ST_LAUNCH_TRY("grid_mean_bpc_kernel");
  return ST_OK;
}

int st_attn_relayout(const float* f, int B, int C, int P, void* F, void* FT, int ldft, int out_bf16,
                     float* mean_f, st_stream_t stream) {
  return relayout_any(f, 0, B, C, P, F, FT, ldft, out_bf16, mean_f, stream);
}

int st_attn_relayout_bf16in(const void* f_bf16, int B, int C, int P, void* F, void* FT, int ldft, int out_bf16,
                            float* mean_f, st_stream_t stream) {
  return relayout_any(f_bf16, 1, B, C, P, F, FT, ldft, out_bf16, mean_f, stream);
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
                     float* datt2, void* datt2_bf16, float* gt, int act, st_stream_t stream) {
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
    sp.gt_out = gt;
    const int r = attn_stream_try(true, rows, in_bf16, act, sp, as_stream(stream));
    if (r != 0) return r < 0 ? r : ST_OK;
  }
  const size_t smem = sizeof(float) * (2 * (size_t)A + E + P);
  ST_REQUIRE(smem <= 48 * 1024, ST_ERR_BAD_SHAPE, "st_attn_step_bwd: A=%d E=%d P=%d too large", A, E, P);
  cudaStream_t s = as_stream(stream);
#define ST_LAUNCH_BWD(T, ACT)                                                                         \
  attn_step_bwd_kernel<T, ACT><<<rows, NT, smem, s>>>(P, A, E, (const T*)att1, (const T*)Fe, att2, wf, alphas, \
                                                      alpha_stride, dalpha, dalpha_stride, dctx, ld_dctx, de_out, datt2, \
                                                      (__nv_bfloat16*)datt2_bf16, gt)
  if (in_bf16) { if (act == 0) ST_LAUNCH_BWD(__nv_bfloat16, 0); else ST_LAUNCH_BWD(__nv_bfloat16, 1); }
  else         { if (act == 0) ST_LAUNCH_BWD(float, 0); else ST_LAUNCH_BWD(float, 1); }
#undef ST_LAUNCH_BWD
  ST_LAUNCH_TRY("attn_step_bwd_kernel");
  return ST_OK;
}

int st_attn_hoist_bwd(int nsteps, const int* batch_sizes_host, int P, int A, const void* att1, int in_bf16,
                      const float* att2, const float* de, const float* wf, void* datt1, void* datt1T, int ldt,
                      int out_bf16, float* dwf, const float* gt, int act, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(att1 && att2 && de && wf && datt1 && dwf, ST_ERR_NULL, "st_attn_hoist_bwd: NULL pointer");
  ST_REQUIRE(act == 0 || act == 1, ST_ERR_UNSUPPORTED, "st_attn_hoist_bwd: act=%d", act);
  ST_REQUIRE(in_bf16 == out_bf16, ST_ERR_UNSUPPORTED, "st_attn_hoist_bwd: mixed storage types");
  const int B = tab.bs[0];
  ST_REQUIRE(!datt1T || ldt >= B * P, ST_ERR_BAD_SHAPE, "st_attn_hoist_bwd: ldt=%d", ldt);
  dim3 grid((P + HB_P - 1) / HB_P, B);
  ST_REQUIRE(grid.y <= 65535, ST_ERR_BAD_SHAPE, "st_attn_hoist_bwd: batch too large");
  cudaStream_t s = as_stream(stream);
  ST_CUDA_TRY(cudaMemsetAsync(dwf, 0, sizeof(float) * A, s));
  const size_t smem = sizeof(float) * HB_P * HB_T + (datt1T ? (size_t)A * (HB_P + 1) * (in_bf16 ? 2 : 4) : 0);
  ST_REQUIRE(smem <= 200 * 1024, ST_ERR_BAD_SHAPE, "st_attn_hoist_bwd: A=%d too large", A);
  if (act == 0 && !datt1T) {     // LeakyReLU without the transposed copy: the compare + predicated-add form
    const dim3 g3(grid.x, grid.y, (A + NT - 1) / NT);
    const int Ntok = tab.off[tab.nsteps];
    const bool full = P % HB_P == 0 && A % NT == 0 && (long long)B * P * A < (1LL << 31) && (long long)Ntok * A < (1LL << 31) &&
                      (long long)Ntok * P < (1LL << 31);
#define ST_LAUNCH_L(T, FULL, GT) attn_hoist_bwd_lrelu_kernel<T, T, FULL, GT><<<g3, NT, 0, s>>>(tab, P, A, (const T*)att1, att2, de, wf, (T*)datt1, dwf)
#define ST_LAUNCH_LF(T) do { if (full) { if (gt) ST_LAUNCH_L(T, true, true); else ST_LAUNCH_L(T, true, false); } \
                             else { if (gt) ST_LAUNCH_L(T, false, true); else ST_LAUNCH_L(T, false, false); } } while (0)
    if (in_bf16) ST_LAUNCH_LF(__nv_bfloat16); else ST_LAUNCH_LF(float);
#undef ST_LAUNCH_LF
#undef ST_LAUNCH_L
    if (gt) {
      ST_LAUNCH_TRY("attn_hoist_bwd_lrelu_kernel");
      colsum_prod_kernel<<<dim3((A + 255) / 256, (Ntok + CSP_ROWS - 1) / CSP_ROWS), 256, 0, s>>>(dwf, att2, gt, Ntok, A);
    }
    ST_LAUNCH_TRY("attn_hoist_bwd_lrelu_kernel");
    return ST_OK;
  }
#define ST_LAUNCH_H(T, ACT)                                                                                 \
  do {                                                                                                      \
    auto kern = attn_hoist_bwd_kernel<T, T, ACT>;                                                           \
    ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    kern<<<grid, NT, smem, s>>>(tab, P, A, (const T*)att1, att2, de, wf, (T*)datt1, (T*)datt1T, ldt, dwf);   \
  } while (0)
  if (in_bf16) { if (act == 0) ST_LAUNCH_H(__nv_bfloat16, 0); else ST_LAUNCH_H(__nv_bfloat16, 1); }
  else         { if (act == 0) ST_LAUNCH_H(float, 0); else ST_LAUNCH_H(float, 1); }
#undef ST_LAUNCH_H
  ST_LAUNCH_TRY("attn_hoist_bwd_kernel");
  return ST_OK;
}

int st_attn_ctx_all(int nsteps, const int* batch_sizes_host, int P, int C, int T_cap, const void* F, int in_bf16,
                    const float* alphas, void* ctx, void* ctxT, int ldt, int out_bf16, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(F && alphas && (ctx || ctxT), ST_ERR_NULL, "st_attn_ctx_all: NULL pointer");
  ST_REQUIRE(T_cap >= nsteps && (!ctxT || ldt >= tab.off[nsteps]), ST_ERR_BAD_SHAPE,
             "st_attn_ctx_all: T_cap=%d ldt=%d", T_cap, ldt);
  const size_t smem = sizeof(float) * CX_T * (size_t)P;
  ST_REQUIRE(smem <= 48 * 1024, ST_ERR_BAD_SHAPE, "st_attn_ctx_all: P=%d too large", P);
  dim3 grid((C + 4 * NT - 1) / (4 * NT), tab.bs[0]);
  cudaStream_t s = as_stream(stream);
  ST_REQUIRE(in_bf16 || !out_bf16, ST_ERR_UNSUPPORTED, "st_attn_ctx_all: fp32 features with bf16 output");
  if (in_bf16 && out_bf16)
    attn_ctx_all_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, NT, smem, s>>>(
        tab, P, C, T_cap, (const __nv_bfloat16*)F, alphas, (__nv_bfloat16*)ctx, (__nv_bfloat16*)ctxT, ldt);
  else if (in_bf16)
    attn_ctx_all_kernel<__nv_bfloat16, float><<<grid, NT, smem, s>>>(
        tab, P, C, T_cap, (const __nv_bfloat16*)F, alphas, (float*)ctx, (float*)ctxT, ldt);
  else
    attn_ctx_all_kernel<float, float><<<grid, NT, smem, s>>>(tab, P, C, T_cap, (const float*)F, alphas,
                                                             (float*)ctx, (float*)ctxT, ldt);
  ST_LAUNCH_TRY("attn_ctx_all_kernel");
  return ST_OK;
}

int st_attn_embed_q(int nsteps, const int* batch_sizes_host, int P, int E, int T_cap, const float* alphas,
                    const float* dctx, int ld_dctx, void* Q_bf16, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(alphas && dctx && Q_bf16, ST_ERR_NULL, "st_attn_embed_q: NULL pointer");
  ST_REQUIRE(P >= 1 && E >= 1 && T_cap >= nsteps && ld_dctx >= E, ST_ERR_BAD_SHAPE,
             "st_attn_embed_q: P=%d E=%d T_cap=%d ld_dctx=%d", P, E, T_cap, ld_dctx);
  dim3 grid((P + QP - 1) / QP, tab.bs[0]);
  attn_embed_q_kernel<<<grid, NT, 0, as_stream(stream)>>>(tab, P, E, T_cap, alphas, dctx, ld_dctx,
                                                        reinterpret_cast<__nv_bfloat16*>(Q_bf16));
  ST_LAUNCH_TRY("attn_embed_q_kernel");
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
