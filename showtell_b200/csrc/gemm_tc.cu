// bf16 GEMM on the 5th-generation tensor cores: tcgen05.mma (UMMA) with the accumulator in TMEM,
// operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), mbarrier
// pipelines between the roles, persistent CTAs (one per SM) walking a static tile schedule.
//
//     D[M,N] = A[M,K] . B[N,K]^T          A, B bf16 row-major with K contiguous ("K-major")
//
// This is the nn.Linear shape (weights are (out, in)); the transposed products of the backward
// pass are brought to the same form by the caller with transposed bf16 copies.
//
// CTA = 384 threads: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane each), warp 2 =
// TMEM allocator, warps 4..11 = epilogue (warp w reads TMEM lanes 32*(w%4)..+31 = 32 rows and one
// 64-column half of the 128x128 tile; two warps per SM sub-partition hide each other's latency).  6-stage smem ring (A 16 KB + B 16 KB per stage), 2 TMEM accumulator stages
// (2 x 128 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Epilogues (template parameter):
//   EPI_STORE   C = alpha*acc + bias[n]  as fp32 or bf16                     (W_ih / att1 hoists, dX, dW)
//   EPI_CE_FWD  per-row online (max, sum-exp) over this tile's 128 vocabulary columns + the
//               target logit: the (N,V) logits are never written              (rnn.py:33 + main.py:149)
//   EPI_CE_BWD  recomputes the logits tile and writes dlogits = (softmax - onehot)*scale as bf16,
//               row-major and transposed, the operands of the two backward GEMMs
#include <cuda.h>
#include <cuda_bf16.h>

#include <cfloat>

#include "common.cuh"
#include "tc_common.cuh"

namespace st {
namespace {

constexpr int BM = 128, BK = 64, UK = 16;
constexpr int ACC_STAGES = 2, NTHREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr uint32_t A_BYTES = BM * BK * 2;
// Tile width BN (template): 256 where the problem has enough tiles, else 128.  An SM ingests ~64 B/clk
// from L2; a 128x128x64 k-block needs 32 KB for 256 MMA clocks (128 B/clk: port-bound at <= 50 % of the
// tensor pipe), a 128x256x64 k-block 48 KB for 512 clocks (96 B/clk: <= 67 %).
template <int BN> struct TileCfg {
  static constexpr uint32_t B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t TMEM_COLS = ACC_STAGES * BN;
  static constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 256;
};

enum { EPI_STORE = 0, EPI_CE_FWD = 1, EPI_CE_BWD = 2 };

struct TcParams {
  int M, N, K;
  // EPI_STORE
  void* C;
  int ldc, c_bf16;
  const float* bias;  // [N] or null (all epilogues)
  float alpha, beta;   // EPI_STORE fp32: C = alpha*acc + bias + beta*C
  // CE
  const int64_t* target;  // [M]
  float *pmax, *psum, *tlogit;  // fwd: (M, npart) partials, (M) target logit
  int npart;
  const float* lse;  // bwd: [M]
  float scale;
  __nv_bfloat16 *P, *PT;  // bwd: (M, ldp) and (N, ldpt)
  int ldp, ldpt;
};

template <int EPI, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcParams p) {
  constexpr int STAGES = TileCfg<BN>::STAGES;
  constexpr uint32_t STAGE_BYTES = TileCfg<BN>::STAGE_BYTES, TMEM_COLS = TileCfg<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = (p.M + BM - 1) / BM, nt = (p.N + BN - 1) / BN;
  const int ntiles = mt * nt, kb = (p.K + BK - 1) / BK;
  const bool nfast = nt < mt;  // consecutive tiles walk the shorter dimension: the long operand streams once

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {  // ---------------------------------------------------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (nfast ? tile / nt : tile % mt) * BM, n0 = (nfast ? tile % nt : tile / mt) * BN;
        for (int k = 0; k < kb; ++k) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* a = smem + (size_t)stage * STAGE_BYTES;
          tma_load_2d(a, &tmA, k * BK, m0, &full[stage]);
          tma_load_2d(a + A_BYTES, &tmB, k * BK, n0, &full[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = umma_idesc(BM, BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int k = 0; k < kb; ++k) {
          mbar_wait(&full[stage], phase);  // TMA bytes have landed
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / UK; ++kk)
            tc_mma(d_tmem, umma_desc_k128(a_addr + kk * UK * 2), umma_desc_k128(b_addr + kk * UK * 2), idesc,
                   (k | kk) != 0);
          tc_commit(&empty[stage]);  // smem slot free once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull[acc]);  // accumulator complete
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {  // ------------------------------------------------------ epilogue
    // 8 warps: warp w owns TMEM lanes 32*(w%4)..+31 (32 rows) and one half of the tile's BN columns.
    const int ew = warp & 3, half = (warp - 4) >> 2;
    constexpr float LOG2E = 1.4426950408889634f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m0 = (nfast ? tile / nt : tile % mt) * BM, n0 = (nfast ? tile % nt : tile / mt) * BN;
      const int row = m0 + ew * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);

      float run_m = -FLT_MAX, run_s = 0.f, tl = 0.f;  // EPI_CE_FWD
      bool have_tl = false;
      int tgt = -1;
      float lse_l2 = 0.f;
      if (EPI != EPI_STORE && row_ok) tgt = (int)p.target[row];
      if (EPI == EPI_CE_BWD && row_ok) lse_l2 = p.lse[row] * LOG2E;

#pragma unroll 1
      for (int cc = 0; cc < BN / 64; ++cc) {
        const int c = half * (BN / 64) + cc;
        float v[32];
        tmem_ld32(tbase + c * 32, v);
        const int nb = n0 + c * 32;
        const bool full = nb + 32 <= p.N;
        if (EPI != EPI_STORE && nb >= p.N) continue;  // chunk entirely past the vocabulary
        // bias: uniform (same address in every lane) 128-bit loads on full chunks
        if (p.bias) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
              if (EPI == EPI_STORE) {
                v[j] = fmaf(p.alpha, v[j], b4.x); v[j + 1] = fmaf(p.alpha, v[j + 1], b4.y);
                v[j + 2] = fmaf(p.alpha, v[j + 2], b4.z); v[j + 3] = fmaf(p.alpha, v[j + 3], b4.w);
              } else {
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float bj = (nb + j < p.N) ? p.bias[nb + j] : 0.f;
              v[j] = (EPI == EPI_STORE) ? fmaf(p.alpha, v[j], bj) : v[j] + bj;
            }
          }
        } else if (EPI == EPI_STORE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
        }

        if (EPI == EPI_STORE) {
          if (row_ok) {
            if (p.c_bf16) {
              __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + nb;
              if (full && (p.ldc & 7) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint32_t w[4];
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(v[j + 2 * q], v[j + 2 * q + 1]);
                    w[q] = *reinterpret_cast<uint32_t*>(&h);
                  }
                  *reinterpret_cast<uint4*>(out + j) = make_uint4(w[0], w[1], w[2], w[3]);
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (nb + j < p.N) out[j] = __float2bfloat16(v[j]);
              }
            } else {
              float* out = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + nb;
              if (full && (p.ldc & 3) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                  if (p.beta != 0.f) {
                    const float4 c4 = *reinterpret_cast<const float4*>(out + j);
                    o.x = fmaf(p.beta, c4.x, o.x); o.y = fmaf(p.beta, c4.y, o.y);
                    o.z = fmaf(p.beta, c4.z, o.z); o.w = fmaf(p.beta, c4.w, o.w);
                  }
                  *reinterpret_cast<float4*>(out + j) = o;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (nb + j < p.N) out[j] = (p.beta != 0.f) ? fmaf(p.beta, out[j], v[j]) : v[j];
              }
            }
          }
        } else if (EPI == EPI_CE_FWD) {
          if (!full) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j >= p.N) v[j] = -FLT_MAX;
          }
          if (tgt >= nb && tgt < nb + 32) {  // rare: pick the target logit
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j == tgt) tl = v[j];
            have_tl = true;
          }
          float cm = v[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) cm = fmaxf(cm, v[j]);
          const float nm = fmaxf(run_m, cm);
          const float nml2 = nm * LOG2E;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) s += exp2f(fmaf(v[j], LOG2E, -nml2));  // exp(-huge) = 0 on masked columns
          run_s = fmaf(run_s, exp2f(fmaf(run_m, LOG2E, -nml2)), s);
          run_m = nm;
        } else {  // EPI_CE_BWD: d = (softmax - onehot) * scale
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = exp2f(fmaf(v[j], LOG2E, -lse_l2)) * p.scale;
          if (tgt >= nb && tgt < nb + 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j == tgt) v[j] -= p.scale;
          }
          if (row_ok) {
            __nv_bfloat16* out = p.P + (size_t)row * p.ldp + nb;
            if (full && (p.ldp & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(v[j + 2 * q], v[j + 2 * q + 1]);
                  w[q] = *reinterpret_cast<uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(out + j) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) out[j] = __float2bfloat16(v[j]);
            }
            if (p.PT) {  // transposed copy: lanes are consecutive rows -> 64 B contiguous per column
              __nv_bfloat16* outT = p.PT + (size_t)nb * p.ldpt + row;
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; ++j) outT[(size_t)j * p.ldpt] = __float2bfloat16(v[j]);
              } else {
                for (int j = 0; j < 32; ++j)
                  if (nb + j < p.N) outT[(size_t)j * p.ldpt] = __float2bfloat16(v[j]);
              }
            }
          }
        }
      }
      if (EPI == EPI_CE_FWD && row_ok) {
        const int part = (n0 / BN) * 2 + half;
        p.pmax[(size_t)row * p.npart + part] = run_m;
        p.psum[(size_t)row * p.npart + part] = run_s;
        if (have_tl) p.tlogit[row] = tl;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// lse[m] = log sum_v exp(logit) from the per-tile partials; loss_sum += lse - target logit.
__global__ void ce_combine_kernel(int M, int npart, const float* __restrict__ pmax, const float* __restrict__ psum,
                                  const float* __restrict__ tlogit, float* __restrict__ lse,
                                  float* __restrict__ loss_sum) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (m < M) {
    float mx = -FLT_MAX;
    for (int j = 0; j < npart; ++j) mx = fmaxf(mx, pmax[(size_t)m * npart + j]);
    float s = 0.f;
    for (int j = 0; j < npart; ++j) s += psum[(size_t)m * npart + j] * expf(pmax[(size_t)m * npart + j] - mx);
    const float l = mx + logf(s);
    lse[m] = l;
    contrib = l - tlogit[m];
  }
  contrib = warp_sum(contrib);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss_sum, contrib);
}

// fp32 (rows, cols) -> bf16 copy and/or bf16 transpose (cols, rows).  32x32 tiles through smem.
__global__ void cast_bf16_kernel(const float* __restrict__ src, int rows, int cols, int lds,
                                 __nv_bfloat16* __restrict__ dst, int ldd, __nv_bfloat16* __restrict__ dstT,
                                 int lddT) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = (r < rows && c < cols) ? src[(size_t)r * lds + c] : 0.f;
    tile[i][threadIdx.x] = v;
    if (dst && r < rows && c < cols) dst[(size_t)r * ldd + c] = __float2bfloat16(v);
  }
  if (!dstT) return;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dstT[(size_t)c * lddT + r] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

template <int EPI, int BN>
int launch_tc_bn(const TcParams& p, const void* A, int lda, const void* B, int ldb, cudaStream_t s, int sms) {
  CUtensorMap tmA, tmB;
  ST_TRY(make_tmap(&tmA, A, p.M, p.K, lda, BM, "A"));
  ST_TRY(make_tmap(&tmB, B, p.N, p.K, ldb, BN, "B"));
  auto kern = gemm_tc_kernel<EPI, BN>;
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<BN>::SMEM_BYTES));
  const int ntiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = ntiles < sms ? ntiles : sms;
  kern<<<grid, NTHREADS, TileCfg<BN>::SMEM_BYTES, s>>>(tmA, tmB, p);
  ST_LAUNCH_TRY("gemm_tc_kernel");
  return ST_OK;
}

// Tile width by estimated time: rounds of the persistent grid x clocks per k-block (measured ~600 for a
// 128-wide and ~750 for a 256-wide tile, both L2-port-bound).
inline int pick_bn(int M, int N, int sms) {
  if (N <= 128) return 128;
  const long mt = (M + BM - 1) / BM;
  const long r128 = (mt * ((N + 127) / 128) + sms - 1) / sms, r256 = (mt * ((N + 255) / 256) + sms - 1) / sms;
  return (r256 * 750 <= r128 * 600) ? 256 : 128;
}

template <int EPI>
int launch_tc(const TcParams& p, const void* A, int lda, const void* B, int ldb, cudaStream_t s, int bn = 0) {
  ST_REQUIRE(p.M >= 1 && p.N >= 1 && p.K >= 1, ST_ERR_BAD_SHAPE, "gemm_bf16: M=%d N=%d K=%d", p.M, p.N, p.K);
  int sms = 0;
  ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
  if (bn == 0) bn = pick_bn(p.M, p.N, sms);
  return bn == 256 ? launch_tc_bn<EPI, 256>(p, A, lda, B, ldb, s, sms) : launch_tc_bn<EPI, 128>(p, A, lda, B, ldb, s, sms);
}

}  // namespace
}  // namespace st

extern "C" {

int st_gemm_bf16(int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                 int c_is_bf16, const float* bias, float alpha, float beta, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(C != nullptr, ST_ERR_NULL, "st_gemm_bf16: C is NULL");
  ST_REQUIRE(ldc >= N, ST_ERR_BAD_SHAPE, "st_gemm_bf16: ldc=%d < N=%d", ldc, N);
  TcParams p{};
  p.M = M; p.N = N; p.K = K;
  ST_REQUIRE(beta == 0.f || !c_is_bf16, ST_ERR_UNSUPPORTED, "st_gemm_bf16: beta needs an fp32 C");
  p.C = C; p.ldc = ldc; p.c_bf16 = c_is_bf16; p.bias = bias; p.alpha = alpha; p.beta = beta;
  return launch_tc<EPI_STORE>(p, A, lda, B, ldb, as_stream(stream));
}

int st_vocab_ce_fwd(int M, int V, int H, const void* Hs, int ldh, const void* Wv, int ldw, const float* bv,
                    const int64_t* target, float* part_max, float* part_sum, float* tlogit, float* lse,
                    float* loss_sum, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(target && part_max && part_sum && tlogit && lse && loss_sum, ST_ERR_NULL, "st_vocab_ce_fwd: NULL pointer");
  cudaStream_t s = as_stream(stream);
  TcParams p{};
  p.M = M; p.N = V; p.K = H;
  p.bias = bv; p.target = target; p.pmax = part_max; p.psum = part_sum; p.tlogit = tlogit;
  int sms = 0;
  ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
  const int bn = pick_bn(M, V, sms);
  p.npart = 2 * ((V + bn - 1) / bn);   // <= st_vocab_ce_parts(V)
  ST_CUDA_TRY(cudaMemsetAsync(loss_sum, 0, sizeof(float), s));
  ST_TRY(launch_tc<EPI_CE_FWD>(p, Hs, ldh, Wv, ldw, s, bn));
  ce_combine_kernel<<<(M + 127) / 128, 128, 0, s>>>(M, p.npart, part_max, part_sum, tlogit, lse, loss_sum);
  ST_LAUNCH_TRY("ce_combine_kernel");
  return ST_OK;
}

int st_vocab_ce_parts(int V) { return 2 * ((V + 127) / 128); }

int st_vocab_ce_bwd(int M, int V, int H, const void* Hs, int ldh, const void* Wv, int ldw, const float* bv,
                    const int64_t* target, const float* lse, float scale, void* P, int ldp, void* PT, int ldpt,
                    st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(target && lse && P, ST_ERR_NULL, "st_vocab_ce_bwd: NULL pointer");
  ST_REQUIRE(ldp >= V && (!PT || ldpt >= M), ST_ERR_BAD_SHAPE, "st_vocab_ce_bwd: ldp=%d ldpt=%d", ldp, ldpt);
  TcParams p{};
  p.M = M; p.N = V; p.K = H;
  p.bias = bv; p.target = target; p.lse = lse; p.scale = scale;
  p.P = reinterpret_cast<__nv_bfloat16*>(P); p.ldp = ldp;
  p.PT = reinterpret_cast<__nv_bfloat16*>(PT); p.ldpt = ldpt;
  return launch_tc<EPI_CE_BWD>(p, Hs, ldh, Wv, ldw, as_stream(stream));
}

int st_cast_bf16(const float* src, int rows, int cols, int lds, void* dst, int ldd, void* dstT, int lddT,
                 st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(src && (dst || dstT), ST_ERR_NULL, "st_cast_bf16: NULL pointer");
  ST_REQUIRE(rows >= 1 && cols >= 1 && lds >= cols && (!dst || ldd >= cols) && (!dstT || lddT >= rows),
             ST_ERR_BAD_SHAPE, "st_cast_bf16: rows=%d cols=%d lds=%d ldd=%d lddT=%d", rows, cols, lds, ldd, lddT);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  ST_REQUIRE(grid.y <= 65535, ST_ERR_BAD_SHAPE, "st_cast_bf16: too many rows");
  cast_bf16_kernel<<<grid, block, 0, as_stream(stream)>>>(src, rows, cols, lds,
                                                          reinterpret_cast<__nv_bfloat16*>(dst), ldd,
                                                          reinterpret_cast<__nv_bfloat16*>(dstT), lddT);
  ST_LAUNCH_TRY("cast_bf16_kernel");
  return ST_OK;
}

}  // extern "C"
