// bf16 GEMM on the 5th-generation tensor cores: tcgen05.mma (UMMA) with the accumulator in TMEM,
// operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), mbarrier
// pipelines between the roles, persistent CTAs walking a static schedule.
//
//     D[M,N] = A[M,K] . B[N,K]^T          A, B bf16 row-major with K contiguous ("K-major")
//
// This is the nn.Linear shape (weights are (out, in)); the transposed products of the backward
// pass are brought to the same form by the caller with transposed bf16 copies.
//
// Two kernels share one epilogue:
//   gemm_tc2_kernel  (large problems) -- CTA PAIRS (cluster 2x1x1, tcgen05.mma.cta_group::2): one 256x256
//       tile per pair, each CTA holds its 128 rows of A and HALF of the B tile (128 rows), so a 64-deep
//       k-block costs every SM 32 KB of L2 ingress for 512 tensor-pipe clocks (64 B/clk -- the 1-CTA
//       128x256 tile needs 96 B/clk and is port-bound at <= 67 %).  The leader CTA issues the MMAs for
//       both; TMA of both CTAs completes on the leader's mbarrier; tcgen05.commit multicasts "slot free"
//       / "accumulator full" to both.  Schedule: data-parallel tiles, or STREAM-K (the k-blocks of all
//       tiles are cut into equal contiguous ranges, one per pair; tiles cut by a range boundary are
//       accumulated with vector fp32 reductions into a zeroed C) when whole tiles would leave a large
//       part of the GPU idle (dX / dW of the vocabulary projection: 40 / 80 tiles on 74 pairs).
//   gemm_tc_kernel   (small problems: per-step products of the attention loop) -- one CTA per 128x128
//       or 128x256 tile.
// CTA = 384 threads: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane each), warp 2 =
// TMEM allocator, warps 4..11 = epilogue (warp w reads TMEM lanes 32*(w%4)..+31 = 32 rows and one
// half of the tile's columns; two warps per SM sub-partition hide each other's latency).  2 TMEM
// accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Epilogues (template parameter); a warp's 32x32 chunk goes through a private swizzled shared-memory
// staging tile so that every global store instruction writes whole 128-byte lines (the TMEM layout
// has one ROW per lane: storing straight from registers would touch 32 lines per instruction):
//   EPI_STORE   C = alpha*acc + bias[n] (+ beta*C)  as fp32 or bf16              (hoists, dX, dW)
//   EPI_CE_FWD  per-row online (max, sum-exp) over this tile's vocabulary columns + the
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
constexpr uint32_t STG_BYTES = 8 * 4096;       // epilogue staging: 4 KB per epilogue warp
// Tile width BN (template, 1-CTA kernel): 256 where the problem has enough tiles, else 128.
template <int BN> struct TileCfg {
  static constexpr uint32_t B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t TMEM_COLS = ACC_STAGES * BN;
  static constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + STG_BYTES + 256;
};
// CTA-pair kernel: per CTA and stage A 128x64 + B 128x64 (its half of the 256-row B tile).
struct PairCfg {
  static constexpr int BN = 256, STAGES = 5;
  static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES, TMEM_COLS = ACC_STAGES * BN;
  static constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + STG_BYTES + 256;
};

enum { EPI_STORE = 0, EPI_CE_FWD = 1, EPI_CE_BWD = 2, EPI_TOPK = 3, EPI_TOPS = 4 };   // TOPS: screening, ST_SCREEN_SLOTS per part
constexpr int TOPK_SLOTS = 8;   // per-row candidates kept per epilogue warp (K <= 8 on the fused path)
int g_variant = 0;   // st_debug_gemm_variant: 0 = choose, 128 / 256 = single-CTA tile width, 2 = CTA pairs
int g_streamk = 1;   // st_debug_gemm_variant(v | 0x1000) turns stream-K off
int g_tf32_bn192 = 1; // st_debug_gemm_variant(v | 0x20000) turns the 192-column tiles of the 3xTF32 product off
int g_mn3d = 1;      // st_debug_gemm_variant(v | 0x2000): MN-major operands through 2-D boxes (A/B timing)
int g_sm_limit = 0;  // st_gemm_set_sm_limit: cap on the persistent grids (0 = all SMs)
int g_c_zeroed = 0;  // st_gemm_set_c_zeroed: the caller has already cleared C (stream-K launches skip their memset)
inline int gemm_sms(int* sms) {
  ST_TRY(st_device_info(sms, nullptr, nullptr, nullptr));
  if (g_sm_limit > 0 && g_sm_limit < *sms) *sms = g_sm_limit;
  return ST_OK;
}

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
  float* tk_val;          // EPI_TOPK: (M, npart, tk_k) partial top values (descending) ...
  int32_t* tk_idx;        // ... and their column indices (ties: lower index first); npart as for the CE partials
  int tk_k;               // candidates written per (row, part): the caller's K (<= TOPK_SLOTS).  With pmax / psum set the
                          // epilogue also keeps the soft-max partials (running max, sum of exp) of its columns.
  int streamk;            // pair kernel: 1 = stream-K schedule (C zeroed by the launcher)
  float* zero_word;       // cleared by the kernel's first thread (the CE forward's loss accumulator: no memset in front)
  int a3d, b3d;           // MN-major operand given as the 3-D tensor map (make_tmap_mn3d): one TMA operation per k-block
  int mc;                 // 1: launched as clusters of two CTAs that work on vertically adjacent tiles (same columns) and
                          // SHARE the B tile: each loads half of it and multicasts it into both CTAs' shared memory
};

__device__ __forceinline__ float ex2_fast(float x) {   // 2^x, MUFU only (inputs here are <= ~0: no range fix-up needed)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// One epilogue warp's part of an accumulator tile: 32 rows (TMEM lanes, row = row0 + lane) x
// `nchunks` chunks of 32 columns starting at column n0 + 32*chunk0 of the problem.  `stg` = this warp's
// 4 KB staging tile.  `partial`: the accumulator holds only a slice of K (stream-K) -> reduce into C;
// `first_k`: the slice starts at k = 0 (adds the bias).  `part`: index of this warp's CE partial.
template <int EPI>
__device__ __forceinline__ void epilogue_warp(const TcParams& p, uint32_t tbase, int row0, int n0, int chunk0, int nchunks,
                                              uint8_t* stg, int lane, bool partial, bool first_k, int part) {
  constexpr float LOG2E = 1.4426950408889634f;
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  const uint32_t stg_s = smem_u32(stg);
  float run_m = -FLT_MAX, run_s = 0.f, tl = 0.f;  // EPI_CE_FWD
  constexpr bool TOPX = EPI == EPI_TOPK || EPI == EPI_TOPS;
  constexpr int TS = (EPI == EPI_TOPS) ? (int)ST_SCREEN_SLOTS : TOPK_SLOTS;   // candidates kept per (row, part)
  float tkv[TS];                                    // EPI_TOPK / TOPS: this row's best TS logits of the warp's columns,
  int tki[TS];                                      // descending; equal values keep the lower column first
#pragma unroll
  for (int q = 0; q < TS; ++q) { tkv[q] = -FLT_MAX; tki[q] = 0x7fffffff; }
  bool have_tl = false;
  int tgt = -1;
  float lse_l2 = 0.f;
  if ((EPI == EPI_CE_FWD || EPI == EPI_CE_BWD) && row_ok) tgt = (int)p.target[row];
  if (EPI == EPI_CE_BWD && row_ok) lse_l2 = p.lse[row] * LOG2E;
  const float* bias = (p.bias && first_k) ? p.bias : nullptr;
  // 16-byte vector stores need an aligned base and leading dimension (views into wider buffers may have neither)
  const bool c_vec = EPI == EPI_STORE && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (p.ldc & (p.c_bf16 ? 7 : 3)) == 0;
  const bool p_vec = EPI == EPI_CE_BWD && (reinterpret_cast<uintptr_t>(p.P) & 15) == 0 && (p.ldp & 7) == 0;

#pragma unroll 1
  for (int cc = 0; cc < nchunks; ++cc) {
    const int c = chunk0 + cc;
    const int nb = n0 + c * 32;
    if (nb >= p.N) break;                          // chunk entirely past the last column
    float v[32];
    tmem_ld32(tbase + c * 32, v);
    const bool full = nb + 32 <= p.N;
    // bias: uniform (same address in every lane) 128-bit loads on full chunks
    if (bias) {
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + nb + j));
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
          const float bj = (nb + j < p.N) ? bias[nb + j] : 0.f;
          v[j] = (EPI == EPI_STORE) ? fmaf(p.alpha, v[j], bj) : v[j] + bj;
        }
      }
    } else if (EPI == EPI_STORE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
    }

    if (EPI == EPI_STORE) {
      if (p.c_bf16) {
        __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(p.C);
        if (full && c_vec) {
          // staging tile: 32 rows x 64 B, 16-byte groups XOR-swizzled by (row >> 1) & 3
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t a = stg_s + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf2(v[8 * g], v[8 * g + 1])),
                         "r"(pack_bf2(v[8 * g + 2], v[8 * g + 3])), "r"(pack_bf2(v[8 * g + 4], v[8 * g + 5])),
                         "r"(pack_bf2(v[8 * g + 6], v[8 * g + 7]))
                         : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {             // 8 rows x 64 B per instruction
            const int r = i * 8 + (lane >> 2), g = lane & 3;
            uint4 o;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                         : "r"(stg_s + r * 64 + ((g ^ ((r >> 1) & 3)) << 4)));
            if (row0 + r < p.M) *reinterpret_cast<uint4*>(C + (size_t)(row0 + r) * p.ldc + nb + g * 8) = o;
          }
          __syncwarp();
        } else if (row_ok) {
          __nv_bfloat16* out = C + (size_t)row * p.ldc + nb;
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) out[j] = __float2bfloat16(v[j]);
        }
      } else {
        float* C = reinterpret_cast<float*>(p.C);
        if (full && c_vec) {
          // staging tile: 32 rows x 128 B, 16-byte groups XOR-swizzled by row & 7
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint32_t a = stg_s + lane * 128 + ((g ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * g]), "f"(v[4 * g + 1]),
                         "f"(v[4 * g + 2]), "f"(v[4 * g + 3])
                         : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {             // 4 rows x 128 B per instruction
            const int r = i * 4 + (lane >> 3), g = lane & 7;
            float4 o;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                         : "r"(stg_s + r * 128 + ((g ^ (r & 7)) << 4)));
            if (row0 + r < p.M) {
              float* out = C + (size_t)(row0 + r) * p.ldc + nb + g * 4;
              if (partial) {
                red_add_v4(out, o);
              } else {
                if (p.beta != 0.f) {
                  const float4 c4 = *reinterpret_cast<const float4*>(out);
                  o.x = fmaf(p.beta, c4.x, o.x); o.y = fmaf(p.beta, c4.y, o.y);
                  o.z = fmaf(p.beta, c4.z, o.z); o.w = fmaf(p.beta, c4.w, o.w);
                }
                *reinterpret_cast<float4*>(out) = o;
              }
            }
          }
          __syncwarp();
        } else if (row_ok) {
          float* out = C + (size_t)row * p.ldc + nb;
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) {
              if (partial) atomicAdd(out + j, v[j]);
              else out[j] = (p.beta != 0.f) ? fmaf(p.beta, out[j], v[j]) : v[j];
            }
        }
      }
    } else if (TOPX) {
      // rnn.py:51,63,90-91: arg-max / top-K of the vocabulary logits without writing them.  Columns arrive in
      // increasing order, so a strict '>' keeps the lower index among equal values (torch.max's first-maximum rule).
      if (EPI == EPI_TOPK && p.pmax) {   // beam_search.py:84-88 ranks by soft-max probability: keep the normaliser's partials too
        float cm = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || nb + j < p.N) cm = fmaxf(cm, v[j]);
        const float nm = fmaxf(run_m, cm);
        float s0 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || nb + j < p.N) s0 += expf(v[j] - nm);
        run_s = fmaf(run_s, expf(run_m - nm), s0);
        run_m = nm;
      }
      if (EPI == EPI_TOPS) {
        // Screening keeps APPROXIMATE logits (the caller re-scores the survivors exactly and widens its threshold by
        // the perturbation, decode.cu): the column's position inside the 128-column part replaces the low 7 mantissa
        // bits (as 127 - position: among equal values the lower column is the larger number), so the three best are
        // five min / max per column -- no index registers, no divergent insertion (the (value, index) insertion
        // below made this epilogue, not the MMAs, the bound of the screening product: 86 us against 55).
        static_assert(EPI != EPI_TOPS || TS >= 3, "screening keeps 3 per part");
        const uint32_t e = 127u ^ (uint32_t)(nb & 127);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float xr = (full || nb + j < p.N) ? v[j] : -FLT_MAX;
          const float x = __uint_as_float(((__float_as_uint(xr) & 0xffffff80u) | e) ^ (uint32_t)j);
          const float t = fminf(tkv[0], x);
          tkv[2] = fmaxf(tkv[2], fminf(tkv[1], t));
          tkv[1] = fmaxf(tkv[1], t);
          tkv[0] = fmaxf(tkv[0], x);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = (full || nb + j < p.N) ? v[j] : -FLT_MAX;
          if (x > tkv[TS - 1]) {
            tkv[TS - 1] = x; tki[TS - 1] = nb + j;
#pragma unroll
            for (int q = TS - 1; q > 0; --q) {
              if (tkv[q] > tkv[q - 1]) {
                const float tv = tkv[q]; tkv[q] = tkv[q - 1]; tkv[q - 1] = tv;
                const int ti = tki[q]; tki[q] = tki[q - 1]; tki[q - 1] = ti;
              }
            }
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
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {            // exp(-huge) = 0 on masked columns
        s0 += ex2_fast(fmaf(v[j], LOG2E, -nml2));
        s1 += ex2_fast(fmaf(v[j + 1], LOG2E, -nml2));
      }
      run_s = fmaf(run_s, ex2_fast(fmaf(run_m, LOG2E, -nml2)), s0 + s1);
      run_m = nm;
    } else {  // EPI_CE_BWD: d = (softmax - onehot) * scale
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = ex2_fast(fmaf(v[j], LOG2E, -lse_l2)) * p.scale;
      if (tgt >= nb && tgt < nb + 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nb + j == tgt) v[j] -= p.scale;
      }
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = pack_bf2(v[2 * j], v[2 * j + 1]);
      if (full && p_vec) {                        // row-major copy through the staging tile (see EPI_STORE bf16)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t a = stg_s + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * g]), "r"(w[4 * g + 1]),
                       "r"(w[4 * g + 2]), "r"(w[4 * g + 3])
                       : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2), g = lane & 3;
          uint4 o;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                       : "r"(stg_s + r * 64 + ((g ^ ((r >> 1) & 3)) << 4)));
          if (row0 + r < p.M) *reinterpret_cast<uint4*>(p.P + (size_t)(row0 + r) * p.ldp + nb + g * 8) = o;
        }
        __syncwarp();
      } else if (row_ok) {
        __nv_bfloat16* out = p.P + (size_t)row * p.ldp + nb;
        for (int j = 0; j < 32; ++j)
          if (nb + j < p.N) out[j] = __float2bfloat16(v[j]);
      }
      // (Reading the transposed chunk back from the staging tile -- 32 LDS.U16 + 4 STG.128 per lane -- was
      //  measured slower than the shuffle exchange below: 108 vs 94 us for the config-2 kernel.)
      if (p.PT) {
        // transposed copy: lanes are consecutive rows.  Lane pairs swap halves so that every lane stores 4 B:
        // even lanes hold (row, row+1) of the even columns, odd lanes of the odd columns.
        const bool odd = lane & 1;
        const int prow = row & ~1;                  // first row of this lane pair
        const bool pair_ok = (prow + 1 < p.M) && ((p.ldpt & 1) == 0) && (reinterpret_cast<uintptr_t>(p.PT) & 3) == 0;
        if (full && __all_sync(0xffffffffu, pair_ok)) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // w[j] = (col 2j, col 2j+1) of my row.  Send the half the partner needs, keep the other.
            const uint32_t mine = w[j];
            const uint32_t send = odd ? (mine & 0xffffu) : (mine >> 16);          // odd lanes give col 2j, even give col 2j+1
            const uint32_t got = __shfl_xor_sync(0xffffffffu, send, 1);
            // even lane: col 2j of (row, row+1) = (my low, partner low); odd lane: col 2j+1 = (partner high, my high)
            const uint32_t o = odd ? (got | (mine & 0xffff0000u)) : ((mine & 0xffffu) | (got << 16));
            const int col = nb + 2 * j + (odd ? 1 : 0);
            *reinterpret_cast<uint32_t*>(p.PT + (size_t)col * p.ldpt + prow) = o;
          }
        } else if (row_ok) {
          __nv_bfloat16* outT = p.PT + (size_t)nb * p.ldpt + row;
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) {
              const uint32_t h = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xffffu);
              outT[(size_t)j * p.ldpt] = __ushort_as_bfloat16((unsigned short)h);
            }
        }
      }
    }
  }
  if (TOPX && row_ok) {
    float* ov = p.tk_val + ((size_t)row * p.npart + part) * p.tk_k;
    int32_t* oi = p.tk_idx + ((size_t)row * p.npart + part) * p.tk_k;
    if (EPI == EPI_TOPS) {                           // the column rides in the value's low bits
      const int col0 = (n0 + chunk0 * 32) & ~127;    // the 128-aligned block this warp's part lies in
#pragma unroll
      for (int q = 0; q < TS; ++q) tki[q] = col0 + 127 - (int)(__float_as_uint(tkv[q]) & 127u);
    }
#pragma unroll
    for (int q = 0; q < TS; ++q)
      if (q < p.tk_k) { ov[q] = tkv[q]; oi[q] = tki[q]; }
    if (EPI == EPI_TOPK && p.pmax) {
      p.pmax[(size_t)row * p.npart + part] = run_m;
      p.psum[(size_t)row * p.npart + part] = run_s;
    }
  }
  if (EPI == EPI_CE_FWD && row_ok) {
    p.pmax[(size_t)row * p.npart + part] = run_m;
    p.psum[(size_t)row * p.npart + part] = run_s;
    if (have_tl) p.tlogit[row] = tl;
  }
}

// Static schedule of one worker (a CTA, or a CTA pair): a sequence of segments (tile, [k0, k1)).
// Data-parallel: whole tiles, round-robin.  Stream-K: the worker's contiguous slice of the (tile-major)
// k-block sequence; tiles cut by a slice boundary are reduced into C by their epilogues.
struct WorkSched {
  int kb, ntiles, npairs, streamk;
  long cur, end;      // stream-K: global k-block range
  int tile;           // data-parallel: next tile
  __device__ WorkSched(int ntiles_, int kb_, int npairs_, int pair, int streamk_)
      : kb(kb_), ntiles(ntiles_), npairs(npairs_), streamk(streamk_) {
    if (streamk) {
      const long total = (long)ntiles * kb, per = (total + npairs - 1) / npairs;
      cur = per * pair;
      end = cur + per < total ? cur + per : total;
    } else {
      tile = pair; cur = 0; end = 0;
    }
  }
  __device__ bool next(int& t, int& k0, int& k1) {
    if (streamk) {
      if (cur >= end) return false;
      t = (int)(cur / kb);
      k0 = (int)(cur - (long)t * kb);
      const long left = end - cur;
      k1 = (k0 + left < kb) ? (int)(k0 + left) : kb;
      cur += k1 - k0;
      return true;
    }
    if (tile >= ntiles) return false;
    t = tile; k0 = 0; k1 = kb;
    tile += npairs;
    return true;
  }
};

__device__ __forceinline__ uint32_t cluster_rank_() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all_() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads that land at the same CTA-relative address in every CTA of `mask` and complete bytes on each one's mbarrier
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
// completion of this thread's MMAs -> one arrival on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// ------------------------------------------------------------------------------------ 1-CTA kernel
// p.mc = 1 ("B-multicast pairs"): the operand stream of these GEMMs is bound by the L2 -> SM bandwidth (a 128 x 256
// tile pulls 48 KB per k-block; 148 SMs x ~80 GB/s = the measured 12 TB/s L2 ceiling, which is what the 5120 x 10000 x
// 512 and 25088 x 512 x 2048 products ran at).  Launched as clusters of two CTAs that take the tiles (2m, n) and
// (2m + 1, n) in lock step, each CTA loads its own A tile and HALF of the shared B tile, multicast into both CTAs:
// 32 KB instead of 48 KB per CTA and k-block.  A stage is free once BOTH CTAs' MMAs have read it (multicast commit).
// AMN / BMN = 1: that operand is given "MN-major" -- as the (K, M) resp. (K, N) row-major matrix, i.e. the
// transpose of the K-major form -- and is consumed in place: TMA boxes of 64 k-rows x 64 (m|n) columns land
// as 8 KB blocks [k][128 B] (128-byte swizzle), the UMMA descriptor walks them with LBO = 8 KB (next 64 m|n)
// and SBO = 1 KB (next 8 k-rows), and the instruction descriptor flags the operand as MN-major.  This is what
// lets every transposed product of the backward pass (dW = dY^T X, dX = dY W) read the tensors the forward
// pass already has, with no transposed copy in HBM.
template <int EPI, int BN, int AMN = 0, int BMN = 0>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcParams p) {
  constexpr int STAGES = TileCfg<BN>::STAGES;
  constexpr uint32_t STAGE_BYTES = TileCfg<BN>::STAGE_BYTES, TMEM_COLS = TileCfg<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + (size_t)STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + STG_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mc = p.mc;                                             // pairs: "tile" = two vertically adjacent tiles
  const int crank = mc ? (int)cluster_rank_() : 0;
  const int nworkers = mc ? (int)(gridDim.x >> 1) : (int)gridDim.x, worker = mc ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int mt = mc ? ((p.M + 2 * BM - 1) / (2 * BM)) : ((p.M + BM - 1) / BM), nt = (p.N + BN - 1) / BN;
  const int BMS = mc ? 2 * BM : BM;                                // rows per (pair) tile
  const int ntiles = mt * nt, kb = (p.K + BK - 1) / BK;
  const bool nfast = nt < mt;  // consecutive tiles walk the shorter dimension: the long operand streams once

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], mc ? 2 : 1);   // pairs: a stage also holds the peer's half of B -> both CTAs' MMAs release it
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
  if (mc) cluster_sync_all_();            // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // PDL: everything above overlapped the previous kernel's tail; its outputs (our operands, and buffers it
  // still reads that we overwrite) are safe from here on.
  pdl_wait();
  pdl_launch_dependents();
  if (p.zero_word && blockIdx.x == 0 && threadIdx.x == 0) *p.zero_word = 0.f;

  // Producer and MMA warps run their loops with all 32 lanes (uniform control flow keeps addresses and UMMA
  // descriptors in uniform registers); only the TMA / tcgen05 instructions are issued by one elected lane.
  if (warp == 0) {  // ------------------------------------------------------------ TMA producer
    WorkSched sched(ntiles, kb, nworkers, worker, p.streamk);
    int stage = 0, tile, k0, k1;
    uint32_t phase = 0;
    while (sched.next(tile, k0, k1)) {
      const int m0 = (nfast ? tile / nt : tile % mt) * BMS + crank * BM, n0 = (nfast ? tile % nt : tile / mt) * BN;
      for (int k = k0; k < k1; ++k) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* a = smem + (size_t)stage * STAGE_BYTES;
          if (mc) {
            // own A tile; this CTA's half of the B tile into both CTAs (the other half arrives from the peer)
            if (AMN && p.a3d) tma_load_3d(a, &tmA, 0, k * BK, m0 / 64, &full[stage]);
            else if (AMN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(a + j * 8192, &tmA, m0 + 64 * j, k * BK, &full[stage]);
            } else tma_load_2d(a, &tmA, k * BK, m0, &full[stage]);
            uint8_t* bh = a + A_BYTES + crank * (TileCfg<BN>::B_BYTES / 2);
            const int nh = n0 + crank * (BN / 2);
            if (BMN) tma_load_3d_mc(bh, &tmB, 0, k * BK, nh / 64, &full[stage], (uint16_t)3);
            else tma_load_2d_mc(bh, &tmB, k * BK, nh, &full[stage], (uint16_t)3);
          } else {
          if (AMN && p.a3d) {
            tma_load_3d(a, &tmA, 0, k * BK, m0 / 64, &full[stage]);
          } else if (AMN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(a + j * 8192, &tmA, m0 + 64 * j, k * BK, &full[stage]);
          } else {
            tma_load_2d(a, &tmA, k * BK, m0, &full[stage]);
          }
          if (BMN && p.b3d) {
            tma_load_3d(a + A_BYTES, &tmB, 0, k * BK, n0 / 64, &full[stage]);
          } else if (BMN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(a + A_BYTES + j * 8192, &tmB, n0 + 64 * j, k * BK, &full[stage]);
          } else {
            tma_load_2d(a + A_BYTES, &tmB, k * BK, n0, &full[stage]);
          }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {  // ----------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc(BM, BN) | (AMN ? (1u << 15) : 0u) | (BMN ? (1u << 16) : 0u);
    const uint64_t adesc0 = AMN ? umma_desc_mn128(smem_u32(smem)) : umma_desc_k128(smem_u32(smem));
    const uint64_t bdesc0 = BMN ? umma_desc_mn128(smem_u32(smem) + A_BYTES) : umma_desc_k128(smem_u32(smem) + A_BYTES);
    // descriptor start-address step per 16-deep k-step, in 16-byte units: K-major 32 B, MN-major two 8-row groups
    constexpr uint64_t astep = AMN ? 128 : 2, bstep = BMN ? 128 : 2;
    WorkSched sched(ntiles, kb, nworkers, worker, p.streamk);
    int stage = 0, acc = 0, tile, k0, k1;
    uint32_t phase = 0, acc_phase = 0;
    while (sched.next(tile, k0, k1)) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int k = k0; k < k1; ++k) {
        mbar_wait(&full[stage], phase);  // TMA bytes have landed
        tc_fence_after();
        // descriptor address fields are in 16-byte units: + stage offset, + 2 per 16-element k-step
        const uint64_t so = (uint64_t)(stage * (STAGE_BYTES >> 4));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / UK; ++kk)
            tc_mma(d_tmem, adesc0 + so + astep * kk, bdesc0 + so + bstep * kk, idesc, (k > k0) || (kk != 0));
          if (mc) tc_commit_mc(&empty[stage], (uint16_t)3);   // ... in both CTAs: each holds half of the other's B tile
          else tc_commit(&empty[stage]);  // smem slot free once these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(&tfull[acc]);  // accumulator complete
      __syncwarp();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {  // ------------------------------------------------------ epilogue
    // 8 warps: warp w owns TMEM lanes 32*(w%4)..+31 (32 rows) and one half of the tile's BN columns.
    const int ew = warp & 3, half = (warp - 4) >> 2;
    WorkSched sched(ntiles, kb, nworkers, worker, p.streamk);
    int acc = 0, tile, k0, k1;
    uint32_t acc_phase = 0;
    while (sched.next(tile, k0, k1)) {
      const int m0 = (nfast ? tile / nt : tile % mt) * BMS + crank * BM, n0 = (nfast ? tile % nt : tile / mt) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
      epilogue_warp<EPI>(p, tbase, m0 + ew * 32, n0, half * (BN / 64), BN / 64, stg + (warp - 4) * 4096, lane,
                         !(k0 == 0 && k1 == kb), k0 == 0, (n0 / BN) * 2 + half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (mc) cluster_sync_all_();            // the peer may still be signalling this CTA's barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------ fp32-accurate GEMM (3xTF32)
// C = A . B^T to fp32 accuracy on the tensor cores: every fp32 operand is split into hi = its upper 11 mantissa
// bits (a valid tf32) and lo = a - hi (exact in fp32), and  a.b ~= hi_a.hi_b + hi_a.lo_b + lo_a.hi_b  with fp32
// accumulation; the dropped lo.lo term and the tf32 rounding of lo are ~2^-22 relative, below the rounding noise
// of an fp32 dot product of length K = 512.  Three kind::tf32 MMAs (K = 8 each) per 8 columns of K: one sixth of
// the bf16 tensor rate, ~5x the FFMA rate of the SM.  Used by the decoding loops (decode.cu), whose token ids are
// defined against fp32 arithmetic.  Same roles / pipeline / epilogue as gemm_tc_kernel; a stage holds the four
// 32-column k-blocks A_hi | A_lo | B_hi | B_lo.
template <int BN> struct Tf32Cfg {
  static constexpr uint32_t A_B = BM * 128, B_B = BN * 128, STAGE_BYTES = 2 * (A_B + B_B);
  static constexpr int STAGES = (BN == 128) ? 3 : 2;
  static constexpr uint32_t TMEM_COLS = (ACC_STAGES * BN <= 256) ? 256 : 512;   // power of two (192-column tiles: 384 -> 512)
  static constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + STG_BYTES + 256;
};

template <int BN, int EPI = EPI_STORE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const TcParams p) {
  constexpr int STAGES = Tf32Cfg<BN>::STAGES, BK32 = 32;
  constexpr uint32_t STAGE_BYTES = Tf32Cfg<BN>::STAGE_BYTES, TMEM_COLS = Tf32Cfg<BN>::TMEM_COLS;
  constexpr uint32_t A_B = Tf32Cfg<BN>::A_B, B_B = Tf32Cfg<BN>::B_B;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + (size_t)STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + STG_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = (p.M + BM - 1) / BM, nt = (p.N + BN - 1) / BN;
  const int ntiles = mt * nt, kb = (p.K + BK32 - 1) / BK32;
  const bool nfast = nt < mt;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
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
  pdl_wait();
  pdl_launch_dependents();
  if (p.zero_word && blockIdx.x == 0 && threadIdx.x == 0) *p.zero_word = 0.f;

  if (warp == 0) {  // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m0 = (nfast ? tile / nt : tile % mt) * BM, n0 = (nfast ? tile % nt : tile / mt) * BN;
      for (int k = 0; k < kb; ++k) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* a = smem + (size_t)stage * STAGE_BYTES;
          tma_load_2d(a, &tmAh, k * BK32, m0, &full[stage]);
          tma_load_2d(a + A_B, &tmAl, k * BK32, m0, &full[stage]);
          tma_load_2d(a + 2 * A_B, &tmBh, k * BK32, n0, &full[stage]);
          tma_load_2d(a + 2 * A_B + B_B, &tmBl, k * BK32, n0, &full[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {  // ----------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(BM, BN);
    const uint32_t s0 = smem_u32(smem);
    const uint64_t ah0 = umma_desc_k128(s0), al0 = umma_desc_k128(s0 + A_B), bh0 = umma_desc_k128(s0 + 2 * A_B),
                   bl0 = umma_desc_k128(s0 + 2 * A_B + B_B);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int k = 0; k < kb; ++k) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t so = (uint64_t)(stage * (STAGE_BYTES >> 4));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {   // 8 tf32 = 32 bytes = 2 descriptor units per k-step
            tc_mma_tf32(d_tmem, al0 + so + 2 * kk, bh0 + so + 2 * kk, idesc, (k != 0) || (kk != 0));   // small terms first
            tc_mma_tf32(d_tmem, ah0 + so + 2 * kk, bl0 + so + 2 * kk, idesc, 1);
            tc_mma_tf32(d_tmem, ah0 + so + 2 * kk, bh0 + so + 2 * kk, idesc, 1);
          }
          tc_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(&tfull[acc]);
      __syncwarp();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {  // ------------------------------------------------------ epilogue
    const int ew = warp & 3, half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m0 = (nfast ? tile / nt : tile % mt) * BM, n0 = (nfast ? tile % nt : tile / mt) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
      epilogue_warp<EPI>(p, tbase, m0 + ew * 32, n0, half * (BN / 64), BN / 64, stg + (warp - 4) * 4096, lane, false,
                         true, (n0 / BN) * 2 + half);
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

// hi = x with the low 13 mantissa bits cleared (exactly representable as tf32), lo = x - hi (exact).
__global__ void split_tf32_kernel(const float* __restrict__ src, int rows, int cols, int lds, float* __restrict__ hi,
                                  float* __restrict__ lo, int ldd) {
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float x = src[(size_t)r * lds + c];
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    hi[(size_t)r * ldd + c] = h;
    lo[(size_t)r * ldd + c] = x - h;
  }
}

template <int BN, int EPI = EPI_STORE>
int launch_tf32x3(const TcParams& p, const float* Ah, const float* Al, int lda, const float* Bh, const float* Bl, int ldb,
                  cudaStream_t s, int sms) {
  CUtensorMap tmAh, tmAl, tmBh, tmBl;
  ST_TRY(make_tmap_f32(&tmAh, Ah, p.M, p.K, lda, BM, "A_hi"));
  ST_TRY(make_tmap_f32(&tmAl, Al, p.M, p.K, lda, BM, "A_lo"));
  ST_TRY(make_tmap_f32(&tmBh, Bh, p.N, p.K, ldb, BN, "B_hi"));
  ST_TRY(make_tmap_f32(&tmBl, Bl, p.N, p.K, ldb, BN, "B_lo"));
  auto kern = gemm_tf32x3_kernel<BN, EPI>;
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tf32Cfg<BN>::SMEM_BYTES));
  const int ntiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = ntiles < sms ? ntiles : sms;
  ST_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(NTHREADS), Tf32Cfg<BN>::SMEM_BYTES, s, tmAh, tmAl, tmBh, tmBl, p));
  note_launch();
  return ST_OK;
}

// --------------------------------------------------------------------------------- CTA-pair kernel
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(0u));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into THIS CTA's shared memory; the bytes complete on the LEADER CTA's mbarrier (peer bit cleared).
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_at(uint32_t cluster_addr, uint32_t bytes) {
  // default semantics (release at CTA scope), as CUTLASS's ClusterTransactionBarrier does for the leader's barrier: with
  // .release.cluster every call became MEMBAR.ALL.CTA + ERRBAR on the producer thread -- once per k-block, which made the
  // PRODUCER the bottleneck of the pair kernel (tensor pipe 29 %, ncu)
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_at(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
// D[tmem of both CTAs] (+)= A . B^T over the CTA pair: M = 256 (128 rows per CTA), N = 256 (128 B rows per CTA)
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// all prior MMAs of this thread retired -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

// AMN / BMN: that operand MN-major (see gemm_tc_kernel), given as the 3-D tensor map of make_tmap_mn3d with two 64-wide
// blocks per box: a CTA's 128 rows of A / its 128-column half of B are one operation each either way.
template <int EPI, int AMN = 0, int BMN = 0>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const TcParams p) {
  constexpr int STAGES = PairCfg::STAGES, BN = PairCfg::BN;
  constexpr uint32_t STAGE_BYTES = PairCfg::STAGE_BYTES, TMEM_COLS = PairCfg::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + (size_t)STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + STG_BYTES);   // used in the leader: 2 producers arrive (+ their bytes)
  uint64_t* empty = full + STAGES;                                  // per CTA: multicast commit
  uint64_t* tfull = empty + STAGES;                                 // per CTA: multicast commit
  uint64_t* tempty = tfull + ACC_STAGES;                            // used in the leader: 16 epilogue warps arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int mt = (p.M + 2 * BM - 1) / (2 * BM), nt = (p.N + BN - 1) / BN;
  const int ntiles = mt * nt, kb = (p.K + BK - 1) / BK;
  const bool nfast = nt < mt;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  if (p.zero_word && blockIdx.x == 0 && threadIdx.x == 0) *p.zero_word = 0.f;
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {  // --------------------------------------------------- TMA producer (both CTAs)
    WorkSched sched(ntiles, kb, npairs, pair, p.streamk);
    int stage = 0, tile, k0, k1;
    uint32_t phase = 0;
    while (sched.next(tile, k0, k1)) {
      const int m0 = (nfast ? tile / nt : tile % mt) * 2 * BM + (int)rank * BM;
      const int n0 = (nfast ? tile % nt : tile / mt) * BN + (int)rank * (BN / 2);
      for (int k = k0; k < k1; ++k) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx_at(mapa_rank0(smem_u32(&full[stage])), STAGE_BYTES);
          uint8_t* a = smem + (size_t)stage * STAGE_BYTES;
          if (AMN) tma_load_3d_pair(a, &tmA, 0, k * BK, m0 / 64, &full[stage]);
          else tma_load_2d_pair(a, &tmA, k * BK, m0, &full[stage]);
          if (BMN) tma_load_3d_pair(a + A_BYTES, &tmB, 0, k * BK, n0 / 64, &full[stage]);
          else tma_load_2d_pair(a + A_BYTES, &tmB, k * BK, n0, &full[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {  // ------------------------------------------------- MMA issuer (leader CTA only)
      constexpr uint32_t idesc = umma_idesc(2 * BM, BN) | (AMN ? (1u << 15) : 0u) | (BMN ? (1u << 16) : 0u);
      const uint64_t adesc0 = AMN ? umma_desc_mn128(smem_u32(smem)) : umma_desc_k128(smem_u32(smem));
      const uint64_t bdesc0 = BMN ? umma_desc_mn128(smem_u32(smem) + A_BYTES) : umma_desc_k128(smem_u32(smem) + A_BYTES);
      constexpr uint64_t astep = AMN ? 128 : 2, bstep = BMN ? 128 : 2;
      WorkSched sched(ntiles, kb, npairs, pair, p.streamk);
      int stage = 0, acc = 0, tile, k0, k1;
      uint32_t phase = 0, acc_phase = 0;
      while (sched.next(tile, k0, k1)) {
        mbar_wait_cluster(&tempty[acc], acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int k = k0; k < k1; ++k) {
          mbar_wait(&full[stage], phase);  // both CTAs' TMA bytes have landed
          tc_fence_after();
          const uint64_t so = (uint64_t)(stage * (STAGE_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / UK; ++kk)
              tc_mma_pair(d_tmem, adesc0 + so + astep * kk, bdesc0 + so + bstep * kk, idesc, (k > k0) || (kk != 0));
            tc_commit_pair(&empty[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit_pair(&tfull[acc]);
        __syncwarp();
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {  // ------------------------------------------------------ epilogue (both CTAs)
    const int ew = warp & 3, half = (warp - 4) >> 2;
    WorkSched sched(ntiles, kb, npairs, pair, p.streamk);
    int acc = 0, tile, k0, k1;
    uint32_t acc_phase = 0;
    while (sched.next(tile, k0, k1)) {
      const int m0 = (nfast ? tile / nt : tile % mt) * 2 * BM + (int)rank * BM;
      const int n0 = (nfast ? tile % nt : tile / mt) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
      epilogue_warp<EPI>(p, tbase, m0 + ew * 32, n0, half * (BN / 64), BN / 64, stg + (warp - 4) * 4096, lane,
                         !(k0 == 0 && k1 == kb), k0 == 0, (n0 / BN) * 2 + half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_at(mapa_rank0(smem_u32(&tempty[acc])));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading this CTA's operands / signalling its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// lse[m] = log sum_v exp(logit) from the per-tile partials; loss_sum += lse - target logit.
// One warp per row: the row's npart partials are contiguous, so the lanes read them coalesced and combine with
// shuffles (a thread per row walked 2 x npart strided words: 13 us for 5120 rows, now a few).
__global__ void __launch_bounds__(256) ce_combine_kernel(int M, int npart, const float* __restrict__ pmax,
                                                         const float* __restrict__ psum, const float* __restrict__ tlogit,
                                                         float* __restrict__ lse, float* __restrict__ loss_sum) {
  __shared__ float wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * 8 + warp;
  float contrib = 0.f;
  if (m < M) {
    const float* pm = pmax + (size_t)m * npart;
    const float* ps = psum + (size_t)m * npart;
    float mx = -FLT_MAX;
    for (int j = lane; j < npart; j += 32) mx = fmaxf(mx, pm[j]);
    mx = warp_max(mx);
    float sacc = 0.f;
    for (int j = lane; j < npart; j += 32) sacc += ps[j] * expf(pm[j] - mx);
    sacc = warp_sum(sacc);
    const float l = mx + logf(sacc);
    if (lane == 0) {
      lse[m] = l;
      contrib = l - tlogit[m];
    }
  }
  if (lane == 0) wsum[warp] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += wsum[w];
    atomicAdd(loss_sum, t);
  }
}

// Per row: the K best of the npart * K candidates the epilogues kept (value descending, lower column index first
// among equal values).  One warp per row, K rounds of warp arg-max with exclusion.  Writes (val, idx) with row stride
// `out_stride` and/or the best index as int64 (greedy decoding: token matrix column).  With pmax / psum: also the row's
// soft-max normaliser from the epilogues' partials: row_max, row_sum = sum_j exp(x_j - row_max).
__global__ void __launch_bounds__(256) topk_merge_kernel(int M, int npart, int K, const float* __restrict__ cv,
                                                         const int32_t* __restrict__ ci, float* __restrict__ val,
                                                         int32_t* __restrict__ idx, int out_stride,
                                                         int64_t* __restrict__ tok, int tok_stride,
                                                         const float* __restrict__ pmax, const float* __restrict__ psum,
                                                         float* __restrict__ row_max, float* __restrict__ row_sum) {
  const int lane = threadIdx.x & 31, m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int ncand = npart * K;
  const float* v = cv + (size_t)m * ncand;
  const int32_t* ix = ci + (size_t)m * ncand;
  float pv = FLT_MAX;            // the previous winner: candidates must rank strictly after it
  int pi = -1;
  for (int k = 0; k < K; ++k) {
    float bv = -FLT_MAX;
    int bi = 0x7fffffff;
    for (int c = lane; c < ncand; c += 32) {
      const float x = v[c];
      const int i = ix[c];
      const bool after_prev = (x < pv) || (x == pv && i > pi);
      const bool better = (x > bv) || (x == bv && i < bi);
      if (after_prev && better) { bv = x; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    pv = bv; pi = bi;
    if (lane == 0) {
      if (val) val[(size_t)m * out_stride + k] = bv;
      if (idx) idx[(size_t)m * out_stride + k] = bi;
      if (tok && k == 0) tok[(size_t)m * tok_stride] = bi;
    }
  }
  if (pmax) {
    float mx = -FLT_MAX;
    for (int c = lane; c < npart; c += 32) mx = fmaxf(mx, pmax[(size_t)m * npart + c]);
    mx = warp_max(mx);
    float sm = 0.f;
    for (int c = lane; c < npart; c += 32) sm += psum[(size_t)m * npart + c] * expf(pmax[(size_t)m * npart + c] - mx);
    sm = warp_sum(sm);
    if (lane == 0) { row_max[m] = mx; row_sum[m] = sm; }
  }
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

// Stream-K pays (a zero fill of C + fp32 reductions of the cut tiles) when whole tiles would leave a good part
// of the workers idle and K is long enough to cut: dX / dW of the vocabulary projection (80 tiles on 148 SMs),
// dW_hh / dW_ih (32 tiles), dW_enc of the attention (32 tiles, 392 k-blocks).  Cost model fitted on B200
// (tools/time_kernels.py gemmv): a k-block costs ~0.42 us of main loop; cutting every tile costs ~12 us of
// un-overlapped reductions at the end + ~0.15 us per MB of C.
template <int EPI>
inline bool want_streamk(const TcParams& p, int ntiles, int kb, int workers) {
  if (g_streamk == 0) return false;
  if (!(EPI == EPI_STORE && !p.c_bf16 && p.beta == 0.f && (p.ldc & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && kb >= 16))
    return false;
  const long rounds = (ntiles + workers - 1) / workers;
  const long per = ((long)ntiles * kb + workers - 1) / workers;
  const double saved_us = 0.42 * (double)(rounds * kb - per);
  const double cost_us = 12.0 + 0.15 * ((double)p.M * p.N * 4.0 / 1e6);
  return saved_us > cost_us;
}

int g_mc = 0;    // st_debug_gemm_variant(v | 0x4000 / 0x8000): B-multicast CTA pairs forced on / off (0 = choose)

template <int EPI, int BN, int AMN = 0, int BMN = 0>
int launch_tc_bn(TcParams p, const void* A, int lda, const void* B, int ldb, cudaStream_t s, int sms) {
  CUtensorMap tmA, tmB;
  auto kern = gemm_tc_kernel<EPI, BN, AMN, BMN>;
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<BN>::SMEM_BYTES));
  const int mt1 = (p.M + BM - 1) / BM, ntn = (p.N + BN - 1) / BN, kb = (p.K + BK - 1) / BK;
  // B-multicast pairs (see the kernel).  Measured on B200: no gain where the tile grid is large (multicast over two CTAs
  // does not lower the L2 traffic enough to matter: 5120 x 10000 x 512 and 25088 x 512 x 2048 run at the same 55-58 us),
  // but the long-K weight-gradient products with a handful of row tiles -- dW_enc / dW_embed = X^T F, 512 x 2048 x 25088,
  // 32 tiles cut into stream-K ranges -- go from 87 to 64 us.  MN-major B: its columns in whole 64-blocks (3-D map).
  const bool mc_ok = BN >= 128 && mt1 >= 2 && sms >= 2 && (!BMN || (p.N % 64 == 0 && g_mn3d));
  p.mc = (mc_ok && g_mc >= 0 && (g_mc > 0 || (mt1 <= 8 && kb >= 128))) ? 1 : 0;
  // MN-major operands whose (m|n) extent is a multiple of 64: one 3-D box per k-block instead of BM/64 (BN/64) 2-D boxes
  p.a3d = (AMN && p.M % 64 == 0 && g_mn3d) ? 1 : 0;
  p.b3d = (BMN && p.N % 64 == 0 && g_mn3d) ? 1 : 0;
  if (AMN && p.a3d) ST_TRY(make_tmap_mn3d(&tmA, A, p.K, p.M, lda, BM / 64, "A (MN-major)"));
  else if (AMN) ST_TRY(make_tmap(&tmA, A, p.K, p.M, lda, 64, "A (MN-major)"));
  else ST_TRY(make_tmap(&tmA, A, p.M, p.K, lda, BM, "A"));
  const int bdiv = p.mc ? 2 : 1;                                   // pairs load B in halves
  if (BMN && p.b3d) ST_TRY(make_tmap_mn3d(&tmB, B, p.K, p.N, ldb, BN / 64 / bdiv, "B (MN-major)"));
  else if (BMN) ST_TRY(make_tmap(&tmB, B, p.K, p.N, ldb, 64, "B (MN-major)"));
  else ST_TRY(make_tmap(&tmB, B, p.N, p.K, ldb, BN / bdiv, "B"));
  const int workers = p.mc ? sms / 2 : sms;
  const int ntiles = (p.mc ? (mt1 + 1) / 2 : mt1) * ntn;
  int grid = ntiles < workers ? ntiles : workers;
  p.streamk = want_streamk<EPI>(p, ntiles, kb, workers) ? 1 : 0;
  if (p.streamk) {
    grid = workers;
    if (!g_c_zeroed) ST_CUDA_TRY(cudaMemset2DAsync(p.C, (size_t)p.ldc * 4, 0, (size_t)p.N * 4, p.M, s));
  }
  if (!p.mc) {
    ST_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(NTHREADS), TileCfg<BN>::SMEM_BYTES, s, tmA, tmB, p));
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = TileCfg<BN>::SMEM_BYTES; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 2 : 1;
    ST_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
  }
  note_launch();
  return ST_OK;
}

// Tile width without stream-K: estimated time = rounds of the persistent grid x clocks per k-block.
inline int pick_bn(int M, int N, int sms) {
  if (N <= 128) return 128;
  const long mt = (M + BM - 1) / BM;
  const long r128 = (mt * ((N + 127) / 128) + sms - 1) / sms, r256 = (mt * ((N + 255) / 256) + sms - 1) / sms;
  return (r256 * 750 <= r128 * 600) ? 256 : 128;
}

template <int EPI, int AMN = 0, int BMN = 0>
int launch_tc_pair(TcParams p, const void* A, int lda, const void* B, int ldb, cudaStream_t s, int sms) {
  CUtensorMap tmA, tmB;
  if (AMN) ST_TRY(make_tmap_mn3d(&tmA, A, p.K, p.M, lda, 2, "A (MN-major)"));
  else ST_TRY(make_tmap(&tmA, A, p.M, p.K, lda, BM, "A"));
  if (BMN) ST_TRY(make_tmap_mn3d(&tmB, B, p.K, p.N, ldb, 2, "B (MN-major)"));
  else ST_TRY(make_tmap(&tmB, B, p.N, p.K, ldb, BM, "B"));
  auto kern = gemm_tc2_kernel<EPI, AMN, BMN>;
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PairCfg::SMEM_BYTES));
  const int pairs = sms / 2;
  const int ntiles = ((p.M + 255) / 256) * ((p.N + 255) / 256), kb = (p.K + BK - 1) / BK;
  p.streamk = want_streamk<EPI>(p, ntiles, kb, pairs) ? 1 : 0;
  int npairs = ntiles < pairs ? ntiles : pairs;
  if (p.streamk) {
    npairs = pairs;
    if (!g_c_zeroed) ST_CUDA_TRY(cudaMemset2DAsync(p.C, (size_t)p.ldc * 4, 0, (size_t)p.N * 4, p.M, s));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = PairCfg::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  ST_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
  note_launch();
  return ST_OK;
}

// Kernel + tile choice for a problem: 128 / 256 = single-CTA kernel with that tile width; 2 = CTA-pair kernel
// (cta_group::2, 256 x 256 tiles over two SMs: 32 KB instead of 48 KB of operands per CTA and k-block, which is what
// bounds the big products -- 51 against 56-58 us on 25088 x 512 x 2048 and 5120 x 10000 x 512).  K-major operands only.
int g_pair_auto = 1;   // st_debug_gemm_variant(v | 0x10000): never choose the pair kernel
template <int EPI>
inline int pick_variant(const TcParams& p, int sms) {
  if (g_variant) return g_variant;
  if (g_pair_auto && p.M >= 512 && p.N >= 256 && p.K >= 256) {
    // enough 256 x 256 tiles to keep every pair busy for several rounds (or stream-K over the pairs)
    const long t2 = (long)((p.M + 255) / 256) * ((p.N + 255) / 256);
    if (t2 >= 2L * (sms / 2)) return 2;
    // (few tiles + long K, i.e. stream-K over the pairs: measured equal or slower than the single-CTA kernel -- 66.6 vs
    //  64.5 us for 512 x 2048 x 25088, 29.4 vs 25.4 us for 2048 x 512 x 5120 -- so those stay where they were)
  }
  if (p.N > 128) {
    const int ntiles = ((p.M + BM - 1) / BM) * ((p.N + 255) / 256), kb = (p.K + BK - 1) / BK;
    if (want_streamk<EPI>(p, ntiles, kb, sms)) return 256;
  }
  return pick_bn(p.M, p.N, sms);
}

// The per-step products of the attention loop (M = live batch rows, N = K = 512) are a latency chain, not a
// throughput problem: with 128-wide tiles only 4-16 CTAs work and each pulls a 16 KB B block per k-step; 64-wide
// tiles double the CTAs and shorten the per-k-step TMA time.  Plain stores only (the CE epilogues index their
// partials by 128-column tile).
template <int EPI>
inline bool want_narrow(const TcParams& p, int sms) {
  if (EPI != EPI_STORE || g_variant || p.N <= 64) return false;
  const long t128 = (long)((p.M + BM - 1) / BM) * ((p.N + 127) / 128);
  return t128 * 8 <= sms;
}

// EPI_STORE with operand-major flags (the backward products).  Tile width 128 or 256 (64-wide tiles are a
// K-major latency-path special).
template <int AMN, int BMN>
int launch_tc_major(const TcParams& p, const void* A, int lda, const void* B, int ldb, cudaStream_t s) {
  ST_REQUIRE(p.M >= 1 && p.N >= 1 && p.K >= 1, ST_ERR_BAD_SHAPE, "gemm_bf16: M=%d N=%d K=%d", p.M, p.N, p.K);
  int sms = 0;
  ST_TRY(gemm_sms(&sms));
  int bn = (g_variant == 128 || g_variant == 256) ? g_variant : pick_variant<EPI_STORE>(p, sms);
  // the pair kernel takes MN-major operands as 3-D maps of whole 64-blocks
  if (bn == 2 && ((AMN && p.M % 64 != 0) || (BMN && p.N % 64 != 0) || !g_mn3d)) bn = (g_variant == 2) ? 256 : pick_bn(p.M, p.N, sms);
  if (bn == 2) return launch_tc_pair<EPI_STORE, AMN, BMN>(p, A, lda, B, ldb, s, sms);
  return bn == 256 ? launch_tc_bn<EPI_STORE, 256, AMN, BMN>(p, A, lda, B, ldb, s, sms)
                   : launch_tc_bn<EPI_STORE, 128, AMN, BMN>(p, A, lda, B, ldb, s, sms);
}

template <int EPI>
int launch_tc(const TcParams& p, const void* A, int lda, const void* B, int ldb, cudaStream_t s, int bn = 0) {
  ST_REQUIRE(p.M >= 1 && p.N >= 1 && p.K >= 1, ST_ERR_BAD_SHAPE, "gemm_bf16: M=%d N=%d K=%d", p.M, p.N, p.K);
  int sms = 0;
  ST_TRY(gemm_sms(&sms));
  if (bn == 0 && want_narrow<EPI>(p, sms)) return launch_tc_bn<EPI_STORE, 64>(p, A, lda, B, ldb, s, sms);
  if (bn == 0) bn = pick_variant<EPI>(p, sms);
  if (bn == 2) return launch_tc_pair<EPI>(p, A, lda, B, ldb, s, sms);
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

int st_gemm_bf16_ex(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                    int ldc, int c_is_bf16, const float* bias, float alpha, float beta, st_stream_t stream) {
  using namespace st;
  if (!a_mn && !b_mn) return st_gemm_bf16(M, N, K, A, lda, B, ldb, C, ldc, c_is_bf16, bias, alpha, beta, stream);
  ST_REQUIRE(C != nullptr, ST_ERR_NULL, "st_gemm_bf16_ex: C is NULL");
  ST_REQUIRE(ldc >= N, ST_ERR_BAD_SHAPE, "st_gemm_bf16_ex: ldc=%d < N=%d", ldc, N);
  ST_REQUIRE(beta == 0.f || !c_is_bf16, ST_ERR_UNSUPPORTED, "st_gemm_bf16_ex: beta needs an fp32 C");
  TcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc; p.c_bf16 = c_is_bf16; p.bias = bias; p.alpha = alpha; p.beta = beta;
  if (a_mn && b_mn) return launch_tc_major<1, 1>(p, A, lda, B, ldb, as_stream(stream));
  if (b_mn) return launch_tc_major<0, 1>(p, A, lda, B, ldb, as_stream(stream));
  return launch_tc_major<1, 0>(p, A, lda, B, ldb, as_stream(stream));
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
  const int bn = pick_variant<EPI_CE_FWD>(p, sms);
  p.npart = 2 * ((V + (bn == 128 ? 127 : 255)) / (bn == 128 ? 128 : 256));   // <= st_vocab_ce_parts(V)
  p.zero_word = loss_sum;   // cleared by the GEMM kernel itself; ce_combine_kernel (stream-ordered behind it) accumulates
  ST_TRY(launch_tc<EPI_CE_FWD>(p, Hs, ldh, Wv, ldw, s, bn));
  ce_combine_kernel<<<(M + 7) / 8, 256, 0, s>>>(M, p.npart, part_max, part_sum, tlogit, lse, loss_sum);
  ST_LAUNCH_TRY("ce_combine_kernel");
  return ST_OK;
}

int st_split_tf32(const float* src, int rows, int cols, int lds, float* hi, float* lo, int ldd, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(src && hi && lo, ST_ERR_NULL, "st_split_tf32: NULL pointer");
  ST_REQUIRE(rows >= 1 && cols >= 1 && lds >= cols && ldd >= cols, ST_ERR_BAD_SHAPE,
             "st_split_tf32: rows=%d cols=%d lds=%d ldd=%d", rows, cols, lds, ldd);
  const long long n = (long long)rows * cols;
  const long long blocks = (n + 255) / 256;
  split_tf32_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, as_stream(stream)>>>(src, rows, cols, lds, hi, lo, ldd);
  ST_LAUNCH_TRY("split_tf32_kernel");
  return ST_OK;
}

int st_gemm_tf32x3(int M, int N, int K, const float* A_hi, const float* A_lo, int lda, const float* B_hi,
                   const float* B_lo, int ldb, float* C, int ldc, const float* bias, float alpha, float beta,
                   st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(C != nullptr, ST_ERR_NULL, "st_gemm_tf32x3: C is NULL");
  ST_REQUIRE(M >= 1 && N >= 1 && K >= 1 && ldc >= N, ST_ERR_BAD_SHAPE, "st_gemm_tf32x3: M=%d N=%d K=%d ldc=%d", M, N, K, ldc);
  TcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc; p.c_bf16 = 0; p.bias = bias; p.alpha = alpha; p.beta = beta;
  int sms = 0;
  ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
  // Tile width by modelled time: waves of tiles x operand bytes per tile (this product streams (BM + BN) x K x 8 bytes
  // per tile and is bound by that stream).  4096 x 1536 (the decode step's W_hh product): 192 tiles of 256 columns are
  // 1.3 waves on 148 SMs and run as 2; 256 tiles of 192 columns are 2 shorter ones.
  const long mt = (M + BM - 1) / BM;
  long best = -1;
  int bn = 256;
  for (int c : {256, 192, 128}) {
    if (c == 192 && (g_tf32_bn192 == 0 || N < 384)) continue;
    const long tiles = mt * ((N + c - 1) / c), waves = (tiles + sms - 1) / sms, cost = waves * (BM + c);
    if (best < 0 || cost < best) { best = cost; bn = c; }
  }
  if (N <= 128) bn = 128;
  if (bn == 256) return launch_tf32x3<256>(p, A_hi, A_lo, lda, B_hi, B_lo, ldb, as_stream(stream), sms);
  if (bn == 192) return launch_tf32x3<192>(p, A_hi, A_lo, lda, B_hi, B_lo, ldb, as_stream(stream), sms);
  return launch_tf32x3<128>(p, A_hi, A_lo, lda, B_hi, B_lo, ldb, as_stream(stream), sms);
}

int st_topk_parts(int N) { return 2 * ((N + 127) / 128) * st::TOPK_SLOTS; }

int st_gemm_tf32x3_topk(int M, int N, int K, const float* A_hi, const float* A_lo, int lda, const float* B_hi,
                        const float* B_lo, int ldb, const float* bias, int topk, float* cand_val, int32_t* cand_idx,
                        float* val, int32_t* idx, int out_stride, int64_t* tok, int tok_stride, float* part_stats,
                        float* row_max, float* row_sum, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(cand_val && cand_idx && (val || idx || tok), ST_ERR_NULL, "st_gemm_tf32x3_topk: NULL pointer");
  ST_REQUIRE((!row_max && !row_sum) || (row_max && row_sum && part_stats), ST_ERR_NULL,
             "st_gemm_tf32x3_topk: row_max, row_sum and part_stats go together");
  ST_REQUIRE(M >= 1 && N >= 1 && K >= 1 && topk >= 1 && topk <= TOPK_SLOTS && topk <= N && (!(val || idx) || out_stride >= topk) &&
                 (!tok || tok_stride >= 1),
             ST_ERR_BAD_SHAPE, "st_gemm_tf32x3_topk: M=%d N=%d K=%d topk=%d (max %d)", M, N, K, topk, (int)TOPK_SLOTS);
  TcParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.alpha = 1.f;
  p.tk_val = cand_val; p.tk_idx = cand_idx; p.tk_k = topk;
  int sms = 0;
  ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
  cudaStream_t s = as_stream(stream);
  const int bn = pick_bn(M, N, sms);
  p.npart = 2 * ((N + bn - 1) / bn);
  if (row_max) { p.pmax = part_stats; p.psum = part_stats + (size_t)M * p.npart; }
  if (bn == 256) ST_TRY((launch_tf32x3<256, EPI_TOPK>(p, A_hi, A_lo, lda, B_hi, B_lo, ldb, s, sms)));
  else ST_TRY((launch_tf32x3<128, EPI_TOPK>(p, A_hi, A_lo, lda, B_hi, B_lo, ldb, s, sms)));
  topk_merge_kernel<<<(M + 7) / 8, 256, 0, s>>>(M, p.npart, topk, cand_val, cand_idx, val, idx, out_stride, tok, tok_stride,
                                               p.pmax, p.psum, row_max, row_sum);
  ST_LAUNCH_TRY("topk_merge_kernel");
  return ST_OK;
}

int st_gemm_set_c_zeroed(int on) {
  st::g_c_zeroed = on ? 1 : 0;
  return ST_OK;
}

int st_gemm_set_sm_limit(int n) {
  st::g_sm_limit = n > 0 ? n : 0;
  return ST_OK;
}

int st_gemm_bf16_screen(int M, int N, int K, const void* A, int lda, const void* B, int ldb, const float* bias,
                        float* cand_val, int32_t* cand_idx, int* npart_out, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(A && B && cand_val && cand_idx && npart_out, ST_ERR_NULL, "st_gemm_bf16_screen: NULL pointer");
  ST_REQUIRE(M >= 1 && N >= 1 && K >= 1, ST_ERR_BAD_SHAPE, "st_gemm_bf16_screen: M=%d N=%d K=%d", M, N, K);
  TcParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.alpha = 1.f;
  p.tk_val = cand_val; p.tk_idx = cand_idx; p.tk_k = ST_SCREEN_SLOTS;
  const int bn = N > 128 ? 256 : 128;
  p.npart = 2 * ((N + bn - 1) / bn);
  *npart_out = p.npart;
  // the CTA-pair kernel has the same 256-column tiles (same parts) and is the faster one on large problems
  int sms = 0;
  ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
  const long t2 = (long)((M + 255) / 256) * ((N + 255) / 256);
  const int variant = (g_variant == 0 && g_pair_auto && M >= 512 && N >= 256 && K >= 256 && t2 >= 2L * (sms / 2)) ? 2 : bn;
  return launch_tc<EPI_TOPS>(p, A, lda, B, ldb, as_stream(stream), variant);
}

int st_debug_gemm_variant(int variant) {
  st::g_streamk = (variant & 0x1000) ? 0 : 1;
  st::g_tf32_bn192 = (variant & 0x20000) ? 0 : 1;
  st::g_mn3d = (variant & 0x2000) ? 0 : 1;
  st::g_mc = (variant & 0x4000) ? 1 : ((variant & 0x8000) ? -1 : 0);
  st::g_pair_auto = (variant & 0x10000) ? 0 : 1;
  variant &= 0xfff;
  st::g_variant = (variant == 2 || variant == 128 || variant == 256) ? variant : 0;
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
