// Persistent recurrent kernels on the tensor cores (bf16 mode): nn.GRU / nn.LSTM over a
// PackedSequence (rnn.py:32, LSTM/rnn_lstm.py:30) and its backward through time.
//
// Forward, per CTA (x = tile of 16 hidden units, y = tile of 128 batch rows), for ALL time steps:
//   * the G*16 rows of W_hh (bf16) that produce its units' gates stay resident in shared memory
//     (TMA-loaded once, 128B-swizzled K-major = the UMMA B operand);
//   * each step the 128 x H tile of h_{t-1} (bf16, written by all unit tiles in the previous step)
//     is TMA-loaded k-block by k-block and multiplied on tcgen05.mma into a 128 x (G*16) fp32
//     accumulator in TMEM;
//   * 8 epilogue warps read the accumulator (tcgen05.ld), add the hoisted input pre-activations,
//     apply the gate non-linearities and the state update with c / h carried in REGISTERS across
//     steps, and write h_t (fp32 + bf16), c_t and the saved gates;
//   * the unit tiles of one batch tile meet at a device-wide barrier (one counter per batch tile).
// Backward mirrors it: phase 1 (thread-local) turns dh_t into gate gradients, written as bf16
// row-major and transposed (the operands of the hoisted weight-gradient GEMMs) with the bias
// gradients accumulated in registers; after the barrier phase 2 streams the 128 x (G*H) tile of
// dGh_t through a TMA ring against the resident W_hh^T slice to form dh_{t-1} for the CTA's units.
//
// Generic-proxy global stores of step t are read by TMA (async proxy) in step t+1 / phase 2:
// writers and the reader both issue fence.proxy.async around the device-wide barrier.
#include "common.cuh"
#include "tc_common.cuh"

namespace st {
namespace {

constexpr int UT = 16, BT = 128, HALF = 8, NTH = 320, MAXKB = 8;
constexpr uint32_t KBLK_A = BT * 128;  // one 64-wide k-block of a 128-row bf16 tile, bytes

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void proxy_fence_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8cg(const float* p, float (&v)[8]) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8bf(__nv_bfloat16* p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
    w[q] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

struct TcFwdParams {
  int H, nsteps, has_h0;
  const float *Gx, *bhh, *h0, *c0;
  float *Hs, *Cs, *gates, *ghn;
  __nv_bfloat16* Hsb;
  int* barrier;
};

template <int G>
__global__ void __launch_bounds__(NTH, 1)
rnn_seq_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
                      const __grid_constant__ CUtensorMap tmH0, const __grid_constant__ StepTable tab,
                      const TcFwdParams p) {
  constexpr int NC = G * UT;                    // accumulator columns
  constexpr uint32_t KBLK_W = NC * 128;         // bytes of one k-block of the W slice
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, KB = (H + 63) / 64;
  uint8_t* sA = smem;                            // [KB][128 rows][128 B]
  uint8_t* sW = smem + (size_t)KB * KBLK_A;      // [KB][NC rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)KB * KBLK_W);
  uint64_t* wbar = bars;
  uint64_t* accbar = bars + 1;
  uint64_t* hfull = bars + 2;                    // [MAXKB]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + MAXKB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UT, r0 = blockIdx.y * BT;

  if (warp == 0 && lane == 0) {
    mbar_init(wbar, 1);
    mbar_init(accbar, 1);
    for (int i = 0; i < MAXKB; ++i) mbar_init(&hfull[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(64u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0 && lane == 0) {  // resident W_hh slice: G boxes of 16 rows per k-block
    mbar_expect_tx(wbar, (uint32_t)KB * KBLK_W);
    for (int kb = 0; kb < KB; ++kb)
      for (int g = 0; g < G; ++g)
        tma_load_2d(sW + (size_t)kb * KBLK_W + g * UT * 128, &tmW, kb * 64, g * H + u0, wbar);
  }

  // epilogue role: warps 2..9; lane quarter q = warp % 4, unit half hf
  const bool is_epi = warp >= 2;
  const int q = warp & 3, hf = (warp - 2) >> 2;
  const int row = q * 32 + lane;                 // row inside the batch tile == TMEM lane
  const int uu = u0 + hf * HALF;                 // first of this thread's 8 hidden units
  float hreg[HALF], creg[HALF], bh[G][HALF];
#pragma unroll
  for (int j = 0; j < HALF; ++j) { hreg[j] = 0.f; creg[j] = 0.f; }
  if (is_epi) {
#pragma unroll
    for (int g = 0; g < G; ++g) ld8(p.bhh + g * H + uu, bh[g]);
    if (r0 + row < tab.bs[0]) {
      if (p.h0) ld8(p.h0 + (size_t)(r0 + row) * H + uu, hreg);
      if (G == 4 && p.c0) ld8(p.c0 + (size_t)(r0 + row) * H + uu, creg);
    }
  }

  uint32_t ph = 0;       // parity of hfull[] / accbar uses
  bool w_ready = false;
  int nbar = 0;
  for (int t = 0; t < p.nsteps; ++t) {
    const int nr = min(BT, tab.bs[t] - r0);
    if (nr <= 0) break;
    const bool use_mma = (t > 0) || p.has_h0;

    if (warp == 0 && lane == 0 && use_mma) {
      proxy_fence_global();
      const CUtensorMap* src = (t == 0) ? &tmH0 : &tmH;
      const int rbase = (t == 0) ? r0 : tab.off[t - 1] + r0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_expect_tx(&hfull[kb], KBLK_A);
        tma_load_2d(sA + (size_t)kb * KBLK_A, src, kb * 64, rbase, &hfull[kb]);
      }
      if (!w_ready) { mbar_wait(wbar, 0); w_ready = true; }
      constexpr uint32_t idesc = umma_idesc(BT, NC);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&hfull[kb], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + (size_t)kb * KBLK_A);
        const uint32_t b_addr = smem_u32(sW + (size_t)kb * KBLK_W);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma(tmem_base, umma_desc_k128(a_addr + kk * 32), umma_desc_k128(b_addr + kk * 32), idesc, (kb | kk) != 0);
      }
      tc_commit(accbar);
    }

    if (is_epi) {
      const bool r_ok = row < nr;
      const size_t n = (size_t)tab.off[t] + r0 + row;
      float gx[G][HALF];
      if (r_ok) {
#pragma unroll
        for (int g = 0; g < G; ++g) ld8(p.Gx + n * (size_t)(G * H) + g * H + uu, gx[g]);
      }
      float acc[G][HALF];
      if (use_mma) {
        mbar_wait(accbar, ph);
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < G; ++g) tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + g * UT + hf * HALF, acc[g]);
      } else {
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int j = 0; j < HALF; ++j) acc[g][j] = 0.f;
      }
      if (r_ok) {
        float go[G][HALF], ghn[HALF];
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          if (G == 4) {
            const float ig = sigmoid_fast(gx[0][j] + acc[0][j] + bh[0][j]);
            const float fg = sigmoid_fast(gx[1][j] + acc[1][j] + bh[1][j]);
            const float gg = tanh_fast(gx[2][j] + acc[2][j] + bh[2][j]);
            const float og = sigmoid_fast(gx[G - 1][j] + acc[G - 1][j] + bh[G - 1][j]);
            creg[j] = fmaf(fg, creg[j], ig * gg);
            hreg[j] = og * tanh_fast(creg[j]);
            go[0][j] = ig; go[1][j] = fg; go[2][j] = gg; go[G - 1][j] = og;
          } else {
            ghn[j] = acc[2][j] + bh[2][j];
            const float rr = sigmoid_fast(gx[0][j] + acc[0][j] + bh[0][j]);
            const float zz = sigmoid_fast(gx[1][j] + acc[1][j] + bh[1][j]);
            const float nn = tanh_fast(fmaf(rr, ghn[j], gx[2][j]));
            hreg[j] = fmaf(zz, hreg[j] - nn, nn);
            go[0][j] = rr; go[1][j] = zz; go[2][j] = nn;
          }
        }
        st8(p.Hs + n * H + uu, hreg);
        st8bf(p.Hsb + n * H + uu, hreg);
        if (G == 4) st8(p.Cs + n * H + uu, creg);
        if (p.gates) {
#pragma unroll
          for (int g = 0; g < G; ++g) st8(p.gates + n * (size_t)(G * H) + g * H + uu, go[g]);
          if (G == 3) st8(p.ghn + n * H + uu, ghn);
        }
        proxy_fence_global();
      }
      tc_fence_before();
    }
    if (use_mma) ph ^= 1;
    if (t + 1 < p.nsteps) {
      ++nbar;
      grid_barrier(p.barrier + blockIdx.y, nbar * (int)gridDim.x);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0 && lane == 0 && !w_ready) mbar_wait(wbar, 0);  // never exit with a TMA in flight
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

// --------------------------------------------------------------------------------------- backward
constexpr int BSTAGES = 6;

struct TcBwdParams {
  int H, nsteps;
  const float *h0, *c0, *Hs, *Cs, *gates, *ghn, *dHs;
  __nv_bfloat16 *dG, *dGT, *dGh, *dGhT;  // (N, GH), (GH, ldt); dGh* == dG* for LSTM
  int ldt;
  float *dbih, *dbhh;                    // (GH) zero-initialised, accumulated atomically
  float* dstate;                         // (2, B0, H) out: dh0, dc0
  int* barrier;
};

template <int G>
__global__ void __launch_bounds__(NTH, 1)
rnn_seq_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmWT, const __grid_constant__ CUtensorMap tmD,
                      const __grid_constant__ StepTable tab, const TcBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, GH = G * H, KB = (GH + 63) / 64;
  constexpr uint32_t KBLK_W = UT * 128;          // 16 rows x 128 B
  uint8_t* sW = smem;                            // [KB][16 rows][128 B]   W_hh^T slice (B operand)
  uint8_t* sA = smem + (size_t)KB * KBLK_W;      // [BSTAGES][128 rows][128 B]  dGh ring (A operand)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)BSTAGES * KBLK_A);
  uint64_t* wbar = bars;
  uint64_t* accbar = bars + 1;
  uint64_t* full = bars + 2;
  uint64_t* empty = full + BSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + BSTAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UT, r0 = blockIdx.y * BT;

  if (warp == 0 && lane == 0) {
    mbar_init(wbar, 1);
    mbar_init(accbar, 1);
    for (int i = 0; i < BSTAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(32u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0 && lane == 0) {
    mbar_expect_tx(wbar, (uint32_t)KB * KBLK_W);
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + (size_t)kb * KBLK_W, &tmWT, kb * 64, u0, wbar);
  }

  const bool is_epi = warp >= 2;
  const int q = warp & 3, hf = (warp - 2) >> 2;
  const int row = q * 32 + lane, b = r0 + row;
  const int uu = u0 + hf * HALF;
  float dhrec[HALF], dcrec[HALF], dbi[G][HALF], dbn[HALF], direct[HALF];
#pragma unroll
  for (int j = 0; j < HALF; ++j) {
    dhrec[j] = 0.f; dcrec[j] = 0.f; dbn[j] = 0.f; direct[j] = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) dbi[g][j] = 0.f;
  }

  int stage_p = 0, stage_c = 0;   // ring positions of the producer (warp 0) / MMA issuer (warp 1)
  uint32_t phase_p = 0, phase_c = 0, aph = 0;
  bool w_ready = false;
  int nbar = 0;
  for (int t = p.nsteps - 1; t >= 0; --t) {
    const int nr = min(BT, tab.bs[t] - r0);
    if (nr <= 0) continue;
    const bool r_ok = row < nr;

    // ---------------- phase 1: gate gradients of this CTA's (row, unit) pairs (rnn.py:32 autograd)
    if (is_epi && r_ok) {
      const size_t n = (size_t)tab.off[t] + b;
      float dh[HALF];
      ld8(p.dHs + n * H + uu, dh);
#pragma unroll
      for (int j = 0; j < HALF; ++j) dh[j] += dhrec[j];   // dhrec is 0 for rows not live at t+1
      float da[G][HALF], dan_r[HALF];
      if (G == 4) {
        float ig[HALF], fg[HALF], gg[HALF], og[HALF], ct[HALF], cp[HALF];
        const float* gs = p.gates + n * (size_t)(4 * H) + uu;
        ld8(gs, ig); ld8(gs + H, fg); ld8(gs + 2 * H, gg); ld8(gs + 3 * H, og);
        ld8(p.Cs + n * H + uu, ct);
        if (t > 0) ld8(p.Cs + ((size_t)tab.off[t - 1] + b) * H + uu, cp);
        else if (p.c0) ld8(p.c0 + (size_t)b * H + uu, cp);
        else {
#pragma unroll
          for (int j = 0; j < HALF; ++j) cp[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          const float tc = tanh_fast(ct[j]);
          const float dc = fmaf(dh[j] * og[j], 1.f - tc * tc, dcrec[j]);
          da[0][j] = dc * gg[j] * ig[j] * (1.f - ig[j]);
          da[1][j] = dc * cp[j] * fg[j] * (1.f - fg[j]);
          da[2][j] = dc * ig[j] * (1.f - gg[j] * gg[j]);
          da[G - 1][j] = dh[j] * tc * og[j] * (1.f - og[j]);
          dcrec[j] = dc * fg[j];
          direct[j] = 0.f;
        }
      } else {
        float rr[HALF], zz[HALF], nn[HALF], gn[HALF], hp[HALF];
        const float* gs = p.gates + n * (size_t)(3 * H) + uu;
        ld8(gs, rr); ld8(gs + H, zz); ld8(gs + 2 * H, nn);
        ld8(p.ghn + n * H + uu, gn);
        if (t > 0) ld8(p.Hs + ((size_t)tab.off[t - 1] + b) * H + uu, hp);
        else if (p.h0) ld8(p.h0 + (size_t)b * H + uu, hp);
        else {
#pragma unroll
          for (int j = 0; j < HALF; ++j) hp[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          da[2][j] = dh[j] * (1.f - zz[j]) * (1.f - nn[j] * nn[j]);
          da[1][j] = dh[j] * (hp[j] - nn[j]) * zz[j] * (1.f - zz[j]);
          da[0][j] = da[2][j] * gn[j] * rr[j] * (1.f - rr[j]);
          dan_r[j] = da[2][j] * rr[j];
          direct[j] = dh[j] * zz[j];
          dbn[j] += dan_r[j];
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        st8bf(p.dG + n * (size_t)GH + g * H + uu, da[g]);
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          dbi[g][j] += da[g][j];
          p.dGT[(size_t)(g * H + uu + j) * p.ldt + n] = __float2bfloat16(da[g][j]);
        }
      }
      if (G == 3) {
        st8bf(p.dGh + n * (size_t)GH + uu, da[0]);
        st8bf(p.dGh + n * (size_t)GH + H + uu, da[1]);
        st8bf(p.dGh + n * (size_t)GH + 2 * H + uu, dan_r);
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          p.dGhT[(size_t)(uu + j) * p.ldt + n] = __float2bfloat16(da[0][j]);
          p.dGhT[(size_t)(H + uu + j) * p.ldt + n] = __float2bfloat16(da[1][j]);
          p.dGhT[(size_t)(2 * H + uu + j) * p.ldt + n] = __float2bfloat16(dan_r[j]);
        }
      }
      proxy_fence_global();
    }
    ++nbar;
    grid_barrier(p.barrier + blockIdx.y, nbar * (int)gridDim.x);

    // ---------------- phase 2: dh_{t-1}[rows, own units] = dGh_t[rows, :] . W_hh[:, own units]
    if (warp == 0 && lane == 0) {
      proxy_fence_global();
      const int rbase = tab.off[t] + r0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty[stage_p], phase_p ^ 1);
        mbar_expect_tx(&full[stage_p], KBLK_A);
        tma_load_2d(sA + (size_t)stage_p * KBLK_A, &tmD, kb * 64, rbase, &full[stage_p]);
        if (++stage_p == BSTAGES) { stage_p = 0; phase_p ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      if (!w_ready) { mbar_wait(wbar, 0); w_ready = true; }
      constexpr uint32_t idesc = umma_idesc(BT, UT);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full[stage_c], phase_c);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + (size_t)stage_c * KBLK_A);
        const uint32_t b_addr = smem_u32(sW + (size_t)kb * KBLK_W);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma(tmem_base, umma_desc_k128(a_addr + kk * 32), umma_desc_k128(b_addr + kk * 32), idesc, (kb | kk) != 0);
        tc_commit(&empty[stage_c]);
        if (++stage_c == BSTAGES) { stage_c = 0; phase_c ^= 1; }
      }
      tc_commit(accbar);
    } else if (is_epi) {
      mbar_wait(accbar, aph);
      tc_fence_after();
      float acc[HALF];
      tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + hf * HALF, acc);
      if (r_ok) {
#pragma unroll
        for (int j = 0; j < HALF; ++j) dhrec[j] = acc[j] + direct[j];
      }
      tc_fence_before();
    }
    aph ^= 1;
  }

  if (is_epi) {
    if (b < tab.bs[0]) {
      st8(p.dstate + (size_t)b * H + uu, dhrec);
      if (G == 4) st8(p.dstate + (size_t)(tab.bs[0] + b) * H + uu, dcrec);
    }
    // bias gradients: sum the per-row register accumulators over the 32 rows of the warp
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int j = 0; j < HALF; ++j) {
        const float s = warp_sum(dbi[g][j]);
        if (lane == 0) {
          atomicAdd(p.dbih + g * H + uu + j, s);
          if (G == 4 || g < 2) atomicAdd(p.dbhh + g * H + uu + j, s);
        }
      }
    if (G == 3) {
#pragma unroll
      for (int j = 0; j < HALF; ++j) {
        const float s = warp_sum(dbn[j]);
        if (lane == 0) atomicAdd(p.dbhh + 2 * H + uu + j, s);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    if (lane == 0 && !w_ready) mbar_wait(wbar, 0);
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u) : "memory");
  }
}

int coresident(const void* kern, size_t smem, int* out) {
  int dev = 0, sms = 0, per_sm = 0;
  ST_CUDA_TRY(cudaGetDevice(&dev));
  ST_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NTH, smem));
  *out = sms * per_sm;
  return ST_OK;
}

template <int G>
int launch_tc_fwd(const StepTable& tab, TcFwdParams p, const void* Whh_bf16, const void* h0_bf16, cudaStream_t s) {
  const int H = p.H, KB = (H + 63) / 64, N = tab.off[tab.nsteps], B0 = tab.bs[0];
  const size_t smem = 1024 + (size_t)KB * KBLK_A + (size_t)KB * (G * UT * 128) + 256;
  CUtensorMap tmW, tmH, tmH0;
  ST_TRY(make_tmap(&tmW, Whh_bf16, G * H, H, H, UT, "Whh_bf16"));
  ST_TRY(make_tmap(&tmH, p.Hsb, N, H, H, BT, "Hs_bf16"));
  ST_TRY(make_tmap(&tmH0, p.has_h0 ? h0_bf16 : (const void*)p.Hsb, p.has_h0 ? B0 : N, H, H, BT, "h0_bf16"));
  auto kern = rnn_seq_tc_fwd_kernel<G>;
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(H / UT, (B0 + BT - 1) / BT);
  int cores = 0;
  ST_TRY(coresident((const void*)kern, smem, &cores));
  ST_REQUIRE((int)(grid.x * grid.y) <= cores && grid.y <= 64, ST_ERR_UNSUPPORTED,
             "rnn_seq_tc_fwd: grid %ux%u is not co-resident (%d CTAs fit)", grid.x, grid.y, cores);
  ST_CUDA_TRY(cudaMemsetAsync(p.barrier, 0, sizeof(int) * 64, s));
  void* args[] = {(void*)&tmW, (void*)&tmH, (void*)&tmH0, (void*)&tab, (void*)&p};
  ST_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(NTH), args, smem, s));
  note_launch();
  return ST_OK;
}

template <int G>
int launch_tc_bwd(const StepTable& tab, TcBwdParams p, const void* WhhT_bf16, cudaStream_t s) {
  const int H = p.H, GH = G * H, KB = (GH + 63) / 64, N = tab.off[tab.nsteps], B0 = tab.bs[0];
  const size_t smem = 1024 + (size_t)KB * (UT * 128) + (size_t)BSTAGES * KBLK_A + 256;
  CUtensorMap tmWT, tmD;
  ST_TRY(make_tmap(&tmWT, WhhT_bf16, H, GH, GH, UT, "WhhT_bf16"));
  ST_TRY(make_tmap(&tmD, p.dGh, N, GH, GH, BT, "dGh_bf16"));
  auto kern = rnn_seq_tc_bwd_kernel<G>;
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(H / UT, (B0 + BT - 1) / BT);
  int cores = 0;
  ST_TRY(coresident((const void*)kern, smem, &cores));
  ST_REQUIRE((int)(grid.x * grid.y) <= cores && grid.y <= 64, ST_ERR_UNSUPPORTED,
             "rnn_seq_tc_bwd: grid %ux%u is not co-resident (%d CTAs fit)", grid.x, grid.y, cores);
  ST_CUDA_TRY(cudaMemsetAsync(p.barrier, 0, sizeof(int) * 64, s));
  ST_CUDA_TRY(cudaMemsetAsync(p.dbih, 0, sizeof(float) * GH, s));
  ST_CUDA_TRY(cudaMemsetAsync(p.dbhh, 0, sizeof(float) * GH, s));
  void* args[] = {(void*)&tmWT, (void*)&tmD, (void*)&tab, (void*)&p};
  ST_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(NTH), args, smem, s));
  note_launch();
  return ST_OK;
}

}  // namespace
}  // namespace st

extern "C" {

int st_rnn_seq_tc_supported(int kind, int H) {
  (void)kind;
  return (H % 16 == 0 && H >= 16 && H <= 64 * st::MAXKB) ? 1 : 0;
}

int st_rnn_seq_tc_fwd(int kind, int H, int nsteps, const int* batch_sizes_host, const float* Gx,
                      const void* Whh_bf16, const float* bhh, const float* h0, const void* h0_bf16,
                      const float* c0, float* Hs, void* Hs_bf16, float* Cs, float* gates, float* ghn,
                      int* barrier, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(kind == ST_GRU || kind == ST_LSTM, ST_ERR_UNSUPPORTED, "st_rnn_seq_tc_fwd: kind=%d", kind);
  ST_REQUIRE(st_rnn_seq_tc_supported(kind, H), ST_ERR_UNSUPPORTED,
             "st_rnn_seq_tc_fwd: H=%d must be a multiple of 16 and <= %d", H, 64 * MAXKB);
  ST_REQUIRE(Gx && Whh_bf16 && bhh && Hs && Hs_bf16 && barrier, ST_ERR_NULL, "st_rnn_seq_tc_fwd: NULL pointer");
  ST_REQUIRE(kind == ST_GRU || Cs, ST_ERR_NULL, "st_rnn_seq_tc_fwd: LSTM needs Cs");
  ST_REQUIRE(kind == ST_LSTM || !gates || ghn, ST_ERR_NULL, "st_rnn_seq_tc_fwd: GRU gates need ghn");
  ST_REQUIRE((h0 == nullptr) == (h0_bf16 == nullptr), ST_ERR_NULL, "st_rnn_seq_tc_fwd: h0 needs both copies");
  TcFwdParams p{H, nsteps, h0 != nullptr, Gx, bhh, h0, c0, Hs, Cs, gates, ghn,
                reinterpret_cast<__nv_bfloat16*>(Hs_bf16), barrier};
  return kind == ST_LSTM ? launch_tc_fwd<4>(tab, p, Whh_bf16, h0_bf16, as_stream(stream))
                         : launch_tc_fwd<3>(tab, p, Whh_bf16, h0_bf16, as_stream(stream));
}

int st_rnn_seq_tc_bwd(int kind, int H, int nsteps, const int* batch_sizes_host, const void* WhhT_bf16,
                      const float* h0, const float* c0, const float* Hs, const float* Cs, const float* gates,
                      const float* ghn, const float* dHs, void* dG, void* dGT, void* dGh, void* dGhT, int ldt,
                      float* dbih, float* dbhh, float* dstate, int* barrier, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(kind == ST_GRU || kind == ST_LSTM, ST_ERR_UNSUPPORTED, "st_rnn_seq_tc_bwd: kind=%d", kind);
  ST_REQUIRE(st_rnn_seq_tc_supported(kind, H), ST_ERR_UNSUPPORTED,
             "st_rnn_seq_tc_bwd: H=%d must be a multiple of 16 and <= %d", H, 64 * MAXKB);
  ST_REQUIRE(WhhT_bf16 && Hs && gates && dHs && dG && dGT && dbih && dbhh && dstate && barrier, ST_ERR_NULL,
             "st_rnn_seq_tc_bwd: NULL pointer");
  ST_REQUIRE(kind == ST_GRU || Cs, ST_ERR_NULL, "st_rnn_seq_tc_bwd: LSTM needs Cs");
  ST_REQUIRE(kind == ST_LSTM || (ghn && dGh && dGhT), ST_ERR_NULL, "st_rnn_seq_tc_bwd: GRU needs ghn, dGh, dGhT");
  ST_REQUIRE(ldt >= tab.off[nsteps] && ldt % 8 == 0, ST_ERR_BAD_SHAPE, "st_rnn_seq_tc_bwd: ldt=%d", ldt);
  if (kind == ST_LSTM) { dGh = dG; dGhT = dGT; }
  TcBwdParams p{H, nsteps, h0, c0, Hs, Cs, gates, ghn, dHs,
                reinterpret_cast<__nv_bfloat16*>(dG), reinterpret_cast<__nv_bfloat16*>(dGT),
                reinterpret_cast<__nv_bfloat16*>(dGh), reinterpret_cast<__nv_bfloat16*>(dGhT), ldt,
                dbih, dbhh, dstate, barrier};
  return kind == ST_LSTM ? launch_tc_bwd<4>(tab, p, WhhT_bf16, as_stream(stream))
                         : launch_tc_bwd<3>(tab, p, WhhT_bf16, as_stream(stream));
}

}  // extern "C"
