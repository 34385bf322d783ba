// One recurrent step of the attention decoders on the tensor cores with the context half of the input
// projection folded in (rnn_attn.py:70, rnn_attn_LSTM.py:72: unit([emb_t | embed(ctx_t)], h_{t-1})).
//
// The attention loop feeds the recurrence an input whose second half, embed(ctx_t), only exists after the
// attention of the same step, so its projection W_ih[:, E:] . embed(ctx_t) cannot be hoisted; as a separate
// small GEMM it costs a dependent kernel per step (~11 us of launch / fill / drain for 0.2 GFLOP).  Here it is
// the same accumulation as the recurrent product: the CTA's gate rows see the K-concatenated operand
//     [ h_{t-1} | embed(ctx_t) ]  .  [ W_hh | W_ih[:, E:] ]^T
// with both weight slices resident in shared memory and both activation tiles streamed through one TMA ring.
// LSTM: all four gates accumulate over the whole concatenated K.  GRU keeps W_hn h apart from the input side
// (n = tanh(gi_n + r * (W_hn h + b_hn))): accumulator columns [r | z | gh_n | gi_n(ctx)], the ctx k-blocks go to
// r, z (N = 32) and to the fourth group (N = 16) with two MMAs per k-step.
//
// CTA = 16 hidden units x BT batch rows (as rnn_seq_tc.cu); warp 0 = TMA producer, warp 1 = TMEM + MMA issue,
// warps 2..9 = epilogue (gate math, state update, stores).  One step has no inter-CTA dependency: an ordinary
// launch chained to the previous kernel with programmatic dependent launch; the prologue (barriers, TMEM,
// the two weight slices, which were cast long before the loop) overlaps that kernel's tail.
#include "common.cuh"
#include "tc_common.cuh"

namespace st {
namespace {

constexpr int UT = 16, HALF = 8, NTH = 320, MAXKB = 8, MAXST = 16;

struct StepXParams {
  int H, EX, t, nstages;
  const float *Gx, *bhh, *h0, *c0;      // Gx (N, G*H): hoisted W_ih[:, :E] emb + b_ih
  float *Hs, *Cs, *gates, *ghn;
  __nv_bfloat16* Hsb;
};

template <int G, int BT>
__global__ void __launch_bounds__(NTH, 1)
rnn_step_x_tc_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWx,
                     const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ StepTable tab, const StepXParams p) {
  constexpr int NG = G * UT;                     // gate rows of one weight slice k-block (48 / 64)
  constexpr uint32_t KBLK_W = NG * 128;          // bytes of one k-block of a weight slice
  constexpr uint32_t KBLK_A = BT * 128;          // bytes of one 64-wide k-block of an activation tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, KBH = (H + 63) / 64, KBX = (p.EX + 63) / 64, KT = KBH + KBX, NST = p.nstages;
  uint8_t* sW = smem;                            // [KT][NG rows][128 B]: W_hh slice k-blocks, then W_x slice k-blocks
  uint8_t* sA = smem + (size_t)KT * KBLK_W;      // [NST][BT rows][128 B] ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)NST * KBLK_A);
  uint64_t* wbar = bars;
  uint64_t* accbar = bars + 1;
  uint64_t* full = bars + 2;                     // [MAXST]
  uint64_t* empty = full + MAXST;                // [MAXST]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + MAXST);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UT, r0 = blockIdx.y * BT;
  const int t = p.t;
  const int nr = min(BT, tab.bs[t] - r0);        // > 0: the grid covers live rows only

  if (warp == 0 && lane == 0) {
    mbar_init(wbar, 1);
    mbar_init(accbar, 1);
    for (int i = 0; i < MAXST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
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

  if (warp == 0 && lane == 0) {  // resident weight slices: G boxes of 16 rows per k-block (constant during the loop)
    mbar_expect_tx(wbar, (uint32_t)KT * KBLK_W);
    for (int kb = 0; kb < KBH; ++kb)
      for (int g = 0; g < G; ++g) tma_load_2d(sW + (size_t)kb * KBLK_W + g * UT * 128, &tmWh, kb * 64, g * H + u0, wbar);
    for (int kb = 0; kb < KBX; ++kb)
      for (int g = 0; g < G; ++g)
        tma_load_2d(sW + (size_t)(KBH + kb) * KBLK_W + g * UT * 128, &tmWx, kb * 64, g * H + u0, wbar);
  }
  // everything above overlapped the previous kernel's tail; its outputs (ctx, h_{t-1}) are read from here on
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {            // ---- TMA producer: h_{t-1} k-blocks, then embed(ctx_t) k-blocks
    const int hbase = (t == 0) ? r0 : tab.off[t - 1] + r0;     // tmH maps h0 (t == 0) or the packed bf16 states
    const int xbase = tab.off[t] + r0;
    for (int k = 0; k < KT; ++k) {
      const int stage = k % NST;
      if (k >= NST) mbar_wait(&empty[stage], ((k / NST) - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(&full[stage], KBLK_A);
        if (k < KBH) tma_load_2d(sA + (size_t)stage * KBLK_A, &tmH, k * 64, hbase, &full[stage]);
        else tma_load_2d(sA + (size_t)stage * KBLK_A, &tmX, (k - KBH) * 64, xbase, &full[stage]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {     // ---- MMA issue (whole warp, uniform control flow; one elected lane issues)
    mbar_wait(wbar, 0);
    const uint64_t adesc0 = umma_desc_k128(smem_u32(sA)), bdesc0 = umma_desc_k128(smem_u32(sW));
    for (int k = 0; k < KT; ++k) {
      const int stage = k % NST;
      mbar_wait(&full[stage], (k / NST) & 1);
      tc_fence_after();
      const uint64_t ad = adesc0 + (uint64_t)(stage * (KBLK_A >> 4)), bd = bdesc0 + (uint64_t)(k * (KBLK_W >> 4));
      if (elect_one()) {
        if (G == 4 || k < KBH) {
          // LSTM: every k-block feeds all four gates.  GRU, h k-blocks: [r | z | gh_n] = columns 0..47
          constexpr uint32_t idesc = umma_idesc(BT, NG);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0);
        } else {
          // GRU, ctx k-blocks: r, z rows (N = 32) accumulate into columns 0..31; the n rows (N = 16, 32 rows = 4096 B
          // into the k-block) start / continue gi_n(ctx) in columns 48..63
          constexpr uint32_t idesc_rz = umma_idesc(BT, 2 * UT), idesc_n = umma_idesc(BT, UT);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            tc_mma(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc_rz, 1u);
            tc_mma(tmem_base + 3 * UT, ad + 2 * kk, bd + (uint64_t)((2 * UT * 128) >> 4) + 2 * kk, idesc_n,
                   (uint32_t)(((k - KBH) | kk) != 0));
          }
        }
        tc_commit(&empty[stage]);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(accbar);
    __syncwarp();
  } else {                    // ---- epilogue: warps 2..9; lane quarter q = warp % 4, unit half hf
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const bool lane_ok = (BT == 128) || lane < 16;
    const int row = (BT == 128) ? q * 32 + lane : q * 16 + lane;   // row of the batch tile held by this TMEM lane
    const int uu = u0 + hf * HALF;
    const bool r_ok = lane_ok && row < nr;
    const size_t n = (size_t)tab.off[t] + r0 + row;
    float hreg[HALF], creg[HALF], bh[G][HALF], gx[G][HALF];
#pragma unroll
    for (int j = 0; j < HALF; ++j) { hreg[j] = 0.f; creg[j] = 0.f; }
#pragma unroll
    for (int g = 0; g < G; ++g) ld8(p.bhh + g * H + uu, bh[g]);
    if (r_ok) {
      if (t == 0) {
        if (G == 3) ld8(p.h0 + (size_t)(r0 + row) * H + uu, hreg);
        if (G == 4 && p.c0) ld8(p.c0 + (size_t)(r0 + row) * H + uu, creg);
      } else {
        const size_t np = (size_t)tab.off[t - 1] + r0 + row;
        if (G == 3) ld8(p.Hs + np * H + uu, hreg);
        if (G == 4) ld8(p.Cs + np * H + uu, creg);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) ld8(p.Gx + n * (size_t)(G * H) + g * H + uu, gx[g]);
    }
    mbar_wait(accbar, 0);
    tc_fence_after();
    float acc[4][HALF];
#pragma unroll
    for (int g = 0; g < 4; ++g) tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + g * UT + hf * HALF, acc[g]);
    tc_fence_before();
    if (r_ok) {
      float go[G][HALF], ghn[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) {
        if (G == 4) {
          const float ig = sigmoid_fast(gx[0][j] + acc[0][j] + bh[0][j]);
          const float fg = sigmoid_fast(gx[1][j] + acc[1][j] + bh[1][j]);
          const float gg = tanh_fast(gx[2][j] + acc[2][j] + bh[2][j]);
          const float og = sigmoid_fast(gx[G - 1][j] + acc[3][j] + bh[G - 1][j]);
          creg[j] = fmaf(fg, creg[j], ig * gg);
          hreg[j] = og * tanh_fast(creg[j]);
          go[0][j] = ig; go[1][j] = fg; go[2][j] = gg; go[G - 1][j] = og;
        } else {
          ghn[j] = acc[2][j] + bh[2][j];
          const float rr = sigmoid_fast(gx[0][j] + acc[0][j] + bh[0][j]);
          const float zz = sigmoid_fast(gx[1][j] + acc[1][j] + bh[1][j]);
          const float nn = tanh_fast(fmaf(rr, ghn[j], gx[2][j] + acc[3][j]));
          hreg[j] = fmaf(zz, hreg[j] - nn, nn);
          go[0][j] = rr; go[1][j] = zz; go[2][j] = nn;
        }
      }
      st8bf(p.Hsb + n * H + uu, hreg);
      st8(p.Hs + n * H + uu, hreg);
      if (G == 4) st8(p.Cs + n * H + uu, creg);
      if (p.gates) {
#pragma unroll
        for (int g = 0; g < G; ++g) st8(p.gates + n * (size_t)(G * H) + g * H + uu, go[g]);
        if (G == 3) st8(p.ghn + n * H + uu, ghn);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ backward step
// One step of BPTT through the same fused step: phase 1 (thread-local) turns dh_t into the gate gradients of
// the CTA's 16 units; after the batch tile's CTAs have published them, phase 2 streams the 64/128-row gate-gradient
// tile ONCE against a resident [W_hh^T slice ; W_ih[:, E:]^T slice] pair and forms, side by side in one
// accumulator, dh_{t-1} for the CTA's 16 units (columns 0..15) and d embed(ctx_t) for the CTA's 16 context
// columns (columns 16..31) -- the latter used to be a separate dependent GEMM per step.  LSTM: one N = 32 MMA
// per k-step.  GRU: dG and dGh differ only in the n gate (dGh_n = r * dG_n), so the r, z k-blocks are streamed
// once (N = 32) and the n k-blocks twice (dGh_n -> columns 0..15, dG_n -> columns 16..31, N = 16 each).
// Requires EX == H (one grid of H/16 unit tiles covers both outputs) and H % 64 == 0.
constexpr int XB_STAGES = 9;

struct StepXBwdParams {
  int H, t, nsteps, nstages;
  const float *h0, *c0, *Hs, *Cs, *gates, *ghn, *dHs;
  __nv_bfloat16 *dG, *dGT, *dGh, *dGhT;  // (N, GH), (GH, ldt); dGh* == dG* for LSTM
  int ldt;
  float* dstate;                         // (2, B0, H) in/out: carried dh, dc
  float* dX;                             // (N, ldx) out: rows of step t of d embed(ctx)
  int ldx;
  int* barrier;
  int KQ;                                // phase 0: k-blocks of the attention-query gradient (0 = off)
};

template <int G, int BT>
__global__ void __launch_bounds__(NTH, 1)
rnn_step_x_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmWT, const __grid_constant__ CUtensorMap tmXT,
                         const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmDg,
                         const __grid_constant__ CUtensorMap tmQT, const __grid_constant__ CUtensorMap tmDq,
                         const __grid_constant__ StepTable tab, const StepXBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, GH = G * H, KB = GH / 64, KBN = H / 64, KBRZ = KB - KBN, NST = p.nstages;
  constexpr uint32_t KBLK_W = 2 * UT * 128;      // [W_hh^T rows of the 16 units ; W_x^T rows of the 16 ctx columns]
  constexpr uint32_t KBLK_A = BT * 128;
  // phase 0 (optional, p.KQ > 0): the attention of step t+1 read the PRE-step hidden state as its query
  // (rnn_attn.py:69), so dh_t also receives datt2_{t+1} . W_dec.  Those KQ k-blocks of the datt2 tile go through
  // the same ring first, against a resident W_dec^T slice, into accumulator columns 32..47.
  const int KQ = p.KQ;
  uint8_t* sW = smem;                            // [KB][32 rows][128 B]
  uint8_t* sQ = smem + (size_t)KB * KBLK_W;      // [KQ][16 rows][128 B]  W_dec^T slice
  uint8_t* sA = sQ + (size_t)KQ * (UT * 128);    // [NST][BT rows][128 B] ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)NST * KBLK_A);
  uint64_t* wbar = bars;
  uint64_t* accbar = bars + 1;
  uint64_t* qbar = bars + 2 + 2 * MAXST;         // phase-0 accumulator ready
  uint64_t* full = bars + 2;
  uint64_t* empty = full + MAXST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + MAXST + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UT, r0 = blockIdx.y * BT;
  const int t = p.t;
  const int nr = min(BT, tab.bs[t] - r0);        // > 0 by construction of the grid
  // rows of this tile that were live at step t+1 carry a gradient (and a query gradient) from it
  const bool have_next = (t + 1 < p.nsteps) && (tab.bs[t + 1] > r0);
  const bool do_q = KQ > 0 && have_next;
  // ring loads of phase 2: LSTM one per k-block; GRU the n-gate k-blocks twice (dGh, then dG)
  const int NL = (G == 4) ? KB : KB + KBN;

  if (warp == 0 && lane == 0) {
    mbar_init(wbar, 1);
    mbar_init(accbar, 1);
    mbar_init(qbar, 1);
    for (int i = 0; i < MAXST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
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

  if (warp == 0 && lane == 0) {                  // resident transposed weight slices (cast long before the loop)
    mbar_expect_tx(wbar, (uint32_t)KB * KBLK_W + (uint32_t)KQ * (UT * 128));
    for (int kb = 0; kb < KQ; ++kb) tma_load_2d(sQ + (size_t)kb * (UT * 128), &tmQT, kb * 64, u0, wbar);
    for (int kb = 0; kb < KB; ++kb) {
      tma_load_2d(sW + (size_t)kb * KBLK_W, &tmWT, kb * 64, u0, wbar);
      tma_load_2d(sW + (size_t)kb * KBLK_W + UT * 128, &tmXT, kb * 64, u0, wbar);
    }
  }
  pdl_wait();                 // the prologue overlapped the previous kernel's tail
  pdl_launch_dependents();

  // ---------------- phase 0: dh_t[rows, own units] += datt2_{t+1}[rows, :] . W_dec[:, own units]  (ring slots 0..KQ-1)
  if (do_q) {
    if (warp == 0) {
      const int qbase = tab.off[t + 1] + r0;
      for (int k = 0; k < KQ; ++k) {
        const int stage = k % NST;
        if (k >= NST) mbar_wait(&empty[stage], ((k / NST) - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], KBLK_A);
          tma_load_2d(sA + (size_t)stage * KBLK_A, &tmDq, k * 64, qbase, &full[stage]);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      mbar_wait(wbar, 0);
      const uint64_t adesc0 = umma_desc_k128(smem_u32(sA)), qdesc0 = umma_desc_k128(smem_u32(sQ));
      constexpr uint32_t idesc16 = umma_idesc(BT, UT);
      for (int k = 0; k < KQ; ++k) {
        const int stage = k % NST;
        mbar_wait(&full[stage], (k / NST) & 1);
        tc_fence_after();
        const uint64_t ad = adesc0 + (uint64_t)(stage * (KBLK_A >> 4)), bd = qdesc0 + (uint64_t)(k * ((UT * 128) >> 4));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma(tmem_base + 2 * UT, ad + 2 * kk, bd + 2 * kk, idesc16, (k | kk) != 0);
          tc_commit(&empty[stage]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(qbar);
      __syncwarp();
    }
  }
  const int C0 = do_q ? KQ : 0;   // ring uses so far: phase 2 continues the slot / parity sequence from here

  const bool is_epi = warp >= 2;
  const int q = warp & 3, hf = (warp - 2) >> 2;
  const bool lane_ok = (BT == 128) || lane < 16;
  const int row = (BT == 128) ? q * 32 + lane : q * 16 + lane, b = r0 + row;
  const int uu = u0 + hf * HALF;
  const bool r_ok = is_epi && lane_ok && row < nr;
  const size_t n = (size_t)tab.off[t] + b;
  float dhrec[HALF], dcrec[HALF], direct[HALF];
#pragma unroll
  for (int j = 0; j < HALF; ++j) { dhrec[j] = 0.f; dcrec[j] = 0.f; direct[j] = 0.f; }

  // ---------------- phase 1: gate gradients of this CTA's (row, unit) pairs (as rnn_seq_tc.cu)
  if (is_epi && do_q) {        // all lanes of the epilogue warps: tcgen05.ld is warp-collective
    mbar_wait(qbar, 0);
    tc_fence_after();
    tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + 2 * UT + hf * HALF, direct);   // `direct` reused as scratch
    tc_fence_before();
  }
  if (r_ok) {
    if (t + 1 < p.nsteps && b < tab.bs[t + 1]) {               // carried gradient of rows live at t+1
      ld8(p.dstate + (size_t)b * H + uu, dhrec);
      if (G == 4) ld8(p.dstate + (size_t)(tab.bs[0] + b) * H + uu, dcrec);
      if (do_q) {
#pragma unroll
        for (int j = 0; j < HALF; ++j) dhrec[j] += direct[j];
      }
    }
    float dh[HALF], gsv[G][HALF], pa[HALF], pb[HALF], da[G][HALF], dan_r[HALF];
    ld8(p.dHs + n * H + uu, dh);
#pragma unroll
    for (int j = 0; j < HALF; ++j) dh[j] += dhrec[j];
    const float* gs = p.gates + n * (size_t)GH + uu;
#pragma unroll
    for (int g = 0; g < G; ++g) ld8(gs + g * H, gsv[g]);
    ld8(((G == 4) ? p.Cs : p.ghn) + n * H + uu, pa);
    const float* hist = (G == 4) ? p.Cs : p.Hs;
    const float* init = (G == 4) ? p.c0 : p.h0;
    if (t > 0) ld8(hist + ((size_t)tab.off[t - 1] + b) * H + uu, pb);
    else if (init) ld8(init + (size_t)b * H + uu, pb);
    else {
#pragma unroll
      for (int j = 0; j < HALF; ++j) pb[j] = 0.f;
    }
    if (G == 4) {
#pragma unroll
      for (int j = 0; j < HALF; ++j) {
        const float ig = gsv[0][j], fg = gsv[1][j], gg = gsv[2][j], og = gsv[G - 1][j];
        const float tc = tanh_fast(pa[j]);
        const float dc = fmaf(dh[j] * og, 1.f - tc * tc, dcrec[j]);
        da[0][j] = dc * gg * ig * (1.f - ig);
        da[1][j] = dc * pb[j] * fg * (1.f - fg);
        da[2][j] = dc * ig * (1.f - gg * gg);
        da[G - 1][j] = dh[j] * tc * og * (1.f - og);
        dcrec[j] = dc * fg;
        direct[j] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < HALF; ++j) {
        const float rr = gsv[0][j], zz = gsv[1][j], nn = gsv[2][j];
        da[2][j] = dh[j] * (1.f - zz) * (1.f - nn * nn);
        da[1][j] = dh[j] * (pb[j] - nn) * zz * (1.f - zz);
        da[0][j] = da[2][j] * pa[j] * rr * (1.f - rr);
        dan_r[j] = da[2][j] * rr;
        direct[j] = dh[j] * zz;
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) st8bf(p.dG + n * (size_t)GH + g * H + uu, da[g]);
    if (G == 3) {
      st8bf(p.dGh + n * (size_t)GH + uu, da[0]);
      st8bf(p.dGh + n * (size_t)GH + H + uu, da[1]);
      st8bf(p.dGh + n * (size_t)GH + 2 * H + uu, dan_r);
    }
    // optional (NULL: the weight-gradient GEMMs read dG in place, MN-major): transposed copies
    if (p.dGT) {
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < HALF; ++j) p.dGT[(size_t)(g * H + uu + j) * p.ldt + n] = __float2bfloat16(da[g][j]);
    }
    if (G == 3 && p.dGT) {
#pragma unroll
      for (int j = 0; j < HALF; ++j) {
        p.dGhT[(size_t)(uu + j) * p.ldt + n] = __float2bfloat16(da[0][j]);
        p.dGhT[(size_t)(H + uu + j) * p.ldt + n] = __float2bfloat16(da[1][j]);
        p.dGhT[(size_t)(2 * H + uu + j) * p.ldt + n] = __float2bfloat16(dan_r[j]);
      }
    }
  }
  if (is_epi) {  // publish this warp's gate gradients of step t (8 arrivals per CTA)
    __syncwarp();
    if (lane == 0) red_release_gpu_add(p.barrier + blockIdx.y, 1);
  }

  // ---------------- phase 2
  if (warp == 0) {            // producer (whole warp, uniform; one elected lane issues)
    // arrivals the tile's counter holds from the earlier steps of this reverse pass (zeroed at t = nsteps - 1)
    int prior = 0;
    for (int tt = t + 1; tt < p.nsteps; ++tt) prior += (tab.bs[tt] > r0) ? 1 : 0;
    wait_counter_geq(p.barrier + blockIdx.y, (prior + 1) * 8 * (int)gridDim.x);
    proxy_fence_global();
    const int rbase = tab.off[t] + r0;
    for (int i = 0; i < NL; ++i) {
      const int c = C0 + i, stage = c % NST;
      if (c >= NST) mbar_wait(&empty[stage], ((c / NST) - 1) & 1);
      int kb = i;
      bool from_dg = false;
      if (G == 3 && i >= KBRZ) { kb = KBRZ + ((i - KBRZ) >> 1); from_dg = ((i - KBRZ) & 1) != 0; }
      if (elect_one()) {
        mbar_expect_tx(&full[stage], KBLK_A);
        tma_load_2d(sA + (size_t)stage * KBLK_A, from_dg ? &tmDg : &tmDh, kb * 64, rbase, &full[stage]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {     // MMA issue
    mbar_wait(wbar, 0);
    const uint64_t adesc0 = umma_desc_k128(smem_u32(sA)), bdesc0 = umma_desc_k128(smem_u32(sW));
    constexpr uint32_t idesc32 = umma_idesc(BT, 2 * UT), idesc16 = umma_idesc(BT, UT);
    for (int i = 0; i < NL; ++i) {
      const int c = C0 + i, stage = c % NST;
      mbar_wait(&full[stage], (c / NST) & 1);
      tc_fence_after();
      int kb = i;
      bool from_dg = false;
      if (G == 3 && i >= KBRZ) { kb = KBRZ + ((i - KBRZ) >> 1); from_dg = ((i - KBRZ) & 1) != 0; }
      const uint64_t ad = adesc0 + (uint64_t)(stage * (KBLK_A >> 4)), bd = bdesc0 + (uint64_t)(kb * (KBLK_W >> 4));
      if (elect_one()) {
        if (G == 4 || i < KBRZ) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc32, (i | kk) != 0);
        } else if (!from_dg) {   // dGh_n . W_hn^T  -> dh columns
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc16, 1u);
        } else {                 // dG_n . W_x,n^T -> d ctx columns
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma(tmem_base + UT, ad + 2 * kk, bd + (uint64_t)((UT * 128) >> 4) + 2 * kk, idesc16, 1u);
        }
        tc_commit(&empty[stage]);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(accbar);
    __syncwarp();
  } else {                    // epilogue
    mbar_wait(accbar, 0);
    tc_fence_after();
    float acc[HALF], accx[HALF];
    tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + hf * HALF, acc);
    tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + UT + hf * HALF, accx);
    tc_fence_before();
    if (r_ok) {
#pragma unroll
      for (int j = 0; j < HALF; ++j) dhrec[j] = acc[j] + direct[j];
      st8(p.dstate + (size_t)b * H + uu, dhrec);
      if (G == 4) st8(p.dstate + (size_t)(tab.bs[0] + b) * H + uu, dcrec);
      st8(p.dX + n * (size_t)p.ldx + uu, accx);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

template <int G, int BT>
int try_step_x_bwd(const StepTable& tab, StepXBwdParams p, const void* WhhT, const void* WxT, int ldwxt, const void* WqT,
                   int ldwqt, const void* dQ, int ldq, int AQ, cudaStream_t s, bool* launched) {
  const int H = p.H, GH = G * H, KB = GH / 64, N = tab.off[tab.nsteps];
  p.KQ = (WqT != nullptr) ? AQ / 64 : 0;
  int dev = 0, optin = 0, sms = 0;
  ST_CUDA_TRY(cudaGetDevice(&dev));
  ST_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  ST_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t fixed = 1024 + (size_t)KB * (2 * UT * 128) + (size_t)p.KQ * (UT * 128) + (4 + 2 * MAXST) * 8 + 64;
  *launched = false;
  if ((size_t)optin <= fixed) return ST_OK;
  int nst = (int)(((size_t)optin - fixed) / (BT * 128));
  nst = nst > XB_STAGES ? XB_STAGES : nst;
  if (nst < 2) return ST_OK;
  // the tile height is chosen from the FULL batch so that every launch of a reverse pass agrees on it
  // (the barrier counters carry over between the launches), and all CTAs must be co-resident (one per SM)
  const int gy_full = (tab.bs[0] + BT - 1) / BT;
  if ((H / UT) * gy_full > sms || gy_full > 64) return ST_OK;
  dim3 grid(H / UT, (tab.bs[p.t] + BT - 1) / BT);
  p.nstages = nst;
  const size_t smem = fixed + (size_t)nst * (BT * 128);
  auto kern = rnn_step_x_tc_bwd_kernel<G, BT>;
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap tmWT, tmXT, tmDh, tmDg, tmQT, tmDq;
  ST_TRY(make_tmap(&tmWT, WhhT, H, GH, GH, UT, "WhhT_bf16"));
  ST_TRY(make_tmap(&tmXT, WxT, H, GH, ldwxt, UT, "WxT_bf16"));
  ST_TRY(make_tmap(&tmDh, p.dGh, N, GH, GH, BT, "dGh_bf16"));
  ST_TRY(make_tmap(&tmDg, p.dG, N, GH, GH, BT, "dG_bf16"));
  if (p.KQ > 0) {
    ST_TRY(make_tmap(&tmQT, WqT, H, AQ, ldwqt, UT, "WdecT_bf16"));
    ST_TRY(make_tmap(&tmDq, dQ, N, AQ, ldq, BT, "datt2_bf16"));
  } else {
    tmQT = tmWT;
    tmDq = tmDh;
  }
  if (p.t == p.nsteps - 1) ST_CUDA_TRY(cudaMemsetAsync(p.barrier, 0, sizeof(int) * 64, s));
  ST_CUDA_TRY(launch_pdl(kern, grid, dim3(NTH), smem, s, tmWT, tmXT, tmDh, tmDg, tmQT, tmDq, tab, p));
  note_launch();
  *launched = true;
  return ST_OK;
}

template <int G, int BT>
int try_step_x(const StepTable& tab, StepXParams p, const void* Whh, const void* Wx, int ldwx, const void* hprev,
               int hprev_rows, const void* X, int ldx, int x_rows, cudaStream_t s, bool* launched) {
  const int H = p.H, KBH = (H + 63) / 64, KBX = (p.EX + 63) / 64, KT = KBH + KBX;
  int dev = 0, optin = 0, sms = 0;
  ST_CUDA_TRY(cudaGetDevice(&dev));
  ST_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  ST_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t fixed = 1024 + (size_t)KT * (G * UT * 128) + (2 + 2 * MAXST) * 8 + 64;
  *launched = false;
  if ((size_t)optin <= fixed) return ST_OK;
  int nst = (int)(((size_t)optin - fixed) / (BT * 128));
  nst = nst > KT ? KT : nst;
  nst = nst > MAXST ? MAXST : nst;
  if (nst < 2) return ST_OK;
  dim3 grid(H / UT, (tab.bs[p.t] + BT - 1) / BT);
  if (BT == 64 && (int)(grid.x * grid.y) > sms) return ST_OK;   // 64-row tiles only while every CTA gets its own SM
  p.nstages = nst;
  const size_t smem = fixed + (size_t)nst * (BT * 128);
  auto kern = rnn_step_x_tc_kernel<G, BT>;
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap tmWh, tmWx, tmH, tmX;
  ST_TRY(make_tmap(&tmWh, Whh, G * H, H, H, UT, "Whh_bf16"));
  ST_TRY(make_tmap(&tmWx, Wx, G * H, p.EX, ldwx, UT, "Wx_bf16"));
  ST_TRY(make_tmap(&tmH, hprev, hprev_rows, H, H, BT, "hprev_bf16"));
  ST_TRY(make_tmap(&tmX, X, x_rows, p.EX, ldx, BT, "x_bf16"));
  ST_CUDA_TRY(launch_pdl(kern, grid, dim3(NTH), smem, s, tmWh, tmWx, tmH, tmX, tab, p));
  note_launch();
  *launched = true;
  return ST_OK;
}

template <int G>
int launch_step_x(const StepTable& tab, StepXParams p, const void* Whh, const void* Wx, int ldwx, const void* hprev,
                  int hprev_rows, const void* X, int ldx, int x_rows, cudaStream_t s) {
  bool ok = false;
  ST_TRY((try_step_x<G, 64>(tab, p, Whh, Wx, ldwx, hprev, hprev_rows, X, ldx, x_rows, s, &ok)));
  if (!ok) ST_TRY((try_step_x<G, 128>(tab, p, Whh, Wx, ldwx, hprev, hprev_rows, X, ldx, x_rows, s, &ok)));
  ST_REQUIRE(ok, ST_ERR_UNSUPPORTED, "rnn_step_x_tc_fwd: H=%d EX=%d does not fit shared memory", p.H, p.EX);
  return ST_OK;
}

}  // namespace
}  // namespace st

extern "C" {

int st_rnn_step_x_tc_supported(int kind, int H, int EX) {
  (void)kind;
  return (H % 16 == 0 && H >= 16 && H <= 64 * st::MAXKB && EX % 8 == 0 && EX >= 8 && EX <= 64 * st::MAXKB) ? 1 : 0;
}

int st_rnn_step_x_tc_bwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t, const void* WhhT_bf16,
                         const void* WxT_bf16, int ldwxt, const float* h0, const float* c0, const float* Hs,
                         const float* Cs, const float* gates, const float* ghn, const float* dHs, void* dG, void* dGT,
                         void* dGh, void* dGhT, int ldt, float* dstate, float* dX, int ldx, int* barrier,
                         const void* WdecT_bf16, int ldwdt, const void* datt2_bf16, int ldq, int A, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(kind == ST_GRU || kind == ST_LSTM, ST_ERR_UNSUPPORTED, "st_rnn_step_x_tc_bwd: kind=%d", kind);
  ST_REQUIRE((WdecT_bf16 == nullptr) == (datt2_bf16 == nullptr), ST_ERR_NULL,
             "st_rnn_step_x_tc_bwd: WdecT_bf16 and datt2_bf16 go together");
  ST_REQUIRE(!WdecT_bf16 || (A % 64 == 0 && A >= 64 && A <= 64 * MAXKB && ldwdt >= A && ldwdt % 8 == 0 && ldq >= A && ldq % 8 == 0),
             ST_ERR_UNSUPPORTED, "st_rnn_step_x_tc_bwd: attention width A=%d must be a multiple of 64 and <= %d", A, 64 * MAXKB);
  ST_REQUIRE(H % 64 == 0 && H >= 64 && H <= 64 * MAXKB, ST_ERR_UNSUPPORTED,
             "st_rnn_step_x_tc_bwd: H=%d must be a multiple of 64 and <= %d (and equal the context width)", H, 64 * MAXKB);
  ST_REQUIRE(WhhT_bf16 && WxT_bf16 && Hs && gates && dHs && dG && dstate && dX && barrier, ST_ERR_NULL,
             "st_rnn_step_x_tc_bwd: NULL pointer");
  ST_REQUIRE(kind == ST_GRU || Cs, ST_ERR_NULL, "st_rnn_step_x_tc_bwd: LSTM needs Cs");
  ST_REQUIRE(kind == ST_LSTM || (ghn && dGh && (dGhT || !dGT)), ST_ERR_NULL, "st_rnn_step_x_tc_bwd: GRU needs ghn, dGh (and dGhT with dGT)");
  ST_REQUIRE(0 <= t && t < nsteps, ST_ERR_BAD_SHAPE, "st_rnn_step_x_tc_bwd: step %d outside [0,%d)", t, nsteps);
  ST_REQUIRE((!dGT || (ldt >= tab.off[nsteps] && ldt % 8 == 0)) && ldx >= H && ldx % 4 == 0 && ldwxt >= (kind == ST_LSTM ? 4 : 3) * H &&
                 ldwxt % 8 == 0,
             ST_ERR_BAD_SHAPE, "st_rnn_step_x_tc_bwd: ldt=%d ldx=%d ldwxt=%d", ldt, ldx, ldwxt);
  if (kind == ST_LSTM) { dGh = dG; dGhT = dGT; }
  StepXBwdParams p{H, t, nsteps, 0, h0, c0, Hs, Cs, gates, ghn, dHs,
                   reinterpret_cast<__nv_bfloat16*>(dG), reinterpret_cast<__nv_bfloat16*>(dGT),
                   reinterpret_cast<__nv_bfloat16*>(dGh), reinterpret_cast<__nv_bfloat16*>(dGhT), ldt, dstate, dX, ldx, barrier, 0};
  bool ok = false;
  cudaStream_t s = as_stream(stream);
  if (kind == ST_LSTM) {
    ST_TRY((try_step_x_bwd<4, 64>(tab, p, WhhT_bf16, WxT_bf16, ldwxt, WdecT_bf16, ldwdt, datt2_bf16, ldq, A, s, &ok)));
    if (!ok) ST_TRY((try_step_x_bwd<4, 128>(tab, p, WhhT_bf16, WxT_bf16, ldwxt, WdecT_bf16, ldwdt, datt2_bf16, ldq, A, s, &ok)));
  } else {
    ST_TRY((try_step_x_bwd<3, 64>(tab, p, WhhT_bf16, WxT_bf16, ldwxt, WdecT_bf16, ldwdt, datt2_bf16, ldq, A, s, &ok)));
    if (!ok) ST_TRY((try_step_x_bwd<3, 128>(tab, p, WhhT_bf16, WxT_bf16, ldwxt, WdecT_bf16, ldwdt, datt2_bf16, ldq, A, s, &ok)));
  }
  ST_REQUIRE(ok, ST_ERR_UNSUPPORTED, "st_rnn_step_x_tc_bwd: batch %d x H %d is not co-resident", tab.bs[0], H);
  return ST_OK;
}

int st_rnn_step_x_tc_fwd(int kind, int H, int EX, int nsteps, const int* batch_sizes_host, int t, const float* Gx,
                         const void* X_bf16, int ldx, const void* Whh_bf16, const void* Wx_bf16, int ldwx,
                         const float* bhh, const float* h0, const void* h0_bf16, const float* c0, float* Hs,
                         void* Hs_bf16, float* Cs, float* gates, float* ghn, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(kind == ST_GRU || kind == ST_LSTM, ST_ERR_UNSUPPORTED, "st_rnn_step_x_tc_fwd: kind=%d", kind);
  ST_REQUIRE(st_rnn_step_x_tc_supported(kind, H, EX), ST_ERR_UNSUPPORTED,
             "st_rnn_step_x_tc_fwd: H=%d (multiple of 16) / EX=%d (multiple of 8) must be <= %d", H, EX, 64 * MAXKB);
  ST_REQUIRE(Gx && X_bf16 && Whh_bf16 && Wx_bf16 && bhh && h0 && h0_bf16 && Hs && Hs_bf16, ST_ERR_NULL,
             "st_rnn_step_x_tc_fwd: NULL pointer");
  ST_REQUIRE(kind == ST_GRU || Cs, ST_ERR_NULL, "st_rnn_step_x_tc_fwd: LSTM needs Cs");
  ST_REQUIRE(kind == ST_LSTM || !gates || ghn, ST_ERR_NULL, "st_rnn_step_x_tc_fwd: GRU gates need ghn");
  ST_REQUIRE(0 <= t && t < nsteps, ST_ERR_BAD_SHAPE, "st_rnn_step_x_tc_fwd: step %d outside [0,%d)", t, nsteps);
  ST_REQUIRE(ldx >= EX && ldx % 8 == 0 && ldwx >= EX && ldwx % 8 == 0, ST_ERR_BAD_SHAPE,
             "st_rnn_step_x_tc_fwd: ldx=%d ldwx=%d must be multiples of 8 and >= EX=%d", ldx, ldwx, EX);
  const int N = tab.off[nsteps], B0 = tab.bs[0];
  StepXParams p{H, EX, t, 0, Gx, bhh, h0, c0, Hs, Cs, gates, ghn, reinterpret_cast<__nv_bfloat16*>(Hs_bf16)};
  const void* hprev = (t == 0) ? h0_bf16 : Hs_bf16;
  const int hrows = (t == 0) ? B0 : N;
  return kind == ST_LSTM
             ? launch_step_x<4>(tab, p, Whh_bf16, Wx_bf16, ldwx, hprev, hrows, X_bf16, ldx, N, as_stream(stream))
             : launch_step_x<3>(tab, p, Whh_bf16, Wx_bf16, ldwx, hprev, hrows, X_bf16, ldx, N, as_stream(stream));
}

}  // extern "C"
