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
// Generic-proxy global stores of step t are read by TMA (async proxy) in step t+1 / phase 2: the
// writers publish with red.release.gpu on the batch tile's arrival counter, the single TMA-issuing
// thread acquires it and issues fence.proxy.async before the loads.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

extern "C" int st_rowsum_bf16(float* out, const void* M, int rows, int cols, int ld, st_stream_t stream);

namespace st {
long long* g_timeline = nullptr;   // development aid, see st_debug_set_timeline
namespace {

constexpr int UT = 16, HALF = 8, NTH = 320, MAXKB = 8;
// Batch-tile height BT (template): 128 rows (UMMA M=128, accumulator row i in TMEM lane i) or 64 rows
// (UMMA M=64: accumulator row i in lane 32*(i/16) + i%16, so each epilogue warp has 16 live lanes).
// The per-step operand stream is bound by what ONE SM can pull from L2 (~30 B/clk measured), so the
// launcher picks 64-row tiles whenever twice as many CTAs are still co-resident: twice the SMs pull.

// ---- thread-block-cluster helpers of the K-split BPTT kernel
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 bytes into a peer CTA's shared memory; completes 16 bytes on the peer's mbarrier (no fences, no remote arrive)
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster, float a, float b, float c, float d, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(dst_cluster), "f"(a), "f"(b), "f"(c), "f"(d), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void lds8(uint32_t addr, float (&v)[8]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(addr + 16));
}

__device__ __forceinline__ void ld8cg(const float* p, float (&v)[8]) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// Optional per-step timeline of CTA (0,0) (development aid; NULL in production): 8 slots per step.
__device__ __forceinline__ void stamp(long long* tl, int step, int slot) {
  if (tl != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tl[step * 8 + slot] = t;
  }
}

struct TcFwdParams {
  long long* tl;
  int H, t_begin, t_end, has_h0;
  const float *Gx, *bhh, *h0, *c0;
  float *Hs, *Cs, *gates, *ghn;
  __nv_bfloat16* Hsb;
  int* barrier;
};

template <int G, int BT>
__global__ void __launch_bounds__(NTH, 1)
rnn_seq_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
                      const __grid_constant__ CUtensorMap tmH0, const __grid_constant__ StepTable tab,
                      const TcFwdParams p) {
  constexpr int NC = G * UT;                    // accumulator columns
  constexpr uint32_t KBLK_W = NC * 128;         // bytes of one k-block of the W slice
  constexpr uint32_t KBLK_A = BT * 128;         // bytes of one 64-wide k-block of the h tile
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

  // PDL (single-step launches of the attention loop): everything above -- barrier init, TMEM allocation, the
  // W_hh slice (cast long before the loop) -- overlapped the previous kernel's tail; its outputs (Gx, h_{t-1})
  // are read only from here on.  No-op under a cooperative launch.
  pdl_wait();
  pdl_launch_dependents();

  // epilogue role: warps 2..9; lane quarter q = warp % 4, unit half hf
  const bool is_epi = warp >= 2;
  const int q = warp & 3, hf = (warp - 2) >> 2;
  const bool lane_ok = (BT == 128) || lane < 16;
  const int row = (BT == 128) ? q * 32 + lane : q * 16 + lane;   // row of the batch tile held by this TMEM lane
  const int uu = u0 + hf * HALF;                 // first of this thread's 8 hidden units
  float hreg[HALF], creg[HALF], bh[G][HALF];
#pragma unroll
  for (int j = 0; j < HALF; ++j) { hreg[j] = 0.f; creg[j] = 0.f; }
  if (is_epi) {
#pragma unroll
    for (int g = 0; g < G; ++g) ld8(p.bhh + g * H + uu, bh[g]);
    if (p.t_begin == 0) {
      if (lane_ok && r0 + row < tab.bs[0]) {
        if (p.h0) ld8(p.h0 + (size_t)(r0 + row) * H + uu, hreg);
        if (G == 4 && p.c0) ld8(p.c0 + (size_t)(r0 + row) * H + uu, creg);
      }
    } else if (lane_ok && r0 + row < tab.bs[p.t_begin]) {   // resume: carried state = rows of step t_begin-1
      const size_t n = (size_t)tab.off[p.t_begin - 1] + r0 + row;
      if (G == 3) ld8cg(p.Hs + n * H + uu, hreg);
      if (G == 4) ld8cg(p.Cs + n * H + uu, creg);
    }
  }

  uint32_t ph = 0;       // parity of hfull[] / accbar uses
  bool w_ready = false;
  for (int t = p.t_begin; t < p.t_end; ++t) {
    const int nr = min(BT, tab.bs[t] - r0);
    if (nr <= 0) break;
    const bool use_mma = (t > 0) || p.has_h0;

    if (warp == 0 && use_mma) {   // whole warp, uniform control flow; one elected lane issues (see elect_one)
      // h_{t-1} is complete once every epilogue warp of every unit tile of this batch tile has
      // arrived for step t-1 (8 arrivals per CTA per step, release/acquire on the counter; the 32
      // lanes poll one address = one request)
      stamp(p.tl, t, 0);
      if (t > p.t_begin) wait_counter_geq(p.barrier + blockIdx.y, (t - p.t_begin) * 8 * (int)gridDim.x);
      stamp(p.tl, t, 1);
      proxy_fence_global();
      const CUtensorMap* src = (t == 0) ? &tmH0 : &tmH;
      const int rbase = (t == 0) ? r0 : tab.off[t - 1] + r0;
      if (elect_one()) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_expect_tx(&hfull[kb], KBLK_A);
          tma_load_2d(sA + (size_t)kb * KBLK_A, src, kb * 64, rbase, &hfull[kb]);
        }
      }
      __syncwarp();
      if (!w_ready) { mbar_wait(wbar, 0); w_ready = true; }
      constexpr uint32_t idesc = umma_idesc(BT, NC);
      const uint64_t adesc0 = umma_desc_k128(smem_u32(sA)), bdesc0 = umma_desc_k128(smem_u32(sW));
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&hfull[kb], ph);
        if (kb == 0) stamp(p.tl, t, 2);
        if (kb == KB - 1) stamp(p.tl, t, 3);
        tc_fence_after();
        // descriptor address fields are in 16-byte units: + k-block offset, + 2 per 16-element k-step
        const uint64_t ad = adesc0 + (uint64_t)(kb * (KBLK_A >> 4)), bd = bdesc0 + (uint64_t)(kb * (KBLK_W >> 4));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc, (kb | kk) != 0);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(accbar);
      __syncwarp();
    }

    if (is_epi) {
      const bool r_ok = lane_ok && row < nr;
      const size_t n = (size_t)tab.off[t] + r0 + row;
      float gx[G][HALF];
      if (r_ok) {
#pragma unroll
        for (int g = 0; g < G; ++g) ld8(p.Gx + n * (size_t)(G * H) + g * H + uu, gx[g]);
      }
      float acc[G][HALF];
      if (use_mma) {
        mbar_wait(accbar, ph);
        if (threadIdx.x == 64) stamp(p.tl, t, 4);
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < G; ++g) tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + g * UT + hf * HALF, acc[g]);
        if (threadIdx.x == 64) stamp(p.tl, t, 5);
      } else {
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int j = 0; j < HALF; ++j) acc[g][j] = 0.f;
      }
      float go[G][HALF], ghn[HALF];
      if (r_ok) {
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
        st8bf(p.Hsb + n * H + uu, hreg);      // the only store the next step depends on
      }
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 64) stamp(p.tl, t, 6);
      if (lane == 0) red_release_gpu_add(p.barrier + blockIdx.y, 1);  // this warp's part of h_t is published
      if (threadIdx.x == 64) stamp(p.tl, t, 7);
      if (r_ok) {                             // off the critical path: fp32 state and saved gates
        st8(p.Hs + n * H + uu, hreg);
        if (G == 4) st8(p.Cs + n * H + uu, creg);
        if (p.gates) {
#pragma unroll
          for (int g = 0; g < G; ++g) st8(p.gates + n * (size_t)(G * H) + g * H + uu, go[g]);
          if (G == 3) st8(p.ghn + n * H + uu, ghn);
        }
      }
    }
    if (use_mma) ph ^= 1;
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
// Phase-2 operand ring.  Measured with tools/probe_stream.cu on B200: the stream of the 64 x (G*H) gate-gradient
// tile is bound by the NUMBER of TMA operations, not by their bytes -- every operation costs ~95-190 ns of serial
// mbarrier wait + issue in the producing thread and again in the MMA-issuing thread, whatever its size (a 64-row
// k-block of 8 KB and a 64 KB box take the same time; one SM ingests > 300 GB/s when asked in 64 KB pieces).  So a
// ring stage holds KP consecutive 64-wide k-blocks fetched by ONE 3-D tensor-map box {64 columns, BT rows, KP
// k-blocks} (the row-major matrix viewed as (GH/64, N, 64)); the k-blocks land one after the other in the swizzled
// K-major layout the UMMA descriptors expect.  KP = 4 cut the stream from 6.2 to ~1.5 us per step at B = 256.
// KP = 1 keeps the 2-D boxes (any G*H; the 3-D view needs G*H % 64 == 0).
template <int KP, int BT> struct RingCfg { static constexpr int STAGES = (KP == 1) ? 9 : 131072 / (KP * BT * 128); };
struct TcBwdParams {
  long long* tl;
  int H, t_hi, t_lo, nsteps, t_zero;
  const float *h0, *c0, *Hs, *Cs, *gates, *ghn, *dHs;
  __nv_bfloat16 *dG, *dGT, *dGh, *dGhT;  // (N, GH), (GH, ldt); dGh* == dG* for LSTM
  int ldt;
  float *dbih, *dbhh;                    // (GH) zero-initialised, accumulated atomically
  float* dstate;                         // (2, B0, H) out: dh0, dc0
  int* barrier;
};

// KS = 1: every CTA streams the whole K = G*H extent of its batch tile's gate gradients.
// KS = 4: the K extent is split over a thread-block cluster of 4 unit tiles of the same batch tile.  CTA j of
// the cluster streams k-blocks [j KB/4, (j+1) KB/4) only -- a quarter of the TMA bytes and a quarter of the MMA
// instructions per step, both of which bound the step (a 64 x N x 16 MMA costs ~23 clocks whatever N is; tools/
// probe_mma.cu) -- against the W_hh^T rows of ALL 64 units of the cluster, so its accumulator is a K-partial of the
// cluster's 64 x 64 output tile.  The partials of the 16 units each CTA owns are sent to it through distributed
// shared memory (st.async, completing bytes on the owner's mbarrier) and summed there in a fixed order.
template <int G, int BT, int KP, int KS>
__global__ void __launch_bounds__(NTH, 1)
rnn_seq_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmWT, const __grid_constant__ CUtensorMap tmD,
                      const __grid_constant__ StepTable tab, const TcBwdParams p) {
  constexpr int BSTAGES = RingCfg<KP, BT>::STAGES;
  constexpr int NU = UT * KS;                    // units of the MMA's N dimension
  static_assert(KS == 1 || KS == 4, "K split: clusters of 4 unit tiles");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, GH = G * H, KB = (GH + 63) / 64;
  const int KBC = KB / KS;                       // k-blocks this CTA streams per step (launcher: KB % KS == 0, KBC % KP == 0)
  const int krank = (KS == 1) ? 0 : (int)(blockIdx.x % KS);   // rank in the cluster (cluster = KS consecutive x)
  const int kb0 = krank * KBC;
  constexpr uint32_t KBLK_W = NU * 128;          // NU rows x 128 B
  constexpr uint32_t KBLK_A = BT * 128;          // one k-block of the dGh tile
  constexpr uint32_t STAGE_A = KP * KBLK_A;      // one ring stage: KP k-blocks
  const int NOPS = (KBC + KP - 1) / KP;          // TMA operations (= ring stages consumed) per step
  uint8_t* sW = smem;                            // [KBC][NU rows][128 B]   W_hh^T slice (B operand)
  uint8_t* sA = smem + (size_t)KBC * KBLK_W;     // [BSTAGES][KP][BT rows][128 B]  dGh ring (A operand)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)BSTAGES * STAGE_A);
  uint64_t* wbar = bars;
  uint64_t* accbar = bars + 1;
  uint64_t* xbar = bars + 2;                     // K split: the peers' partials have landed
  uint64_t* full = bars + 3;
  uint64_t* empty = full + BSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + BSTAGES);
  float* xbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // K split: [KS-1 sources][BT rows][UT units]
  constexpr uint32_t XBYTES = (KS - 1) * BT * UT * 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UT, r0 = blockIdx.y * BT;

  if (warp == 0 && lane == 0) {
    mbar_init(wbar, 1);
    mbar_init(accbar, 1);
    mbar_init(xbar, 1);
    for (int i = 0; i < BSTAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // K split: arm the exchange barrier for the first step (bytes may arrive before or after; each later phase is
    // armed by the thread that saw the previous one complete -- never two arrivals in one phase)
    if (KS > 1) mbar_expect_tx(xbar, (KS - 1) * BT * UT * 4);
  }
  constexpr uint32_t TCOLS = NU < 32 ? 32u : (uint32_t)NU;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (KS > 1) cluster_sync_();    // every peer's exchange barrier exists before anyone can send to it
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0 && lane == 0) {
    mbar_expect_tx(wbar, (uint32_t)KBC * KBLK_W);
    for (int kb = 0; kb < KBC; ++kb) tma_load_2d(sW + (size_t)kb * KBLK_W, &tmWT, (kb0 + kb) * 64, u0 - krank * UT, wbar);
  }

  pdl_wait();                 // see the forward kernel: the prologue overlapped the previous kernel's tail
  pdl_launch_dependents();

  const bool is_epi = warp >= 2;
  const int q = warp & 3, hf = (warp - 2) >> 2;
  const bool lane_ok = (BT == 128) || lane < 16;
  const int row = (BT == 128) ? q * 32 + lane : q * 16 + lane, b = r0 + row;
  const int uu = u0 + hf * HALF;
  float dhrec[HALF], dcrec[HALF], direct[HALF];
#pragma unroll
  for (int j = 0; j < HALF; ++j) { dhrec[j] = 0.f; dcrec[j] = 0.f; direct[j] = 0.f; }
  if (is_epi && lane_ok && p.t_hi < p.nsteps && b < tab.bs[p.t_hi]) {   // resume: carried gradient of rows live at t_hi
    ld8cg(p.dstate + (size_t)b * H + uu, dhrec);
    if (G == 4) ld8cg(p.dstate + (size_t)(tab.bs[0] + b) * H + uu, dcrec);
  }
  // phase-1 operands of one step, prefetched into registers while the previous phase 2 streams:
  // pf_dh = dHs row, pf_g = saved gates, pf_a = c_t (LSTM) / gh_n (GRU), pf_b = c_{t-1} / h_{t-1}
  float pf_dh[HALF], pf_g[G][HALF], pf_a[HALF], pf_b[HALF];
  auto load_step = [&](int ts) {
    const int nrs = min(BT, tab.bs[ts] - r0);
    if (!is_epi || !lane_ok || row >= nrs) return;
    const size_t n = (size_t)tab.off[ts] + b;
    ld8(p.dHs + n * H + uu, pf_dh);
    const float* gs = p.gates + n * (size_t)GH + uu;
#pragma unroll
    for (int g = 0; g < G; ++g) ld8(gs + g * H, pf_g[g]);
    const float* cur = (G == 4) ? p.Cs : p.ghn;
    const float* hist = (G == 4) ? p.Cs : p.Hs;
    const float* init = (G == 4) ? p.c0 : p.h0;
    ld8(cur + n * H + uu, pf_a);
    if (ts > 0) ld8(hist + ((size_t)tab.off[ts - 1] + b) * H + uu, pf_b);
    else if (init) ld8(init + (size_t)b * H + uu, pf_b);
    else {
#pragma unroll
      for (int j = 0; j < HALF; ++j) pf_b[j] = 0.f;
    }
  };

  int stage_p = 0, stage_c = 0;   // ring positions of the producer (warp 0) / MMA issuer (warp 1)
  uint32_t phase_p = 0, phase_c = 0, aph = 0;
  bool w_ready = false, have_pf = false;
  // arrivals this batch tile's counter already holds: the steps [t_hi, t_zero) processed by earlier launches
  // since the counters were last zeroed (8 per unit tile per step in which the tile had live rows)
  int nbar = 0;
  for (int tt = p.t_hi; tt < p.t_zero; ++tt) nbar += (tab.bs[tt] > r0) ? 1 : 0;
  for (int t = p.t_hi - 1; t >= p.t_lo; --t) {
    const int nr = min(BT, tab.bs[t] - r0);
    if (nr <= 0) continue;
    const bool r_ok = lane_ok && row < nr;
    if (!have_pf) { load_step(t); have_pf = true; }

    // ---------------- phase 1: gate gradients of this CTA's (row, unit) pairs (rnn.py:32 autograd)
    if (threadIdx.x == 64) stamp(p.tl, t, 0);
    float da[G][HALF], dan_r[HALF];
    if (is_epi && r_ok) {
      const size_t n = (size_t)tab.off[t] + b;
      float dh[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) dh[j] = pf_dh[j] + dhrec[j];   // dhrec is 0 for rows not live at t+1
      if (G == 4) {
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          const float ig = pf_g[0][j], fg = pf_g[1][j], gg = pf_g[2][j], og = pf_g[G - 1][j];
          const float tc = tanh_fast(pf_a[j]);
          const float dc = fmaf(dh[j] * og, 1.f - tc * tc, dcrec[j]);
          da[0][j] = dc * gg * ig * (1.f - ig);
          da[1][j] = dc * pf_b[j] * fg * (1.f - fg);
          da[2][j] = dc * ig * (1.f - gg * gg);
          da[G - 1][j] = dh[j] * tc * og * (1.f - og);
          dcrec[j] = dc * fg;
          direct[j] = 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          const float rr = pf_g[0][j], zz = pf_g[1][j], nn = pf_g[2][j];
          da[2][j] = dh[j] * (1.f - zz) * (1.f - nn * nn);
          da[1][j] = dh[j] * (pf_b[j] - nn) * zz * (1.f - zz);
          da[0][j] = da[2][j] * pf_a[j] * rr * (1.f - rr);
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
    }
    ++nbar;
    if (is_epi) {  // publish this warp's row-major gate gradients of step t (8 arrivals per CTA per step)
      __syncwarp();
      if (threadIdx.x == 64) stamp(p.tl, t, 1);
      if (lane == 0) red_release_gpu_add(p.barrier + blockIdx.y, 1);
    }
    if (is_epi && r_ok && p.dGT) {  // optional (NULL: the weight-gradient GEMMs read dG in place, MN-major): transposed copies
      const size_t n = (size_t)tab.off[t] + b;
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < HALF; ++j) p.dGT[(size_t)(g * H + uu + j) * p.ldt + n] = __float2bfloat16(da[g][j]);
      if (G == 3) {
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          p.dGhT[(size_t)(uu + j) * p.ldt + n] = __float2bfloat16(da[0][j]);
          p.dGhT[(size_t)(H + uu + j) * p.ldt + n] = __float2bfloat16(da[1][j]);
          p.dGhT[(size_t)(2 * H + uu + j) * p.ldt + n] = __float2bfloat16(dan_r[j]);
        }
      }
    }
    if (t > p.t_lo) load_step(t - 1);  // overlaps with the phase-2 stream below

    // ---------------- phase 2: dh_{t-1}[rows, own units] = dGh_t[rows, :] . W_hh[:, own units]
    if (warp == 0) {          // whole warp, uniform control flow; one elected lane issues (see elect_one)
      wait_counter_geq(p.barrier + blockIdx.y, nbar * 8 * (int)gridDim.x);  // all unit tiles wrote dGh_t
      stamp(p.tl, t, 2);
      proxy_fence_global();
      const int rbase = tab.off[t] + r0;
      for (int op = 0; op < NOPS; ++op) {
        mbar_wait(&empty[stage_p], phase_p ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage_p], STAGE_A);   // k-blocks past the last one are zero-filled, bytes still count
          if (KP == 1) tma_load_2d(sA + (size_t)stage_p * STAGE_A, &tmD, (kb0 + op) * 64, rbase, &full[stage_p]);
          else tma_load_3d(sA + (size_t)stage_p * STAGE_A, &tmD, 0, rbase, kb0 + op * KP, &full[stage_p]);
        }
        __syncwarp();
        if (++stage_p == BSTAGES) { stage_p = 0; phase_p ^= 1; }
      }
    } else if (warp == 1) {   // whole warp, uniform control flow; one elected lane issues (see elect_one)
      if (!w_ready) { mbar_wait(wbar, 0); w_ready = true; }
      constexpr uint32_t idesc = umma_idesc(BT, NU);
      const uint64_t adesc0 = umma_desc_k128(smem_u32(sA)), bdesc0 = umma_desc_k128(smem_u32(sW));
      for (int op = 0; op < NOPS; ++op) {
        mbar_wait(&full[stage_c], phase_c);
        if (op == 0) stamp(p.tl, t, 3);
        if (op == NOPS - 1) stamp(p.tl, t, 4);
        tc_fence_after();
        // descriptor address fields are in 16-byte units: + stage / k-block offset, + 2 per 16-element k-step
        const uint64_t ad = adesc0 + (uint64_t)(stage_c * (STAGE_A >> 4));
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            const int kb = op * KP + j;
            if (kb < KBC) {
              const uint64_t bd = bdesc0 + (uint64_t)(kb * (KBLK_W >> 4));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                tc_mma(tmem_base, ad + (uint64_t)(j * (KBLK_A >> 4)) + 2 * kk, bd + 2 * kk, idesc, (kb | kk) != 0);
            }
          }
          tc_commit(&empty[stage_c]);
        }
        __syncwarp();
        if (++stage_c == BSTAGES) { stage_c = 0; phase_c ^= 1; }
      }
      if (elect_one()) tc_commit(accbar);
      __syncwarp();
    } else if (is_epi) {
      mbar_wait(accbar, aph);
      if (threadIdx.x == 64) stamp(p.tl, t, 5);
      tc_fence_after();
      float acc[HALF];
      if (KS == 1) {
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + hf * HALF, acc);
      } else {
        // K-partials of the cluster's 64 x 64 tile: columns [16 d, 16 d + 16) belong to CTA d.  Keep mine, send the
        // others' (every live lane sends, masked rows included: the owner counts bytes).
        const uint32_t xb = smem_u32(xbuf), xr = smem_u32(xbar);
#pragma unroll
        for (int d = 0; d < KS; ++d) {
          float v[HALF];
          tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + d * UT + hf * HALF, v);
          if (d == krank) {
#pragma unroll
            for (int j = 0; j < HALF; ++j) acc[j] = v[j];
          } else if (lane_ok) {
            const int slot = (krank + KS - d - 1) % KS;                      // 0 .. KS-2, one per source
            const uint32_t dst = mapa_u32(xb + (uint32_t)(((slot * BT + row) * UT + hf * HALF) * 4), (uint32_t)d);
            const uint32_t bar = mapa_u32(xr, (uint32_t)d);
            st_async_v4(dst, v[0], v[1], v[2], v[3], bar);
            st_async_v4(dst + 16, v[4], v[5], v[6], v[7], bar);
          }
        }
        mbar_wait(xbar, aph);
        if (warp == 2 && lane == 0) mbar_expect_tx(xbar, XBYTES);            // next step's phase
        if (lane_ok) {
#pragma unroll
          for (int sidx = 0; sidx < KS - 1; ++sidx) {                        // fixed order: deterministic sums
            float v[HALF];
            lds8(xb + (uint32_t)(((sidx * BT + row) * UT + hf * HALF) * 4), v);
#pragma unroll
            for (int j = 0; j < HALF; ++j) acc[j] += v[j];
          }
        }
      }
      if (r_ok) {
#pragma unroll
        for (int j = 0; j < HALF; ++j) dhrec[j] = acc[j] + direct[j];
      }
      tc_fence_before();
    }
    aph ^= 1;
  }

  if (is_epi && lane_ok && b < tab.bs[p.t_lo]) {
    st8(p.dstate + (size_t)b * H + uu, dhrec);
    if (G == 4) st8(p.dstate + (size_t)(tab.bs[0] + b) * H + uu, dcrec);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    if (lane == 0 && !w_ready) mbar_wait(wbar, 0);
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
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

template <int G, int BT>
int try_tc_fwd(const StepTable& tab, TcFwdParams p, const void* Whh_bf16, const void* h0_bf16, cudaStream_t s,
               bool* launched) {
  const int H = p.H, KB = (H + 63) / 64, N = tab.off[tab.nsteps], B0 = tab.bs[0];
  const size_t smem = 1024 + (size_t)KB * (BT * 128) + (size_t)KB * (G * UT * 128) + 256;
  auto kern = rnn_seq_tc_fwd_kernel<G, BT>;
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(H / UT, (tab.bs[p.t_begin] + BT - 1) / BT);
  int cores = 0;
  ST_TRY(coresident((const void*)kern, smem, &cores));
  *launched = (int)(grid.x * grid.y) <= cores && grid.y <= 64;
  if (!*launched) return ST_OK;
  CUtensorMap tmW, tmH, tmH0;
  ST_TRY(make_tmap(&tmW, Whh_bf16, G * H, H, H, UT, "Whh_bf16"));
  ST_TRY(make_tmap(&tmH, p.Hsb, N, H, H, BT, "Hs_bf16"));
  ST_TRY(make_tmap(&tmH0, p.has_h0 ? h0_bf16 : (const void*)p.Hsb, p.has_h0 ? B0 : N, H, H, BT, "h0_bf16"));
  if (p.t_end - p.t_begin == 1) {
    // one step has no inter-CTA dependency: an ordinary launch, programmatically chained to the previous kernel
    ST_CUDA_TRY(launch_pdl(kern, grid, dim3(NTH), smem, s, tmW, tmH, tmH0, tab, p));
    note_launch();
    return ST_OK;
  }
  ST_CUDA_TRY(cudaMemsetAsync(p.barrier, 0, sizeof(int) * 64, s));
  void* args[] = {(void*)&tmW, (void*)&tmH, (void*)&tmH0, (void*)&tab, (void*)&p};
  ST_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(NTH), args, smem, s));
  note_launch();
  return ST_OK;
}

template <int G>
int launch_tc_fwd(const StepTable& tab, TcFwdParams p, const void* Whh_bf16, const void* h0_bf16, cudaStream_t s) {
  bool ok = false;
  ST_TRY((try_tc_fwd<G, 64>(tab, p, Whh_bf16, h0_bf16, s, &ok)));   // twice the CTAs when they fit
  if (!ok) ST_TRY((try_tc_fwd<G, 128>(tab, p, Whh_bf16, h0_bf16, s, &ok)));
  ST_REQUIRE(ok, ST_ERR_UNSUPPORTED, "rnn_seq_tc_fwd: batch %d x H %d is not co-resident", tab.bs[0], p.H);
  return ST_OK;
}

// gate-gradient matrix (N, GH) bf16 viewed as (GH/64 k-blocks, N rows, 64 columns): box {64, box_rows, kp}, SW128
int make_tmap_kblocks(CUtensorMap* map, const void* ptr, int rows, int GH, int box_rows, int kp) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn enc = nullptr;
  if (!enc) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    ST_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    ST_REQUIRE(sym != nullptr && q == cudaDriverEntryPointSuccess, ST_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    enc = reinterpret_cast<EncodeFn>(sym);
  }
  ST_REQUIRE(ptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && GH % 64 == 0, ST_ERR_BAD_SHAPE,
             "rnn_seq_tc_bwd: the k-block view needs a 16-byte aligned matrix with G*H %% 64 == 0");
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(GH / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)GH * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)kp};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ST_REQUIRE(r == CUDA_SUCCESS, ST_ERR_CUDA, "cuTensorMapEncodeTiled(dGh k-block view) failed with CUresult %d", (int)r);
  return ST_OK;
}

int g_coop = 0;     // st_debug_set_coop: 1 = cooperative launches for the multi-step recurrent kernels
int g_bwd_kp = 0;   // st_debug_set_bwd_kp: 0 = choose, 1 / 2 / 4 = k-blocks per TMA operation (A/B timing)
int g_bwd_ks = 0;   // st_debug_set_bwd_ks: 0 = choose, 1 = no K split, 4 = K split over clusters of 4 unit tiles

template <int G, int BT, int KP, int KS>
int try_tc_bwd(const StepTable& tab, TcBwdParams p, const void* WhhT_bf16, cudaStream_t s, bool* launched) {
  constexpr int BSTAGES = RingCfg<KP, BT>::STAGES;
  const int H = p.H, GH = G * H, KB = (GH + 63) / 64, N = tab.off[tab.nsteps];
  *launched = false;
  if (KS > 1 && (H % (UT * KS) != 0 || GH % 64 != 0 || KB % KS != 0 || (KB / KS) % KP != 0)) return ST_OK;
  const size_t smem = 1024 + (size_t)(KB / KS) * (UT * KS * 128) + (size_t)BSTAGES * KP * (BT * 128) + 512 +
                      (size_t)(KS - 1) * BT * UT * 4 + 64;
  auto kern = rnn_seq_tc_bwd_kernel<G, BT, KP, KS>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return ST_OK;
  }
  dim3 grid(H / UT, (tab.bs[p.t_lo] + BT - 1) / BT);
  // the tile height is chosen from the FULL batch so that every launch of a reverse pass agrees on it
  // (the barrier counters carry over between the launches)
  const int gy_full = (tab.bs[0] + BT - 1) / BT;
  if (gy_full > 64) return ST_OK;
  const bool one_step = p.t_hi - p.t_lo == 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(NTH); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (KS > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = KS; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  if (KS > 1) {
    int nclusters = 0;
    cfg.gridDim = dim3(H / UT, gy_full);
    const cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    if (getenv("ST_DEBUG")) fprintf(stderr, "rnn_seq_tc_bwd K split: occupancy query %s, %d clusters of %d for %d CTAs\n",
                                    cudaGetErrorString(oe), nclusters, KS, (int)(grid.x * gy_full));
    if (oe != cudaSuccess) { cudaGetLastError(); return ST_OK; }
    if (nclusters * KS < (int)(grid.x * gy_full)) return ST_OK;
    cfg.gridDim = grid;
  } else {
    int cores = 0;
    ST_TRY(coresident((const void*)kern, smem, &cores));
    if ((int)(grid.x * gy_full) > cores) return ST_OK;
  }
  *launched = true;
  CUtensorMap tmWT, tmD;
  ST_TRY(make_tmap(&tmWT, WhhT_bf16, H, GH, GH, UT * KS, "WhhT_bf16"));
  if (KP == 1) ST_TRY(make_tmap(&tmD, p.dGh, N, GH, GH, BT, "dGh_bf16"));
  else ST_TRY(make_tmap_kblocks(&tmD, p.dGh, N, GH, BT, KP));
  // Barrier counters are zeroed by the launch that starts a reverse pass (t_hi == nsteps); later launches
  // of the pass continue them (TcBwdParams::t_zero), which spares a memset node per step.
  if (p.t_hi == p.nsteps) ST_CUDA_TRY(cudaMemsetAsync(p.barrier, 0, sizeof(int) * 64, s));
  // The CTAs meet at device-wide barriers and all fit on the device (checked above).  An ordinary launch,
  // programmatically chained to the previous kernel: a COOPERATIVE launch starts only once the GPU has drained --
  // measured 17 us of idle SMs per training step while the vocabulary bias-gradient sums of the side stream finish --
  // whereas ordinary CTAs start as SMs free up.  Nothing resident can wait on this kernel (its programmatic dependents
  // are scheduled only after every CTA of this grid is resident), so all its CTAs do become resident; the one
  // thing the caller must not do is run TWO persistent recurrent kernels on one device at the same time
  // (SHOWTELL_COOP=1 / st_debug_set_coop(1) restores cooperative launches for that case).
  if (one_step || !g_coop) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    if (g_pdl) ++na;
  } else {
    at[na].id = cudaLaunchAttributeCooperative;
    at[na].val.cooperative = 1;
    ++na;
  }
  cfg.numAttrs = na;
  ST_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmWT, tmD, tab, p));
  note_launch();
  if (p.t_lo == 0 && p.dbih != nullptr && p.dGT != nullptr) {
    // bias gradients = row sums of the transposed gate gradients (contiguous per gate row)
    ST_TRY(st_rowsum_bf16(p.dbih, p.dGT, GH, N, p.ldt, s));
    if (p.dGhT != p.dGT) ST_TRY(st_rowsum_bf16(p.dbhh, p.dGhT, GH, N, p.ldt, s));
    else ST_CUDA_TRY(cudaMemcpyAsync(p.dbhh, p.dbih, sizeof(float) * GH, cudaMemcpyDeviceToDevice, s));
  }
  return ST_OK;
}

int g_bwd_bt = 0;   // st_debug_set_bwd_ks(ks | bt << 8): batch-tile height of the K-split kernel, 0 = choose

// Batch-tile height of the K-split kernel: 64 rows.  (128-row tiles halve the CTAs but measured 8.9 us instead of
// 5.8 us per step at B = 256, H = 512: twice the stores in front of every release, twice the exchange bytes.)
inline int ksplit_bt(int B, int H) {
  (void)B; (void)H;
  return g_bwd_bt == 128 ? 128 : 64;
}

template <int G, int BT>
int try_tc_bwd_ks(const StepTable& tab, TcBwdParams p, const void* WhhT_bf16, cudaStream_t s, bool* ok) {
  const int kbc = ((G * p.H + 63) / 64) / 4;
  // 128-row tiles: 64 KB ring stages at 4 k-blocks per operation -- two stages are one step's operands
  if (kbc % 4 == 0) return try_tc_bwd<G, BT, 4, 4>(tab, p, WhhT_bf16, s, ok);
  if (kbc % 2 == 0) return try_tc_bwd<G, BT, 2, 4>(tab, p, WhhT_bf16, s, ok);
  return try_tc_bwd<G, BT, 1, 4>(tab, p, WhhT_bf16, s, ok);
}

template <int G>
int launch_tc_bwd(const StepTable& tab, TcBwdParams p, const void* WhhT_bf16, cudaStream_t s) {
  bool ok = false;
  const int GH = G * p.H, KB = (GH + 63) / 64;
  // K split over clusters of 4 unit tiles (see the kernel): a quarter of the bytes and MMAs per CTA and step
  if (g_bwd_ks != 1 && GH % 64 == 0 && KB % 4 == 0) {
    if (ksplit_bt(tab.bs[0], p.H) == 128) ST_TRY((try_tc_bwd_ks<G, 128>(tab, p, WhhT_bf16, s, &ok)));
    if (!ok) ST_TRY((try_tc_bwd_ks<G, 64>(tab, p, WhhT_bf16, s, &ok)));
    if (ok) return ST_OK;
  }
  // k-blocks per TMA operation: 4 when they tile the K extent (G*H % 256 == 0), else 2 (% 128), else 2-D boxes
  int kp = (GH % 256 == 0) ? 4 : ((GH % 128 == 0) ? 2 : 1);
  if (g_bwd_kp == 1 || (g_bwd_kp == 2 && GH % 128 == 0) || (g_bwd_kp == 4 && GH % 256 == 0)) kp = g_bwd_kp;
  if (kp == 4) ST_TRY((try_tc_bwd<G, 64, 4, 1>(tab, p, WhhT_bf16, s, &ok)));
  else if (kp == 2) ST_TRY((try_tc_bwd<G, 64, 2, 1>(tab, p, WhhT_bf16, s, &ok)));
  else ST_TRY((try_tc_bwd<G, 64, 1, 1>(tab, p, WhhT_bf16, s, &ok)));
  if (!ok && kp >= 2) ST_TRY((try_tc_bwd<G, 128, 2, 1>(tab, p, WhhT_bf16, s, &ok)));
  if (!ok) ST_TRY((try_tc_bwd<G, 128, 1, 1>(tab, p, WhhT_bf16, s, &ok)));
  ST_REQUIRE(ok, ST_ERR_UNSUPPORTED, "rnn_seq_tc_bwd: batch %d x H %d is not co-resident", tab.bs[0], p.H);
  return ST_OK;
}

}  // namespace
}  // namespace st

extern "C" {

int st_debug_set_bwd_kp(int kp) {
  st::g_bwd_kp = (kp == 1 || kp == 2 || kp == 4) ? kp : 0;
  return ST_OK;
}

int st_debug_set_bwd_ks(int ks) {
  const int bt = ks >> 8;
  ks &= 0xff;
  st::g_bwd_ks = (ks == 1 || ks == 4) ? ks : 0;
  st::g_bwd_bt = (bt == 64 || bt == 128) ? bt : 0;
  return ST_OK;
}

int st_debug_set_coop(int on) {
  st::g_coop = on ? 1 : 0;
  return ST_OK;
}

int st_rnn_seq_tc_bwd_ctas(int kind, int H, int B) {
  using namespace st;
  const int G = kind == ST_LSTM ? 4 : 3, GH = G * H, KB = (GH + 63) / 64;
  if (g_bwd_ks != 1 && GH % 64 == 0 && KB % 4 == 0 && H % 64 == 0) return (H / UT) * ((B + ksplit_bt(B, H) - 1) / ksplit_bt(B, H));
  return (H / UT) * ((B + 63) / 64);
}

/* development aid: device buffer of (nsteps * 8) int64 receiving %globaltimer stamps of CTA (0,0) */
void st_debug_set_timeline(void* dev_ptr) { st::g_timeline = reinterpret_cast<long long*>(dev_ptr); }

int st_rnn_seq_tc_supported(int kind, int H) {
  (void)kind;
  return (H % 16 == 0 && H >= 16 && H <= 64 * st::MAXKB) ? 1 : 0;
}

int st_rnn_seq_tc_fwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_begin, int t_end,
                      const float* Gx,
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
  ST_REQUIRE(0 <= t_begin && t_begin < t_end && t_end <= nsteps, ST_ERR_BAD_SHAPE,
             "st_rnn_seq_tc_fwd: step range [%d,%d) outside [0,%d)", t_begin, t_end, nsteps);
  TcFwdParams p{g_timeline, H, t_begin, t_end, h0 != nullptr, Gx, bhh, h0, c0, Hs, Cs, gates, ghn,
                reinterpret_cast<__nv_bfloat16*>(Hs_bf16), barrier};
  return kind == ST_LSTM ? launch_tc_fwd<4>(tab, p, Whh_bf16, h0_bf16, as_stream(stream))
                         : launch_tc_fwd<3>(tab, p, Whh_bf16, h0_bf16, as_stream(stream));
}

int st_rnn_seq_tc_bwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_hi, int t_lo,
                      const void* WhhT_bf16,
                      const float* h0, const float* c0, const float* Hs, const float* Cs, const float* gates,
                      const float* ghn, const float* dHs, void* dG, void* dGT, void* dGh, void* dGhT, int ldt,
                      float* dbih, float* dbhh, float* dstate, int* barrier, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(kind == ST_GRU || kind == ST_LSTM, ST_ERR_UNSUPPORTED, "st_rnn_seq_tc_bwd: kind=%d", kind);
  ST_REQUIRE(st_rnn_seq_tc_supported(kind, H), ST_ERR_UNSUPPORTED,
             "st_rnn_seq_tc_bwd: H=%d must be a multiple of 16 and <= %d", H, 64 * MAXKB);
  ST_REQUIRE(WhhT_bf16 && Hs && gates && dHs && dG && dstate && barrier, ST_ERR_NULL,
             "st_rnn_seq_tc_bwd: NULL pointer");
  ST_REQUIRE(dGT || !dbih, ST_ERR_NULL, "st_rnn_seq_tc_bwd: bias gradients are formed from dGT (pass both or neither)");
  ST_REQUIRE((dbih == nullptr) == (dbhh == nullptr), ST_ERR_NULL, "st_rnn_seq_tc_bwd: dbih/dbhh go together");
  ST_REQUIRE(0 <= t_lo && t_lo < t_hi && t_hi <= nsteps, ST_ERR_BAD_SHAPE,
             "st_rnn_seq_tc_bwd: step range [%d,%d) outside [0,%d)", t_lo, t_hi, nsteps);
  ST_REQUIRE(kind == ST_GRU || Cs, ST_ERR_NULL, "st_rnn_seq_tc_bwd: LSTM needs Cs");
  ST_REQUIRE(kind == ST_LSTM || (ghn && dGh && (dGhT || !dGT)), ST_ERR_NULL, "st_rnn_seq_tc_bwd: GRU needs ghn, dGh (and dGhT with dGT)");
  ST_REQUIRE(!dGT || (ldt >= tab.off[nsteps] && ldt % 8 == 0), ST_ERR_BAD_SHAPE, "st_rnn_seq_tc_bwd: ldt=%d", ldt);
  if (kind == ST_LSTM) { dGh = dG; dGhT = dGT; }
  TcBwdParams p{g_timeline, H, t_hi, t_lo, nsteps, nsteps, h0, c0, Hs, Cs, gates, ghn, dHs,
                reinterpret_cast<__nv_bfloat16*>(dG), reinterpret_cast<__nv_bfloat16*>(dGT),
                reinterpret_cast<__nv_bfloat16*>(dGh), reinterpret_cast<__nv_bfloat16*>(dGhT), ldt,
                dbih, dbhh, dstate, barrier};
  return kind == ST_LSTM ? launch_tc_bwd<4>(tab, p, WhhT_bf16, as_stream(stream))
                         : launch_tc_bwd<3>(tab, p, WhhT_bf16, as_stream(stream));
}

}  // extern "C"
