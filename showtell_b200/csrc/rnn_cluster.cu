// Cluster-resident recurrent kernels (bf16 mode): nn.GRU / nn.LSTM over a PackedSequence
// (rnn.py:32, LSTM/rnn_lstm.py:30) with the whole recurrence of a 32-row batch slice kept inside one
// thread-block cluster.
//
// The device-wide formulation (rnn_seq_tc.cu) spreads the hidden units of ONE batch tile over 32 SMs
// and pays an L2 round trip per step: h_t is written to global memory, a release/acquire counter is
// polled, and the next step's operand is TMA-loaded back (~7 us per step at B = 256).  Here a cluster
// of CS = H/32 CTAs (16 for H = 512; non-portable size, 7 such clusters are co-resident on a B200)
// owns a slice of RB = 16/32/48 batch rows and ALL hidden units; CTA r owns units [32r, 32r+32):
//   * its G*32 rows of W_hh (bf16, 128 KB for the LSTM at H = 512) are copied once into TENSOR MEMORY
//     (tcgen05.st, 256 columns) and are the UMMA *A* operand from there (M = 128 gate-unit rows: TMEM
//     lane quadrant q = gate q), so a step does not re-read the weights from shared memory at all;
//   * h_{t-1} of the slice (RB x H bf16) sits in every CTA's shared memory as the UMMA *B* operand
//     (N = RB batch rows), K-major without swizzle and k-chunk-major ([H/8][RB][16 B]) so that the 32
//     units a CTA produces are ONE contiguous block in every peer's tile;
//   * per step: H/16 tcgen05.mma (128 x RB x 16) into an RB-column TMEM accumulator; warp q applies
//     gate q's non-linearity for its 32 units x RB rows (all four SM sub-partitions busy, every
//     lane live); the gates meet in shared memory; c_t / h_t are formed with c (LSTM) or h (GRU)
//     carried in registers; h_t is staged as bf16 and sent with ONE bulk async copy per peer
//     (cp.async.bulk.shared::cluster.shared::cta) that completes bytes on the peer's mbarrier -- no
//     per-thread remote stores, no release fences; the MMA-issuing thread of each CTA waits on its
//     local mbarrier and issues the next step.  The h tile is single-buffered: "my MMAs of this step
//     have retired" reaches every peer as a multicast tcgen05.commit on a second mbarrier before they
//     overwrite it (that wait is over long before the gate math is);
//   * the hoisted input pre-activations Gx of step t+1 are TMA-prefetched into shared memory during
//     step t.
// No global memory is on the step-to-step dependency chain; Hs / Cs / saved gates are streamed out
// beside it.  Slices are independent, so there is no cooperative launch and any batch size works.
#include <cooperative_groups.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace st {
extern long long* g_timeline;       // rnn_seq_tc.cu (development aid: per-step %globaltimer stamps of CTA (0,0))
namespace {

constexpr int UC = 32;            // hidden units per CTA (= lanes of a warp)
// Warps: NEW = 4*(RB/16) epilogue warps (warp w: TMEM quadrant w%4 = gate, 16-column group w/4), then one
// TMA + MMA warp and one TMEM-allocator warp.
constexpr int CMAXKB = 8;         // H <= 512

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Shared-memory accesses through 32-bit shared addresses: the compiler cannot see that the carved-up dynamic
// buffer is shared memory (it would emit generic LD/ST and order every load behind the previous store).
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_b16(uint32_t a, __nv_bfloat16 v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(*reinterpret_cast<unsigned short*>(&v)) : "memory");
}
template <int NT> __device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void stamp(long long* tl, int step, int slot) {
  if (tl != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tl[step * 16 + slot] = t;
  }
}

struct ClFwdParams {
  long long* tl;
  int H, t_begin, t_end, has_h0;
  const float *bhh, *h0, *c0;
  const __nv_bfloat16 *Whh, *h0b;     // (G*H, H) row-major; (B0, H)
  float *Hs, *Cs, *gates, *ghn;
  __nv_bfloat16* Hsb;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]^T, kind::f16: the A operand (W_hh slice) lives in tensor memory.
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Completion of all prior MMAs of this thread -> one arrival on the mbarrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void tc_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// K-major operand without swizzle: 8-row x 16-byte core matrices; SBO between 8-row groups, LBO between
// the two 16-byte k-chunks of one MMA.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// shared memory of this CTA -> shared memory of a peer CTA; completes `bytes` on the peer's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}

// G = 4 (LSTM, gate order i|f|g|o) or 3 (GRU, r|z|n); RB = batch rows per cluster (= UMMA N).
// grid = (CS, nslices), cluster = (CS,1,1).
template <int G, int RB>
__global__ void __launch_bounds__((4 * (RB / 16) + 2) * 32, 1)
rnn_cluster_fwd_kernel(const __grid_constant__ CUtensorMap tmGx, const __grid_constant__ StepTable tab,
                       const ClFwdParams p) {
  constexpr int NEW = 4 * (RB / 16), NET = NEW * 32, CTH = NET + 64, MW = NEW, AW = NEW + 1;
  constexpr uint32_t TCOLS = 512, DCOL = 256, NACC = 4, ASTR = 64;   // TMEM: W_hh slice in columns [0, H/2); four
                                                       // partial accumulators (independent MMA chains) from column 256
  constexpr uint32_t CHB = RB * 16;                    // bytes of one 16-byte k-chunk column of the h tile
  constexpr uint32_t GXB = G * RB * UC * 4;            // bytes of one step's Gx tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, KB = H / 64;
  uint8_t* sH = smem;                                                    // [H/8 chunks][RB rows][16 B]
  uint8_t* hst = sH + (size_t)(H / 8) * CHB;                             // [2][4 chunks][RB rows][16 B]
  float* gxs = reinterpret_cast<float*>(hst + 2 * 4 * CHB);              // [2][G][RB][UC]
  float* exch = gxs + 2 * G * RB * UC;                                   // [4][RB][UC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(exch + 4 * RB * UC);
  uint64_t* accbar = bars;          // accumulator complete
  uint64_t* hbar = bars + 1;        // h tile complete: one local expect_tx arrival + RB*H*2 bytes from the peers
  uint64_t* fbar = bars + 2;        // h tile free: every peer's MMAs of the step have retired (multicast commits)
  uint64_t* gxbar = bars + 3;       // [2] Gx tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = blockIdx.x, CS = gridDim.x;   // cluster spans the x dimension
  const int r0 = blockIdx.y * RB;                      // first batch row of this cluster's slice
  const int u0 = rank * UC;

  if (tid == 0) {
    mbar_init(accbar, 1);
    mbar_init(hbar, 1);
    mbar_init(fbar, CS);
    mbar_init(&gxbar[0], 1);
    mbar_init(&gxbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == AW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();               // every CTA's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const bool first_from_mem = (p.t_begin > 0) || p.has_h0;   // the first step's h_{t-1} comes from global memory
  auto load_gx = [&](int t, int buf) {                 // TMA: G boxes of RB rows x 32 units (fp32)
    mbar_expect_tx(&gxbar[buf], GXB);
    for (int g = 0; g < G; ++g)
      tma_load_2d(gxs + ((size_t)buf * G + g) * RB * UC, &tmGx, g * H + u0, tab.off[t] + r0, &gxbar[buf]);
  };
  if (warp == MW && lane == 0) load_gx(p.t_begin, 0);

  // ---- W_hh slice -> tensor memory: thread (gate = warp, unit = lane) owns TMEM lane 32*warp + lane and copies
  // its weight row (H bf16, two per 32-bit column)
  if (warp < G) {
    const uint4* wrow = reinterpret_cast<const uint4*>(p.Whh + (size_t)(warp * H + u0 + lane) * H);
    for (int kb = 0; kb < KB; ++kb) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 v = __ldg(wrow + kb * 8 + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + kb * 32, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // ---- first step's h_{t-1} from global memory (h0 or the packed rows of step t_begin-1), k-chunk-major
  if (first_from_mem) {
    const __nv_bfloat16* src = (p.t_begin == 0) ? p.h0b + (size_t)r0 * H : p.Hsb + ((size_t)tab.off[p.t_begin - 1] + r0) * H;
    const int live = tab.bs[p.t_begin] - r0;
    for (int i = tid; i < RB * (H / 8); i += CTH) {
      const int b = i % RB, kc = i / RB;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (b < live) v = *reinterpret_cast<const uint4*>(src + (size_t)b * H + kc * 8);
      *reinterpret_cast<uint4*>(sH + (size_t)kc * CHB + b * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- epilogue thread state.  Stage 1 (gate g = warp, unit = lane): the row's hidden-bias.
  // Stage 2 (unit = lane, rows warp + 4k): carried c (LSTM) / h (GRU).
  constexpr int RPW = RB / NEW;     // stage-2 rows per warp (= 4)
  const bool is_epi = warp < NEW;
  const int q = warp & 3, cg = warp >> 2;   // stage 1: gate (TMEM quadrant) and 16-column group of this warp
  float carry[RPW], bh = 0.f;
#pragma unroll
  for (int k = 0; k < RPW; ++k) carry[k] = 0.f;
  if (is_epi) {
    if (q < G) bh = p.bhh[q * H + u0 + lane];
    const float* init = (G == 4) ? p.c0 : p.h0;
    const float* hist = (G == 4) ? p.Cs : p.Hs;
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
      const int b = r0 + warp + NEW * k;
      if (p.t_begin == 0) {
        if (init && b < tab.bs[0]) carry[k] = init[(size_t)b * H + u0 + lane];
      } else if (b < tab.bs[p.t_begin]) {
        carry[k] = __ldcg(hist + ((size_t)tab.off[p.t_begin - 1] + b) * H + u0 + lane);
      }
    }
  }

  uint32_t acc_ph = 0, fph = 0;
  for (int t = p.t_begin; t < p.t_end; ++t) {
    const int nr = min(RB, tab.bs[t] - r0);
    if (nr <= 0) break;                                // uniform over the cluster: the slice is finished
    const int s = t - p.t_begin;
    const bool use_mma = (s > 0) || first_from_mem;
    const bool publish = (t + 1 < p.t_end) && (tab.bs[t + 1] - r0 > 0);   // uniform: a next step exists for this slice

    if (warp == MW) {   // whole warp, uniform control flow; one elected lane issues (see elect_one in tc_common.cuh)
      stamp(p.tl, t, 0);
      if (s > 0) mbar_wait(hbar, (s - 1) & 1);         // every peer's part of h_{t-1} has landed (async-proxy writes)
      stamp(p.tl, t, 1);
      if (publish && elect_one()) {
        mbar_expect_tx(hbar, (uint32_t)RB * H * 2);    // arm the next phase before anyone can send h_t
        load_gx(t + 1, (s + 1) & 1);                   // that buffer was last read in step t-1, which is over
      }
      if (use_mma) {
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc(128, RB);
        // k-steps round-robin over NACC accumulators; fully unrolled with a running descriptor so the issue
        // loop is nothing but the MMAs
        const uint64_t d0 = umma_desc_nosw(smem_u32(sH), CHB, 128);
        const int nks = H / 16;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < CMAXKB * 4; ++ks)
            if (ks < nks)
              tc_mma_ts(tmem_base + DCOL + (ks % NACC) * ASTR, tmem_base + ks * 8, d0 + (uint64_t)(ks * ((2 * CHB) >> 4)), idesc,
                        ks >= (int)NACC);
          tc_commit(accbar);
          if (publish) tc_commit_multicast(fbar, (uint16_t)((1u << CS) - 1));
        }
        stamp(p.tl, t, 2);
      }
      __syncwarp();
    }

    if (is_epi) {
      // ---- stage 1: gate `warp` of unit `lane` for the slice's rows (accumulator lane = gate-unit row)
      const size_t nbase = (size_t)tab.off[t] + r0;
      const uint32_t gx_s = smem_u32(gxs + (size_t)(s & 1) * G * RB * UC), ex_s = smem_u32(exch);
      mbar_wait(&gxbar[s & 1], (s >> 1) & 1);
      if (q < G) {
        float v[16], gxr[16];
        const bool add_gx = !(G == 3 && q == 2);        // GRU: the n-gate's input part is added in stage 2
#pragma unroll
        for (int j = 0; j < 16; ++j) gxr[j] = add_gx ? lds_f32(gx_s + (((q * RB + cg * 16 + j) * UC + lane) << 2)) : 0.f;
        if (use_mma) {
          mbar_wait(accbar, acc_ph);
          tc_fence_after();
          if (tid == 0) stamp(p.tl, t, 4);
          const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + DCOL + cg * 16;
          tmem_ld16(tq, v);
          const int nacc = min((int)NACC, H / 16);
          for (int a = 1; a < nacc; ++a) {
            float w[16];
            tmem_ld16(tq + a * ASTR, w);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += w[j];
          }
          if (tid == 0) stamp(p.tl, t, 5);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int b = cg * 16 + j;
          const float a = v[j] + gxr[j] + bh;
          float g;
          if (G == 4) g = (q == 2) ? tanh_fast(a) : sigmoid_fast(a);
          else g = (q == 2) ? a : sigmoid_fast(a);      // GRU quadrant 2 carries gh_n = W_hn h + b_hn
          sts_f32(ex_s + (((q * RB + b) * UC + lane) << 2), g);
          if (b < nr && p.gates) {
            if (G == 4 || q < 2) p.gates[(nbase + b) * (size_t)(G * H) + q * H + u0 + lane] = g;
            else p.ghn[(nbase + b) * H + u0 + lane] = g;
          }
        }
      }
      tc_fence_before();
      if (tid == 0) stamp(p.tl, t, 6);
      epi_sync<NET>();
      // ---- stage 2: state update for unit `lane`, rows warp + 4k; h_t staged k-chunk-major as bf16
      uint8_t* hs = hst + (size_t)(s & 1) * 4 * CHB;
      const uint32_t hs_s = smem_u32(hs);
      float ge[RPW][4];
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        const int b = warp + NEW * k;
#pragma unroll
        for (int g = 0; g < 3; ++g) ge[k][g] = lds_f32(ex_s + (((g * RB + b) * UC + lane) << 2));
        ge[k][3] = (G == 4) ? lds_f32(ex_s + (((3 * RB + b) * UC + lane) << 2))      // o gate
                            : lds_f32(gx_s + (((2 * RB + b) * UC + lane) << 2));     // GRU: gx_n
      }
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        const int b = warp + NEW * k;
        float hv;
        if (G == 4) {
          carry[k] = fmaf(ge[k][1], carry[k], ge[k][0] * ge[k][2]);
          hv = ge[k][3] * tanh_fast(carry[k]);
          if (b < nr) p.Cs[(nbase + b) * H + u0 + lane] = carry[k];
        } else {
          const float nn = tanh_fast(fmaf(ge[k][0], ge[k][2], ge[k][3]));
          carry[k] = fmaf(ge[k][1], carry[k] - nn, nn);
          hv = carry[k];
          if (b < nr && p.gates) p.gates[(nbase + b) * (size_t)(G * H) + 2 * H + u0 + lane] = nn;
        }
        if (b >= nr) hv = 0.f;
        sts_b16(hs_s + (lane >> 3) * CHB + b * 16 + (lane & 7) * 2, __float2bfloat16(hv));
        if (b < nr) p.Hs[(nbase + b) * H + u0 + lane] = hv;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged h_t is read by bulk copies
      if (tid == 0) stamp(p.tl, t, 8);
      epi_sync<NET>();
      // ---- publish h_t: one bulk copy per peer into its h tile (this CTA's 4 k-chunk columns are contiguous there)
      if (publish && lane == 0) {                      // peers are spread over the epilogue warps' elected lanes
        if (use_mma) mbar_wait(fbar, fph);             // every peer's MMAs of this step have retired
        if (tid == 0) stamp(p.tl, t, 10);
        for (uint32_t j = warp; j < CS; j += NEW) {
          const uint32_t peer = (rank + j) % CS;
          bulk_copy_to_peer(mapa(smem_u32(sH) + (u0 >> 3) * CHB, peer), smem_u32(hs), 4 * CHB, mapa(smem_u32(hbar), peer));
        }
        if (tid == 0) stamp(p.tl, t, 11);
      }
      for (int i = tid; i < RB * 4; i += NET) {        // bf16 copy of h_t for the GEMMs that follow the recurrence
        const int b = i % RB, c = i / RB;
        if (b < nr) *reinterpret_cast<uint4*>(p.Hsb + (nbase + b) * H + u0 + c * 8) = *reinterpret_cast<const uint4*>(hs + (size_t)c * CHB + b * 16);
      }
    }
    if (use_mma) {
      acc_ph ^= 1;
      if (publish) fph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();               // peers may still be copying into / signalling this CTA's shared memory
  if (warp == AW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

// fp32 (rows, cols) matrix: box of box_rows x 32 columns (128 B), no swizzle.
int make_tmap_f32_box32(CUtensorMap* map, const float* ptr, int rows, int cols, int ld, int box_rows) {
  static PFN_encode enc = nullptr;
  if (!enc) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    ST_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    ST_REQUIRE(sym != nullptr && q == cudaDriverEntryPointSuccess, ST_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    enc = reinterpret_cast<PFN_encode>(sym);
  }
  ST_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 4 == 0, ST_ERR_BAD_SHAPE,
             "rnn_cluster: Gx must be 16-byte aligned with ld %% 4 == 0");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ST_REQUIRE(r == CUDA_SUCCESS, ST_ERR_CUDA, "cuTensorMapEncodeTiled(Gx) failed with CUresult %d", (int)r);
  return ST_OK;
}

template <int G, int RB>
int launch_cluster_fwd(const StepTable& tab, const ClFwdParams& p, const float* Gx, cudaStream_t s, bool probe_only,
                       int* max_clusters, bool* launched) {
  const int H = p.H, CS = H / UC, N = tab.off[tab.nsteps];
  const size_t smem = 1024 + (size_t)RB * H * 2 + 2 * 4 * RB * 16 + sizeof(float) * (2 * G + 4) * RB * UC + 64;
  auto kern = rnn_cluster_fwd_kernel<G, RB>;
  *launched = false;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      (CS > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
    cudaGetLastError();
    return ST_OK;
  }
  const int nsl = (tab.bs[p.t_begin] + RB - 1) / RB;
  if (nsl > 65535) return ST_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS, nsl);
  cfg.blockDim = dim3((4 * (RB / 16) + 2) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) {
    cudaGetLastError();
    return ST_OK;                                      // this GPU cannot co-schedule a cluster of CS CTAs
  }
  *max_clusters = nclusters;
  if (probe_only) return ST_OK;
  CUtensorMap tmGx;
  ST_TRY(make_tmap_f32_box32(&tmGx, Gx, N, G * H, G * H, RB));
  ST_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmGx, tab, p));
  note_launch();
  *launched = true;
  return ST_OK;
}

}  // namespace
}  // namespace st

extern "C" {

int st_rnn_cluster_supported(int kind, int H) {
  (void)kind;
  return (H == 64 || H == 128 || H == 256 || H == 512) ? 1 : 0;
}

int st_rnn_cluster_fwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_begin, int t_end,
                       const float* Gx, const void* Whh_bf16, const float* bhh, const float* h0, const void* h0_bf16,
                       const float* c0, float* Hs, void* Hs_bf16, float* Cs, float* gates, float* ghn,
                       st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(kind == ST_GRU || kind == ST_LSTM, ST_ERR_UNSUPPORTED, "st_rnn_cluster_fwd: kind=%d", kind);
  ST_REQUIRE(st_rnn_cluster_supported(kind, H), ST_ERR_UNSUPPORTED, "st_rnn_cluster_fwd: H=%d (need 64/128/256/512)", H);
  ST_REQUIRE(Gx && Whh_bf16 && bhh && Hs && Hs_bf16, ST_ERR_NULL, "st_rnn_cluster_fwd: NULL pointer");
  ST_REQUIRE(kind == ST_GRU || Cs, ST_ERR_NULL, "st_rnn_cluster_fwd: LSTM needs Cs");
  ST_REQUIRE(kind == ST_LSTM || !gates || ghn, ST_ERR_NULL, "st_rnn_cluster_fwd: GRU gates need ghn");
  ST_REQUIRE((h0 == nullptr) == (h0_bf16 == nullptr), ST_ERR_NULL, "st_rnn_cluster_fwd: h0 needs both copies");
  ST_REQUIRE(0 <= t_begin && t_begin < t_end && t_end <= nsteps, ST_ERR_BAD_SHAPE,
             "st_rnn_cluster_fwd: step range [%d,%d) outside [0,%d)", t_begin, t_end, nsteps);
  ClFwdParams p{g_timeline, H, t_begin, t_end, h0 != nullptr, bhh, h0, c0,
                reinterpret_cast<const __nv_bfloat16*>(Whh_bf16), reinterpret_cast<const __nv_bfloat16*>(h0_bf16),
                Hs, Cs, gates, ghn, reinterpret_cast<__nv_bfloat16*>(Hs_bf16)};
  ST_REQUIRE((reinterpret_cast<uintptr_t>(Whh_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(Hs_bf16) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(h0_bf16) & 15) == 0,
             ST_ERR_BAD_SHAPE, "st_rnn_cluster_fwd: bf16 operands must be 16-byte aligned");
  // Rows per cluster: the smallest slice height whose slices all fit in one wave of co-resident clusters.
  cudaStream_t s = as_stream(stream);
  bool ok = false;
  int maxc = 0;
  const int B = tab.bs[t_begin];
  int rb = 48;
#define ST_CL(G, RBV, probe) launch_cluster_fwd<G, RBV>(tab, p, Gx, s, probe, &maxc, &ok)
  if (kind == ST_LSTM) ST_TRY(ST_CL(4, 16, true)); else ST_TRY(ST_CL(3, 16, true));
  ST_REQUIRE(maxc >= 1, ST_ERR_UNSUPPORTED, "st_rnn_cluster_fwd: clusters of %d CTAs cannot be scheduled on this GPU",
             H / UC);
  if ((B + 15) / 16 <= maxc) rb = 16;
  else if ((B + 31) / 32 <= maxc) rb = 32;
  // more slices than co-resident clusters would run in waves: slower than the device-wide kernel (measured)
  ST_REQUIRE((B + rb - 1) / rb <= maxc, ST_ERR_UNSUPPORTED, "st_rnn_cluster_fwd: %d rows need more than %d co-resident clusters",
             B, maxc);
  if (kind == ST_LSTM) {
    if (rb == 16) ST_TRY(ST_CL(4, 16, false)); else if (rb == 32) ST_TRY(ST_CL(4, 32, false)); else ST_TRY(ST_CL(4, 48, false));
  } else {
    if (rb == 16) ST_TRY(ST_CL(3, 16, false)); else if (rb == 32) ST_TRY(ST_CL(3, 32, false)); else ST_TRY(ST_CL(3, 48, false));
  }
#undef ST_CL
  ST_REQUIRE(ok, ST_ERR_UNSUPPORTED, "st_rnn_cluster_fwd: clusters of %d CTAs cannot be scheduled on this GPU", H / UC);
  return ST_OK;
}

}  // extern "C"
