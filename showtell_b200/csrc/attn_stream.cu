// Streaming soft-attention step kernels (the per-step hot kernels of RNN_Attn training,
// Attention/rnn_attn.py:21-31 inside the loop of :66-74, and their backward).
//
// Per step and batch row b the work is two passes over row-b slabs that are contiguous in HBM:
//   forward   (1) e_p   = w_f . act(att1[b,p,:] + att2[b,:]) + b_f        "row dots"   over att1[b]  (P x A)
//             (2) ctx_e = sum_p alpha_p Fe[b,p,:] + b_embed               "column sums" over Fe[b]   (P x E)
//   backward  (1) dalpha_p = <dctx_e[b,:], Fe[b,p,:]> + dpen_p            "row dots"   over Fe[b]
//             (2) datt2[a] = w_f[a] sum_p de_p act'(att1[b,p,a]+att2[b,a]) "column sums" over att1[b]
// with a softmax (forward) / softmax-backward (backward) over the P scores in between.  Both are
// bound by the bytes of the two slabs (B*P*(A+E)*sizeof per step; at B=128, P=196 bf16 they stay
// L2-resident across the 20 steps).
//
// Persistent CTAs (one per SM) walk the batch rows; 8 consumer warps + 1 producer warp.  The producer's elected lane streams
// the two slabs back to back through a 12-stage, 16 KB/stage shared-memory ring with 1-D bulk
// async copies (cp.async.bulk ... mbarrier::complete_tx), so up to 192 KB are in flight per SM regardless
// of what the consumers are doing (the second slab, and then the next row's first slab, are already arriving during the softmax).
// Consumers read 16-byte vectors from shared memory (conflict-free), keep the per-column constants
// (w_f, att2 / dctx) in registers, and hand slots back through "empty" mbarriers.
//
// Shape requirements of this path: A*sizeof(T) and E*sizeof(T) in {512, 1024, 2048} bytes, 16-byte
// aligned slabs; anything else takes the generic kernels in attn.cu (same arithmetic).
#include <cuda_bf16.h>

#include <cfloat>

#include "common.cuh"
#include "tc_common.cuh"
#include "attn_stream.cuh"

namespace st {
namespace {

// Tunables, measured on B200 (tools/attn_variants.sh A/B builds): with at most one batch row
// per SM (rows <= SM count, config 3: B = 128) the step is a latency chain and wants many consumer warps and big
// chunks -- 24 warps, 4 x 48 KB; with several rows per SM (config 4: B = 512) it is bound by the bytes and two
// smaller CTAs per SM (8 warps, 3 x 32 KB each) overlap one row's softmax with the other's stream: 85 -> 47 us.
template <int NCW_, int STG_, int CHUNK_, int PER_SM_>
struct StreamCfg {
  static constexpr int NCW = NCW_, NCT = NCW_ * 32, SNT = NCW_ * 32 + 32;   // consumer warps / threads, + producer warp
  static constexpr int STG = STG_, PER_SM = PER_SM_;
  static constexpr uint32_t CHUNK = CHUNK_;
};
#ifdef ST_ATTN_NCW      // pinned A/B build
using CfgWide = StreamCfg<ST_ATTN_NCW, ST_ATTN_STG, ST_ATTN_CHUNK, ST_ATTN_CTAS_PER_SM>;
using CfgDual = CfgWide;
#else
using CfgWide = StreamCfg<24, 4, 49152, 1>;
using CfgDual = StreamCfg<8, 3, 32768, 2>;
#endif

template <typename T> struct V16;
template <> struct V16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
  }
};
template <> struct V16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
};

template <int ACT> __device__ __forceinline__ float actf(float s) {
  return ACT == 0 ? fmaxf(s, 0.2f * s) : tanhf(s);
}

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
template <int NCT> __device__ __forceinline__ void cons_sync_n() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

template <int NCW> __device__ __forceinline__ float cons_sum_n(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  cons_sync_n<NCW * 32>();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < NCW; ++i) r += red[i];
  cons_sync_n<NCW * 32>();
  return r;
}
template <int NCW> __device__ __forceinline__ float cons_max_n(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  cons_sync_n<NCW * 32>();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < NCW; ++i) r = fmaxf(r, red[i]);
  cons_sync_n<NCW * 32>();
  return r;
}


// VPL1 = 16-byte vectors per lane of a pass-1 row (row bytes / 512); RV2 = vectors per pass-2 row.
template <typename T, int ACT, bool BWD, int VPL1, int RV2, typename C>
__global__ void __launch_bounds__(C::SNT) attn_stream_kernel(const int nrows, const StreamParams p) {
  constexpr int NCW = C::NCW, NCT = C::NCT, STG = C::STG;
  constexpr uint32_t CHUNK = C::CHUNK;
  auto cons_sync = [] { cons_sync_n<NCT>(); };
  auto cons_sum = [](float v, float* red) { return cons_sum_n<NCW>(v, red); };
  auto cons_max = [](float v, float* red) { return cons_max_n<NCW>(v, red); };
  constexpr int EPV = V16<T>::N;
  constexpr int NG = NCT / RV2;                   // pass-2 row groups
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* s_e = reinterpret_cast<float*>(ring + STG * CHUNK);     // [P]
  __shared__ float red[NCW];
  __shared__ float s_part[NCT * 8];               // pass-2 cross-group reduction: [NG][RV2*EPV] <= 2048 floats
  __shared__ uint64_t full[STG], empty[STG];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = p.P;
  const int R1 = BWD ? p.E : p.A, R2 = BWD ? p.A : p.E;          // row elements of pass 1 / pass 2
  const uint32_t rb1 = R1 * sizeof(T), rb2 = R2 * sizeof(T);
  const int cp1 = CHUNK / rb1, cp2 = CHUNK / rb2;                // rows per chunk
  const int nc1 = (P + cp1 - 1) / cp1, nc2 = (P + cp2 - 1) / cp2;

  if (tid == 0) {
    for (int i = 0; i < STG; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW) {  // ------------------------------------------------------------- producer
    // PDL: the slabs (att1, Fe) are written once per iteration, long before the loop -- the producer does not
    // wait for the previous kernel and fills the ring while that kernel (a small GEMM) is still running.
    pdl_launch_dependents();
    if (lane == 0) {
      int c = 0;
      for (int b = blockIdx.x; b < nrows; b += gridDim.x) {
        const uint8_t* slab1 = reinterpret_cast<const uint8_t*>(BWD ? p.Fe : p.att1) + (size_t)b * P * rb1;
        const uint8_t* slab2 = reinterpret_cast<const uint8_t*>(BWD ? p.att1 : p.Fe) + (size_t)b * P * rb2;
        for (int k = 0; k < nc1 + nc2; ++k, ++c) {
          const int stage = c % STG, use = c / STG;
          if (use > 0) mbar_wait(&empty[stage], (use - 1) & 1);
          const bool first = k < nc1;
          const int ci = first ? k : k - nc1, cp = first ? cp1 : cp2;
          const uint32_t rb = first ? rb1 : rb2;
          const int rows = min(cp, P - ci * cp);
          const uint8_t* src = (first ? slab1 : slab2) + (size_t)ci * cp * rb;
          mbar_expect_tx(&full[stage], rows * rb);
          bulk_load(ring + stage * CHUNK, src, rows * rb, &full[stage]);
        }
      }
    }
    return;
  }

  // --------------------------------------------------------------------------------- consumers
  pdl_wait();               // att2 / dctx / alphas come from (or are still read by) the previous kernels
  pdl_launch_dependents();
  int c = 0;
  for (int b = blockIdx.x; b < nrows; b += gridDim.x) {
  // pass-1 per-lane constants: lane owns vectors lane, lane+32, ... of every row
  float w1[VPL1][EPV], a1[VPL1][EPV];
#pragma unroll
  for (int k = 0; k < VPL1; ++k)
#pragma unroll
    for (int j = 0; j < EPV; ++j) {
      const int el = (lane + 32 * k) * EPV + j;
      if (BWD) { w1[k][j] = p.dctx[(size_t)b * p.ld_dctx + el]; a1[k][j] = 0.f; }
      else     { w1[k][j] = p.wf[el]; a1[k][j] = p.att2[(size_t)b * p.A + el]; }
    }

  for (int ci = 0; ci < nc1; ++ci, ++c) {
    const int stage = c % STG;
    mbar_wait(&full[stage], (c / STG) & 1);
    const uint8_t* buf = ring + stage * CHUNK;
    const int rows = min(cp1, P - ci * cp1);
    for (int r = warp; r < rows; r += NCW) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < VPL1; ++k) {
        const uint4 u = *reinterpret_cast<const uint4*>(buf + (size_t)r * rb1 + (lane + 32 * k) * 16);
        float x[EPV];
        V16<T>::unpack(u, x);
#pragma unroll
        for (int j = 0; j < EPV; ++j)
          acc = BWD ? fmaf(w1[k][j], x[j], acc) : fmaf(w1[k][j], actf<ACT>(x[j] + a1[k][j]), acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) s_e[ci * cp1 + r] = acc;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }
  cons_sync();

  // softmax over the P locations (nn.Softmax(dim=1), rnn_attn.py:26) / its backward
  if (!BWD) {
    const float bf = p.bfp[0];
    float m = -FLT_MAX;
    for (int q = tid; q < P; q += NCT) m = fmaxf(m, s_e[q] + bf);
    m = cons_max(m, red);
    float s = 0.f;
    for (int q = tid; q < P; q += NCT) {
      const float w = expf(s_e[q] + bf - m);
      s_e[q] = w;
      s += w;
    }
    s = cons_sum(s, red);
    const float inv = 1.f / s;
    for (int q = tid; q < P; q += NCT) {
      const float al = s_e[q] * inv;
      s_e[q] = al;
      p.alphas_w[(size_t)b * p.alpha_stride + q] = al;
      p.S[(size_t)b * P + q] += al;
    }
  } else {
    float dot = 0.f;
    for (int q = tid; q < P; q += NCT) {
      const float da = s_e[q] + (p.dal ? p.dal[(size_t)b * p.dal_stride + q] : 0.f);
      s_e[q] = da;
      dot += p.alphas_r[(size_t)b * p.alpha_stride + q] * da;
    }
    dot = cons_sum(dot, red);
    for (int q = tid; q < P; q += NCT) {
      const float de = p.alphas_r[(size_t)b * p.alpha_stride + q] * (s_e[q] - dot);
      s_e[q] = de;
      p.de_out[(size_t)b * P + q] = de;
    }
  }
  cons_sync();

  // pass 2: thread owns vector column v of group g; groups interleave the rows of a chunk
  const int v = tid % RV2, g = tid / RV2;
  float a2[EPV], acc2[EPV];
#pragma unroll
  for (int j = 0; j < EPV; ++j) {
    acc2[j] = 0.f;
    a2[j] = BWD ? p.att2[(size_t)b * p.A + v * EPV + j] : 0.f;
  }
  for (int ci = 0; ci < nc2; ++ci, ++c) {
    const int stage = c % STG;
    mbar_wait(&full[stage], (c / STG) & 1);
    const uint8_t* buf = ring + stage * CHUNK;
    const int rows = min(cp2, P - ci * cp2);
#pragma unroll 4
    for (int r = g; r < rows; r += NG) {
      const uint4 u = *reinterpret_cast<const uint4*>(buf + (size_t)r * rb2 + v * 16);
      const float sc = s_e[ci * cp2 + r];
      float x[EPV];
      V16<T>::unpack(u, x);
      if (!BWD) {
#pragma unroll
        for (int j = 0; j < EPV; ++j) acc2[j] = fmaf(sc, x[j], acc2[j]);
      } else if (ACT == 0) {
        const float sc2 = 0.2f * sc;
#pragma unroll
        for (int j = 0; j < EPV; ++j) acc2[j] += (x[j] + a2[j] > 0.f) ? sc : sc2;
      } else {
#pragma unroll
        for (int j = 0; j < EPV; ++j) {
          const float t = tanhf(x[j] + a2[j]);
          acc2[j] = fmaf(sc, 1.f - t * t, acc2[j]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }
  // reduce the NG row groups
#pragma unroll
  for (int j = 0; j < EPV; ++j) s_part[g * (RV2 * EPV) + v * EPV + j] = acc2[j];
  cons_sync();
  for (int col = tid; col < R2; col += NCT) {
    float s = 0.f;
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) s += s_part[gg * (RV2 * EPV) + col];
    if (!BWD) {
      s += p.b_embed[col];
      p.ctx_out[(size_t)b * p.ld_ctx + col] = s;
      if (p.ctx_bf16) p.ctx_bf16[(size_t)b * p.ld_ctx_bf16 + col] = __float2bfloat16(s);
    } else {
      if (p.gt_out) p.gt_out[(size_t)b * p.A + col] = s;
      s *= p.wf[col];
      p.datt2[(size_t)b * p.A + col] = s;
      if (p.datt2_bf16) p.datt2_bf16[(size_t)b * p.A + col] = __float2bfloat16(s);
    }
  }
  }  // rows
}

template <typename T, int ACT, bool BWD, int VPL1, typename C>
int launch_rv2(int rows, const StreamParams& p, int rv2, cudaStream_t s, int sms) {
  const size_t smem = 128 + (size_t)C::STG * C::CHUNK + sizeof(float) * (size_t)p.P;
  const int cap = sms * C::PER_SM;
#define ST_GO(RV2)                                                                                         \
  do {                                                                                                     \
    auto kern = attn_stream_kernel<T, ACT, BWD, VPL1, RV2, C>;                                             \
    ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    ST_CUDA_TRY(launch_pdl(kern, dim3(rows < cap ? rows : cap), dim3(C::SNT), smem, s, rows, p));          \
  } while (0)
  if (rv2 == 32) ST_GO(32);
  else if (rv2 == 64) ST_GO(64);
  else ST_GO(128);
#undef ST_GO
  note_launch();
  return ST_OK;
}

template <typename T, int ACT, bool BWD>
int launch_vpl(int rows, const StreamParams& p, cudaStream_t s) {
  const int R1 = BWD ? p.E : p.A, R2 = BWD ? p.A : p.E;
  const int vpl1 = R1 * (int)sizeof(T) / 512, rv2 = R2 * (int)sizeof(T) / 16;
  int sms = 0;
  ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
  if (rows <= sms) {     // at most one row per SM: the wide configuration
    if (vpl1 == 1) return launch_rv2<T, ACT, BWD, 1, CfgWide>(rows, p, rv2, s, sms);
    if (vpl1 == 2) return launch_rv2<T, ACT, BWD, 2, CfgWide>(rows, p, rv2, s, sms);
    return launch_rv2<T, ACT, BWD, 4, CfgWide>(rows, p, rv2, s, sms);
  }
  if (vpl1 == 1) return launch_rv2<T, ACT, BWD, 1, CfgDual>(rows, p, rv2, s, sms);
  if (vpl1 == 2) return launch_rv2<T, ACT, BWD, 2, CfgDual>(rows, p, rv2, s, sms);
  return launch_rv2<T, ACT, BWD, 4, CfgDual>(rows, p, rv2, s, sms);
}

bool row_ok(int elems, int esz) {
  const int bytes = elems * esz;
  return bytes == 512 || bytes == 1024 || bytes == 2048;
}

}  // namespace

// Returns 1 if the streaming path handles this shape (and launches it), 0 if the caller must use
// the generic kernels, < 0 on error.
int attn_stream_try(bool bwd, int rows, int in_bf16, int act, const StreamParams& p, cudaStream_t s) {
  const int esz = in_bf16 ? 2 : 4;
  if (!row_ok(p.A, esz) || !row_ok(p.E, esz) || p.P > 4096) return 0;
  if ((reinterpret_cast<uintptr_t>(p.att1) | reinterpret_cast<uintptr_t>(p.Fe)) & 15) return 0;
  int st;
#define ST_PICK(T)                                                                       \
  st = bwd ? (act == 0 ? launch_vpl<T, 0, true>(rows, p, s) : launch_vpl<T, 1, true>(rows, p, s))   \
           : (act == 0 ? launch_vpl<T, 0, false>(rows, p, s) : launch_vpl<T, 1, false>(rows, p, s))
  if (in_bf16) ST_PICK(__nv_bfloat16); else ST_PICK(float);
#undef ST_PICK
  return st == ST_OK ? 1 : st;
}

}  // namespace st
