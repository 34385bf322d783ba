// Decoding loops, run entirely on the stream (no host synchronisation inside a call):
//   st_decode_greedy      RNN.sentence_index(beam_size=0)      rnn.py:44-58, rnn_lstm.py:35-57
//   st_decode_beam_chain  RNN.sentence_index(beam_size=K)      rnn.py:60-108   ("chain" beam)
//   st_decode_beam_tree   beam_search.beam_search()            beam_search.py:45-97 ("tree" beam)
// Images are independent, so every step is batched over the images of the call.  Two arithmetic modes:
//   gemm_mode 0 (fp32, CUDA cores): per step {embedding gather, input GEMM, rnn_seq single-step kernel, vocabulary
//     GEMM into (rows, V) logits, row-wise arg-max / top-K}.
//   gemm_mode 1 (3xTF32, tensor cores; fp32-accurate): every product of the loop on tcgen05 and the logits are NEVER
//     written: per dependent step {W_hh GEMM, gate kernel, vocabulary GEMM with the top-K (+ soft-max normaliser)
//     kept in its epilogue, candidate merge} = 4 launches.  The input projection of a fed-back word is a row of the
//     table  EP = emb . W_ih^T + b_ih  (V rows, built once per call; per-step GEMM for short calls), which the gate
//     kernel gathers by token id.
// A per-image bookkeeping kernel reproduces the reference's ranking rules (documented at each kernel).
#include <cuda_bf16.h>
#include <type_traits>

#include <cfloat>

#include "common.cuh"

extern "C" int st_rnn_seq_fwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_begin,
                              int t_end, const float* Gx, const float* Whh, const float* bhh,
                              const float* h0, const float* c0, float* Hs, float* Cs, float* gates,
                              float* ghn, int* barrier, st_stream_t stream);
extern "C" int st_argmax_rows(const float* X, int ld, int rows, int cols, int64_t* idx, int idx_stride,
                              st_stream_t stream);
extern "C" int st_split_tf32(const float* src, int rows, int cols, int lds, float* hi, float* lo, int ldd,
                             st_stream_t stream);
extern "C" int st_gemm_tf32x3(int M, int N, int K, const float* A_hi, const float* A_lo, int lda, const float* B_hi,
                              const float* B_lo, int ldb, float* C, int ldc, const float* bias, float alpha,
                              float beta, st_stream_t stream);
extern "C" int st_topk_rows(const float* X, int ld, int rows, int cols, int K, float* val,
                            int32_t* idx, int out_stride, st_stream_t stream);
extern "C" int st_topk_parts(int N);
extern "C" int st_gemm_tf32x3_topk(int M, int N, int K, const float* A_hi, const float* A_lo, int lda, const float* B_hi,
                                   const float* B_lo, int ldb, const float* bias, int topk, float* cand_val,
                                   int32_t* cand_idx, float* val, int32_t* idx, int out_stride, int64_t* tok,
                                   int tok_stride, float* part_stats, float* row_max, float* row_sum,
                                   st_stream_t stream);

extern "C" int st_gemm_bf16_screen(int M, int N, int K, const void* A, int lda, const void* B, int ldb, const float* bias,
                                 float* cand_val, int32_t* cand_idx, int* npart_out, st_stream_t stream);
extern "C" int st_cast_bf16(const float* src, int rows, int cols, int lds, void* dst, int ldd, void* dstT, int lddT,
                            st_stream_t stream);

extern "C" int st_row_norm_max(const float* W, int rows, int cols, float* out, st_stream_t stream);
extern "C" int st_vocab_topk_screen(int M, int V, int H, const float* h, const void* h_bf16, const float* Wv, const void* Wv_bf16,
                                    const float* bv, const float* wmax, int K, float* cand_val, int32_t* cand_idx, float* val,
                                    int32_t* idx, int out_stride, int64_t* tok, int tok_stride, st_stream_t stream);

namespace st {
namespace {

constexpr int MAXL = 16;

struct Bump {
  char* base;
  int64_t cap, used;
  template <typename T>
  T* take(int64_t n) {
    used = (used + 255) & ~int64_t(255);
    T* p = reinterpret_cast<T*>(base + used);
    used += n * (int64_t)sizeof(T);
    return p;
  }
};

int check_weights(const st_rnn_weights* w) {
  ST_REQUIRE(w != nullptr, ST_ERR_NULL, "decode: weights struct is NULL");
  ST_REQUIRE(w->kind == ST_GRU || w->kind == ST_LSTM, ST_ERR_UNSUPPORTED, "decode: kind=%d", w->kind);
  ST_REQUIRE(w->L >= 1 && w->L <= MAXL, ST_ERR_BAD_SHAPE, "decode: L=%d outside [1,%d]", w->L, MAXL);
  ST_REQUIRE(w->E >= 1 && w->H >= 4 && w->H % 4 == 0 && w->V >= 1, ST_ERR_BAD_SHAPE,
             "decode: E=%d H=%d V=%d", w->E, w->H, w->V);
  ST_REQUIRE(w->emb && w->Wih_host && w->Whh_host && w->bih_host && w->bhh_host && w->Wv && w->bv,
             ST_ERR_NULL, "decode: NULL weight pointer");
  for (int l = 0; l < w->L; ++l)
    ST_REQUIRE(w->Wih_host[l] && w->Whh_host[l] && w->bih_host[l] && w->bhh_host[l], ST_ERR_NULL,
               "decode: NULL weight pointer in layer %d", l);
  return ST_OK;
}

// The nn.Linear products of the loops (input projection, vocabulary projection) run either on the CUDA
// cores in fp32 (gemm_mode 0) or on the tensor cores as 3xTF32 (gemm_mode 1: fp32-accurate, see
// st_gemm_tf32x3); the latter needs rows of 4 floats (E, H multiples of 4).
inline bool use_tc(const st_rnn_weights* w) { return w->gemm_mode == 1 && w->E % 4 == 0 && w->H % 4 == 0; }

constexpr int FUSED_TOPK_MAX = 8;   // TOPK_SLOTS of the fused epilogue (gemm_tc.cu)
int g_force_table = 0;              // st_debug_decode_table: 0 = by size, 1 = always, -1 = never
int g_screen = 0;                   // st_debug_decode_screen: 0 / 1 = bf16 screening + exact re-scoring, -1 = 3xTF32 fused top-K

__device__ __forceinline__ void split_store(float x, float* hi, float* lo) {
  const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);   // as st_split_tf32
  *hi = h;
  *lo = x - h;
}

template <typename I>
__global__ void gather_emb_kernel(float* __restrict__ X, const float* __restrict__ emb, int E,
                                  const I* __restrict__ tok, int stride) {
  const int64_t id = (int64_t)tok[(size_t)blockIdx.x * stride];
  const float* src = emb + id * E;
  float* dst = X + (size_t)blockIdx.x * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = src[e];
}

// Where a step's input projection Gx = W_ih x + b_ih comes from: a (rows, G*H) matrix, or -- the fed-back word of
// each row given by tok[row * stride] (int32 or int64) -- rows of the projected-embedding table EP (V, G*H).
struct GxSrc {
  const float* gx;
  const void* tok;
  int tok64, stride;
};

// One recurrent step from the two projections (tensor-core path of the decoding loops): Gx (see GxSrc) and
// Gh = W_hh h + b_hh (NULL on the first step, h = 0: Gh = b_hh) -> gates, state update (the formulas of
// rnn_seq.cu, i.e. nn.GRU / nn.LSTM), h' written as fp32 AND as its (hi, lo) tf32 split -- the operand of the next
// W_hh / vocabulary / next-layer products, so no separate split pass exists in the loop.
template <int G>
__global__ void __launch_bounds__(256) decode_gate_kernel(int rows, int H, GxSrc src, const float* __restrict__ Gh,
                                                          const float* __restrict__ bhh, const float* __restrict__ h_prev,
                                                          const float* __restrict__ c_prev, float* __restrict__ h_out,
                                                          float* __restrict__ c_out, float* __restrict__ h_hi,
                                                          float* __restrict__ h_lo, __nv_bfloat16* __restrict__ h_bf) {
  // a thread owns 4 adjacent hidden units of one row (H % 4 == 0: every access is a 16-byte vector)
  const int H4 = H >> 2;
  const long long n4 = (long long)rows * H4;
  for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i4 / H4), u = (int)(i4 - (long long)r * H4) * 4;
    int64_t gr = r;
    if (src.tok)
      gr = src.tok64 ? reinterpret_cast<const int64_t*>(src.tok)[(size_t)r * src.stride]
                     : (int64_t) reinterpret_cast<const int32_t*>(src.tok)[(size_t)r * src.stride];
    const float* gx = src.gx + (size_t)gr * G * H + u;
    const float* gh = (Gh ? Gh + (size_t)r * G * H : bhh) + u;
    float x[G][4], y[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float4 a = *reinterpret_cast<const float4*>(gx + g * H), b = *reinterpret_cast<const float4*>(gh + g * H);
      x[g][0] = a.x; x[g][1] = a.y; x[g][2] = a.z; x[g][3] = a.w;
      y[g][0] = b.x; y[g][1] = b.y; y[g][2] = b.z; y[g][3] = b.w;
    }
    const size_t i = (size_t)r * H + u;
    float prev[4] = {0.f, 0.f, 0.f, 0.f};
    const float* pp = (G == 4) ? c_prev : h_prev;
    if (pp) { const float4 v = *reinterpret_cast<const float4*>(pp + i); prev[0] = v.x; prev[1] = v.y; prev[2] = v.z; prev[3] = v.w; }
    float hv[4], cv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (G == 4) {
        const float ig = sigmoidf_(x[0][k] + y[0][k]);
        const float fg = sigmoidf_(x[1][k] + y[1][k]);
        const float gg = tanhf(x[2][k] + y[2][k]);
        const float og = sigmoidf_(x[G - 1][k] + y[G - 1][k]);
        cv[k] = fmaf(fg, prev[k], ig * gg);
        hv[k] = og * tanhf(cv[k]);
      } else {
        const float rr = sigmoidf_(x[0][k] + y[0][k]);
        const float zz = sigmoidf_(x[1][k] + y[1][k]);
        const float nn = tanhf(fmaf(rr, y[2][k], x[2][k]));
        hv[k] = fmaf(zz, prev[k] - nn, nn);
      }
    }
    if (G == 4) *reinterpret_cast<float4*>(c_out + i) = make_float4(cv[0], cv[1], cv[2], cv[3]);
    *reinterpret_cast<float4*>(h_out + i) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    float hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) split_store(hv[k], &hi[k], &lo[k]);
    *reinterpret_cast<float4*>(h_hi + i) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(h_lo + i) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    if (h_bf) {      // operand of the screening GEMM (top layer)
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(hv[0], hv[1]), p1 = __floats2bfloat162_rn(hv[2], hv[3]);
      *reinterpret_cast<uint2*>(h_bf + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Vocabulary projection of the tensor-core decoding loops as "screen, then re-score" (rnn.py:50-51, 62-63, 88-91).
// The ranking of 10 000 logits needs fp32 accuracy only among the few columns that can be in the top K at all:
//   1. screening: the bf16 tensor-core product (one MMA per k-step instead of the three of 3xTF32) with the 8 largest
//      approximate logits of every 128-column part kept in the epilogue (st_gemm_bf16_screen; logits never written);
//   2. error bound of a row: both operands are rounded to bf16 (relative 2^-9 each), so
//        |approx_v - exact_v| <= 2^-8 (1 + 2^-10) sum_i |h_i w_vi| (+ fp32 accumulation)  <=  eps := c |h| max_v |w_v|
//      (Cauchy-Schwarz; c = 2^-7.9).  With a_K = the K-th largest KEPT approximation (a lower bound of the K-th largest
//      approximation overall), K columns have exact logits >= a_K - eps, hence so has the K-th largest exact logit, and
//      every column of the exact top K has an approximation >= tau := a_K - 2 eps (the kept values also carry their
//      column in the low 7 mantissa bits -- gemm_tc.cu -- a further 2^-16 relative, added to eps below);
//   3. a part whose last kept value is < tau has kept every column >= tau; otherwise ALL of its columns are re-scored;
//   4. the survivors (typically 5-20 of 10 000) get their exact logit b_v + <h, w_v> in fp32 (one warp per row, fixed
//      summation order) and the top K of those is the result: value descending, lower column first among equals.
// One warp per row.  wmax = max_v |w_v|_2 (one scalar per call).
// ---------------------------------------------------------------------------------------------------------------
constexpr float SCREEN_C = 0.0041874f;   // 2^-7.9

__global__ void __launch_bounds__(256) row_norm_max_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) { const float v = W[(size_t)r * cols + c]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(s)));   // non-negative floats order as ints
}

template <int KMAX>
__device__ __forceinline__ void topk_insert(float (&bv)[KMAX], int (&bi)[KMAX], float v, int i) {
  // sorted list (value descending, lower index first among equals); an index already present is ignored
#pragma unroll
  for (int q = 0; q < KMAX; ++q)
    if (bi[q] == i) return;
  if (!(v > bv[KMAX - 1] || (v == bv[KMAX - 1] && i < bi[KMAX - 1]))) return;
  bv[KMAX - 1] = v; bi[KMAX - 1] = i;
#pragma unroll
  for (int q = KMAX - 1; q > 0; --q) {
    if (bv[q] > bv[q - 1] || (bv[q] == bv[q - 1] && bi[q] < bi[q - 1])) {
      const float tv = bv[q]; bv[q] = bv[q - 1]; bv[q - 1] = tv;
      const int ti = bi[q]; bi[q] = bi[q - 1]; bi[q - 1] = ti;
    }
  }
}

constexpr int SC_CAND = 64, SC_OVER = 96;  // per row: survivors re-scored one by one / parts re-scored as a whole
constexpr int SC_PPL = 3;                  // fast path: a lane owns the kept values of up to 3 parts (npart <= 96)

// KMAX: list length (4 for K <= 4: the insertion is a third of the instructions of a re-scored column).
// A kept value carries its column: (128-aligned base of its part) + 127 - (low 7 mantissa bits)  (gemm_tc.cu, EPI_TOPS);
// cand_idx is what the screening GEMM wrote for the general path (npart > 96).
template <int KMAX>
__global__ void __launch_bounds__(256, 4) screen_select_kernel(int M, int H, int V, int npart, int part_cols, int K,
                                                            const float* __restrict__ cand_val, const int32_t* __restrict__ cand_idx,
                                                            const float* __restrict__ h, const float* __restrict__ Wv,
                                                            const float* __restrict__ bv, const float* __restrict__ wmax,
                                                            float* __restrict__ val, int32_t* __restrict__ idx, int out_stride,
                                                            int64_t* __restrict__ tok, int tok_stride) {
  extern __shared__ __align__(16) float s_h[];      // [8 warps][H]
  __shared__ int s_cand[8][SC_CAND];                // survivors of each warp's row
  __shared__ unsigned short s_over[8][SC_OVER];     // parts of each warp's row that are re-scored as a whole (the fast
  __shared__ int s_nover[8];                        // path has at most SC_OVER parts: the list cannot overflow there)
  __shared__ float s_res[128];                      // exact logits of one such part (phase B)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, m = blockIdx.x * 8 + wid;
  const bool valid = m < M;
  float* hs = s_h + (size_t)wid * H;
  constexpr int SL = ST_SCREEN_SLOTS;
  // 128-bit path of the exact dot products (H = 512 at the bench dims)
  const bool vec = (H & 127) == 0 && ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(Wv)) & 15) == 0;
  float bestv[KMAX];
  int besti[KMAX];
#pragma unroll
  for (int q = 0; q < KMAX; ++q) { bestv[q] = -FLT_MAX; besti[q] = 0x7fffffff - q; }
  // exact logits of two columns of a row whose state sits in `hrow` (shared memory).  A fixed summation order: lane j
  // sums its elements in increasing order (vec: the float4s j, j + 32, ... of the row, x y z w in turn; else the
  // elements i = j mod 32), the 32 partial sums meet in the xor-shuffle tree -- a column's value does not depend on
  // who computes it; every lane ends with both values.  All loads of a 512-element block of both rows are in flight
  // before the first multiply.
  auto exact2 = [&](const float* hrow, int va, int vb, float& sa, float& sb) {
    const float* wa = Wv + (size_t)va * H;
    const float* wb = Wv + (size_t)vb * H;
    const float ba = bv[va], bb = bv[vb];              // issued with the rows, not after the reduction
    sa = 0.f; sb = 0.f;
    if (vec) {
      const float4* wa4 = reinterpret_cast<const float4*>(wa);
      const float4* wb4 = reinterpret_cast<const float4*>(wb);
      const float4* h4 = reinterpret_cast<const float4*>(hrow);
      const int n4 = H >> 2;
      if ((n4 & 127) == 0) {                           // H a multiple of 512: four 128-bit loads per column in flight
        for (int i0 = 0; i0 < n4; i0 += 128) {
          float4 xa[4], xb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            xa[j] = __ldg(wa4 + i0 + lane + 32 * j);
            xb[j] = __ldg(wb4 + i0 + lane + 32 * j);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 hv = h4[i0 + lane + 32 * j];
            sa = fmaf(hv.x, xa[j].x, sa); sa = fmaf(hv.y, xa[j].y, sa); sa = fmaf(hv.z, xa[j].z, sa); sa = fmaf(hv.w, xa[j].w, sa);
            sb = fmaf(hv.x, xb[j].x, sb); sb = fmaf(hv.y, xb[j].y, sb); sb = fmaf(hv.z, xb[j].z, sb); sb = fmaf(hv.w, xb[j].w, sb);
          }
        }
      } else {                                         // same order of additions, one load per column at a time
        for (int i = lane; i < n4; i += 32) {
          const float4 xa = __ldg(wa4 + i), xb = __ldg(wb4 + i), hv = h4[i];
          sa = fmaf(hv.x, xa.x, sa); sa = fmaf(hv.y, xa.y, sa); sa = fmaf(hv.z, xa.z, sa); sa = fmaf(hv.w, xa.w, sa);
          sb = fmaf(hv.x, xb.x, sb); sb = fmaf(hv.y, xb.y, sb); sb = fmaf(hv.z, xb.z, sb); sb = fmaf(hv.w, xb.w, sb);
        }
      }
    } else {
      for (int i0 = 0; i0 < H; i0 += 512) {
        float xa[16], xb[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int i = i0 + lane + 32 * j;
          xa[j] = i < H ? wa[i] : 0.f;
          xb[j] = i < H ? wb[i] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int i = i0 + lane + 32 * j;
          if (i < H) { const float hv = hrow[i]; sa = fmaf(hv, xa[j], sa); sb = fmaf(hv, xb[j], sb); }
        }
      }
    }
    sa = warp_sum(sa) + ba;
    sb = warp_sum(sb) + bb;
  };
  auto part_first = [&](int pt) { return (pt >> 1) * (2 * part_cols) + (pt & 1) * part_cols; };
  int ncnd = 0, nover = 0;
  // one 32-part group of the threshold pass: lane's part `part` kept xv[] (descending) at columns xi[]
  auto threshold_group = [&](auto may_overflow, int p0, int part, const float (&xv)[SL], const int (&xi)[SL], float tau) {
    int npass = 0;
#pragma unroll
    for (int q = 0; q < SL; ++q)
      if (part < npart && xv[q] >= tau && xi[q] >= 0 && xi[q] < V) npass = q + 1;
    // a part whose LAST kept value still passes may have dropped columns that pass: all of its columns are re-scored
    // (so are parts whose survivors no longer fit the list)
    bool whole = npass == SL;
    const int mine = whole ? 0 : npass;
    int before = mine;                              // inclusive prefix sum over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, before, o);
      if (lane >= o) before += t;
    }
    const int total = __shfl_sync(0xffffffffu, before, 31);
    if (ncnd + total > SC_CAND) whole = whole || npass > 0;
    else {
#pragma unroll
      for (int q = 0; q < SL - 1; ++q)
        if (q < mine) s_cand[wid][ncnd + before - mine + q] = xi[q];
    }
    if (ncnd + total <= SC_CAND) ncnd += total;
    unsigned over = __ballot_sync(0xffffffffu, whole);
    while (over) {
      const int b = __ffs(over) - 1;
      over &= over - 1;
      if (!decltype(may_overflow)::value || (nover < SC_OVER && p0 + b < 65536)) {
        if (lane == 0) s_over[wid][nover] = (unsigned short)(p0 + b);
        ++nover;
      } else if constexpr (decltype(may_overflow)::value) {   // (list full: re-scored right here)
        const int v0 = part_first(p0 + b), v1 = min(V, v0 + part_cols);
        for (int v = v0; v < v1; v += 2) {
          float sa, sb;
          exact2(hs, v, min(v + 1, v1 - 1), sa, sb);
          topk_insert<KMAX>(bestv, besti, sa, v);
          if (v + 1 < v1) topk_insert<KMAX>(bestv, besti, sb, v + 1);
        }
      }
    }
  };
  if (valid) {
    // ---- phase A: this warp's row.  State and its norm; the bound; a_K; survivors and parts to re-score as a whole
    const int ncand = npart * SL;
    const float* cv = cand_val + (size_t)m * ncand;
    const bool fast = npart <= 32 * SC_PPL;
    float kv[SC_PPL][SL];                            // fast path: kept values of the parts lane, lane + 32, lane + 64
    if (fast) {
#pragma unroll
      for (int a = 0; a < SC_PPL; ++a) {
        const int part = lane + 32 * a;
#pragma unroll
        for (int q = 0; q < SL; ++q) kv[a][q] = part < npart ? cv[part * SL + q] : -FLT_MAX;
      }
    }
    float hh = 0.f;
    if (vec) {
      const float4* src = reinterpret_cast<const float4*>(h + (size_t)m * H);
      for (int i = lane; i < (H >> 2); i += 32) {
        const float4 v = src[i];
        reinterpret_cast<float4*>(hs)[i] = v;
        hh = fmaf(v.x, v.x, hh); hh = fmaf(v.y, v.y, hh); hh = fmaf(v.z, v.z, hh); hh = fmaf(v.w, v.w, hh);
      }
    } else {
      for (int i = lane; i < H; i += 32) { const float v = h[(size_t)m * H + i]; hs[i] = v; hh = fmaf(v, v, hh); }
    }
    hh = warp_sum(hh);
    __syncwarp();
    const float eps = SCREEN_C * sqrtf(hh) * (*wmax);
    // a lower bound a_K of the K-th largest approximation: the K-th largest KEPT one (a part may have dropped some of
    // the overall K largest, which only lowers it).  K rounds of {warp maximum, one instance of it removed}.
    float pv = FLT_MAX, a1 = 0.f;
    if (fast) {
      float w[SC_PPL * SL];
#pragma unroll
      for (int a = 0; a < SC_PPL; ++a)
#pragma unroll
        for (int q = 0; q < SL; ++q) w[a * SL + q] = kv[a][q];
      for (int k = 0; k < K; ++k) {
        float lm = w[0];
#pragma unroll
        for (int j = 1; j < SC_PPL * SL; ++j) lm = fmaxf(lm, w[j]);
        const float gm = warp_max(lm);
        const int owner = __ffs(__ballot_sync(0xffffffffu, lm == gm)) - 1;
        bool open = lane == owner;
#pragma unroll
        for (int j = 0; j < SC_PPL * SL; ++j) {
          const bool hit = open && w[j] == gm;
          w[j] = hit ? -FLT_MAX : w[j];
          open = open && !hit;
        }
        pv = gm;
        if (k == 0) a1 = gm;
      }
    } else {
      const int32_t* ci = cand_idx + (size_t)m * ncand;
      int pi = -1;
      for (int k = 0; k < K; ++k) {
        float b = -FLT_MAX;
        int bi = 0x7fffffff;
        for (int c = lane; c < ncand; c += 32) {
          const float x = cv[c];
          const int i = ci[c];
          const bool after_prev = (x < pv) || (x == pv && i > pi);
          if (after_prev && ((x > b) || (x == b && i < bi))) { b = x; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, b, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > b || (ov == b && oi < bi)) { b = ov; bi = oi; }
        }
        pv = b; pi = bi;
        if (k == 0) a1 = b;
      }
    }
    // the screening epilogue stores the column's position in the low 7 mantissa bits of a kept value: a relative
    // perturbation below 2^-16 of values that lie between tau and a_1, bounded here with a factor 2 to spare
    const float delta = 3.0517578125e-5f * (fmaxf(fabsf(a1), fabsf(pv)) + 2.f * eps);
    const float tau = pv - 2.f * (eps + delta);
    if (fast) {
#pragma unroll
      for (int a = 0; a < SC_PPL; ++a) {
        if (32 * a >= npart) break;                   // warp-uniform
        const int part = lane + 32 * a;
        const int base = part_first(part) & ~127;
        int xi[SL];
#pragma unroll
        for (int q = 0; q < SL; ++q) xi[q] = base + 127 - (int)(__float_as_uint(kv[a][q]) & 127u);
        threshold_group(std::false_type{}, 32 * a, part, kv[a], xi, tau);
      }
    } else {
      const int32_t* ci = cand_idx + (size_t)m * ncand;
      for (int p0 = 0; p0 < npart; p0 += 32) {       // a lane looks at one part's kept entries (descending)
        const int part = p0 + lane;
        float xv[SL];
        int xi[SL];
#pragma unroll
        for (int q = 0; q < SL; ++q) {
          xv[q] = part < npart ? cv[part * SL + q] : -FLT_MAX;
          xi[q] = part < npart ? ci[part * SL + q] : -1;
        }
        threshold_group(std::true_type{}, p0, part, xv, xi, tau);
      }
    }
    __syncwarp();
    for (int j = 0; j < ncnd; j += 2) {               // the survivors, two columns per round trip
      const int va = s_cand[wid][j], vb = s_cand[wid][min(j + 1, ncnd - 1)];
      float sa, sb;
      exact2(hs, va, vb, sa, sb);
      topk_insert<KMAX>(bestv, besti, sa, va);
      if (j + 1 < ncnd) topk_insert<KMAX>(bestv, besti, sb, vb);
    }
  }
  if (lane == 0) s_nover[wid] = nover;
  __syncthreads();
  // ---- phase B: whole parts, by the whole block (rare: a handful of rows per call): warp w scores 16 of the columns
  for (int wo = 0; wo < 8; ++wo) {
    const int n = s_nover[wo];                        // block-uniform
    for (int e = 0; e < n; ++e) {
      const int pt = s_over[wo][e];
      const int v0 = part_first(pt), v1 = min(V, v0 + part_cols);
      const int per = (part_cols + 7) / 8;
      for (int j = 0; j < per; j += 2) {
        const int va = v0 + wid * per + j, vb = va + 1;
        if (va < v1) {
          float sa, sb;
          exact2(s_h + (size_t)wo * H, va, min(vb, v1 - 1), sa, sb);
          if (lane == 0) { s_res[va - v0] = sa; if (vb < v1 && j + 1 < per) s_res[vb - v0] = sb; }
        }
      }
      __syncthreads();
      if (wid == wo)
        for (int v = v0; v < v1; ++v) topk_insert<KMAX>(bestv, besti, s_res[v - v0], v);
      __syncthreads();
    }
  }
  if (valid && lane == 0) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        if (val) val[(size_t)m * out_stride + k] = bestv[k];
        if (idx) idx[(size_t)m * out_stride + k] = besti[k];
      }
    }
    if (tok) tok[(size_t)m * tok_stride] = besti[0];
  }
}

// X (rows, E) = emb[tok[row * stride]] written directly as its (hi, lo) tf32 split (Embedding lookup, rnn.py:53,85)
__global__ void gather_emb_split_kernel(float* __restrict__ X_hi, float* __restrict__ X_lo, const float* __restrict__ emb,
                                        int E, const void* __restrict__ tok, int tok64, int stride) {
  const int64_t id = tok64 ? reinterpret_cast<const int64_t*>(tok)[(size_t)blockIdx.x * stride]
                           : (int64_t) reinterpret_cast<const int32_t*>(tok)[(size_t)blockIdx.x * stride];
  const float* src = emb + id * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    split_store(src[e], X_hi + (size_t)blockIdx.x * E + e, X_lo + (size_t)blockIdx.x * E + e);
}

// Per-call decoder state: double-buffered h/c per layer for `rows` rows, plus GEMM scratch.
struct Rig {
  const st_rnn_weights* w;
  int rows, G;
  bool tc;      // gemm_mode 1: all products on the tensor cores
  bool fused;   // tc and K <= FUSED_TOPK_MAX: vocabulary projection fused with the top-K, no logits buffer
  bool table;   // tc and enough row-steps to amortise it: input projections of fed-back words from the EP table
  bool screen;  // fused: vocabulary projection as bf16 screening + exact re-scoring (see screen_select_kernel)
  float *X, *Gx0, *GxL, *logits;
  float* h[2][MAXL];
  float* c[2][MAXL];
  float *Wih_hi[MAXL], *Wih_lo[MAXL], *Wv_hi, *Wv_lo, *act_hi, *act_lo;   // tc: tf32 splits
  // tc: W_hh splits, the hidden projection Gh, the (hi, lo) split of each layer's state (written by the gate kernel),
  // top-K candidates / soft-max partials of the fused vocabulary projection, the projected-embedding table
  float *Whh_hi[MAXL], *Whh_lo[MAXL], *Gh, *hs_hi[2][MAXL], *hs_lo[2][MAXL], *cand_val, *part_stats, *EP, *emb_hi, *emb_lo;
  int32_t* cand_idx;
  __nv_bfloat16 *Wv_b, *hb[2];   // screen: bf16 vocabulary weights, bf16 copy of the top layer's state
  float* wmax;                   // screen: max_v |w_v|_2
  int* barrier;
  int cur;
  cudaStream_t s;

  // `steps`: time steps of the call (table decision); K: widest top-K the loop asks for
  void carve(Bump& b, const st_rnn_weights* w_, int rows_, int K, int steps) {
    w = w_;
    rows = rows_;
    G = (w->kind == ST_LSTM) ? 4 : 3;
    tc = use_tc(w);
    fused = tc && K <= FUSED_TOPK_MAX;
    table = tc && (int64_t)rows * steps >= (int64_t)w->V && g_force_table >= 0;
    if (tc && g_force_table > 0) table = true;
    screen = fused && g_screen >= 0 && w->H % 8 == 0 && w->H <= 1536;
    X = b.take<float>((int64_t)rows * w->E);
    Gx0 = b.take<float>((int64_t)rows * G * w->H);
    GxL = b.take<float>((int64_t)rows * G * w->H);
    logits = fused ? nullptr : b.take<float>((int64_t)rows * w->V);
    for (int i = 0; i < 2; ++i)
      for (int l = 0; l < w->L; ++l) {
        h[i][l] = b.take<float>((int64_t)rows * w->H);
        c[i][l] = (w->kind == ST_LSTM) ? b.take<float>((int64_t)rows * w->H) : nullptr;
      }
    if (tc) {
      for (int l = 0; l < w->L; ++l) {
        const int64_t n = (int64_t)G * w->H * (l == 0 ? w->E : w->H);
        Wih_hi[l] = b.take<float>(n);
        Wih_lo[l] = b.take<float>(n);
        Whh_hi[l] = b.take<float>((int64_t)G * w->H * w->H);
        Whh_lo[l] = b.take<float>((int64_t)G * w->H * w->H);
        for (int i = 0; i < 2; ++i) {
          hs_hi[i][l] = b.take<float>((int64_t)rows * w->H);
          hs_lo[i][l] = b.take<float>((int64_t)rows * w->H);
        }
      }
      Wv_hi = b.take<float>((int64_t)w->V * w->H);
      Wv_lo = b.take<float>((int64_t)w->V * w->H);
      const int64_t wide = w->E > w->H ? w->E : w->H;
      act_hi = b.take<float>((int64_t)rows * wide);
      act_lo = b.take<float>((int64_t)rows * wide);
      Gh = b.take<float>((int64_t)rows * G * w->H);
      const int parts = st_topk_parts(w->V);
      cand_val = b.take<float>((int64_t)rows * parts);
      cand_idx = b.take<int32_t>((int64_t)rows * parts);
      part_stats = b.take<float>((int64_t)rows * (parts / 4));
      Wv_b = hb[0] = hb[1] = nullptr;
      wmax = nullptr;
      if (screen) {
        Wv_b = b.take<__nv_bfloat16>((int64_t)w->V * w->H);
        hb[0] = b.take<__nv_bfloat16>((int64_t)rows * w->H);
        hb[1] = b.take<__nv_bfloat16>((int64_t)rows * w->H);
        wmax = b.take<float>(4);
      }
      EP = emb_hi = emb_lo = nullptr;
      if (table) {
        EP = b.take<float>((int64_t)w->V * G * w->H);
        emb_hi = b.take<float>((int64_t)w->V * w->E);
        emb_lo = b.take<float>((int64_t)w->V * w->E);
      }
    }
    barrier = b.take<int>(64);
    cur = 0;
  }
  // tc: split the (constant) weights once per call; build the projected-embedding table
  int prepare() {
    if (!tc) return ST_OK;
    for (int l = 0; l < w->L; ++l) {
      const int in = l == 0 ? w->E : w->H;
      ST_TRY(st_split_tf32(w->Wih_host[l], G * w->H, in, in, Wih_hi[l], Wih_lo[l], in, s));
      ST_TRY(st_split_tf32(w->Whh_host[l], G * w->H, w->H, w->H, Whh_hi[l], Whh_lo[l], w->H, s));
    }
    ST_TRY(st_split_tf32(w->Wv, w->V, w->H, w->H, Wv_hi, Wv_lo, w->H, s));
    if (screen) {
      ST_TRY(st_cast_bf16(w->Wv, w->V, w->H, w->H, Wv_b, w->H, nullptr, 0, s));
      ST_TRY(st_row_norm_max(w->Wv, w->V, w->H, wmax, s));
    }
    if (table) {   // EP[v] = W_ih0 emb[v] + b_ih0: the same product, row for row, the per-step input GEMM would form
      ST_TRY(st_split_tf32(w->emb, w->V, w->E, w->E, emb_hi, emb_lo, w->E, s));
      ST_TRY(st_gemm_tf32x3(w->V, G * w->H, w->E, emb_hi, emb_lo, w->E, Wih_hi[0], Wih_lo[0], w->E, EP, G * w->H,
                            w->bih_host[0], 1.f, 0.f, s));
    }
    return ST_OK;
  }
  // ---- tensor-core loop
  int gate(int l, GxSrc src, bool first, int nxt) {
    const int H = w->H;
    const long long n = (long long)rows * (H / 4);          // a thread per 4 units (use_tc: H % 4 == 0)
    int sms = 0;
    ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
    const long long blocks = (n + 255) / 256;
    const unsigned grid = (unsigned)(blocks < 8LL * sms ? blocks : 8LL * sms);
    const float* gh = first ? nullptr : Gh;
    __nv_bfloat16* hbf = (screen && l == w->L - 1) ? hb[nxt] : nullptr;
    if (w->kind == ST_LSTM)
      decode_gate_kernel<4><<<grid, 256, 0, s>>>(rows, H, src, gh, w->bhh_host[l], first ? nullptr : h[cur][l],
                                                 first ? nullptr : c[cur][l], h[nxt][l], c[nxt][l], hs_hi[nxt][l], hs_lo[nxt][l], hbf);
    else
      decode_gate_kernel<3><<<grid, 256, 0, s>>>(rows, H, src, gh, w->bhh_host[l], first ? nullptr : h[cur][l], nullptr,
                                                 h[nxt][l], nullptr, hs_hi[nxt][l], hs_lo[nxt][l], hbf);
    ST_LAUNCH_TRY("decode_gate_kernel");
    return ST_OK;
  }
  // One time step through all layers on the tensor cores.  Input of layer 0: the projection in Gx0 (tok == NULL: the
  // image feature, rnn.py:47-49) or the embedding of each row's fed-back word (rnn.py:53,85).
  int step_tc(const void* tok, int tok64, int stride, bool first) {
    const int H = w->H, nxt = cur ^ 1;
    GxSrc src{Gx0, nullptr, 0, 0};
    if (tok && table) {
      src = GxSrc{EP, tok, tok64, stride};
    } else if (tok) {
      gather_emb_split_kernel<<<rows, 128, 0, s>>>(act_hi, act_lo, w->emb, w->E, tok, tok64, stride);
      ST_LAUNCH_TRY("gather_emb_split_kernel");
      ST_TRY(st_gemm_tf32x3(rows, G * H, w->E, act_hi, act_lo, w->E, Wih_hi[0], Wih_lo[0], w->E, Gx0, G * H, w->bih_host[0],
                            1.f, 0.f, s));
    }
    for (int l = 0; l < w->L; ++l) {
      if (l > 0) {
        ST_TRY(st_gemm_tf32x3(rows, G * H, H, hs_hi[nxt][l - 1], hs_lo[nxt][l - 1], H, Wih_hi[l], Wih_lo[l], H, GxL, G * H,
                              w->bih_host[l], 1.f, 0.f, s));
        src = GxSrc{GxL, nullptr, 0, 0};
      }
      if (!first)
        ST_TRY(st_gemm_tf32x3(rows, G * H, H, hs_hi[cur][l], hs_lo[cur][l], H, Whh_hi[l], Whh_lo[l], H, Gh, G * H,
                              w->bhh_host[l], 1.f, 0.f, s));
      ST_TRY(gate(l, src, first, nxt));
    }
    cur = nxt;
    return ST_OK;
  }
  // top-K of the vocabulary logits of the current top state, logits never written (rnn.py:50-51, 88-91); optionally the
  // rows' soft-max normaliser (beam_search.py:85-88)
  int vocab_topk(int K, float* val, int32_t* idx, int out_stride, int64_t* tok, int tok_stride, float* row_max = nullptr,
                 float* row_sum = nullptr) {
    if (screen && !row_max) {   // (the tree beam's soft-max normaliser needs every logit at fp32 accuracy: 3xTF32 path)
      return st_vocab_topk_screen(rows, w->V, w->H, h[cur][w->L - 1], hb[cur], w->Wv, Wv_b, w->bv, wmax, K, cand_val, cand_idx,
                                  val, idx, out_stride, tok, tok_stride, s);
    }
    return st_gemm_tf32x3_topk(rows, w->V, w->H, hs_hi[cur][w->L - 1], hs_lo[cur][w->L - 1], w->H, Wv_hi, Wv_lo, w->H, w->bv, K,
                               cand_val, cand_idx, val, idx, out_stride, tok, tok_stride, part_stats, row_max, row_sum, s);
  }
  // ---- generic pieces (fp32 mode; tc mode for the feature projection and for K > FUSED_TOPK_MAX)
  // out (rows, N) = in (rows, K) . W^T + bias
  int linear(const float* in, int K, const float* W, const float* W_hi, const float* W_lo, const float* bias, int N,
             float* out) {
    if (!tc) return st_sgemm(0, 1, rows, N, K, 1.f, in, K, W, K, 0.f, out, N, bias, s);
    ST_TRY(st_split_tf32(in, rows, K, K, act_hi, act_lo, K, s));
    return st_gemm_tf32x3(rows, N, K, act_hi, act_lo, K, W_hi, W_lo, K, out, N, bias, 1.f, 0.f, s);
  }
  // Gx0 = Xin (rows, E) . Wih_0^T + bih_0
  int input_proj(const float* Xin) {
    return linear(Xin, w->E, w->Wih_host[0], Wih_hi[0], Wih_lo[0], w->bih_host[0], G * w->H, Gx0);
  }
  // One time step through all layers (fp32 mode); reads state `cur` (zeros when first), writes and flips.
  int step(bool first) {
    const int H = w->H, nxt = cur ^ 1;
    for (int l = 0; l < w->L; ++l) {
      const float* gx = Gx0;
      if (l > 0) {
        ST_TRY(linear(h[nxt][l - 1], H, w->Wih_host[l], Wih_hi[l], Wih_lo[l], w->bih_host[l], G * H, GxL));
        gx = GxL;
      }
      ST_TRY(st_rnn_seq_fwd(w->kind, H, 1, &rows, 0, 1, gx, w->Whh_host[l], w->bhh_host[l],
                            first ? nullptr : h[cur][l], first ? nullptr : c[cur][l], h[nxt][l],
                            c[nxt][l], nullptr, nullptr, barrier, s));
    }
    cur = nxt;
    return ST_OK;
  }
  // One time step in the call's arithmetic mode: the fed-back words tok (NULL: Gx0 already holds the projection)
  int advance(const void* tok, int tok64, int stride, bool first) {
    if (tc) return step_tc(tok, tok64, stride, first);
    if (tok) {
      if (tok64) gather_emb_kernel<int64_t><<<rows, 128, 0, s>>>(X, w->emb, w->E, (const int64_t*)tok, stride);
      else gather_emb_kernel<int32_t><<<rows, 128, 0, s>>>(X, w->emb, w->E, (const int32_t*)tok, stride);
      ST_LAUNCH_TRY("gather_emb_kernel");
      ST_TRY(input_proj(X));
    }
    return step(first);
  }
  // top-K of the vocabulary logits of the current top state -> val / idx (rows, out_stride) and / or tok (int64)
  int top_words(int K, float* val, int32_t* idx, int out_stride, int64_t* tok, int tok_stride) {
    if (fused) return vocab_topk(K, val, idx, out_stride, tok, tok_stride);
    ST_TRY(vocab_logits());
    if (tok) ST_TRY(st_argmax_rows(logits, w->V, rows, w->V, tok, tok_stride, s));
    if (val || idx) ST_TRY(st_topk_rows(logits, w->V, rows, w->V, K, val, idx, out_stride, s));
    return ST_OK;
  }
  float* top() { return h[cur][w->L - 1]; }
  int vocab_logits() { return linear(top(), w->H, w->Wv, Wv_hi, Wv_lo, w->bv, w->V, logits); }
};

int64_t rig_bytes(const st_rnn_weights* w, int64_t rows, int K, int steps) {
  Bump dry{nullptr, 0, 0};   // the same carve on a null base: only the offsets are computed
  Rig r;
  r.carve(dry, w, (int)rows, K, steps);
  return dry.used + 256;
}


// ----------------------------------------------------------------------------- chain beam
__global__ void chain_init_kernel(int n, int K, int max_len, const int32_t* __restrict__ words,
                                  const float* __restrict__ vals, int32_t* __restrict__ sent,
                                  float* trace_scores, int32_t* trace_words) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * K) return;
  const int img = i / K, k = i - img * K;
  sent[((size_t)img * K + k) * max_len] = words[i];  // rnn.py:66-74
  if (trace_scores) trace_scores[(size_t)img * K + k] = vals[i];
  if (trace_words) trace_words[(size_t)img * K + k] = words[i];
}

// rnn.py:102-103: two independent descending sorts over the K*K candidates of one image,
//   (score, sentence-so-far + word)  -> surviving sentences      (ties: lexicographic, larger first)
//   (score, word)                    -> surviving words          (ties: larger word id first)
// both stable w.r.t. generation order (k outer, j inner).  One thread per image; K passes of a
// linear scan each (K*K <= 1024 candidates).
__global__ void chain_select_kernel(int n, int K, int max_len, int pos, const float* __restrict__ cand_val,
                                    const int32_t* __restrict__ cand_idx,
                                    const int32_t* __restrict__ sent_old, int32_t* __restrict__ sent_new,
                                    int32_t* __restrict__ words, float* trace_scores,
                                    int32_t* trace_words) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n) return;
  const int KK = K * K;
  const float* cv = cand_val + (size_t)img * KK;
  const int32_t* ci = cand_idx + (size_t)img * KK;
  const int32_t* so = sent_old + (size_t)img * K * max_len;
  int32_t* sn = sent_new + (size_t)img * K * max_len;
  uint32_t taken[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) taken[i] = 0u;

  for (int s = 0; s < K; ++s) {  // sort by (score, sentence)
    int best = -1;
    for (int c = 0; c < KK; ++c) {
      if (taken[c >> 5] & (1u << (c & 31))) continue;
      if (best < 0) { best = c; continue; }
      bool better;
      if (cv[c] != cv[best]) {
        better = cv[c] > cv[best];
      } else {
        const int32_t* sa = so + (size_t)(c / K) * max_len;
        const int32_t* sb = so + (size_t)(best / K) * max_len;
        int q = 0;
        while (q < pos && sa[q] == sb[q]) ++q;
        if (q < pos) better = sa[q] > sb[q];
        else better = ci[c] > ci[best];  // equal -> keep the earlier candidate (stable)
      }
      if (better) best = c;
    }
    taken[best >> 5] |= 1u << (best & 31);
    const int32_t* sp = so + (size_t)(best / K) * max_len;
    for (int q = 0; q < pos; ++q) sn[(size_t)s * max_len + q] = sp[q];
    sn[(size_t)s * max_len + pos] = ci[best];
    if (trace_scores) trace_scores[((size_t)pos * n + img) * K + s] = cv[best];
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) taken[i] = 0u;
  for (int s = 0; s < K; ++s) {  // sort by (score, word)
    int best = -1;
    for (int c = 0; c < KK; ++c) {
      if (taken[c >> 5] & (1u << (c & 31))) continue;
      if (best < 0) { best = c; continue; }
      bool better = (cv[c] != cv[best]) ? (cv[c] > cv[best]) : (ci[c] > ci[best]);
      if (better) best = c;
    }
    taken[best >> 5] |= 1u << (best & 31);
    words[(size_t)img * K + s] = ci[best];
    if (trace_words) trace_words[((size_t)pos * n + img) * K + s] = ci[best];
  }
}

__global__ void chain_finish_kernel(int n, int K, int max_len, const int32_t* __restrict__ sent,
                                    int64_t* __restrict__ tokens) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * max_len) return;
  const int img = i / max_len, q = i - img * max_len;
  tokens[i] = sent[(size_t)img * K * max_len + q];  // old_beam_sentence[0], rnn.py:106
}

// ----------------------------------------------------------------------------- tree beam
struct TreeBufs {
  int n, K, num_hyp, SL;  // SL = max_length + 1 tokens per sequence
  int32_t *node_val, *node_slot, *live_cnt;  // (n,K), (n,K), (n)
  float* node_cost;                          // (n,K)
  int32_t *seq[2];                           // (n,K,SL)
  int32_t *fr_cnt, *fr_pos;                  // (n), (n,K): live-list positions still in the fringe
  int32_t *parent_slot;                      // (n,K) state slot each new node inherits
  int32_t *hyp_tok, *hyp_len, *hyp_cnt;      // (n,num_hyp,SL), (n,num_hyp), (n)
  float* hyp_cost;                           // (n,num_hyp)
  float *row_max, *row_sum;                  // (n*K)
  float* tk_val;                             // (n*K, K)
  int32_t* tk_idx;                           // (n*K, K)
};

__global__ void tree_init_kernel(TreeBufs b, int start_id) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= b.n) return;
  for (int k = 0; k < b.K; ++k) {
    b.node_val[img * b.K + k] = (k == 0) ? start_id : 0;
    b.node_slot[img * b.K + k] = k;
    b.node_cost[img * b.K + k] = 0.f;
  }
  b.live_cnt[img] = 1;  // the root, beam_search.py:66
  b.seq[0][(size_t)img * b.K * b.SL] = start_id;
  b.hyp_cnt[img] = 0;
  for (int j = 0; j < b.num_hyp; ++j) {
    b.hyp_len[img * b.num_hyp + j] = 0;
    b.hyp_cost[img * b.num_hyp + j] = 0.f;
    for (int q = 0; q < b.SL; ++q) b.hyp_tok[((size_t)img * b.num_hyp + j) * b.SL + q] = -1;
  }
}

// beam_search.py:70-78: finished nodes (value == end_id) leave the live list for `hypotheses`
// (kept here as the num_hyp cheapest, inserted stably so that the final stable sort at :96 is
// reproduced); the rest form the fringe.  len = tokens in each live sequence.
__global__ void tree_retire_kernel(TreeBufs b, int cur_seq, int len, int end_id) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= b.n) return;
  const int K = b.K, NH = b.num_hyp, SL = b.SL;
  int cnt = 0;
  for (int i = 0; i < b.live_cnt[img]; ++i) {
    if (b.node_val[img * K + i] != end_id) {
      b.fr_pos[img * K + cnt++] = i;
      continue;
    }
    const float cost = b.node_cost[img * K + i];
    int hc = b.hyp_cnt[img];
    int at = hc;  // stable: after every kept hypothesis with cost <= this one
    while (at > 0 && b.hyp_cost[img * NH + at - 1] > cost) --at;
    if (at >= NH) continue;
    const int last = min(hc, NH - 1);
    for (int j = last; j > at; --j) {
      b.hyp_cost[img * NH + j] = b.hyp_cost[img * NH + j - 1];
      b.hyp_len[img * NH + j] = b.hyp_len[img * NH + j - 1];
      for (int q = 0; q < SL; ++q)
        b.hyp_tok[((size_t)img * NH + j) * SL + q] = b.hyp_tok[((size_t)img * NH + j - 1) * SL + q];
    }
    b.hyp_cost[img * NH + at] = cost;
    b.hyp_len[img * NH + at] = len;
    const int32_t* src = b.seq[cur_seq] + ((size_t)img * K + i) * SL;
    for (int q = 0; q < SL; ++q) b.hyp_tok[((size_t)img * NH + at) * SL + q] = (q < len) ? src[q] : -1;
    b.hyp_cnt[img] = min(hc + 1, NH);
  }
  b.fr_cnt[img] = cnt;
}

__global__ void tree_tokens_kernel(TreeBufs b, int32_t* __restrict__ tok_rows) {
  // token fed to each state slot this round (slots outside the fringe get a harmless 0)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n * b.K) return;
  tok_rows[i] = 0;
  const int img = i / b.K, slot = i - img * b.K;
  for (int f = 0; f < b.fr_cnt[img]; ++f) {
    const int pos = b.fr_pos[img * b.K + f];
    if (b.node_slot[img * b.K + pos] == slot) tok_rows[i] = b.node_val[img * b.K + pos];
  }
}

__global__ void row_softmax_stats_kernel(const float* __restrict__ X, int ld, int cols,
                                         float* __restrict__ rmax, float* __restrict__ rsum) {
  __shared__ float red[8];
  const float* row = X + (size_t)blockIdx.x * ld;
  float m = -FLT_MAX;
  for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, row[c]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) s += expf(row[c] - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    rmax[blockIdx.x] = m;
    rsum[blockIdx.x] = t;
  }
}

// beam_search.py:84-94: per fringe node the beam_width most probable tokens in ASCENDING order of
// probability (np.argsort(...)[:, -K:]), cost = -log p with p = softmax (float32, as numpy >= 2
// keeps python-float + float32 in float32), cum_cost additive, then the K cheapest by a stable
// sort in generation order.
__global__ void tree_expand_kernel(TreeBufs b, int cur_seq, int len) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= b.n) return;
  const int K = b.K, SL = b.SL;
  const int fc = b.fr_cnt[img];
  if (fc == 0) {
    b.live_cnt[img] = 0;
    return;
  }
  const int ncand = fc * K;
  float new_cost[32];
  int32_t new_val[32], new_par[32];
  uint32_t taken[32];
  for (int i = 0; i < 32; ++i) taken[i] = 0u;
  const int keep = min(K, ncand);
  for (int s = 0; s < keep; ++s) {
    int best = -1;
    float best_cost = 0.f;
    for (int c = 0; c < ncand; ++c) {
      if (taken[c >> 5] & (1u << (c & 31))) continue;
      const int f = c / K, jj = c - f * K;
      const int pos = b.fr_pos[img * K + f];
      const int row = img * K + b.node_slot[img * K + pos];
      const float logit = b.tk_val[(size_t)row * K + (K - 1 - jj)];
      const float p = expf(logit - b.row_max[row]) / b.row_sum[row];
      const float cost = b.node_cost[img * K + pos] + (-logf(p));
      if (best < 0 || cost < best_cost) {
        best = c;
        best_cost = cost;
      }
    }
    taken[best >> 5] |= 1u << (best & 31);
    const int f = best / K, jj = best - f * K;
    const int pos = b.fr_pos[img * K + f];
    const int slot = b.node_slot[img * K + pos];
    new_cost[s] = best_cost;
    new_val[s] = b.tk_idx[(size_t)(img * K + slot) * K + (K - 1 - jj)];
    new_par[s] = pos;
  }
  int32_t par_slot[32];
  for (int s = 0; s < keep; ++s) par_slot[s] = b.node_slot[img * K + new_par[s]];
  const int32_t* so = b.seq[cur_seq] + (size_t)img * K * SL;
  int32_t* sn = b.seq[cur_seq ^ 1] + (size_t)img * K * SL;
  for (int s = 0; s < keep; ++s) {
    for (int q = 0; q < len; ++q) sn[(size_t)s * SL + q] = so[(size_t)new_par[s] * SL + q];
    sn[(size_t)s * SL + len] = new_val[s];
  }
  for (int s = 0; s < keep; ++s) {
    b.node_val[img * K + s] = new_val[s];
    b.node_cost[img * K + s] = new_cost[s];
    b.node_slot[img * K + s] = s;
    b.parent_slot[img * K + s] = par_slot[s];
  }
  for (int s = keep; s < K; ++s) b.parent_slot[img * K + s] = s;
  b.live_cnt[img] = keep;
}

__global__ void gather_state_kernel(float* __restrict__ dst, const float* __restrict__ src, int H, int K,
                                    const int32_t* __restrict__ parent_slot, float* __restrict__ dst_hi,
                                    float* __restrict__ dst_lo) {
  const int row = blockIdx.x, img = row / K;
  const float* s = src + (size_t)(img * K + parent_slot[row]) * H;
  float* d = dst + (size_t)row * H;
  for (int e = threadIdx.x; e < H; e += blockDim.x) {
    d[e] = s[e];
    if (dst_hi) split_store(s[e], dst_hi + (size_t)row * H + e, dst_lo + (size_t)row * H + e);   // tc: next W_hh operand
  }
}

__global__ void spread_root_kernel(float* __restrict__ dst, const float* __restrict__ src, int H, int K) {
  // dst (n*K, H): slot 0 of each image <- src (n, H); other slots zero
  const int row = blockIdx.x, img = row / K, slot = row - img * K;
  float* d = dst + (size_t)row * H;
  for (int e = threadIdx.x; e < H; e += blockDim.x) d[e] = slot == 0 ? src[(size_t)img * H + e] : 0.f;
}

__global__ void tree_finish_kernel(TreeBufs b, int32_t* __restrict__ out_tok, int32_t* __restrict__ out_len,
                                   float* __restrict__ out_cost) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n * b.num_hyp) return;
  const int img = i / b.num_hyp, j = i - img * b.num_hyp;
  const bool ok = j < b.hyp_cnt[img];
  out_len[i] = ok ? b.hyp_len[i] : 0;
  out_cost[i] = ok ? b.hyp_cost[i] : 0.f;
  for (int q = 0; q < b.SL; ++q) out_tok[(size_t)i * b.SL + q] = ok ? b.hyp_tok[(size_t)i * b.SL + q] : -1;
}

}  // namespace
}  // namespace st

extern "C" {

int st_debug_decode_table(int mode) {
  st::g_force_table = mode;
  return ST_OK;
}

int st_row_norm_max(const float* W, int rows, int cols, float* out, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(W && out, ST_ERR_NULL, "st_row_norm_max: NULL pointer");
  ST_REQUIRE(rows >= 1 && cols >= 1, ST_ERR_BAD_SHAPE, "st_row_norm_max: rows=%d cols=%d", rows, cols);
  cudaStream_t s = as_stream(stream);
  ST_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float), s));
  row_norm_max_kernel<<<(rows + 7) / 8, 256, 0, s>>>(W, rows, cols, out);
  ST_LAUNCH_TRY("row_norm_max_kernel");
  return ST_OK;
}

int st_vocab_topk_screen(int M, int V, int H, const float* h, const void* h_bf16, const float* Wv, const void* Wv_bf16,
                         const float* bv, const float* wmax, int K, float* cand_val, int32_t* cand_idx, float* val,
                         int32_t* idx, int out_stride, int64_t* tok, int tok_stride, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(h && h_bf16 && Wv && Wv_bf16 && bv && wmax && cand_val && cand_idx && (val || idx || tok), ST_ERR_NULL,
             "st_vocab_topk_screen: NULL pointer");
  ST_REQUIRE(M >= 1 && V >= 1 && H >= 8 && H % 8 == 0 && H <= 1536 && K >= 1 && K <= 8 && K <= V &&
                 (!(val || idx) || out_stride >= K) && (!tok || tok_stride >= 1),
             ST_ERR_BAD_SHAPE, "st_vocab_topk_screen: M=%d V=%d H=%d K=%d", M, V, H, K);
  int npart = 0;
  ST_TRY(st_gemm_bf16_screen(M, V, H, h_bf16, H, Wv_bf16, H, bv, cand_val, cand_idx, &npart, stream));
  const int part_cols = V > 128 ? 128 : 64;
  if (K <= 4)
    screen_select_kernel<4><<<(M + 7) / 8, 256, (size_t)8 * H * sizeof(float), as_stream(stream)>>>(
        M, H, V, npart, part_cols, K, cand_val, cand_idx, h, Wv, bv, wmax, val, idx, out_stride, tok, tok_stride);
  else
    screen_select_kernel<8><<<(M + 7) / 8, 256, (size_t)8 * H * sizeof(float), as_stream(stream)>>>(
        M, H, V, npart, part_cols, K, cand_val, cand_idx, h, Wv, bv, wmax, val, idx, out_stride, tok, tok_stride);
  ST_LAUNCH_TRY("screen_select_kernel");
  return ST_OK;
}

int st_debug_decode_screen(int mode) {
  st::g_screen = mode;
  return ST_OK;
}

int64_t st_decode_workspace_bytes(const st_rnn_weights* w, int n_img, int K, int max_len) {
  using namespace st;
  if (check_weights(w) != ST_OK || n_img < 1 || max_len < 1) return -1;
  const int64_t k = K < 1 ? 1 : K;
  const int64_t rows = (int64_t)n_img * k;  // tree beam keeps K states per image
  const int64_t SL = max_len + 1;
  int64_t extra = rows * SL * 8          // sentences / sequences, double buffered
                  + rows * k * 8         // candidate / top-K values and ids
                  + rows * 40            // per-node scalars
                  + (int64_t)n_img * 32 * SL * 4 + (int64_t)n_img * 32 * 8 + (int64_t)n_img * 16  // hypotheses
                  + (int64_t)n_img * w->H * 16   // initial-state scratch of the tree beam
                  + 64 * 256;
  return rig_bytes(w, rows, (int)k, max_len) + extra;
}

int st_decode_greedy(const st_rnn_weights* w, const float* feature, int n_img, int max_len,
                     int64_t* tokens, void* workspace, int64_t workspace_bytes, st_stream_t stream) {
  using namespace st;
  ST_TRY(check_weights(w));
  ST_REQUIRE(feature && tokens && workspace, ST_ERR_NULL, "st_decode_greedy: NULL pointer");
  ST_REQUIRE(n_img >= 1 && max_len >= 1, ST_ERR_BAD_SHAPE, "st_decode_greedy: n_img=%d max_len=%d", n_img, max_len);
  Bump b{(char*)workspace, workspace_bytes, 0};
  Rig rig;
  rig.s = as_stream(stream);
  rig.carve(b, w, n_img, 1, max_len);
  ST_REQUIRE(b.used <= b.cap, ST_ERR_WORKSPACE, "st_decode_greedy: workspace %lld < %lld",
             (long long)b.cap, (long long)b.used);
  ST_TRY(rig.prepare());
  ST_TRY(rig.input_proj(feature));                                           // rnn.py:41,49
  for (int step = 0; step < max_len; ++step) {
    // rnn.py:49 (step 0: the image feature) / rnn.py:53 (the embedding of the previous arg-max)
    ST_TRY(rig.advance(step ? tokens + step - 1 : nullptr, 1, max_len, step == 0));
    ST_TRY(rig.top_words(1, nullptr, nullptr, 0, tokens + step, max_len));    // rnn.py:50-51
  }
  return ST_OK;
}

int st_decode_beam_chain(const st_rnn_weights* w, const float* feature, int n_img, int K, int max_len,
                         int64_t* tokens, float* trace_scores, int32_t* trace_words, void* workspace,
                         int64_t workspace_bytes, st_stream_t stream) {
  using namespace st;
  ST_TRY(check_weights(w));
  ST_REQUIRE(w->kind == ST_GRU, ST_ERR_UNSUPPORTED,
             "st_decode_beam_chain: the reference beam search exists for the GRU decoder only (rnn.py:60)");
  ST_REQUIRE(feature && tokens && workspace, ST_ERR_NULL, "st_decode_beam_chain: NULL pointer");
  ST_REQUIRE(n_img >= 1 && max_len >= 1 && K >= 1 && K <= 32 && K <= w->V, ST_ERR_BAD_SHAPE,
             "st_decode_beam_chain: n_img=%d K=%d max_len=%d", n_img, K, max_len);
  Bump b{(char*)workspace, workspace_bytes, 0};
  Rig rig;
  rig.s = as_stream(stream);
  rig.carve(b, w, n_img, K, max_len);
  int32_t* words = b.take<int32_t>((int64_t)n_img * K);
  float* vals0 = b.take<float>((int64_t)n_img * K);
  int32_t* sent[2] = {b.take<int32_t>((int64_t)n_img * K * max_len),
                      b.take<int32_t>((int64_t)n_img * K * max_len)};
  float* cand_val = b.take<float>((int64_t)n_img * K * K);
  int32_t* cand_idx = b.take<int32_t>((int64_t)n_img * K * K);
  ST_REQUIRE(b.used <= b.cap, ST_ERR_WORKSPACE, "st_decode_beam_chain: workspace %lld < %lld",
             (long long)b.cap, (long long)b.used);
  ST_TRY(rig.prepare());
  cudaStream_t s = rig.s;
  const int tpb = 128;

  ST_TRY(rig.input_proj(feature));
  ST_TRY(rig.advance(nullptr, 0, 0, true));                                  // rnn.py:61
  ST_TRY(rig.top_words(K, vals0, words, K, nullptr, 0));                     // rnn.py:62-63
  chain_init_kernel<<<(n_img * K + tpb - 1) / tpb, tpb, 0, s>>>(n_img, K, max_len, words, vals0, sent[0],
                                                               trace_scores, trace_words);
  ST_LAUNCH_TRY("chain_init_kernel");
  int cur = 0;
  for (int pos = 1; pos < max_len; ++pos) {                                  // rnn.py:77-79
    for (int k = 0; k < K; ++k) {                                            // rnn.py:83
      ST_TRY(rig.advance(words + k, 0, K, false));   // rnn.py:85-87: embedding of beam k's word, the ONE chained state
      ST_TRY(rig.top_words(K, cand_val + k * K, cand_idx + k * K, K * K, nullptr, 0));   // rnn.py:88-91
    }
    chain_select_kernel<<<(n_img + 63) / 64, 64, 0, s>>>(n_img, K, max_len, pos, cand_val, cand_idx,
                                                         sent[cur], sent[cur ^ 1], words, trace_scores,
                                                         trace_words);
    ST_LAUNCH_TRY("chain_select_kernel");
    cur ^= 1;
  }
  chain_finish_kernel<<<(n_img * max_len + tpb - 1) / tpb, tpb, 0, s>>>(n_img, K, max_len, sent[cur], tokens);
  ST_LAUNCH_TRY("chain_finish_kernel");
  return ST_OK;
}

int st_decode_beam_tree(const st_rnn_weights* w, const float* feature, int n_img, int beam_width,
                        int num_hyp, int max_length, int start_id, int end_id, int32_t* out_tokens,
                        int32_t* out_len, float* out_cost, void* workspace, int64_t workspace_bytes,
                        st_stream_t stream) {
  using namespace st;
  ST_TRY(check_weights(w));
  ST_REQUIRE(w->kind == ST_GRU && w->L == 1, ST_ERR_UNSUPPORTED,
             "st_decode_beam_tree: single-layer GRU only (beam_search.py:23 keeps one flattened state)");
  ST_REQUIRE(feature && out_tokens && out_len && out_cost && workspace, ST_ERR_NULL,
             "st_decode_beam_tree: NULL pointer");
  const int K = beam_width;
  ST_REQUIRE(n_img >= 1 && K >= 1 && K <= 32 && K <= w->V && num_hyp >= 1 && num_hyp <= 32 &&
                 max_length >= 1 && start_id >= 0 && start_id < w->V,
             ST_ERR_BAD_SHAPE, "st_decode_beam_tree: n_img=%d beam_width=%d num_hyp=%d max_length=%d",
             n_img, K, num_hyp, max_length);
  Bump b{(char*)workspace, workspace_bytes, 0};
  Rig rig;
  rig.s = as_stream(stream);
  const int rows = n_img * K;
  rig.carve(b, w, rows, K, max_length);
  TreeBufs tb;
  tb.n = n_img; tb.K = K; tb.num_hyp = num_hyp; tb.SL = max_length + 1;
  tb.node_val = b.take<int32_t>(rows);
  tb.node_slot = b.take<int32_t>(rows);
  tb.live_cnt = b.take<int32_t>(n_img);
  tb.node_cost = b.take<float>(rows);
  tb.seq[0] = b.take<int32_t>((int64_t)rows * tb.SL);
  tb.seq[1] = b.take<int32_t>((int64_t)rows * tb.SL);
  tb.fr_cnt = b.take<int32_t>(n_img);
  tb.fr_pos = b.take<int32_t>(rows);
  tb.parent_slot = b.take<int32_t>(rows);
  tb.hyp_tok = b.take<int32_t>((int64_t)n_img * num_hyp * tb.SL);
  tb.hyp_len = b.take<int32_t>((int64_t)n_img * num_hyp);
  tb.hyp_cnt = b.take<int32_t>(n_img);
  tb.hyp_cost = b.take<float>((int64_t)n_img * num_hyp);
  tb.row_max = b.take<float>(rows);
  tb.row_sum = b.take<float>(rows);
  tb.tk_val = b.take<float>((int64_t)rows * K);
  tb.tk_idx = b.take<int32_t>((int64_t)rows * K);
  int32_t* tok_rows = b.take<int32_t>(rows);
  float* h_init = b.take<float>((int64_t)n_img * w->H);
  float* gx_init = b.take<float>((int64_t)n_img * 3 * w->H);
  ST_REQUIRE(b.used <= b.cap, ST_ERR_WORKSPACE, "st_decode_beam_tree: workspace %lld < %lld",
             (long long)b.cap, (long long)b.used);
  ST_TRY(rig.prepare());
  cudaStream_t s = rig.s;
  const int tpb = 64, gi = (n_img + tpb - 1) / tpb;
  const int H = w->H;

  // initial state = GRU state after consuming the image feature from zeros (rnn.py:47-49, step 0)
  ST_TRY(st_sgemm(0, 1, n_img, 3 * H, w->E, 1.f, feature, w->E, w->Wih_host[0], w->E, 0.f, gx_init,
                  3 * H, w->bih_host[0], stream));
  ST_TRY(st_rnn_seq_fwd(ST_GRU, H, 1, &n_img, 0, 1, gx_init, w->Whh_host[0], w->bhh_host[0], nullptr,
                        nullptr, h_init, nullptr, nullptr, nullptr, rig.barrier, stream));
  spread_root_kernel<<<rows, 128, 0, s>>>(rig.h[rig.cur][0], h_init, H, K);
  ST_LAUNCH_TRY("spread_root_kernel");
  if (rig.tc) ST_TRY(st_split_tf32(rig.h[rig.cur][0], rows, H, H, rig.hs_hi[rig.cur][0], rig.hs_lo[rig.cur][0], H, s));
  tree_init_kernel<<<gi, tpb, 0, s>>>(tb, start_id);
  ST_LAUNCH_TRY("tree_init_kernel");

  int cur_seq = 0;
  for (int round = 0; round < max_length; ++round) {                          // beam_search.py:68
    const int len = round + 1;
    tree_retire_kernel<<<gi, tpb, 0, s>>>(tb, cur_seq, len, end_id);          // :70-78
    ST_LAUNCH_TRY("tree_retire_kernel");
    tree_tokens_kernel<<<(rows + 127) / 128, 128, 0, s>>>(tb, tok_rows);
    ST_LAUNCH_TRY("tree_tokens_kernel");
    ST_TRY(rig.advance(tok_rows, 0, 1, false));                               // generate_function, :83
    if (rig.fused) {                                                          // :84 top-K + soft-max normaliser, no logits
      ST_TRY(rig.vocab_topk(K, tb.tk_val, tb.tk_idx, K, nullptr, 0, tb.row_max, tb.row_sum));
    } else {
      ST_TRY(rig.vocab_logits());
      row_softmax_stats_kernel<<<rows, 256, 0, s>>>(rig.logits, w->V, w->V, tb.row_max, tb.row_sum);
      ST_LAUNCH_TRY("row_softmax_stats_kernel");
      ST_TRY(st_topk_rows(rig.logits, w->V, rows, w->V, K, tb.tk_val, tb.tk_idx, K, stream));
    }
    tree_expand_kernel<<<gi, tpb, 0, s>>>(tb, cur_seq, len);                  // :86-94
    ST_LAUNCH_TRY("tree_expand_kernel");
    // new node s inherits the post-step state of its parent: gather h[cur] -> h[cur^1], flip
    gather_state_kernel<<<rows, 128, 0, s>>>(rig.h[rig.cur ^ 1][0], rig.h[rig.cur][0], H, K, tb.parent_slot,
                                             rig.tc ? rig.hs_hi[rig.cur ^ 1][0] : nullptr,
                                             rig.tc ? rig.hs_lo[rig.cur ^ 1][0] : nullptr);
    ST_LAUNCH_TRY("gather_state_kernel");
    rig.cur ^= 1;
    cur_seq ^= 1;
  }
  tree_finish_kernel<<<(n_img * num_hyp + 127) / 128, 128, 0, s>>>(tb, out_tokens, out_len, out_cost);
  ST_LAUNCH_TRY("tree_finish_kernel");                                        // :96-97
  return ST_OK;
}

}  // extern "C"
