// Packing / gather / scatter kernels around the recurrence (all HBM-bound row copies).
//   rnn.py:29-31    Embedding + cat(feature) + pack_padded_sequence   -> st_pack_inputs
//   rnn_attn.py:70  caption_embedding[:b_t, t]                         -> st_pack_inputs(with_feature=0)
//   main.py:145     pack_padded_sequence(caption)[0]                   -> st_pack_targets
//   autograd of nn.Embedding / torch.cat                               -> st_pack_inputs_bwd
#include <cuda_bf16.h>

#include "common.cuh"

namespace st {
namespace {

// packed row n -> (t, b): binary search over the (<=128 entry) offset table held in param space.
__device__ __forceinline__ void row_to_tb(const StepTable& tab, int n, int& t, int& b) {
  int lo = 0, hi = tab.nsteps - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (tab.off[mid] <= n) lo = mid; else hi = mid - 1;
  }
  t = lo;
  b = n - tab.off[lo];
}

// Token ids index the embedding table and the logits row: an id outside [0, V) (vocabulary / checkpoint
// mismatch, a bad padding value) must not become an out-of-bounds access.  The reference's nn.Embedding /
// CrossEntropyLoss device-assert in that case; here the id is clamped for the access and reported through a
// host-mapped status word {count, first bad id} that the host side reads at its next call (st_token_error).
__device__ __forceinline__ int64_t checked_token(int64_t tok, int V, long long* status) {
  if (tok >= 0 && tok < V) return tok;
  if (status) {
    if (atomicAdd_system(reinterpret_cast<unsigned long long*>(status), 1ull) == 0ull) status[1] = tok;
    __threadfence_system();
  }
  return tok < 0 ? 0 : V - 1;
}

long long* g_token_status_host = nullptr;   // pinned, mapped; [0] = number of bad ids seen, [1] = the first one
long long* g_token_status_dev = nullptr;

long long* token_status() {
  if (!g_token_status_dev) {
    long long* h = nullptr;
    if (cudaHostAlloc(&h, 2 * sizeof(long long), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;                         // no status word: ids are still clamped
    }
    h[0] = h[1] = 0;
    long long* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) {
      cudaGetLastError();
      cudaFreeHost(h);
      return nullptr;
    }
    g_token_status_host = h;
    g_token_status_dev = d;
  }
  return g_token_status_dev;
}

__device__ __forceinline__ void store_as(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_as(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <typename TX>
__global__ void pack_inputs_kernel(const __grid_constant__ StepTable tab, TX* __restrict__ X, int ldx,
                                   const float* __restrict__ emb, int E, int V,
                                   const float* __restrict__ feature,
                                   const int64_t* __restrict__ caption, int T_cap, int with_feature,
                                   long long* status) {
  const int n = blockIdx.x;
  int t, b;
  row_to_tb(tab, n, t, b);
  const float* src;
  if (with_feature && t == 0) {
    src = feature + (size_t)b * E;
  } else {
    int64_t tok = caption[(size_t)b * T_cap + (with_feature ? t - 1 : t)];
    if (threadIdx.x == 0) checked_token(tok, V, status);
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    src = emb + (size_t)tok * E;
  }
  TX* dst = X + (size_t)n * ldx;
  for (int e = threadIdx.x; e < E; e += blockDim.x) store_as(dst + e, src[e]);
}

__global__ void pack_inputs_bwd_kernel(const __grid_constant__ StepTable tab, const float* __restrict__ dX,
                                       int ldx, float* __restrict__ dEmb, int E,
                                       float* __restrict__ dfeature,
                                       const int64_t* __restrict__ caption, int T_cap,
                                       int with_feature, int V) {
  const int n = blockIdx.x;
  int t, b;
  row_to_tb(tab, n, t, b);
  const float* src = dX + (size_t)n * ldx;
  if (with_feature && t == 0) {
    if (dfeature) {
      float* dst = dfeature + (size_t)b * E;
      for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = src[e];
    }
    return;
  }
  int64_t tok = caption[(size_t)b * T_cap + (with_feature ? t - 1 : t)];
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);      // reported by the forward pack (same captions)
  float* dst = dEmb + (size_t)tok * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(dst + e, src[e]);
}

__global__ void pack_targets_kernel(const __grid_constant__ StepTable tab, int64_t* __restrict__ out,
                                    const int64_t* __restrict__ caption, int T_cap, int N, int V,
                                    long long* status) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int t, b;
  row_to_tb(tab, n, t, b);
  out[n] = checked_token(caption[(size_t)b * T_cap + t], V, status);
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// out[c] (+)= sum_r M[r,c].  CTA = 32 columns x 8 row-stripes over a CS_SLAB-row slab, four loads in flight per thread.
constexpr int CS_SLAB = 128;
template <typename T>
__global__ void colsum_kernel(float* __restrict__ out, const T* __restrict__ M, int rows, int cols,
                              int ld) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * CS_SLAB;
  float s = 0.f;
  if (c < cols) {
    const int r1 = min(rows, r0 + CS_SLAB);
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      const float v0 = to_f32(M[(size_t)r * ld + c]), v1 = to_f32(M[(size_t)(r + 8) * ld + c]);
      const float v2 = to_f32(M[(size_t)(r + 16) * ld + c]), v3 = to_f32(M[(size_t)(r + 24) * ld + c]);
      s += (v0 + v1) + (v2 + v3);
    }
    for (; r < r1; r += 8) s += to_f32(M[(size_t)r * ld + c]);
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x];
    atomicAdd(out + c, v);
  }
}

// bf16 matrix, 16-byte loads: CTA = 256 columns (32 lanes x 8) x 8 row-stripes over a 256-row slab.  Column sums of
// the (N, V) dlogits (db_v) and of the (N, G*H) gate gradients (db_ih / db_hh) read the rows the GEMMs read.
__global__ void __launch_bounds__(256) colsum_bf16v_kernel(float* __restrict__ out, const __nv_bfloat16* __restrict__ M,
                                                           int rows, int cols, int ld) {
  __shared__ float red[8][256 + 8];
  const int c0 = blockIdx.x * 256 + threadIdx.x * 8;
  const int r0 = blockIdx.y * 256, r1 = min(rows, r0 + 256);
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (c0 + 8 <= cols) {
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {                 // four independent 16-byte loads in flight per thread
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldcs(reinterpret_cast<const uint4*>(M + (size_t)(r + 8 * u) * ld + c0));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q[u]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          s[2 * i] += f.x; s[2 * i + 1] += f.y;
        }
      }
    }
    for (; r < r1; r += 8) {
      const uint4 q = __ldcs(reinterpret_cast<const uint4*>(M + (size_t)r * ld + c0));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        s[2 * i] += f.x; s[2 * i + 1] += f.y;
      }
    }
  } else if (c0 < cols) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8)
      for (int i = 0; i < 8 && c0 + i < cols; ++i) s[i] += __bfloat162float(M[(size_t)r * ld + c0 + i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.y][threadIdx.x * 8 + i] = s[i];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x, c = blockIdx.x * 256 + t;
  if (c < cols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][t];
    atomicAdd(out + c, v);
  }
}

// dst[i, :width] = table[idx[i * idx_stride], :]  (nn.Embedding lookup of the decode loops, rnn.py:53)
__global__ void gather_rows_kernel(float* __restrict__ dst, int ld_dst, const float* __restrict__ table, int width,
                                   const int64_t* __restrict__ idx, int idx_stride) {
  const float* src = table + (size_t)idx[(size_t)blockIdx.x * idx_stride] * width;
  float* d = dst + (size_t)blockIdx.x * ld_dst;
  for (int e = threadIdx.x; e < width; e += blockDim.x) d[e] = src[e];
}

// out[r] = sum_c M[r, c] for a bf16 matrix: one warp per row, 16-byte loads (bias gradient of the
// vocabulary projection from the transposed dlogits, which are row-contiguous per vocabulary entry).
__global__ void rowsum_bf16_kernel(float* __restrict__ out, const __nv_bfloat16* __restrict__ M, int rows,
                                   int cols, int ld) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const __nv_bfloat16* row = M + (size_t)r * ld;
  float s = 0.f;
  const int c8 = (((uintptr_t)row & 15) == 0) ? (cols & ~7) : 0;
  for (int c = lane * 8; c < c8; c += 256) {
    const uint4 q = *reinterpret_cast<const uint4*>(row + c);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      s += f.x + f.y;
    }
  }
  for (int c = c8 + lane; c < cols; c += 32) s += __bfloat162float(row[c]);
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}

template <typename T>
__global__ void shift_states_kernel(const __grid_constant__ StepTable tab, T* __restrict__ Hprev,
                                    const T* __restrict__ Hs, const T* __restrict__ h0, int H) {
  const int n = blockIdx.x;
  int t, b;
  row_to_tb(tab, n, t, b);
  T* dst = Hprev + (size_t)n * H;
  if (t == 0) {
    for (int e = threadIdx.x; e < H; e += blockDim.x) dst[e] = h0 ? h0[(size_t)b * H + e] : T(0.f);
  } else {
    const T* src = Hs + (size_t)(tab.off[t - 1] + b) * H;
    for (int e = threadIdx.x; e < H; e += blockDim.x) dst[e] = src[e];
  }
}


// dst_i = src_i * (*g) for up to ST_SCALE_MAX tensors in one launch: the tensors form one index space of
// 4096-element chunks (first[i] = first chunk of tensor i), one chunk per CTA iteration.
struct ScaleTable {
  int n;
  const float* src[ST_SCALE_MAX];
  float* dst[ST_SCALE_MAX];
  long long count[ST_SCALE_MAX];
  int first[ST_SCALE_MAX + 1];
};
constexpr int SCALE_CHUNK = 4096;
__global__ void __launch_bounds__(256) scale_multi_kernel(const ScaleTable tab, const float* __restrict__ g) {
  const float a = __ldg(g);
  for (int c = blockIdx.x; c < tab.first[tab.n]; c += gridDim.x) {
    int i = 0;
    while (c >= tab.first[i + 1]) ++i;
    const long long base = (long long)(c - tab.first[i]) * SCALE_CHUNK;
    const long long n = min((long long)SCALE_CHUNK, tab.count[i] - base);
    const float* __restrict__ src = tab.src[i] + base;
    float* __restrict__ dst = tab.dst[i] + base;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
      const int n4 = (int)(n >> 2);
      for (int j = threadIdx.x; j < n4; j += blockDim.x) {
        float4 v = __ldcs(reinterpret_cast<const float4*>(src) + j);
        v.x *= a; v.y *= a; v.z *= a; v.w *= a;
        reinterpret_cast<float4*>(dst)[j] = v;
      }
      for (int j = (n4 << 2) + threadIdx.x; j < n; j += blockDim.x) dst[j] = src[j] * a;
    } else {
      for (int j = threadIdx.x; j < n; j += blockDim.x) dst[j] = src[j] * a;
    }
  }
}


// ---------------------------------------------------------------------------------------------- collate (utils.py:61-77)
// create_batch sorts the samples by caption length, longest first, with Python's STABLE sort: the position of sample i
// is  #{j : len_j > len_i} + #{j < i : len_j == len_i}.  One thread per sample counts both over all B lengths (staged
// through shared memory); B is a data-loader batch, so the B^2 compares are microseconds.  batch_sizes[t] =
// #{i : len_i > t} is what pack_padded_sequence (rnn.py:31) derives from the sorted lengths.
__global__ void __launch_bounds__(256) collate_rank_kernel(const int64_t* __restrict__ len, int B, int64_t* __restrict__ perm,
                                                           int64_t* __restrict__ sorted_len) {
  __shared__ int64_t tile[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int64_t mine = i < B ? len[i] : 0;
  int rank = 0;
  for (int j0 = 0; j0 < B; j0 += 256) {
    __syncthreads();
    tile[threadIdx.x] = j0 + threadIdx.x < B ? len[j0 + threadIdx.x] : INT64_MIN;
    __syncthreads();
    const int n = min(256, B - j0);
    for (int j = 0; j < n; ++j) {
      const int64_t o = tile[j];
      rank += (o > mine) || (o == mine && j0 + j < i);
    }
  }
  if (i < B) {
    perm[rank] = i;
    sorted_len[rank] = mine;
  }
}

__global__ void __launch_bounds__(256) collate_batch_sizes_kernel(const int64_t* __restrict__ len, int B, int T,
                                                                  int32_t* __restrict__ batch_sizes) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= T) return;
  int n = 0;
  for (int i = 0; i < B; ++i) n += len[i] > t;
  batch_sizes[t] = n;
}

// dst row r = src row perm[r], 16 bytes per thread step (rows are multiples of 16 bytes and 16-byte aligned) or bytewise
__global__ void __launch_bounds__(256) gather_rows_bytes_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                                                const int64_t* __restrict__ perm, long long row_bytes, int vec) {
  const long long r = blockIdx.y;
  const uint8_t* s = src + (size_t)perm[r] * row_bytes;
  uint8_t* d = dst + (size_t)r * row_bytes;
  if (vec) {
    const long long n16 = row_bytes >> 4;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n16; i += (long long)gridDim.x * 256)
      reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
  } else {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < row_bytes; i += (long long)gridDim.x * 256) d[i] = s[i];
  }
}

// sorted, re-padded caption matrix: dst (B, T_out) <- src (B, T_in) rows perm[r], zeros from the caption's length on
// (utils.py:72-75 builds the padded matrix from zeros)
__global__ void __launch_bounds__(256) collate_captions_kernel(int64_t* __restrict__ dst, const int64_t* __restrict__ src,
                                                               const int64_t* __restrict__ perm,
                                                               const int64_t* __restrict__ sorted_len, int T_in, int T_out) {
  const int r = blockIdx.x;
  const int64_t* s = src + (size_t)perm[r] * T_in;
  const int64_t n = sorted_len[r];
  for (int t = threadIdx.x; t < T_out; t += 256) dst[(size_t)r * T_out + t] = (t < n && t < T_in) ? s[t] : 0;
}

}  // namespace
}  // namespace st

extern "C" {

int st_collate_sort(const int64_t* lengths, int B, int T, int64_t* perm, int64_t* sorted_len, int32_t* batch_sizes,
                    st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(lengths && perm && sorted_len, ST_ERR_NULL, "st_collate_sort: NULL pointer");
  ST_REQUIRE(B >= 1 && T >= 0 && (T == 0 || batch_sizes), ST_ERR_BAD_SHAPE, "st_collate_sort: B=%d T=%d", B, T);
  cudaStream_t s = as_stream(stream);
  collate_rank_kernel<<<(B + 255) / 256, 256, 0, s>>>(lengths, B, perm, sorted_len);
  ST_LAUNCH_TRY("collate_rank_kernel");
  if (T > 0) {
    collate_batch_sizes_kernel<<<(T + 255) / 256, 256, 0, s>>>(lengths, B, T, batch_sizes);
    ST_LAUNCH_TRY("collate_batch_sizes_kernel");
  }
  return ST_OK;
}

int st_gather_rows_bytes(void* dst, const void* src, const int64_t* perm, int rows, int64_t row_bytes, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(dst && src && perm, ST_ERR_NULL, "st_gather_rows_bytes: NULL pointer");
  ST_REQUIRE(rows >= 1 && rows <= 65535 && row_bytes >= 1, ST_ERR_BAD_SHAPE, "st_gather_rows_bytes: rows=%d row_bytes=%lld",
             rows, (long long)row_bytes);
  const int vec = (row_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0);
  const long long units = vec ? row_bytes / 16 : row_bytes;
  const unsigned gx = (unsigned)((units + 255) / 256 < 64 ? (units + 255) / 256 : 64);
  gather_rows_bytes_kernel<<<dim3(gx, rows), 256, 0, as_stream(stream)>>>(reinterpret_cast<uint8_t*>(dst),
                                                                          reinterpret_cast<const uint8_t*>(src), perm, row_bytes, vec);
  ST_LAUNCH_TRY("gather_rows_bytes_kernel");
  return ST_OK;
}

int st_collate_captions(int64_t* dst, const int64_t* src, const int64_t* perm, const int64_t* sorted_len, int B, int T_in,
                        int T_out, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(dst && src && perm && sorted_len, ST_ERR_NULL, "st_collate_captions: NULL pointer");
  ST_REQUIRE(B >= 1 && T_in >= 1 && T_out >= 1, ST_ERR_BAD_SHAPE, "st_collate_captions: B=%d T_in=%d T_out=%d", B, T_in, T_out);
  collate_captions_kernel<<<B, 256, 0, as_stream(stream)>>>(dst, src, perm, sorted_len, T_in, T_out);
  ST_LAUNCH_TRY("collate_captions_kernel");
  return ST_OK;
}



int st_pack_inputs(float* X, int ldx, const float* emb, int E, int V, const float* feature,
                   const int64_t* caption, int T_cap, int with_feature, int nsteps,
                   const int* batch_sizes_host, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(X && emb && caption, ST_ERR_NULL, "st_pack_inputs: NULL pointer");
  ST_REQUIRE(!with_feature || feature, ST_ERR_NULL, "st_pack_inputs: feature is NULL");
  ST_REQUIRE(E >= 1 && ldx >= E && V >= 1, ST_ERR_BAD_SHAPE, "st_pack_inputs: E=%d ldx=%d V=%d", E, ldx, V);
  ST_REQUIRE(nsteps <= T_cap + (with_feature ? 1 : 0), ST_ERR_BAD_SHAPE,
             "st_pack_inputs: nsteps=%d exceeds caption length %d", nsteps, T_cap);
  pack_inputs_kernel<float><<<tab.off[nsteps], 128, 0, as_stream(stream)>>>(tab, X, ldx, emb, E, V, feature,
                                                                            caption, T_cap, with_feature, token_status());
  ST_LAUNCH_TRY("pack_inputs_kernel");
  return ST_OK;
}

int st_pack_inputs_bf16(void* X, int ldx, const float* emb, int E, int V, const float* feature,
                        const int64_t* caption, int T_cap, int with_feature, int nsteps,
                        const int* batch_sizes_host, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(X && emb && caption, ST_ERR_NULL, "st_pack_inputs_bf16: NULL pointer");
  ST_REQUIRE(!with_feature || feature, ST_ERR_NULL, "st_pack_inputs_bf16: feature is NULL");
  ST_REQUIRE(E >= 1 && ldx >= E && V >= 1, ST_ERR_BAD_SHAPE, "st_pack_inputs_bf16: E=%d ldx=%d V=%d", E, ldx, V);
  ST_REQUIRE(nsteps <= T_cap + (with_feature ? 1 : 0), ST_ERR_BAD_SHAPE,
             "st_pack_inputs_bf16: nsteps=%d exceeds caption length %d", nsteps, T_cap);
  pack_inputs_kernel<__nv_bfloat16><<<tab.off[nsteps], 128, 0, as_stream(stream)>>>(
      tab, reinterpret_cast<__nv_bfloat16*>(X), ldx, emb, E, V, feature, caption, T_cap, with_feature, token_status());
  ST_LAUNCH_TRY("pack_inputs_kernel");
  return ST_OK;
}

int st_pack_inputs_bwd(const float* dX, int ldx, float* dEmb, int E, int V, float* dfeature,
                       const int64_t* caption, int T_cap, int with_feature, int nsteps,
                       const int* batch_sizes_host, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(dX && dEmb && caption, ST_ERR_NULL, "st_pack_inputs_bwd: NULL pointer");
  ST_REQUIRE(E >= 1 && ldx >= E && V >= 1, ST_ERR_BAD_SHAPE, "st_pack_inputs_bwd: E=%d ldx=%d V=%d", E, ldx, V);
  pack_inputs_bwd_kernel<<<tab.off[nsteps], 128, 0, as_stream(stream)>>>(
      tab, dX, ldx, dEmb, E, dfeature, caption, T_cap, with_feature, V);
  ST_LAUNCH_TRY("pack_inputs_bwd_kernel");
  return ST_OK;
}

int st_token_error(int64_t* first_bad_id, int clear) {
  long long* h = st::g_token_status_host;
  if (!h) return 0;
  const long long n = __atomic_load_n(&h[0], __ATOMIC_ACQUIRE);
  if (first_bad_id) *first_bad_id = (int64_t)h[1];
  if (clear && n) {
    h[1] = 0;
    __atomic_store_n(&h[0], 0, __ATOMIC_RELEASE);
  }
  return n > 0x7fffffff ? 0x7fffffff : (int)n;
}

int st_pack_targets(int64_t* out, const int64_t* caption, int T_cap, int V, int nsteps,
                    const int* batch_sizes_host, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(out && caption, ST_ERR_NULL, "st_pack_targets: NULL pointer");
  ST_REQUIRE(nsteps <= T_cap && V >= 1, ST_ERR_BAD_SHAPE, "st_pack_targets: nsteps=%d > T_cap=%d (V=%d)", nsteps, T_cap, V);
  const int N = tab.off[nsteps];
  pack_targets_kernel<<<(N + 255) / 256, 256, 0, as_stream(stream)>>>(tab, out, caption, T_cap, N, V, token_status());
  ST_LAUNCH_TRY("pack_targets_kernel");
  return ST_OK;
}

int st_colsum(float* out, const void* M, int m_is_bf16, int rows, int cols, int ld, int accumulate,
              st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(out && M, ST_ERR_NULL, "st_colsum: NULL pointer");
  ST_REQUIRE(rows >= 0 && cols >= 1 && ld >= cols, ST_ERR_BAD_SHAPE, "st_colsum: rows=%d cols=%d ld=%d",
             rows, cols, ld);
  cudaStream_t s = as_stream(stream);
  if (!accumulate) ST_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float) * cols, s));
  if (rows == 0) return ST_OK;
  dim3 grid((cols + 31) / 32, (rows + CS_SLAB - 1) / CS_SLAB), block(32, 8);
  ST_REQUIRE(grid.y <= 65535, ST_ERR_BAD_SHAPE, "st_colsum: too many rows (%d)", rows);
  if (m_is_bf16 && (ld & 7) == 0 && (reinterpret_cast<uintptr_t>(M) & 15) == 0 && cols >= 64) {
    dim3 gv((cols + 255) / 256, (rows + 255) / 256);
    colsum_bf16v_kernel<<<gv, block, 0, s>>>(out, reinterpret_cast<const __nv_bfloat16*>(M), rows, cols, ld);
  } else if (m_is_bf16)
    colsum_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(out, reinterpret_cast<const __nv_bfloat16*>(M), rows, cols, ld);
  else
    colsum_kernel<float><<<grid, block, 0, s>>>(out, reinterpret_cast<const float*>(M), rows, cols, ld);
  ST_LAUNCH_TRY("colsum_kernel");
  return ST_OK;
}

int st_gather_rows(float* dst, int ld_dst, const float* table, int width, const int64_t* idx, int idx_stride,
                   int n, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(dst && table && idx, ST_ERR_NULL, "st_gather_rows: NULL pointer");
  ST_REQUIRE(n >= 1 && width >= 1 && ld_dst >= width && idx_stride >= 1, ST_ERR_BAD_SHAPE, "st_gather_rows: bad shape");
  gather_rows_kernel<<<n, 128, 0, as_stream(stream)>>>(dst, ld_dst, table, width, idx, idx_stride);
  ST_LAUNCH_TRY("gather_rows_kernel");
  return ST_OK;
}

int st_rowsum_bf16(float* out, const void* M, int rows, int cols, int ld, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(out && M, ST_ERR_NULL, "st_rowsum_bf16: NULL pointer");
  ST_REQUIRE(rows >= 1 && cols >= 1 && ld >= cols, ST_ERR_BAD_SHAPE, "st_rowsum_bf16: bad shape");
  rowsum_bf16_kernel<<<(rows + 7) / 8, 256, 0, as_stream(stream)>>>(out, reinterpret_cast<const __nv_bfloat16*>(M),
                                                                    rows, cols, ld);
  ST_LAUNCH_TRY("rowsum_bf16_kernel");
  return ST_OK;
}

int st_shift_states(float* Hprev, const float* Hs, const float* h0, int H, int nsteps,
                    const int* batch_sizes_host, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(Hprev && Hs, ST_ERR_NULL, "st_shift_states: NULL pointer");
  ST_REQUIRE(H >= 1, ST_ERR_BAD_SHAPE, "st_shift_states: H=%d", H);
  shift_states_kernel<float><<<tab.off[nsteps], 128, 0, as_stream(stream)>>>(tab, Hprev, Hs, h0, H);
  ST_LAUNCH_TRY("shift_states_kernel");
  return ST_OK;
}

int st_shift_states_bf16(void* Hprev, const void* Hs, const void* h0, int H, int nsteps,
                         const int* batch_sizes_host, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(Hprev && Hs, ST_ERR_NULL, "st_shift_states_bf16: NULL pointer");
  ST_REQUIRE(H >= 1, ST_ERR_BAD_SHAPE, "st_shift_states_bf16: H=%d", H);
  shift_states_kernel<__nv_bfloat16><<<tab.off[nsteps], 128, 0, as_stream(stream)>>>(
      tab, reinterpret_cast<__nv_bfloat16*>(Hprev), reinterpret_cast<const __nv_bfloat16*>(Hs),
      reinterpret_cast<const __nv_bfloat16*>(h0), H);
  ST_LAUNCH_TRY("shift_states_kernel");
  return ST_OK;
}

int st_scale_multi(int n, const float* const* src, float* const* dst, const int64_t* count, const float* g,
                   st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(n >= 0 && n <= ST_SCALE_MAX, ST_ERR_BAD_SHAPE, "st_scale_multi: n=%d (max %d)", n, ST_SCALE_MAX);
  ST_REQUIRE(n == 0 || (src && dst && count && g), ST_ERR_NULL, "st_scale_multi: NULL pointer");
  if (n == 0) return ST_OK;
  ScaleTable tab;
  tab.n = n;
  long long chunks = 0;
  for (int i = 0; i < n; ++i) {
    ST_REQUIRE(count[i] >= 0 && (count[i] == 0 || (src[i] && dst[i])), ST_ERR_NULL, "st_scale_multi: tensor %d is NULL", i);
    tab.src[i] = src[i]; tab.dst[i] = dst[i]; tab.count[i] = count[i];
    tab.first[i] = (int)chunks;
    chunks += (count[i] + SCALE_CHUNK - 1) / SCALE_CHUNK;
    ST_REQUIRE(chunks < (1LL << 30), ST_ERR_BAD_SHAPE, "st_scale_multi: too many elements");
  }
  tab.first[n] = (int)chunks;
  if (chunks == 0) return ST_OK;
  int sms = 0;
  ST_TRY(st_device_info(&sms, nullptr, nullptr, nullptr));
  const long long grid = chunks < 8LL * sms ? chunks : 8LL * sms;
  scale_multi_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(tab, g);
  ST_LAUNCH_TRY("scale_multi_kernel");
  return ST_OK;
}

}  // extern "C"
