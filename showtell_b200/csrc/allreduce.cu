// Gradient exchange of the data-parallel training step: in-place sum all-reduce of fp32
// gradients that live in a SYMMETRIC buffer -- the same allocation on every GPU of the node, each
// rank's copy mapped into every peer's address space over NVLink 5 / NVSwitch, plus (when the fabric
// offers it) one multicast mapping of all copies.
//
// Two-shot, one kernel, no staging copy:
//   barrier A   every rank's producers are done (the kernel is stream-ordered after them on its own
//               rank, so arriving at the barrier says "my gradients are final")
//   reduce      rank r owns the r-th 1/world slice: with multicast it issues multimem.ld_reduce (the
//               NVSwitch pulls the slice from all GPUs and adds in the switch) and multimem.st (the
//               switch broadcasts the sums into every copy); without multicast it loads the slice
//               from each peer, adds in rank order (every rank ends with bit-identical sums) and
//               stores it to each peer
//   barrier B   all slices have landed everywhere
// The barriers are per-CTA channels of one arrival word per (CTA, peer) in the symmetric buffer itself,
// carrying a monotonic epoch, so they need no reset between launches (or CUDA-graph replays).  A few
// CTAs are enough to fill the links (8 of them reach 530-620 GB/s bus bandwidth on 8 GPUs), which
// leaves the SMs to the BPTT kernels this exchange overlaps.  All spins are bounded (trap, never hang).
#include "common.cuh"

namespace st {
namespace {

constexpr int AR_THREADS = 512;
constexpr int AR_UNROLL = 4;

struct ArParams {
  float* peer[ST_AR_MAX_WORLD];
  uint32_t* flag[ST_AR_MAX_WORLD];
  float* mc;
  int rank, world;
  long long count4;   // number of 16-byte units
  unsigned long long timeout_ns;   // wall-clock bound of a cross-GPU barrier wait
  int* error;                      // mapped host word: set to 1 + peer when a wait timed out
};

unsigned long long g_timeout_ns = 300ull * 1000000000ull;   // st_allreduce_set_timeout_ms; default 5 minutes
int* g_error_host = nullptr;                                 // cudaHostAllocMapped word, read by st_allreduce_error()
int* g_error_dev = nullptr;

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Flag area of one rank: per CTA channel `world` arrival words (written by the peers) followed by the
// channel's epoch (local, advanced by 2 per launch).  Barrier k of a launch: every rank posts epoch + k
// into its word on every peer (a plain remote store, no round trip) and waits until all its own words
// reached that value.  A rank can be at most one barrier ahead of a peer, so ">=" on the wrapping
// difference is exact, and the state survives back-to-back launches and CUDA-graph replays.
__device__ __forceinline__ uint32_t* channel(uint32_t* flags, int world) { return flags + (size_t)blockIdx.x * (world + 1); }

__device__ __forceinline__ void peer_barrier(const ArParams& p, uint32_t value) {
  __syncthreads();
  if ((int)threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    fence_sys();                                                     // release what this CTA wrote
    st_flag(channel(p.flag[peer], p.world) + p.rank, value);
    const uint32_t* from = channel(p.flag[p.rank], p.world) + peer;
    // A peer may legitimately be late by seconds (evaluation, a checkpoint, a data-loader stall, a first-time JIT): the
    // wait is bounded by WALL TIME (default 5 minutes), not by a poll count; on timeout the error word is set for the
    // host (st_allreduce_error) and the kernel goes on -- it never hangs and never kills the context of a healthy rank.
    unsigned long long t0 = 0;
    for (unsigned spin = 0; (int32_t)(ld_flag(from) - value) < 0; ++spin) {
      if ((spin & 0x3ff) == 0x3ff) {
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > p.timeout_ns) {
          if (p.error) *reinterpret_cast<volatile int*>(p.error) = 1 + peer;
          break;
        }
      }
    }
    fence_sys();                                                     // acquire what the peers wrote
  }
  __syncthreads();
}

__device__ __forceinline__ float4 mm_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// peer copies are written by other GPUs between launches: never serve them from L1
__device__ __forceinline__ float4 ld_peer(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool MC>
__global__ void __launch_bounds__(AR_THREADS, 1) allreduce_sum_kernel(const ArParams p) {
  uint32_t* my_epoch = channel(p.flag[p.rank], p.world) + p.world;
  const uint32_t epoch = *my_epoch;                                  // written by the previous launch on this stream
  peer_barrier(p, epoch + 1);
  // slice of this rank, in 16-byte units
  const long long per = (p.count4 + p.world - 1) / p.world;
  const long long lo = min(p.count4, per * p.rank), hi = min(p.count4, lo + per);
  const long long stride = (long long)gridDim.x * AR_THREADS;
  for (long long base = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; base < hi; base += stride * AR_UNROLL) {
    float4 acc[AR_UNROLL];
    if (MC) {
#pragma unroll
      for (int u = 0; u < AR_UNROLL; ++u) {
        const long long i = base + u * stride;
        if (i < hi) acc[u] = mm_ld_reduce(p.mc + 4 * i);
      }
#pragma unroll
      for (int u = 0; u < AR_UNROLL; ++u) {
        const long long i = base + u * stride;
        if (i < hi) mm_st(p.mc + 4 * i, acc[u]);
      }
    } else {
#pragma unroll
      for (int u = 0; u < AR_UNROLL; ++u) {
        const long long i = base + u * stride;
        if (i >= hi) continue;
        float4 v[ST_AR_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < ST_AR_MAX_WORLD; ++r)
          if (r < p.world) v[r] = ld_peer(p.peer[r] + 4 * i);
        acc[u] = v[0];
#pragma unroll
        for (int r = 1; r < ST_AR_MAX_WORLD; ++r)
          if (r < p.world) { acc[u].x += v[r].x; acc[u].y += v[r].y; acc[u].z += v[r].z; acc[u].w += v[r].w; }
      }
#pragma unroll
      for (int u = 0; u < AR_UNROLL; ++u) {
        const long long i = base + u * stride;
        if (i >= hi) continue;
#pragma unroll
        for (int r = 0; r < ST_AR_MAX_WORLD; ++r)
          if (r < p.world) st_peer(p.peer[r] + 4 * i, acc[u]);
      }
    }
  }
  peer_barrier(p, epoch + 2);
  if (threadIdx.x == 0) *my_epoch = epoch + 2;
}

// Two GPUs: the switch's multicast reduce tops out at ~300 GB/s bus bandwidth there (it pulls BOTH copies through
// the switch, the requester's own included), plain peer loads / stores move half the bytes over the links.  One peer
// round trip per unit, so the loop keeps AR_UNROLL2 16-byte peer loads in flight per thread (4 gave 115 GB/s with 16
// CTAs: latency-bound).  Sums are formed as (rank 0's value) + (rank 1's value) on both ranks: bit-identical.
constexpr int AR_UNROLL2 = 8;
__global__ void __launch_bounds__(AR_THREADS, 1) allreduce2_kernel(const ArParams p) {
  uint32_t* my_epoch = channel(p.flag[p.rank], p.world) + p.world;
  const uint32_t epoch = *my_epoch;
  peer_barrier(p, epoch + 1);
  const long long per = (p.count4 + 1) / 2;
  const long long lo = min(p.count4, per * p.rank), hi = min(p.count4, lo + per);
  const long long stride = (long long)gridDim.x * AR_THREADS;
  float* mine = p.peer[p.rank];
  float* other = p.peer[p.rank ^ 1];
  for (long long base = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; base < hi; base += stride * AR_UNROLL2) {
    float4 a[AR_UNROLL2], b[AR_UNROLL2];
#pragma unroll
    for (int u = 0; u < AR_UNROLL2; ++u) {
      const long long i = base + u * stride;
      if (i < hi) b[u] = ld_peer(other + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < AR_UNROLL2; ++u) {
      const long long i = base + u * stride;
      if (i < hi) a[u] = ld_peer(mine + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < AR_UNROLL2; ++u) {
      const long long i = base + u * stride;
      if (i >= hi) continue;
      const float4 x = p.rank == 0 ? a[u] : b[u], y = p.rank == 0 ? b[u] : a[u];      // rank order on both sides
      const float4 r = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
      st_peer(other + 4 * i, r);
      st_peer(mine + 4 * i, r);
    }
  }
  peer_barrier(p, epoch + 2);
  if (threadIdx.x == 0) *my_epoch = epoch + 2;
}

int g_pair_p2p = 1;   // st_debug_allreduce_pair_p2p

}  // namespace
}  // namespace st

extern "C" {

int st_allreduce_flag_words(int world) { return ST_AR_MAX_BLOCKS * (world + 1); }

int st_allreduce_set_timeout_ms(int64_t ms) {
  st::g_timeout_ns = ms > 0 ? (unsigned long long)ms * 1000000ull : 300ull * 1000000000ull;
  return ST_OK;
}

int st_debug_allreduce_pair_p2p(int on) {
  st::g_pair_p2p = on ? 1 : 0;
  return ST_OK;
}

int st_allreduce_error(void) {
  if (!st::g_error_host) return 0;
  const int e = *reinterpret_cast<volatile int*>(st::g_error_host);
  return e;
}

int st_allreduce_sum_f32(void* const* peers_host, void* multicast, void* const* flags_host, int rank, int world,
                         int64_t count, int nblocks, st_stream_t stream) {
  using namespace st;
  ST_REQUIRE(peers_host && flags_host, ST_ERR_NULL, "st_allreduce_sum_f32: NULL pointer table");
  ST_REQUIRE(world >= 1 && world <= ST_AR_MAX_WORLD && rank >= 0 && rank < world, ST_ERR_BAD_SHAPE,
             "st_allreduce_sum_f32: rank %d / world %d (max %d)", rank, world, (int)ST_AR_MAX_WORLD);
  ST_REQUIRE(count >= 0 && count % 4 == 0, ST_ERR_BAD_SHAPE,
             "st_allreduce_sum_f32: count=%lld must be a multiple of 4 floats", (long long)count);
  ST_REQUIRE(nblocks >= 1 && nblocks <= ST_AR_MAX_BLOCKS, ST_ERR_BAD_SHAPE,
             "st_allreduce_sum_f32: nblocks=%d outside [1, %d]", nblocks, (int)ST_AR_MAX_BLOCKS);
  if (world == 1 || count == 0) return ST_OK;
  ArParams p;
  for (int r = 0; r < ST_AR_MAX_WORLD; ++r) {
    p.peer[r] = r < world ? reinterpret_cast<float*>(peers_host[r]) : nullptr;
    p.flag[r] = r < world ? reinterpret_cast<uint32_t*>(flags_host[r]) : nullptr;
    ST_REQUIRE(r >= world || (p.peer[r] && p.flag[r]), ST_ERR_NULL, "st_allreduce_sum_f32: peer %d has a NULL mapping", r);
    ST_REQUIRE(r >= world || (reinterpret_cast<uintptr_t>(p.peer[r]) % 16 == 0), ST_ERR_BAD_SHAPE,
               "st_allreduce_sum_f32: peer %d mapping is not 16-byte aligned", r);
  }
  ST_REQUIRE(reinterpret_cast<uintptr_t>(multicast) % 16 == 0, ST_ERR_BAD_SHAPE,
             "st_allreduce_sum_f32: multicast mapping is not 16-byte aligned");
  p.mc = reinterpret_cast<float*>(multicast);
  if (!g_error_host) {
    ST_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&g_error_host), sizeof(int), cudaHostAllocMapped));
    *g_error_host = 0;
    ST_CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_error_dev), g_error_host, 0));
  }
  p.timeout_ns = g_timeout_ns;
  p.error = g_error_dev;
  p.rank = rank;
  p.world = world;
  p.count4 = count / 4;
  if (world == 2 && (g_pair_p2p || !multicast)) allreduce2_kernel<<<nblocks, AR_THREADS, 0, as_stream(stream)>>>(p);
  else if (multicast) allreduce_sum_kernel<true><<<nblocks, AR_THREADS, 0, as_stream(stream)>>>(p);
  else allreduce_sum_kernel<false><<<nblocks, AR_THREADS, 0, as_stream(stream)>>>(p);
  ST_LAUNCH_TRY("allreduce_sum_kernel");
  return ST_OK;
}

}  // extern "C"
