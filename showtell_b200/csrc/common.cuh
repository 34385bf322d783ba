// Shared host/device helpers for libshowtell_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "showtell_b200.h"

namespace st {

void set_error(const char* fmt, ...);
void note_launch(int n = 1);  // counts kernel launches issued by this library (st_launch_count)

// batch_sizes + packed-row offsets of a PackedSequence, passed to kernels by value
// (pack_padded_sequence semantics, rnn.py:31).
struct StepTable {
  int nsteps;
  int bs[ST_MAX_STEPS + 1];   // bs[nsteps] = 0 sentinel
  int off[ST_MAX_STEPS + 1];  // off[t] = sum_{s<t} bs[s]; off[nsteps] = N
};

// Validates (1 <= nsteps <= ST_MAX_STEPS, sizes positive and non-increasing) and fills `tab`.
int make_step_table(StepTable& tab, int nsteps, const int* batch_sizes_host);

inline cudaStream_t as_stream(st_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (st_debug_set_pdl, default on): kernels that call pdl_wait() before they touch
// anything the previous kernel of the stream wrote (or still reads) may be scheduled while that kernel drains,
// so their prologue -- barrier init, TMEM allocation, tensor-map fetch, prefetch of step-invariant operands --
// leaves the dependent chain of the per-step launches.
extern int g_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace st

#define ST_CUDA_TRY(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      st::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,      \
                    __LINE__);                                                             \
      return ST_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define ST_LAUNCH_TRY(what)                                                                \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      st::set_error("launch of %s failed: %s (%s:%d)", what, cudaGetErrorString(_e),       \
                    __FILE__, __LINE__);                                                   \
      return ST_ERR_CUDA;                                                                  \
    }                                                                                      \
    st::note_launch();                                                                     \
  } while (0)

#define ST_REQUIRE(cond, code, ...)                                                        \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      st::set_error(__VA_ARGS__);                                                          \
      return (code);                                                                       \
    }                                                                                      \
  } while (0)

#define ST_TRY(expr)                                                                       \
  do {                                                                                     \
    int _s = (expr);                                                                       \
    if (_s != ST_OK) return _s;                                                            \
  } while (0)

#ifdef __CUDACC__
namespace st {

// PDL, device side: wait until the preceding kernel(s) of the stream have completed and their writes are
// visible (no-op without the launch attribute); allow the next kernel of the stream to be scheduled.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// One-sided wait of a single thread for a monotonic arrival counter (bounded: traps instead of hanging).
__device__ __forceinline__ void wait_counter_geq(const int* counter, int target) {
  for (unsigned spin = 0; ld_acquire_gpu(counter) < target; ++spin)
    if (spin > (1u << 27)) __trap();
}

// Device-wide barrier among the CTAs that share `counter` (monotonic: the k-th barrier waits for
// k * participants arrivals).  Requires all participants co-resident (cooperative launch).
__device__ __forceinline__ void grid_barrier(int* counter, int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1);
    while (ld_acquire_gpu(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace st
#endif
