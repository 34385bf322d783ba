// Dev probe (not part of the library): how long does a chain of small tcgen05.mma instructions take?
// One warp of each CTA issues NM MMAs (kind::f16, M x N x 16, both operands in shared memory, K-major SW128) that
// accumulate round-robin into NACC independent TMEM accumulators, commits, and waits for completion.
// Reports ns per MMA for (M, N, NACC): the recurrent kernels issue 32..128 such MMAs per time step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I showtell_b200/csrc -I include -o tools/probe_mma.bin tools/probe_mma.cu
#include <cstdio>
#include <vector>

#include "tc_common.cuh"
using namespace st;

template <int M, int N, int NACC>
__global__ void __launch_bounds__(128, 1) mma_kernel(int nm, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc(M, N);
    const uint64_t ad = umma_desc_k128(smem_u32(smem)), bd = umma_desc_k128(smem_u32(smem + 16384));
    long long best = 1ll << 60;
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < nm; ++i)
          tc_mma(tmem + (i % NACC) * N, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i >= NACC);
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, ph);
      ph ^= 1;
      long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = best;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int M, int N, int NACC>
void run(int nm) {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  auto k = mma_kernel<M, N, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<<<148, 128, 64 * 1024>>>(nm, 20, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("M=%d N=%d NACC=%d: %s\n", M, N, NACC, cudaGetErrorString(e)); return; }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), out, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
  long long mn = h[0], mx = h[0];
  for (long long v : h) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
  printf("M=%3d N=%3d NACC=%d nm=%3d: %6lld .. %6lld clk = %.1f clk per MMA\n", M, N, NACC, nm, mn, mx, (double)mn / nm);
  cudaFree(out);
}

int main() {
  for (int nm : {32, 64, 128}) {
    run<64, 16, 1>(nm); run<64, 16, 2>(nm); run<64, 16, 4>(nm);
    run<64, 32, 1>(nm); run<64, 32, 4>(nm);
    run<128, 16, 1>(nm); run<128, 16, 4>(nm);
    run<128, 48, 1>(nm); run<128, 48, 4>(nm);
    run<128, 64, 1>(nm); run<128, 128, 1>(nm); run<128, 256, 1>(nm);
  }
  return 0;
}
