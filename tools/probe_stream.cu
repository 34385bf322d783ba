// Dev probe (not part of the library): what bounds the per-step operand stream of the BPTT kernel?
// Each CTA (x, y) of a (X, Y) grid streams the R x 2048 bf16 tile of batch tile y (rows y*R.., row stride 4096 B)
// KB = 32 k-blocks of R rows x 128 B through an S-stage TMA ring, REP times, with nothing consuming it (a thread
// releases each stage as it lands).  Variants: mode 0 = 2-D tensor-map boxes (what the kernel does today),
// mode 1 = one contiguous cp.async.bulk of R*128 B per k-block from a tile-major copy, fresh = 1: before every
// repetition the CTAs of a batch tile rewrite the tile with 16-byte stores (as phase 1 does) and meet at a barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_stream.bin tools/probe_stream.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int GH = 2048, KB = 32, MAXS = 16;

__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* tile_major, __nv_bfloat16* rowmajor,
                                                        int R, int S, int REP, int mode, int fresh, int* counters, long long* out, const __grid_constant__ CUtensorMap tm3, int KP, int ksplit) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[MAXS], empty[MAXS];
  const uint32_t kbytes = (uint32_t)R * 128 * (mode == 2 ? KP : 1);
  const int nops = (mode == 2 ? KB / KP : KB) / ksplit;   // ksplit = 2: the CTAs of a pair stream one half of K each
  const int op0 = (blockIdx.x % ksplit) * nops;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int X = gridDim.x, y = blockIdx.y, x = blockIdx.x;
  long long total = 0;
  int stage_p = 0, stage_c = 0;
  uint32_t ph_p = 0, ph_c = 0;
  for (int rep = 0; rep < REP; ++rep) {
    if (fresh) {
      // rewrite this CTA's share of the tile: columns [x*GH/X, (x+1)*GH/X) of rows y*R.., 16-byte stores
      const int cols = GH / X, c16 = cols / 8;
      for (int i = threadIdx.x; i < R * c16; i += blockDim.x) {
        const int r = i / c16, c = (i % c16) * 8 + x * cols;
        const uint4 v = make_uint4(rep, i, x, y);
        *reinterpret_cast<uint4*>(rowmajor + (size_t)(y * R + r) * GH + c) = v;
        // tile-major copy: [y][kb][r][64]
        const int kb = c / 64, cc = c % 64;
        *reinterpret_cast<uint4*>(tile_major + (((size_t)y * KB + kb) * R + r) * 64 + cc) = v;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counters + y, 1);
        while (atomicAdd(counters + y, 0) < (rep + 1) * X) {}
        __threadfence();
        asm volatile("fence.proxy.async.global;" ::: "memory");
      }
      __syncthreads();
    }
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (int kb = 0; kb < nops; ++kb) {
        mbar_wait(&empty[stage_p], ph_p ^ 1);
        mbar_expect_tx(&full[stage_p], kbytes);
        if (mode == 2) tma_load_3d(smem + (size_t)stage_p * kbytes, &tm3, 0, y * R, (op0 + kb) * KP, &full[stage_p]);
        else if (mode == 0) tma_load_2d(smem + (size_t)stage_p * kbytes, &tm, (op0 + kb) * 64, y * R, &full[stage_p]);
        else bulk_load(smem + (size_t)stage_p * kbytes, tile_major + (((size_t)y * KB + kb) * R) * 64, kbytes, &full[stage_p]);
        if (++stage_p == S) { stage_p = 0; ph_p ^= 1; }
      }
    } else if (threadIdx.x == 32) {
      for (int kb = 0; kb < nops; ++kb) {
        mbar_wait(&full[stage_c], ph_c);
        mbar_arrive(&empty[stage_c]);
        if (++stage_c == S) { stage_c = 0; ph_c ^= 1; }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      total += t1 - t0;
    }
  }
  if (threadIdx.x == 0) out[blockIdx.y * gridDim.x + blockIdx.x] = total;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeFn enc = reinterpret_cast<EncodeFn>(sym);
  const int ROWS = 512;
  __nv_bfloat16 *rowmajor, *tilemajor;
  cudaMalloc(&rowmajor, (size_t)ROWS * GH * 2);
  cudaMalloc(&tilemajor, (size_t)ROWS * GH * 2);
  cudaMemset(rowmajor, 0, (size_t)ROWS * GH * 2);
  cudaMemset(tilemajor, 0, (size_t)ROWS * GH * 2);
  int* counters;
  cudaMalloc(&counters, 64 * sizeof(int));
  long long* out;
  cudaMalloc(&out, 256 * sizeof(long long));
  struct Cfg { int X, Y, R, S, mode, fresh, KP, ksplit; };
  std::vector<Cfg> cfgs;
  cfgs.push_back({32, 4, 64, 9, 0, 1, 1, 1});
  cfgs.push_back({32, 4, 128, 9, 0, 1, 1, 1});      // bigger 2-D boxes: does the time follow the bytes?
  cfgs.push_back({32, 2, 256, 6, 0, 1, 1, 1});
  for (int kp : {1, 2, 4, 8}) cfgs.push_back({32, 4, 64, kp >= 4 ? 3 : 6, 2, 1, kp, 1});   // 3-D boxes: kp k-blocks per op
  for (int kp : {2, 4, 8}) cfgs.push_back({16, 8, 32, 4, 2, 1, kp, 1});
  cfgs.push_back({16, 4, 64, 3, 2, 1, 4, 1});
  for (int kp : {2, 4}) cfgs.push_back({32, 4, 64, 4, 2, 1, kp, 2});   // K split over CTA pairs: half the bytes per SM
  cfgs.push_back({32, 4, 64, 4, 2, 0, 4, 1});   // not rewritten between repetitions
  cfgs.push_back({32, 4, 64, 4, 2, 0, 4, 2});
  const int REP = 20;
  for (const Cfg& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)GH, (cuuint64_t)ROWS};
    cuuint64_t strides[1] = {(cuuint64_t)GH * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)c.R};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rowmajor, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    CUtensorMap tm3;
    {
      cuuint64_t d3[3] = {64, (cuuint64_t)ROWS, (cuuint64_t)(GH / 64)};
      cuuint64_t s3[2] = {(cuuint64_t)GH * 2, 128};
      cuuint32_t b3[3] = {64, (cuuint32_t)c.R, (cuuint32_t)c.KP};
      cuuint32_t e3[3] = {1, 1, 1};
      CUresult r3 = enc(&tm3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, rowmajor, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r3 != CUDA_SUCCESS) { printf("encode3 failed %d\n", (int)r3); return 1; }
    }
    const size_t smem = 1024 + (size_t)c.S * c.R * 128 * (c.mode == 2 ? c.KP : 1);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) {   // second run is the measurement
      cudaMemset(counters, 0, 64 * sizeof(int));
      void* args[] = {(void*)&tm, (void*)&tilemajor, (void*)&rowmajor, (void*)&c.R, (void*)&c.S, (void*)&REP, (void*)&c.mode, (void*)&c.fresh,
                      (void*)&counters, (void*)&out, (void*)&tm3, (void*)&c.KP, (void*)&c.ksplit};
      cudaError_t e = cudaLaunchCooperativeKernel((const void*)stream_kernel, dim3(c.X, c.Y), dim3(128), args, smem, 0);
      if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> h(c.X * c.Y);
    cudaMemcpy(h.data(), out, sizeof(long long) * c.X * c.Y, cudaMemcpyDeviceToHost);
    long long mx = 0, sum = 0;
    for (long long v : h) { mx = v > mx ? v : mx; sum += v; }
    const double bytes = (double)KB * c.R * 128 / c.ksplit;   // per CTA per repetition
    const double us_max = mx / 1e3 / REP, us_mean = sum / 1e3 / REP / h.size();
    printf("grid (%2d,%2d) ksplit=%d R=%3d S=%2d KP=%d mode=%s fresh=%d: per step %.2f us (max CTA) %.2f us (mean); per-SM %.1f GB/s, aggregate %.2f TB/s\n", c.X, c.Y, c.ksplit,
           c.R, c.S, c.KP, c.mode == 2 ? "3-D box" : (c.mode ? "bulk-contiguous" : "tensor-map-box "), c.fresh, us_max, us_mean, bytes / us_mean / 1e3,
           bytes * h.size() / us_max / 1e6);
  }
  return 0;
}
