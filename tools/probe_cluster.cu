// Dev probe: which thread-block-cluster sizes can this GPU co-schedule with ~200 KB of dynamic shared
// memory per CTA, and does a DSMEM store + cluster barrier round trip work at that size?
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void probe(int* out, int steps) {
  extern __shared__ int sm[];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned r = cl.block_rank(), n = cl.num_blocks();
  long long t0 = 0, t1 = 0;
  cl.sync();
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int s = 0; s < steps; ++s) {
    if (threadIdx.x < n) {
      int* remote = cl.map_shared_rank(sm, threadIdx.x);
      remote[r] = s + (int)r;   // every CTA writes its word into every peer
    }
    cl.sync();
  }
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  int ok = 1;
  if (threadIdx.x == 0) {
    for (unsigned i = 0; i < n; ++i) ok &= (sm[i] == steps - 1 + (int)i);
    out[blockIdx.x * 2] = ok;
    out[blockIdx.x * 2 + 1] = (int)(t1 - t0);
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  printf("SMs %d\n", sms);
  for (int cs : {2, 4, 8, 16}) {
    for (size_t smem : {(size_t)64 << 10, (size_t)200 << 10}) {
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (cs > 8) cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 8);
      cfg.blockDim = dim3(320);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclusters = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, probe, &cfg);
      printf("cluster %2d smem %3zu KB: max active clusters %d (%s)\n", cs, smem >> 10, nclusters, cudaGetErrorString(e));
      cudaGetLastError();
      if (e == cudaSuccess && nclusters > 0) {
        int* out;
        const int nblk = cs * (nclusters < 8 ? nclusters : 8);
        cfg.gridDim = dim3(nblk);
        cudaMalloc(&out, sizeof(int) * 2 * nblk);
        cudaMemset(out, 0, sizeof(int) * 2 * nblk);
        int steps = 100;
        e = cudaLaunchKernelEx(&cfg, probe, out, steps);
        cudaError_t e2 = cudaDeviceSynchronize();
        int h[2 * 256];
        cudaMemcpy(h, out, sizeof(int) * 2 * nblk, cudaMemcpyDeviceToHost);
        int ok = 1;
        for (int i = 0; i < nblk; ++i) ok &= h[2 * i];
        printf("   launch %s / %s: %d CTAs, all ok=%d, %.0f ns per {DSMEM broadcast + cluster.sync}\n",
               cudaGetErrorString(e), cudaGetErrorString(e2), nblk, ok, h[1] / (double)steps);
        cudaFree(out);
      }
    }
  }
  return 0;
}
