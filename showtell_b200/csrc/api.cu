// Library-wide state: version, thread-local last error, device facts, step-table validation.
#include <cstring>

#include "common.cuh"

namespace st {

int g_pdl = 1;
static thread_local char g_err[512] = "";
static long long g_launches = 0;

void note_launch(int n) { __atomic_fetch_add(&g_launches, (long long)n, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int make_step_table(StepTable& tab, int nsteps, const int* bs_host) {
  ST_REQUIRE(bs_host != nullptr, ST_ERR_NULL, "batch_sizes_host is NULL");
  ST_REQUIRE(nsteps >= 1 && nsteps <= ST_MAX_STEPS, ST_ERR_BAD_SHAPE,
             "nsteps=%d outside [1, %d]", nsteps, (int)ST_MAX_STEPS);
  int off = 0;
  for (int t = 0; t < nsteps; ++t) {
    ST_REQUIRE(bs_host[t] >= 1, ST_ERR_BAD_SHAPE, "batch_sizes[%d]=%d must be >= 1", t, bs_host[t]);
    ST_REQUIRE(t == 0 || bs_host[t] <= bs_host[t - 1], ST_ERR_UNSORTED,
               "batch_sizes must be non-increasing (lengths sorted descending): bs[%d]=%d > bs[%d]=%d",
               t, bs_host[t], t - 1, bs_host[t - 1]);
    tab.bs[t] = bs_host[t];
    tab.off[t] = off;
    off += bs_host[t];
  }
  for (int t = nsteps; t <= ST_MAX_STEPS; ++t) {
    tab.bs[t] = 0;
    tab.off[t] = off;
  }
  tab.nsteps = nsteps;
  return ST_OK;
}

}  // namespace st

extern "C" {

int st_version(void) { return 100; }

const char* st_last_error(void) { return st::g_err; }

int st_debug_set_pdl(int on) {
  st::g_pdl = on ? 1 : 0;
  return ST_OK;
}

int64_t st_launch_count(void) { return (int64_t)__atomic_load_n(&st::g_launches, __ATOMIC_RELAXED); }

int st_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_optin_bytes) {
  int dev = 0;
  ST_CUDA_TRY(cudaGetDevice(&dev));
  int v = 0;
  if (sm_count) {
    ST_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    *sm_count = v;
  }
  if (cc_major) {
    ST_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    *cc_major = v;
  }
  if (cc_minor) {
    ST_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    *cc_minor = v;
  }
  if (smem_optin_bytes) {
    ST_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    *smem_optin_bytes = v;
  }
  return ST_OK;
}

}  // extern "C"
