// Error string, version and device query of the C-ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void b2n_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int b2n_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

extern "C" const char* b2n_version(void) { return "b200nerf 0.1 (sm_100a)"; }
extern "C" const char* b2n_last_error(void) { return g_err; }

extern "C" int b2n_device_info(int32_t* out4_host) {
  B2N_REQUIRE(out4_host != nullptr, "null output");
  int dev = 0, v = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    b2n_set_error("b2n_device_info: %s", cudaGetErrorString(e));
    return (int)e;
  }
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  out4_host[0] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev);
  out4_host[1] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev);
  out4_host[2] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  out4_host[3] = v;
  return B2N_OK;
}

int b2n_tune_hashgrid(const char* key, int value);
int b2n_tune_mlp_tc(const char* key, int value);
int b2n_tune_density_fused(const char* key, int value);
int b2n_tune_packed(const char* key, int value);

// runtime tuning knobs (kernel launch geometry); returns 1 if the key was recognised
extern "C" int b2n_tune(const char* key, int value) {
  if (!key) return 0;
  return b2n_tune_hashgrid(key, value) || b2n_tune_mlp_tc(key, value) || b2n_tune_packed(key, value) || b2n_tune_density_fused(key, value);
}
