// Shared helpers for libb200nerf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "b200nerf.h"

void b2n_set_error(const char* fmt, ...);

#define B2N_REQUIRE(cond, msg)                       \
  do {                                               \
    if (!(cond)) {                                   \
      b2n_set_error("%s: %s", __func__, msg);        \
      return B2N_E_ARG;                              \
    }                                                \
  } while (0)

#define B2N_UNSUPPORTED(cond, msg)                   \
  do {                                               \
    if (cond) {                                      \
      b2n_set_error("%s: %s", __func__, msg);        \
      return B2N_E_UNSUPPORTED;                      \
    }                                                \
  } while (0)

#define B2N_LAUNCH_CHECK()                                                   \
  do {                                                                       \
    cudaError_t e_ = cudaGetLastError();                                     \
    if (e_ != cudaSuccess) {                                                 \
      b2n_set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e_)); \
      return (int)e_;                                                        \
    }                                                                        \
    return B2N_OK;                                                           \
  } while (0)

static inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }
int b2n_sm_count();

// round-to-nearest, never contracted into FMA: used wherever the reference's result feeds an
// integer decision (cell index, searchsorted) so the index work stays bit-exact with torch's
// separately-rounded elementwise ops.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ float nan_to_num(float v) {
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive scan across the 32 lanes
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ double warp_scan_incl_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
