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
__device__ __forceinline__ int64_t div_up_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }
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

// torch's CPU `sum` over one contiguous fp32 row, reproduced bit for bit (ATen/native/cpu/SumKernel.cpp:
// cascade_sum -> vectorized_inner_sum with 8-lane vectors, ilp factor 4, 4 cascade levels).  The reference's
// PDFSampler normalises by `torch.sum(weights, -1)` (model_components/ray_samplers.py:306) and the golden
// searchsorted indices are only reproducible if that denominator has the same bits.  The 8 lanes x 4 interleaved
// rows are exactly the 32 lanes of a warp: lane L owns elements i*32+L.  x may be shared or global memory;
// every lane returns the sum.  (Checked against torch.sum for S in 5..5000 by tests/test_oracle_golden.py through
// the oracle's restatement of the same order.)
__device__ __forceinline__ int ceil_log2_i(int x) { return x <= 1 ? 0 : 32 - __clz(x - 1); }

__device__ __forceinline__ float torch_cascade_lane(const float* x, int size, int stride, int lane_off) {
  // multi_row_sum for one accumulator column: elements x[i*stride + lane_off], i in [0,size)
  const int level_power = max(4, ceil_log2_i(size) / 4);
  const int level_step = 1 << level_power, level_mask = level_step - 1;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int i = 0;
  while (i + level_step <= size) {
    for (int j = 0; j < level_step; ++j, ++i) a0 = __fadd_rn(a0, x[i * stride + lane_off]);
    a1 = __fadd_rn(a1, a0), a0 = 0.f;
    if ((i & (level_mask << level_power)) == 0) {
      a2 = __fadd_rn(a2, a1), a1 = 0.f;
      if ((i & (level_mask << (2 * level_power))) == 0) a3 = __fadd_rn(a3, a2), a2 = 0.f;
    }
  }
  for (; i < size; ++i) a0 = __fadd_rn(a0, x[i * stride + lane_off]);
  a0 = __fadd_rn(a0, a1), a0 = __fadd_rn(a0, a2), a0 = __fadd_rn(a0, a3);
  return a0;
}

__device__ __forceinline__ float torch_cpu_row_sum(const float* x, int S, int lane) {
  const unsigned full = 0xffffffffu;
  if (S >= 8) {
    const int vec_size = S >> 3, size_ilp = vec_size >> 2;
    float acc = torch_cascade_lane(x, size_ilp, 32, lane);
    for (int v = size_ilp * 4; v < vec_size; ++v)
      if (lane < 8) acc = __fadd_rn(acc, x[v * 8 + lane]);
    const float p1 = __shfl_sync(full, acc, (lane & 7) + 8), p2 = __shfl_sync(full, acc, (lane & 7) + 16),
                p3 = __shfl_sync(full, acc, (lane & 7) + 24);
    acc = __fadd_rn(__fadd_rn(__fadd_rn(acc, p1), p2), p3);  // valid on lanes 0..7
    float f = 0.f;
    for (int k = vec_size * 8; k < S; ++k) f = __fadd_rn(f, x[k]);
#pragma unroll
    for (int l = 0; l < 8; ++l) f = __fadd_rn(f, __shfl_sync(full, acc, l));
    return f;
  }
  // rows shorter than one vector: scalar_inner_sum (4 interleaved scalar accumulators)
  const int size_ilp = S >> 2;
  float acc = lane < 4 ? torch_cascade_lane(x, size_ilp, 4, lane) : 0.f;
  for (int k = size_ilp * 4; k < S; ++k)
    if (lane == 0) acc = __fadd_rn(acc, x[k]);
  float f = __shfl_sync(full, acc, 0);
#pragma unroll
  for (int l = 1; l < 4; ++l) f = __fadd_rn(f, __shfl_sync(full, acc, l));
  return f;
}
