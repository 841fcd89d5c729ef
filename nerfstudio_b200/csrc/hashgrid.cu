// K1/K2 — multiresolution hash-grid gather (forward) and scatter-add (backward) for sm_100a.
//
// Arithmetic follows the reference's torch path bit-for-bit on the index side
// (nerfstudio/field_components/encodings.py:398-458): scaled = x*scale evaluated as a separately rounded
// fp32 multiply, ceil/floor corners, XOR of coordinate*prime products, mod T (T = 2^k => low bits of the
// uint32 wrap equal the reference's int64 result), + l*T.  The trilinear blend uses the reference's pairing
// (f03,f12,f56,f47 -> f0312,f4756 -> out) so values agree to an ulp or two (FMA contraction only).
//
// Mapping: one thread per sample point, consecutive threads = consecutive samples along a ray, so a warp's
// 8 corner loads at a coarse level fall into a handful of 32 B sectors; blockIdx.y selects a group of levels
// (tunable "hash_levels_per_block") so the working set that is live in L1/L2 at any time is a few levels.
// Loads are vector (F*4 bytes) through the read-only path; backward uses vector RED (red.global.add.v2/v4.f32).
#include <string.h>

#include "common.cuh"

struct GridParams {
  int n_levels, log2_T, mode;
  float scale[B2N_MAX_LEVELS];
  uint32_t resolution[B2N_MAX_LEVELS];
  uint32_t offset[B2N_MAX_LEVELS];
  uint32_t size[B2N_MAX_LEVELS];
  uint32_t hashed[B2N_MAX_LEVELS];
};

static int g_levels_per_block_fwd = 0;  // 0 = all levels in one thread
static int g_levels_per_block_bwd = 1;

template <int F>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
};
template <>
struct alignas(8) Vec<2> {
  float v[2];
};
template <>
struct alignas(16) Vec<4> {
  float v[4];
};
template <>
struct alignas(16) Vec<8> {
  float v[8];
};

template <int F>
__device__ __forceinline__ Vec<F> ldg_row(const float* __restrict__ table, uint32_t row) {
  Vec<F> r;
  const float* p = table + (size_t)row * F;
  if constexpr (F == 1) {
    r.v[0] = __ldg(p);
  } else if constexpr (F == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x, r.v[1] = t.y;
  } else if constexpr (F == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x, r.v[1] = t.y, r.v[2] = t.z, r.v[3] = t.w;
  } else {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w;
    r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
  }
  return r;
}

template <int F>
__device__ __forceinline__ void red_row(float* table, uint32_t row, const float* g) {
  float* p = table + (size_t)row * F;
  if constexpr (F == 1) {
    atomicAdd(p, g[0]);
  } else if constexpr (F == 2) {
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(g[0]), "f"(g[1]) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < F; i += 4)
      asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + i), "f"(g[i]), "f"(g[i + 1]),
                   "f"(g[i + 2]), "f"(g[i + 3])
                   : "memory");
  }
}

// corner rows + lerp weights of one (point, level)
struct Corners {
  uint32_t row[8];
  float ox, oy, oz;  // torch: weight of the ceil corner; tcnn: weight of the +1 corner
};

template <int MODE>
__device__ __forceinline__ Corners corners_of(const GridParams& gp, int l, float x, float y, float z) {
  Corners c;
  if constexpr (MODE == B2N_GRID_TORCH) {
    const float s = gp.scale[l];
    const float sx = mul_rn(x, s), sy = mul_rn(y, s), sz = mul_rn(z, s);
    const float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    const uint32_t xf = (uint32_t)(int)fx, yf = (uint32_t)(int)fy * 2654435761u, zf = (uint32_t)(int)fz * 805459861u;
    const uint32_t xc = (uint32_t)(int)ceilf(sx), yc = (uint32_t)(int)ceilf(sy) * 2654435761u,
                   zc = (uint32_t)(int)ceilf(sz) * 805459861u;
    c.ox = sx - fx, c.oy = sy - fy, c.oz = sz - fz;
    const uint32_t mask = (1u << gp.log2_T) - 1u, off = gp.offset[l];
    c.row[0] = ((xc ^ yc ^ zc) & mask) + off;
    c.row[1] = ((xc ^ yf ^ zc) & mask) + off;
    c.row[2] = ((xf ^ yf ^ zc) & mask) + off;
    c.row[3] = ((xf ^ yc ^ zc) & mask) + off;
    c.row[4] = ((xc ^ yc ^ zf) & mask) + off;
    c.row[5] = ((xc ^ yf ^ zf) & mask) + off;
    c.row[6] = ((xf ^ yf ^ zf) & mask) + off;
    c.row[7] = ((xf ^ yc ^ zf) & mask) + off;
  } else {
    const float s = gp.scale[l];
    const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    c.ox = px - fx, c.oy = py - fy, c.oz = pz - fz;
    const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    const uint32_t size = gp.size[l], off = gp.offset[l], res = gp.resolution[l];
    const bool hashed = gp.hashed[l] != 0;
    // same corner order as torch mode with c = +1 corner, f = base corner
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int dx = (k == 0 || k == 1 || k == 4 || k == 5), dy = (k == 0 || k == 3 || k == 4 || k == 7), dz = (k < 4);
      const uint32_t cx = gx + dx, cy = gy + dy, cz = gz + dz;
      uint32_t idx;
      if (hashed)
        idx = (cx ^ (cy * 2654435761u) ^ (cz * 805459861u)) % size;
      else
        idx = (cx + cy * res + cz * res * res) % size;
      c.row[k] = idx + off;
    }
  }
  return c;
}

template <int F, int MODE>
__global__ void __launch_bounds__(256) hashgrid_fwd_kernel(const __grid_constant__ GridParams gp,
                                                           const float* __restrict__ x,
                                                           const float* __restrict__ table, int64_t n,
                                                           float* __restrict__ y, int64_t* __restrict__ idx_out,
                                                           int levels_per_block) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float px = __ldg(x + 3 * i), py = __ldg(x + 3 * i + 1), pz = __ldg(x + 3 * i + 2);
  const int l0 = blockIdx.y * levels_per_block;
  const int l1 = min(gp.n_levels, l0 + levels_per_block);
  float* yrow = y + i * (int64_t)(gp.n_levels * F);
#pragma unroll 2
  for (int l = l0; l < l1; ++l) {
    const Corners c = corners_of<MODE>(gp, l, px, py, pz);
    Vec<F> f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = ldg_row<F>(table, c.row[k]);
    if (idx_out != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) idx_out[(i * gp.n_levels + l) * 8 + k] = (int64_t)c.row[k];
    }
    const float ox = c.ox, oy = c.oy, oz = c.oz;
    const float rx = 1.f - ox, ry = 1.f - oy, rz = 1.f - oz;
    float out[F];
#pragma unroll
    for (int j = 0; j < F; ++j) {
      const float f03 = f[0].v[j] * ox + f[3].v[j] * rx;
      const float f12 = f[1].v[j] * ox + f[2].v[j] * rx;
      const float f56 = f[5].v[j] * ox + f[6].v[j] * rx;
      const float f47 = f[4].v[j] * ox + f[7].v[j] * rx;
      const float f0312 = f03 * oy + f12 * ry;
      const float f4756 = f47 * oy + f56 * ry;
      out[j] = f0312 * oz + f4756 * rz;
    }
    if constexpr (F == 2) {
      *reinterpret_cast<float2*>(yrow + l * 2) = make_float2(out[0], out[1]);
    } else if constexpr (F == 4) {
      *reinterpret_cast<float4*>(yrow + l * 4) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
      for (int j = 0; j < F; ++j) yrow[l * F + j] = out[j];
    }
  }
}

template <int F, int MODE, bool WITH_DX>
__global__ void __launch_bounds__(256) hashgrid_bwd_kernel(const __grid_constant__ GridParams gp,
                                                           const float* __restrict__ x,
                                                           const float* __restrict__ table,
                                                           const float* __restrict__ dy, int64_t n,
                                                           float* __restrict__ dtable, float* __restrict__ dx,
                                                           int levels_per_block) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float px = __ldg(x + 3 * i), py = __ldg(x + 3 * i + 1), pz = __ldg(x + 3 * i + 2);
  const int l0 = blockIdx.y * levels_per_block;
  const int l1 = min(gp.n_levels, l0 + levels_per_block);
  const float* grow = dy + i * (int64_t)(gp.n_levels * F);
  float gx = 0.f, gy = 0.f, gz = 0.f;
  for (int l = l0; l < l1; ++l) {
    const Corners c = corners_of<MODE>(gp, l, px, py, pz);
    float g[F];
    bool any = false;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      g[j] = __ldg(grow + l * F + j);
      any |= (g[j] != 0.f);
    }
    const float ox = c.ox, oy = c.oy, oz = c.oz;
    const float rx = 1.f - ox, ry = 1.f - oy, rz = 1.f - oz;
    if (any) {
      // weight of corner k in the reference's blend
      const float w[8] = {ox * oy * oz, ox * ry * oz, rx * ry * oz, rx * oy * oz,
                          ox * oy * rz, ox * ry * rz, rx * ry * rz, rx * oy * rz};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float t[F];
#pragma unroll
        for (int j = 0; j < F; ++j) t[j] = w[k] * g[j];
        red_row<F>(dtable, c.row[k], t);
      }
    }
    if constexpr (WITH_DX) {
      Vec<F> f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = ldg_row<F>(table, c.row[k]);
      float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
      for (int j = 0; j < F; ++j) {
        const float d03 = f[0].v[j] - f[3].v[j], d12 = f[1].v[j] - f[2].v[j];
        const float d56 = f[5].v[j] - f[6].v[j], d47 = f[4].v[j] - f[7].v[j];
        const float f03 = f[0].v[j] * ox + f[3].v[j] * rx, f12 = f[1].v[j] * ox + f[2].v[j] * rx;
        const float f56 = f[5].v[j] * ox + f[6].v[j] * rx, f47 = f[4].v[j] * ox + f[7].v[j] * rx;
        const float f0312 = f03 * oy + f12 * ry, f4756 = f47 * oy + f56 * ry;
        ax += g[j] * ((d03 * oy + d12 * ry) * oz + (d47 * oy + d56 * ry) * rz);
        ay += g[j] * ((f03 - f12) * oz + (f47 - f56) * rz);
        az += g[j] * (f0312 - f4756);
      }
      const float s = gp.scale[l];
      gx += ax * s, gy += ay * s, gz += az * s;
    }
  }
  if constexpr (WITH_DX) {
    if (gridDim.y == 1) {
      dx[3 * i] = gx, dx[3 * i + 1] = gy, dx[3 * i + 2] = gz;
    } else {
      atomicAdd(dx + 3 * i, gx), atomicAdd(dx + 3 * i + 1, gy), atomicAdd(dx + 3 * i + 2, gz);
    }
  }
}

static int fill_params(const B2nGrid* g, GridParams& gp) {
  if (g->n_levels < 1 || g->n_levels > B2N_MAX_LEVELS) return -1;
  if (g->log2_hashmap_size < 1 || g->log2_hashmap_size > 31) return -1;
  gp.n_levels = g->n_levels, gp.log2_T = g->log2_hashmap_size, gp.mode = g->mode;
  for (int l = 0; l < g->n_levels; ++l) {
    gp.scale[l] = g->scale[l];
    gp.resolution[l] = g->resolution[l];
    gp.offset[l] = g->offset[l];
    gp.size[l] = g->size[l] ? g->size[l] : 1u;
    gp.hashed[l] = g->hashed[l];
  }
  return 0;
}

template <int F>
static void launch_fwd(const GridParams& gp, const float* x, const float* table, int64_t n, float* y, int64_t* idx,
                       cudaStream_t st) {
  int lpb = g_levels_per_block_fwd > 0 ? g_levels_per_block_fwd : gp.n_levels;
  dim3 grid((unsigned)div_up(n, 256), (unsigned)div_up(gp.n_levels, lpb));
  if (gp.mode == B2N_GRID_TORCH)
    hashgrid_fwd_kernel<F, B2N_GRID_TORCH><<<grid, 256, 0, st>>>(gp, x, table, n, y, idx, lpb);
  else
    hashgrid_fwd_kernel<F, B2N_GRID_TCNN><<<grid, 256, 0, st>>>(gp, x, table, n, y, idx, lpb);
}

template <int F, bool DX>
static void launch_bwd(const GridParams& gp, const float* x, const float* table, const float* dy, int64_t n,
                       float* dtable, float* dx, cudaStream_t st) {
  int lpb = g_levels_per_block_bwd > 0 ? g_levels_per_block_bwd : gp.n_levels;
  dim3 grid((unsigned)div_up(n, 256), (unsigned)div_up(gp.n_levels, lpb));
  if (DX && grid.y > 1) cudaMemsetAsync(dx, 0, sizeof(float) * 3 * n, st);
  if (gp.mode == B2N_GRID_TORCH)
    hashgrid_bwd_kernel<F, B2N_GRID_TORCH, DX><<<grid, 256, 0, st>>>(gp, x, table, dy, n, dtable, dx, lpb);
  else
    hashgrid_bwd_kernel<F, B2N_GRID_TCNN, DX><<<grid, 256, 0, st>>>(gp, x, table, dy, n, dtable, dx, lpb);
}

extern "C" int b2n_hashgrid_fwd(const B2nGrid* grid_host, const float* x, const float* table, int64_t n, float* y,
                                int64_t* idx_out, void* stream) {
  if (n == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(grid_host && x && table && y, "null pointer");
  B2N_REQUIRE(n >= 0, "negative n");
  GridParams gp;
  B2N_REQUIRE(fill_params(grid_host, gp) == 0, "bad grid description");
  if (n == 0) return B2N_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (grid_host->n_features) {
    case 1: launch_fwd<1>(gp, x, table, n, y, idx_out, st); break;
    case 2: launch_fwd<2>(gp, x, table, n, y, idx_out, st); break;
    case 4: launch_fwd<4>(gp, x, table, n, y, idx_out, st); break;
    case 8: launch_fwd<8>(gp, x, table, n, y, idx_out, st); break;
    default: B2N_UNSUPPORTED(true, "n_features must be 1, 2, 4 or 8");
  }
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_hashgrid_bwd(const B2nGrid* grid_host, const float* x, const float* table, const float* dy,
                                int64_t n, float* dtable, float* dx, void* stream) {
  if (n == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(grid_host && x && dy && dtable, "null pointer");
  B2N_REQUIRE(dx == nullptr || table != nullptr, "dx needs the table");
  GridParams gp;
  B2N_REQUIRE(fill_params(grid_host, gp) == 0, "bad grid description");
  if (n == 0) return B2N_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define B2N_BWD_CASE(F)                                                  \
  case F:                                                                \
    if (dx)                                                              \
      launch_bwd<F, true>(gp, x, table, dy, n, dtable, dx, st);          \
    else                                                                 \
      launch_bwd<F, false>(gp, x, table, dy, n, dtable, dx, st);         \
    break;
  switch (grid_host->n_features) {
    B2N_BWD_CASE(1)
    B2N_BWD_CASE(2)
    B2N_BWD_CASE(4)
    B2N_BWD_CASE(8)
    default: B2N_UNSUPPORTED(true, "n_features must be 1, 2, 4 or 8");
  }
#undef B2N_BWD_CASE
  B2N_LAUNCH_CHECK();
}

int b2n_tune_hashgrid(const char* key, int value) {
  if (!strcmp(key, "hash_levels_per_block_fwd")) { g_levels_per_block_fwd = value; return 1; }
  if (!strcmp(key, "hash_levels_per_block_bwd")) { g_levels_per_block_bwd = value; return 1; }
  return 0;
}
