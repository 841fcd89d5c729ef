// Device helpers of the multiresolution hash grid shared by hashgrid.cu and the fused density-field kernels.
#pragma once
#include "common.cuh"

struct GridParams {
  int n_levels, log2_T, mode;
  float scale[B2N_MAX_LEVELS];
  uint32_t resolution[B2N_MAX_LEVELS];
  uint32_t offset[B2N_MAX_LEVELS];
  uint32_t size[B2N_MAX_LEVELS];
  uint32_t hashed[B2N_MAX_LEVELS];
};

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
  bool xpair;        // the x-neighbours of every (y,z) corner pair sit in adjacent table rows {2m, 2m+1}
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
    // x enters the hash with prime 1: for an even floor coordinate, row(x+1) = row(x) ^ 1 — the two x-neighbours are
    // one aligned 16-byte pair (level offsets l*T are even), so 4 vector accesses replace 8
    c.xpair = ((xf & 1u) == 0u) && (xc == xf + 1u);
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
    c.xpair = false;
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


// Trilinear blend of the 8 corners in the reference's association (encodings.py:446-456: x pairs f03,f12,f56,f47 -> y
// pairs f0312,f4756 -> z) with the FMA contraction written out, so every gather kernel that uses it produces the same
// bits (the compiler otherwise picks the fused product per call site).
__device__ __forceinline__ float blend8(float f0, float f1, float f2, float f3, float f4, float f5, float f6, float f7,
                                        float ox, float oy, float oz) {
  const float rx = __fsub_rn(1.f, ox), ry = __fsub_rn(1.f, oy), rz = __fsub_rn(1.f, oz);
  const float f03 = __fmaf_rn(f0, ox, __fmul_rn(f3, rx)), f12 = __fmaf_rn(f1, ox, __fmul_rn(f2, rx));
  const float f56 = __fmaf_rn(f5, ox, __fmul_rn(f6, rx)), f47 = __fmaf_rn(f4, ox, __fmul_rn(f7, rx));
  const float f0312 = __fmaf_rn(f03, oy, __fmul_rn(f12, ry)), f4756 = __fmaf_rn(f47, oy, __fmul_rn(f56, ry));
  return __fmaf_rn(f0312, oz, __fmul_rn(f4756, rz));
}

// d(blend8)/d(ox, oy, oz) times g, accumulated into (ax, ay, az): the position gradient of one feature of one level in
// grid units (multiply by the level's scale for unit-cube units).
__device__ __forceinline__ void blend8_dpos(float f0, float f1, float f2, float f3, float f4, float f5, float f6, float f7,
                                            float ox, float oy, float oz, float g, float& ax, float& ay, float& az) {
  const float rx = 1.f - ox, ry = 1.f - oy, rz = 1.f - oz;
  const float d03 = f0 - f3, d12 = f1 - f2, d56 = f5 - f6, d47 = f4 - f7;
  const float f03 = f0 * ox + f3 * rx, f12 = f1 * ox + f2 * rx, f56 = f5 * ox + f6 * rx, f47 = f4 * ox + f7 * rx;
  const float f0312 = f03 * oy + f12 * ry, f4756 = f47 * oy + f56 * ry;
  ax += g * ((d03 * oy + d12 * ry) * oz + (d47 * oy + d56 * ry) * rz);
  ay += g * ((f03 - f12) * oz + (f47 - f56) * rz);
  az += g * (f0312 - f4756);
}

// the 8 corner feature rows of one (point, level), F = 2: 4 x 128-bit loads when the x-neighbours are paired
__device__ __forceinline__ void gather_corners2(const float* __restrict__ table, const Corners& c, Vec<2> (&f)[8]) {
  if (c.xpair) {
    // (ceil-x corner, floor-x corner) index pairs of the reference's corner order
    constexpr int pc[4] = {0, 1, 4, 5}, pf[4] = {3, 2, 7, 6};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t rf = c.row[pf[q]];
      const float4 v = __ldg(reinterpret_cast<const float4*>(table + (size_t)(rf & ~1u) * 2));
      const bool f_hi = (rf & 1u) != 0u;  // which half of the pair is the floor-x corner
      f[pf[q]].v[0] = f_hi ? v.z : v.x, f[pf[q]].v[1] = f_hi ? v.w : v.y;
      f[pc[q]].v[0] = f_hi ? v.x : v.z, f[pc[q]].v[1] = f_hi ? v.y : v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = ldg_row<2>(table, c.row[k]);
  }
}

// scatter counterpart: g[k][0..1] is added to row k; paired x-neighbours go out as one 128-bit RED
__device__ __forceinline__ void scatter_corners2(float* table, const uint32_t (&row)[8], bool xpair, const float (&g)[8][2]) {
  if (xpair) {
    constexpr int pc[4] = {0, 1, 4, 5}, pf[4] = {3, 2, 7, 6};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t rf = row[pf[q]];
      const bool f_hi = (rf & 1u) != 0u;
      float* p = table + (size_t)(rf & ~1u) * 2;
      const float a0 = f_hi ? g[pc[q]][0] : g[pf[q]][0], a1 = f_hi ? g[pc[q]][1] : g[pf[q]][1];
      const float b0 = f_hi ? g[pf[q]][0] : g[pc[q]][0], b1 = f_hi ? g[pf[q]][1] : g[pc[q]][1];
      asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a0), "f"(a1), "f"(b0), "f"(b1)
                   : "memory");
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) red_row<2>(table, row[k], g[k]);
  }
}

static inline int fill_params(const B2nGrid* g, GridParams& gp) {
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
