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

#include "hashgrid.cuh"

static int g_levels_per_block_fwd = 4;  // levels looped per thread (blockIdx.y picks the group); 0 = all
static int g_levels_per_block_bwd = 1;
static int g_fwd_pair = 1;  // F = 2: two lanes per point (hashgrid_fwd_pair_kernel / hashgrid_dx_pair_kernel); 0 = one thread per point
static int g_bwd_chunk = 8;  // samples per thread of the run-length backward kernel (0 = one sample per thread)

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
    if constexpr (F == 2) {
      gather_corners2(table, c, f);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = ldg_row<F>(table, c.row[k]);
    }
    if (idx_out != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) idx_out[(i * gp.n_levels + l) * 8 + k] = (int64_t)c.row[k];
    }
    float out[F];
#pragma unroll
    for (int j = 0; j < F; ++j)
      out[j] = blend8(f[0].v[j], f[1].v[j], f[2].v[j], f[3].v[j], f[4].v[j], f[5].v[j], f[6].v[j], f[7].v[j], c.ox, c.oy, c.oz);
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

// Lane-pair variant of the F = 2 gather (default).  The L1TEX tag stage handles ONE 128-byte line per cycle per SM and
// a warp-wide load costs as many cycles as it touches distinct lines (B300_MICROARCH: nL_j lines -> nL_j wavefronts).
// With one thread per point every corner load of a fine level touches 32 lines, i.e. 8 line-cycles per (point, level)
// — or 4 when the two x-neighbours happen to be an aligned 16-byte pair (x even).  But the x-neighbours ALWAYS sit in
// the same 128-byte line 15 times out of 16: x enters the hash with prime 1, row = (x ^ h(y,z)) & mask, so all x of an
// aligned block of 16 map into one aligned block of 16 rows = 128 B.  Here two adjacent lanes own one point — the even
// lane the floor-x corners, the odd lane the ceil-x corners — so each of the 4 (y,z) loads is ONE instruction in which
// both x-neighbours are requested together: 16 points x 1 line = 16 lines per instruction, 4 line-cycles per (point,
// level) regardless of the parity of x.  The pair then exchanges the feature it does not blend (lane parity = feature
// index), so every output is still formed with the reference's association (x, then y, then z) from all 8 corners —
// bit-identical to the one-thread-per-point kernel.  Outputs of the block's level group are packed into one 16-byte
// store per lane (pairs write 32 contiguous bytes).
template <int MODE>
__global__ void __launch_bounds__(256) hashgrid_fwd_pair_kernel(const __grid_constant__ GridParams gp,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ table, int64_t n,
                                                                float* __restrict__ y, int levels_per_block) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t >> 1;
  const bool live = i < n;
  const int par = (int)(t & 1);  // 0: floor-x corners / feature 0, 1: ceil-x corners / feature 1
  const int64_t ic = live ? i : n - 1;
  const float px = __ldg(x + 3 * ic), py = __ldg(x + 3 * ic + 1), pz = __ldg(x + 3 * ic + 2);
  const int l0 = blockIdx.y * levels_per_block;
  const int l1 = min(gp.n_levels, l0 + levels_per_block);
  float* yrow = y + ic * (int64_t)(gp.n_levels * 2);
  // reference corner order k: x-corner c for k in {0,1,4,5}, f for {3,2,7,6}; (y,z) combos in the order
  // q = 0:(yc,zc) 1:(yf,zc) 2:(yc,zf) 3:(yf,zf)  ->  ceil-x rows {0,1,4,5}, floor-x rows {3,2,7,6}
  float pend[2];
  int n_pend = 0;
  for (int l = l0; l < l1; ++l) {
    const Corners c = corners_of<MODE>(gp, l, px, py, pz);
    uint32_t r4[4];
    r4[0] = par ? c.row[0] : c.row[3], r4[1] = par ? c.row[1] : c.row[2];
    r4[2] = par ? c.row[4] : c.row[7], r4[3] = par ? c.row[5] : c.row[6];
    float2 f4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) f4[q] = __ldg(reinterpret_cast<const float2*>(table) + r4[q]);
    // I blend feature j = par and need the partner's x-corner values of that feature; the partner needs mine of 1 - par
    float mine[4], theirs[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float send = par ? f4[q].x : f4[q].y;  // partner's feature index is 1 - par
      mine[q] = par ? f4[q].y : f4[q].x;
      theirs[q] = __shfl_xor_sync(0xffffffffu, send, 1);
    }
    // fc[q] / ff[q]: ceil-x / floor-x corner of (y,z) combo q, feature par
    float fc[4], ff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) fc[q] = par ? mine[q] : theirs[q], ff[q] = par ? theirs[q] : mine[q];
    // corners: 0 (c,c,c) 1 (c,f,c) 2 (f,f,c) 3 (f,c,c) 4 (c,c,f) 5 (c,f,f) 6 (f,f,f) 7 (f,c,f)
    const float out = blend8(fc[0], fc[1], ff[1], ff[0], fc[2], fc[3], ff[3], ff[2], c.ox, c.oy, c.oz);  // feature `par`
    // pack: two levels -> one float4 per lane.  Even lane stores level pair's first level (o0, o1), odd lane the second.
    pend[n_pend++] = out;
    if (n_pend == 2 || l == l1 - 1) {
      if (n_pend == 2) {
        // even lane keeps (o0[la], o1[la]), odd lane (o0[lb], o1[lb]):  exchange o_par[other level]
        const float send = par ? pend[0] : pend[1];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        const float a = par ? recv : pend[0], b = par ? pend[1] : recv;  // (feature 0, feature 1) of my level
        if (live) *reinterpret_cast<float2*>(yrow + (l - 1 + par) * 2) = make_float2(a, b);
      } else {
        const float recv = __shfl_xor_sync(0xffffffffu, pend[0], 1);
        if (live && par == 0) *reinterpret_cast<float2*>(yrow + l * 2) = make_float2(pend[0], recv);
      }
      n_pend = 0;
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

// Position gradient only (no table scatter): dx[i] = sum over levels and features of dy * d(encoding)/d(x).  Used by the
// captured step when the camera optimiser trains (the table gradient goes through the run-length scatter kernel, which
// has no use for the corner values; this kernel has no use for atomics).  One thread per sample, all levels in-thread.
template <int F, int MODE>
__global__ void __launch_bounds__(256) hashgrid_dx_kernel(const __grid_constant__ GridParams gp, const float* __restrict__ x,
                                                          const float* __restrict__ table, const float* __restrict__ dy,
                                                          int64_t n, float* __restrict__ dx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float px = __ldg(x + 3 * i), py = __ldg(x + 3 * i + 1), pz = __ldg(x + 3 * i + 2);
  const float* grow = dy + i * (int64_t)(gp.n_levels * F);
  float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll 2
  for (int l = 0; l < gp.n_levels; ++l) {
    float g[F];
    bool any = false;
#pragma unroll
    for (int j = 0; j < F; ++j) g[j] = __ldg(grow + l * F + j), any |= g[j] != 0.f;
    if (!any) continue;
    const Corners c = corners_of<MODE>(gp, l, px, py, pz);
    Vec<F> f[8];
    if constexpr (F == 2) {
      gather_corners2(table, c, f);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = ldg_row<F>(table, c.row[k]);
    }
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
    for (int j = 0; j < F; ++j)
      blend8_dpos(f[0].v[j], f[1].v[j], f[2].v[j], f[3].v[j], f[4].v[j], f[5].v[j], f[6].v[j], f[7].v[j], c.ox, c.oy, c.oz, g[j],
                  ax, ay, az);
    const float sc = gp.scale[l];
    gx = fmaf(ax, sc, gx), gy = fmaf(ay, sc, gy), gz = fmaf(az, sc, gz);
  }
  dx[3 * i] = gx, dx[3 * i + 1] = gy, dx[3 * i + 2] = gz;
}

// Lane-pair variant of the position gradient for F = 2 (see hashgrid_fwd_pair_kernel: 4 instead of 6-8 L1 line cycles per
// (point, level)): lane parity = x-corner for the loads and = feature index for the derivative; the pair sums its two
// features' contributions with one shuffle per component.  All levels in-thread: no atomics.
template <int MODE>
__global__ void __launch_bounds__(256) hashgrid_dx_pair_kernel(const __grid_constant__ GridParams gp,
                                                               const float* __restrict__ x, const float* __restrict__ table,
                                                               const float* __restrict__ dy, int64_t n,
                                                               float* __restrict__ dx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t >> 1;
  const bool live = i < n;
  const int par = (int)(t & 1);
  const int64_t ic = live ? i : n - 1;
  const float px = __ldg(x + 3 * ic), py = __ldg(x + 3 * ic + 1), pz = __ldg(x + 3 * ic + 2);
  const float* grow = dy + ic * (int64_t)(gp.n_levels * 2);
  float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll 2
  for (int l = 0; l < gp.n_levels; ++l) {
    const float g = __ldg(grow + 2 * l + par);  // upstream gradient of MY feature
    const Corners c = corners_of<MODE>(gp, l, px, py, pz);
    uint32_t r4[4];
    r4[0] = par ? c.row[0] : c.row[3], r4[1] = par ? c.row[1] : c.row[2];
    r4[2] = par ? c.row[4] : c.row[7], r4[3] = par ? c.row[5] : c.row[6];
    float2 f4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) f4[q] = __ldg(reinterpret_cast<const float2*>(table) + r4[q]);
    float fc[4], ff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float send = par ? f4[q].x : f4[q].y, mine = par ? f4[q].y : f4[q].x;
      const float theirs = __shfl_xor_sync(0xffffffffu, send, 1);
      fc[q] = par ? mine : theirs, ff[q] = par ? theirs : mine;
    }
    float ax = 0.f, ay = 0.f, az = 0.f;
    blend8_dpos(fc[0], fc[1], ff[1], ff[0], fc[2], fc[3], ff[3], ff[2], c.ox, c.oy, c.oz, g, ax, ay, az);
    const float sc = gp.scale[l];
    gx = fmaf(ax, sc, gx), gy = fmaf(ay, sc, gy), gz = fmaf(az, sc, gz);
  }
  gx += __shfl_xor_sync(0xffffffffu, gx, 1), gy += __shfl_xor_sync(0xffffffffu, gy, 1), gz += __shfl_xor_sync(0xffffffffu, gz, 1);
  if (live && par == 0) dx[3 * i] = gx, dx[3 * i + 1] = gy, dx[3 * i + 2] = gz;
}

// Run-length variant of the scatter: a thread walks CH consecutive samples of one level.  Consecutive samples along a
// ray stay in the same grid cell for many steps at the coarse levels, so the 8 corner contributions are accumulated in
// registers while the cell does not change and flushed with one vector RED per corner when it does — this removes
// most of the same-address atomic serialisation at L2 that dominates the proposal-grid backward once gradients
// become dense (measured 0.60 ms -> see profiles/).
template <int F, int MODE, int CH>
__global__ void __launch_bounds__(256) hashgrid_bwd_runs_kernel(const __grid_constant__ GridParams gp,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ dy, int64_t n,
                                                                float* __restrict__ dtable) {
  const int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = chunk * CH;
  if (i0 >= n) return;
  const int l = blockIdx.y;
  const int LF = gp.n_levels * F;
  uint32_t rows[8];
  float acc[8][F];
  bool have = false, pair = false;
  auto flush = [&]() {
    if constexpr (F == 2) {
      scatter_corners2(dtable, rows, pair, acc);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) red_row<F>(dtable, rows[k], acc[k]);
    }
  };
#pragma unroll 1
  for (int s = 0; s < CH; ++s) {
    const int64_t i = i0 + s;
    if (i >= n) break;
    float g[F];
    bool any = false;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      g[j] = __ldg(dy + i * LF + l * F + j);
      any |= (g[j] != 0.f);
    }
    if (!any) continue;
    const Corners c = corners_of<MODE>(gp, l, __ldg(x + 3 * i), __ldg(x + 3 * i + 1), __ldg(x + 3 * i + 2));
    bool same = have;
#pragma unroll
    for (int k = 0; k < 8; ++k) same &= (c.row[k] == rows[k]);
    if (!same) {
      if (have) flush();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        rows[k] = c.row[k];
#pragma unroll
        for (int j = 0; j < F; ++j) acc[k][j] = 0.f;
      }
      have = true, pair = c.xpair;
    }
    const float ox = c.ox, oy = c.oy, oz = c.oz;
    const float rx = 1.f - ox, ry = 1.f - oy, rz = 1.f - oz;
    const float w[8] = {ox * oy * oz, ox * ry * oz, rx * ry * oz, rx * oy * oz,
                        ox * oy * rz, ox * ry * rz, rx * ry * rz, rx * oy * rz};
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < F; ++j) acc[k][j] = fmaf(w[k], g[j], acc[k][j]);
  }
  if (have) flush();
}


template <int F>
static void launch_fwd(const GridParams& gp, const float* x, const float* table, int64_t n, float* y, int64_t* idx,
                       cudaStream_t st) {
  int lpb = g_levels_per_block_fwd > 0 ? g_levels_per_block_fwd : gp.n_levels;
  if (F == 2 && g_fwd_pair && idx == nullptr) {
    dim3 grid((unsigned)div_up(2 * n, 256), (unsigned)div_up(gp.n_levels, lpb));
    if (gp.mode == B2N_GRID_TORCH)
      hashgrid_fwd_pair_kernel<B2N_GRID_TORCH><<<grid, 256, 0, st>>>(gp, x, table, n, y, lpb);
    else
      hashgrid_fwd_pair_kernel<B2N_GRID_TCNN><<<grid, 256, 0, st>>>(gp, x, table, n, y, lpb);
    return;
  }
  dim3 grid((unsigned)div_up(n, 256), (unsigned)div_up(gp.n_levels, lpb));
  if (gp.mode == B2N_GRID_TORCH)
    hashgrid_fwd_kernel<F, B2N_GRID_TORCH><<<grid, 256, 0, st>>>(gp, x, table, n, y, idx, lpb);
  else
    hashgrid_fwd_kernel<F, B2N_GRID_TCNN><<<grid, 256, 0, st>>>(gp, x, table, n, y, idx, lpb);
}

template <int F, int CH>
static void launch_bwd_runs(const GridParams& gp, const float* x, const float* dy, int64_t n, float* dtable, cudaStream_t st) {
  dim3 grid((unsigned)div_up(div_up(n, CH), 256), (unsigned)gp.n_levels);
  if (gp.mode == B2N_GRID_TORCH)
    hashgrid_bwd_runs_kernel<F, B2N_GRID_TORCH, CH><<<grid, 256, 0, st>>>(gp, x, dy, n, dtable);
  else
    hashgrid_bwd_runs_kernel<F, B2N_GRID_TCNN, CH><<<grid, 256, 0, st>>>(gp, x, dy, n, dtable);
}

template <int F, bool DX>
static void launch_bwd(const GridParams& gp, const float* x, const float* table, const float* dy, int64_t n,
                       float* dtable, float* dx, cudaStream_t st) {
  if (!DX && g_bwd_chunk > 0) {
    if (g_bwd_chunk >= 16) launch_bwd_runs<F, 16>(gp, x, dy, n, dtable, st);
    else if (g_bwd_chunk >= 8) launch_bwd_runs<F, 8>(gp, x, dy, n, dtable, st);
    else launch_bwd_runs<F, 4>(gp, x, dy, n, dtable, st);
    return;
  }
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

extern "C" int b2n_hashgrid_dx(const B2nGrid* grid_host, const float* x, const float* table, const float* dy, int64_t n,
                               float* dx, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(grid_host && x && table && dy && dx, "null pointer");
  GridParams gp;
  B2N_REQUIRE(fill_params(grid_host, gp) == 0, "bad grid description");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)div_up(n, 256);
  const bool tm = gp.mode == B2N_GRID_TORCH;
  if (grid_host->n_features == 2 && g_fwd_pair) {
    const unsigned g2 = (unsigned)div_up(2 * n, 256);
    if (tm) hashgrid_dx_pair_kernel<B2N_GRID_TORCH><<<g2, 256, 0, st>>>(gp, x, table, dy, n, dx);
    else hashgrid_dx_pair_kernel<B2N_GRID_TCNN><<<g2, 256, 0, st>>>(gp, x, table, dy, n, dx);
    B2N_LAUNCH_CHECK();
  }
#define B2N_DX_CASE(F)                                                                                  \
  case F:                                                                                               \
    if (tm) hashgrid_dx_kernel<F, B2N_GRID_TORCH><<<grid, 256, 0, st>>>(gp, x, table, dy, n, dx);      \
    else hashgrid_dx_kernel<F, B2N_GRID_TCNN><<<grid, 256, 0, st>>>(gp, x, table, dy, n, dx);          \
    break;
  switch (grid_host->n_features) {
    B2N_DX_CASE(1)
    B2N_DX_CASE(2)
    B2N_DX_CASE(4)
    B2N_DX_CASE(8)
    default: B2N_UNSUPPORTED(true, "n_features must be 1, 2, 4 or 8");
  }
#undef B2N_DX_CASE
  B2N_LAUNCH_CHECK();
}

int b2n_tune_hashgrid(const char* key, int value) {
  if (!strcmp(key, "hash_levels_per_block_fwd")) { g_levels_per_block_fwd = value; return 1; }
  if (!strcmp(key, "hash_levels_per_block_bwd")) { g_levels_per_block_bwd = value; return 1; }
  if (!strcmp(key, "hash_bwd_chunk")) { g_bwd_chunk = value; return 1; }
  if (!strcmp(key, "hash_fwd_pair")) { g_fwd_pair = value; return 1; }
  return 0;
}
