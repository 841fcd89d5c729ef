// Fused proposal density field: ray sample -> unit cube -> hash grid (L levels, F=2) -> 2-layer MLP (H hidden, 1 out)
// -> trunc_exp density, forward and backward, one kernel each.
//
// Replaces HashMLPDensityField.get_density (nerfstudio/fields/density_fields.py:94-117) — the 1.44 M points per
// 4096-ray step that go through the two proposal networks are the largest point count on the path.
// Unfused, each level of the proposal sampler costs 4 forward + 3 backward launches and round-trips the [N,10]
// encoding and [N,16] hidden activations through HBM; the 10->16->1 network is far too small for either the tiled
// SIMT kernel (10 of 256 threads busy in its dW phase) or the tensor cores.  Here one thread owns a sample: the 40
// corner features are gathered straight into registers, the network (193 weights, staged once per CTA in shared
// memory and read as warp-uniform 128-bit broadcasts) is evaluated in FFMA, and nothing but the density leaves the SM.
//
// Backward recomputes the forward from the ray (re-gathering is cheaper than storing), forms d(encoding) in
// registers, scatters it with the run-length accumulation of hashgrid.cu (a thread walks CH consecutive samples of
// a ray and flushes its 8 corner accumulators only when the cell changes), and reduces the weight gradients
// dW1 = dZ1^T Enc through a shared-memory staging tile: thread (j,c) of the CTA sums its product over the CTA's
// samples, keeps the running total in a register across all tiles and issues ONE atomic at the end.
#include <string.h>

#include "hashgrid.cuh"
#include "positions.cuh"

#define DF_MAXL 8
#define DF_H 16
#define DF_MAXIN (2 * DF_MAXL)
#define DF_THREADS 256
#define DF_CH 4  // consecutive samples per thread in the backward kernel

#define DF_W1S 16  // padded row stride of w1 in shared memory (>= 2*DF_MAXL, multiple of 4)

// device pointers to the network of HashMLPDensityField.mlp_base[1] (nn.Linear layout) — they change every
// optimiser step and the launch may be replayed from a CUDA graph, so they are read on the device
struct DensityNet {
  const float* w1;  // [H][in]
  const float* b1;  // [H]
  const float* w2;  // [1][H]
  const float* b2;  // [1]
  float avg_init;
  int in_dim;
};

// shared-memory image: w1 [H][DF_W1S] | b1 [H] | w2 [H] | b2
#define DF_NET_FLOATS (DF_H * DF_W1S + 2 * DF_H + 4)

__device__ __forceinline__ void stage_net(const DensityNet& net, float* ws) {
  for (int idx = threadIdx.x; idx < DF_H * DF_W1S; idx += blockDim.x) {
    const int j = idx / DF_W1S, c = idx - j * DF_W1S;
    ws[idx] = c < net.in_dim ? __ldg(net.w1 + j * net.in_dim + c) : 0.f;
  }
  if (threadIdx.x < DF_H) {
    ws[DF_H * DF_W1S + threadIdx.x] = net.b1 ? __ldg(net.b1 + threadIdx.x) : 0.f;
    ws[DF_H * DF_W1S + DF_H + threadIdx.x] = __ldg(net.w2 + threadIdx.x);
  }
  if (threadIdx.x == 0) ws[DF_H * DF_W1S + 2 * DF_H] = net.b2 ? __ldg(net.b2) : 0.f;
}

struct RayGeom {
  const float* origins;     // [R,3]
  const float* directions;  // [R,3] or nullptr (point form: origins = positions [N,3], S = 1)
  const float* starts;      // [R,S] with row stride
  const float* ends;
  int64_t bin_stride;
  int64_t n_rays;
  int n_samples;
  int sm_count;  // SMs of the device (work distribution of the compacted backward)
};

template <int L>
struct Sample {
  float x[3];
  bool sel;
  float enc[2 * L];
};

template <int L, int MODE>
__device__ __forceinline__ void encode_sample(const GridParams& gp, const PosParams& pp, const RayGeom& rg,
                                              const float* __restrict__ table, int64_t i, Sample<L>& sm) {
  const int64_t r = i / rg.n_samples;
  const int s = (int)(i - r * rg.n_samples);
  if (rg.directions != nullptr)
    frustum_centre(rg.origins + 3 * r, rg.directions + 3 * r, __ldg(rg.starts + r * rg.bin_stride + s),
                   __ldg(rg.ends + r * rg.bin_stride + s), sm.x);
  else {
#pragma unroll
    for (int a = 0; a < 3; ++a) sm.x[a] = __ldg(rg.origins + 3 * i + a);
  }
  sm.sel = unit_cube_point(pp, sm.x);
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const Corners c = corners_of<MODE>(gp, l, sm.x[0], sm.x[1], sm.x[2]);
    Vec<2> f[8];  // plain 64-bit loads: on the coarse proposal grids a warp's corners already share sectors, and the
                  // paired 128-bit path of hashgrid.cuh only adds divergence here (measured 0.42 -> 0.58 ms backward)
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = ldg_row<2>(table, c.row[k]);
    const float ox = c.ox, oy = c.oy, oz = c.oz, rx = 1.f - ox, ry = 1.f - oy, rz = 1.f - oz;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float f03 = f[0].v[j] * ox + f[3].v[j] * rx, f12 = f[1].v[j] * ox + f[2].v[j] * rx;
      const float f56 = f[5].v[j] * ox + f[6].v[j] * rx, f47 = f[4].v[j] * ox + f[7].v[j] * rx;
      sm.enc[2 * l + j] = (f03 * oy + f12 * ry) * oz + (f47 * oy + f56 * ry) * rz;
    }
  }
}

template <int L>
__device__ __forceinline__ float mlp_forward(const float* __restrict__ ws, const float (&enc)[2 * L], float (&z1)[DF_H]) {
  constexpr int IN = 2 * L;
  float z2 = ws[DF_H * DF_W1S + 2 * DF_H];
#pragma unroll
  for (int j = 0; j < DF_H; ++j) {
    float a = ws[DF_H * DF_W1S + j];
    const float4* wr = reinterpret_cast<const float4*>(ws + j * DF_W1S);
#pragma unroll
    for (int q = 0; q < (IN + 3) / 4; ++q) {
      const float4 w = wr[q];
      if (4 * q + 0 < IN) a = fmaf(w.x, enc[4 * q + 0 < IN ? 4 * q + 0 : 0], a);
      if (4 * q + 1 < IN) a = fmaf(w.y, enc[4 * q + 1 < IN ? 4 * q + 1 : 0], a);
      if (4 * q + 2 < IN) a = fmaf(w.z, enc[4 * q + 2 < IN ? 4 * q + 2 : 0], a);
      if (4 * q + 3 < IN) a = fmaf(w.w, enc[4 * q + 3 < IN ? 4 * q + 3 : 0], a);
    }
    z1[j] = a;
    z2 = fmaf(ws[DF_H * DF_W1S + DF_H + j], fmaxf(a, 0.f), z2);
  }
  return z2;
}

template <int L, int MODE>
__global__ void __launch_bounds__(DF_THREADS) density_fused_fwd_kernel(const __grid_constant__ GridParams gp,
                                                                       const __grid_constant__ PosParams pp,
                                                                       const __grid_constant__ DensityNet net,
                                                                       const __grid_constant__ RayGeom rg,
                                                                       const float* __restrict__ table,
                                                                       float* __restrict__ density) {
  __shared__ __align__(16) float ws[DF_NET_FLOATS];
  stage_net(net, ws);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rg.n_rays * rg.n_samples) return;
  Sample<L> sm;
  encode_sample<L, MODE>(gp, pp, rg, table, i, sm);
  float z1[DF_H];
  const float z2 = mlp_forward<L>(ws, sm.enc, z1);
  density[i] = mul_rn(mul_rn(net.avg_init, expf(z2)), sm.sel ? 1.f : 0.f);
}

// Compaction of the samples that carry a gradient: live[0] = count (zeroed by the caller), live[1 + k] = sample index.
// A block owns 4096 consecutive samples and appends its survivors in order (one atomic per block reserves the range), so
// consecutive list entries are mostly consecutive samples of a ray — what the run-length scatter of the backward wants.
__global__ void __launch_bounds__(1024) live_compact_kernel(const float* __restrict__ d_density, int64_t n,
                                                            int32_t* __restrict__ live) {
  __shared__ int warp_cnt[32];
  __shared__ int block_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i0 = ((int64_t)blockIdx.x * 1024 + threadIdx.x) * 4;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (i0 + 3 < n) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(d_density + i0));
    g[0] = q.x, g[1] = q.y, g[2] = q.z, g[3] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = i0 + k < n ? __ldg(d_density + i0 + k) : 0.f;
  }
  const int mine = (g[0] != 0.f) + (g[1] != 0.f) + (g[2] != 0.f) + (g[3] != 0.f);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_cnt[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = warp_cnt[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    warp_cnt[lane] = wi - w;
    if (lane == 31) block_base = wi > 0 ? atomicAdd(live, wi) : 0;
  }
  __syncthreads();
  int dst = 1 + block_base + warp_cnt[warp] + incl - mine;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (g[k] != 0.f) live[dst++] = (int32_t)(i0 + k);
}

template <int L, int MODE, bool DX>
__global__ void __launch_bounds__(DF_THREADS, 2) density_fused_bwd_kernel(const __grid_constant__ GridParams gp,
                                                                       const __grid_constant__ PosParams pp,
                                                                       const __grid_constant__ DensityNet net,
                                                                       const __grid_constant__ RayGeom rg,
                                                                       const float* __restrict__ table,
                                                                       const float* __restrict__ d_density,
                                                                       const int32_t* __restrict__ live, int live_policy,
                                                                       float* __restrict__ dtable, float* __restrict__ dw1,
                                                                       float* __restrict__ db1, float* __restrict__ dw2,
                                                                       float* __restrict__ db2, float* __restrict__ d_origins,
                                                                       float* __restrict__ d_directions) {
  constexpr int IN = 2 * L;
  constexpr int SC = DF_H + IN;                 // staging columns: dz1[16] | enc[IN], stored column-major
  constexpr int CSW = DF_THREADS + 4;           // column stride (+16 B: columns read together sit in different banks)
  constexpr int DW = DF_CH * IN + 1;            // per-thread d(encoding) row in shared memory (odd stride)
  extern __shared__ __align__(16) float dyn_smem[];
  float* stage = dyn_smem;                       // [SC][CSW]
  float* denc = dyn_smem + SC * CSW;             // [DF_THREADS][DW]
  __shared__ __align__(16) float ws[DF_NET_FLOATS];
  __shared__ float red2[DF_H + 1];
  const int t = threadIdx.x;
  // `live` (optional): live[0] = number of samples with a non-zero gradient, live[1..] = their indices (compacted by
  // live_compact_kernel, order preserved inside 4096-sample blocks) — the kernel then only visits those
  const int64_t n = live ? (int64_t)live[0] : rg.n_rays * rg.n_samples;
  // Samples per thread and round: DF_CH when there is work for every CTA; with few live samples (the grid was sized on
  // the host for ALL samples) fewer, so that the live ones spread over as many CTAs as possible — 16 k live samples in
  // tiles of 1024 would keep 16 of 444 CTAs busy for four latency-bound rounds each.
  int ch = DF_CH;
  if (live && live_policy == 1) {         // spread as widely as the launched grid allows
    const int64_t per = (n + (int64_t)gridDim.x * DF_THREADS - 1) / ((int64_t)gridDim.x * DF_THREADS);
    ch = (int)min((int64_t)DF_CH, max((int64_t)1, per));
  } else if (live && live_policy >= 2) {  // the most samples per thread (run-length merged scatter) that still leaves
    const int64_t ctas = min((int64_t)gridDim.x, (int64_t)(live_policy - 1) * rg.sm_count);  // >= (policy-1) CTAs per SM
    ch = (int)min((int64_t)DF_CH, max((int64_t)1, n / (ctas * DF_THREADS)));
  }
  const int64_t n_tiles = (n + DF_THREADS * ch - 1) / (DF_THREADS * ch);
  if ((int64_t)blockIdx.x >= n_tiles) return;  // CTA-uniform: nothing to do, nothing to flush
  stage_net(net, ws);
  // weight-gradient owners (one register accumulator each, kept across all tiles of the CTA):
  //   t in [0, H*IN) owns dW1[j][c], the next H threads own db1[j]; dW2 / db2 are accumulated per thread
  const int oj = t / IN, oc = t - oj * IN;
  const int role = t < DF_H * IN ? 0 : t < DF_H * IN + DF_H ? 1 : 2;
  float own_acc = 0.f;
  float g2[DF_H + 1];
#pragma unroll
  for (int j = 0; j <= DF_H; ++j) g2[j] = 0.f;
  if (t <= DF_H) red2[t] = 0.f;
  for (int q = t; q < SC * CSW; q += DF_THREADS) stage[q] = 0.f;  // skipped warps leave columns untouched: start finite
  __syncthreads();
  float* my_denc = denc + t * DW;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t i0 = (tile * DF_THREADS + t) * ch;
    float xs[DF_CH][3];
#pragma unroll
    for (int s = 0; s < DF_CH; ++s) {
      if (s >= ch) {  // CTA-uniform
#pragma unroll
        for (int c = 0; c < IN; ++c) my_denc[s * IN + c] = 0.f;
        continue;
      }
      const int64_t slot = i0 + s;
      const bool in = slot < n;
      const int64_t i = !in ? 0 : (live ? (int64_t)live[1 + slot] : slot);
      // Every gradient this kernel produces is proportional to d_density of the sample.  The interlevel loss — the only
      // loss that reaches the proposal networks — is zero wherever the proposal histogram already bounds the final
      // weights, and weights_bwd turns that into exact zeros for whole rays / ray tails, so a warp (32 lanes = samples
      // spaced DF_CH apart, i.e. one ray's neighbourhood) whose samples all have zero gradient skips the re-gather, the
      // network and the staging; a CTA round with no live warp also skips the owners' reduction.
      const float g = in ? __ldg(d_density + i) : 0.f;
      const bool warp_live = __any_sync(0xffffffffu, g != 0.f);
      float dz2 = 0.f;
      if (!warp_live) {
#pragma unroll
        for (int j = 0; j < DF_H; ++j) stage[j * CSW + t] = 0.f;  // dz1 = 0: stale encodings in the other columns are harmless
#pragma unroll
        for (int c = 0; c < IN; ++c) my_denc[s * IN + c] = 0.f;
      } else if (in) {
        Sample<L> sm;
        encode_sample<L, MODE>(gp, pp, rg, table, i, sm);
        float z1[DF_H];
        const float z2 = mlp_forward<L>(ws, sm.enc, z1);
        // density = avg * exp(z2) * sel ; trunc_exp backward clamps the exponent (activations.py:36-41)
        if (sm.sel && g != 0.f) dz2 = g * net.avg_init * expf(fminf(fmaxf(z2, -15.f), 15.f));
        float de[IN];
#pragma unroll
        for (int c = 0; c < IN; ++c) de[c] = 0.f;
#pragma unroll
        for (int j = 0; j < DF_H; ++j) {
          g2[j] = fmaf(dz2, fmaxf(z1[j], 0.f), g2[j]);
          const float d = z1[j] > 0.f ? dz2 * ws[DF_H * DF_W1S + DF_H + j] : 0.f;
          stage[j * CSW + t] = d;
          const float4* wr = reinterpret_cast<const float4*>(ws + j * DF_W1S);
#pragma unroll
          for (int q = 0; q < (IN + 3) / 4; ++q) {
            const float4 w = wr[q];
            if (4 * q + 0 < IN) de[4 * q + 0 < IN ? 4 * q + 0 : 0] = fmaf(d, w.x, de[4 * q + 0 < IN ? 4 * q + 0 : 0]);
            if (4 * q + 1 < IN) de[4 * q + 1 < IN ? 4 * q + 1 : 0] = fmaf(d, w.y, de[4 * q + 1 < IN ? 4 * q + 1 : 0]);
            if (4 * q + 2 < IN) de[4 * q + 2 < IN ? 4 * q + 2 : 0] = fmaf(d, w.z, de[4 * q + 2 < IN ? 4 * q + 2 : 0]);
            if (4 * q + 3 < IN) de[4 * q + 3 < IN ? 4 * q + 3 : 0] = fmaf(d, w.w, de[4 * q + 3 < IN ? 4 * q + 3 : 0]);
          }
        }
        g2[DF_H] += dz2;
        if constexpr (DX) {
          // position gradient (camera optimiser): d enc / d x from the 8 corners of every level, through the unit-cube
          // map's Jacobian, added to the ray's d(origin) and d(direction) (ray form only)
          if (dz2 != 0.f && rg.directions != nullptr) {
            float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
            for (int l = 0; l < L; ++l) {
              const Corners c = corners_of<MODE>(gp, l, sm.x[0], sm.x[1], sm.x[2]);
              Vec<2> f[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) f[k] = ldg_row<2>(table, c.row[k]);
              float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
              for (int j = 0; j < 2; ++j)
                blend8_dpos(f[0].v[j], f[1].v[j], f[2].v[j], f[3].v[j], f[4].v[j], f[5].v[j], f[6].v[j], f[7].v[j], c.ox, c.oy,
                            c.oz, de[2 * l + j], ax, ay, az);
              const float sc = gp.scale[l];
              gx = fmaf(ax, sc, gx), gy = fmaf(ay, sc, gy), gz = fmaf(az, sc, gz);
            }
            const int64_t r = i / rg.n_samples;
            const int sidx = (int)(i - r * rg.n_samples);
            const float tm = 0.5f * (__ldg(rg.starts + r * rg.bin_stride + sidx) + __ldg(rg.ends + r * rg.bin_stride + sidx));
            float praw[3], gp3[3] = {gx, gy, gz};
#pragma unroll
            for (int a = 0; a < 3; ++a) praw[a] = fmaf(__ldg(rg.directions + 3 * r + a), tm, __ldg(rg.origins + 3 * r + a));
            unit_cube_point_bwd(pp, praw, gp3);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              if (d_origins) atomicAdd(d_origins + 3 * r + a, gp3[a]);
              if (d_directions) atomicAdd(d_directions + 3 * r + a, tm * gp3[a]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < IN; ++c) stage[(DF_H + c) * CSW + t] = sm.enc[c], my_denc[s * IN + c] = de[c];
#pragma unroll
        for (int a = 0; a < 3; ++a) xs[s][a] = sm.x[a];
      } else {
#pragma unroll
        for (int j = 0; j < SC; ++j) stage[j * CSW + t] = 0.f;
#pragma unroll
        for (int c = 0; c < IN; ++c) my_denc[s * IN + c] = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) xs[s][a] = 0.f;
      }
      if (!__syncthreads_or(warp_live)) continue;  // CTA-uniform: nothing staged this round, nothing to reduce
      // ---- the owners reduce their product over this round's 256 staged samples (128-bit column reads)
      if (role == 0) {
        const float4* zc = reinterpret_cast<const float4*>(stage + oj * CSW);
        const float4* ec = reinterpret_cast<const float4*>(stage + (DF_H + oc) * CSW);
        float a = 0.f;
#pragma unroll 8
        for (int q = 0; q < DF_THREADS / 4; ++q) {
          const float4 z = zc[q], e = ec[q];
          a = fmaf(z.x, e.x, a), a = fmaf(z.y, e.y, a), a = fmaf(z.z, e.z, a), a = fmaf(z.w, e.w, a);
        }
        own_acc += a;
      } else if (role == 1) {
        const float4* zc = reinterpret_cast<const float4*>(stage + (t - DF_H * IN) * CSW);
        float a = 0.f;
#pragma unroll 8
        for (int q = 0; q < DF_THREADS / 4; ++q) {
          const float4 z = zc[q];
          a += (z.x + z.y) + (z.z + z.w);
        }
        own_acc += a;
      }
      __syncthreads();
    }
    // ---- scatter d_enc into the table: per level, run-length accumulation over this thread's CH samples
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      uint32_t rows[8];
      float acc[8][2];
      bool have = false;
#pragma unroll
      for (int s = 0; s < DF_CH; ++s) {
        const float ga = my_denc[s * IN + 2 * l], gb = my_denc[s * IN + 2 * l + 1];
        if (ga == 0.f && gb == 0.f) continue;
        const Corners c = corners_of<MODE>(gp, l, xs[s][0], xs[s][1], xs[s][2]);
        bool same = have;
#pragma unroll
        for (int k = 0; k < 8; ++k) same &= (c.row[k] == rows[k]);
        if (!same) {
          if (have) {
#pragma unroll
            for (int k = 0; k < 8; ++k) red_row<2>(dtable, rows[k], acc[k]);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) rows[k] = c.row[k], acc[k][0] = 0.f, acc[k][1] = 0.f;
          have = true;
        }
        const float ox = c.ox, oy = c.oy, oz = c.oz, rx = 1.f - ox, ry = 1.f - oy, rz = 1.f - oz;
        const float w[8] = {ox * oy * oz, ox * ry * oz, rx * ry * oz, rx * oy * oz,
                            ox * oy * rz, ox * ry * rz, rx * ry * rz, rx * oy * rz};
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k][0] = fmaf(w[k], ga, acc[k][0]), acc[k][1] = fmaf(w[k], gb, acc[k][1]);
      }
      if (have) {
#pragma unroll
        for (int k = 0; k < 8; ++k) red_row<2>(dtable, rows[k], acc[k]);
      }
    }
  }
  // ---- flush the weight gradients: one atomic per entry per CTA
  if (role == 0 && dw1) atomicAdd(dw1 + oj * IN + oc, own_acc);
  if (role == 1 && db1) atomicAdd(db1 + (t - DF_H * IN), own_acc);
#pragma unroll
  for (int j = 0; j <= DF_H; ++j) {
    const float v = warp_sum(g2[j]);
    if ((t & 31) == 0) atomicAdd(&red2[j], v);
  }
  __syncthreads();
  if (t < DF_H && dw2) atomicAdd(dw2 + t, red2[t]);
  if (t == DF_H && db2) atomicAdd(db2, red2[DF_H]);
}

// ------------------------------------------------------------------------------------------------
static int check_shape(const B2nGrid* g, const B2nMlp* m) {
  if (g->n_features != 2 || g->n_levels < 1 || g->n_levels > DF_MAXL) return -1;
  if (m->n_layers != 2 || m->in_dim != 2 * g->n_levels || m->out_dims[0] != DF_H || m->out_dims[1] != 1) return -1;
  if (m->hidden_act != B2N_ACT_RELU || m->out_act != B2N_ACT_NONE || m->skip[0] || m->skip[1]) return -1;
  if (!m->w[0] || !m->w[1]) return -1;
  return 0;
}

// compacted backward: samples per thread and round from the live count (b2n_tune "df_live_policy"): 0 = always DF_CH,
// 1 = spread over the whole launched grid, k >= 2 = the largest count that still leaves (k-1) CTAs per SM busy
static int g_live_policy = 2;
int b2n_tune_density_fused(const char* key, int value) {
  if (strcmp(key, "df_live_policy") == 0 && value >= 0 && value <= 4) {
    g_live_policy = value;
    return 1;
  }
  return 0;
}

template <int L, int MODE>
static void launch_fused_bwd(unsigned grid, cudaStream_t st, const GridParams& gp, const PosParams& pp, const DensityNet& net,
                             const RayGeom& rg, const float* table, const float* d_density, const int32_t* live,
                             float* dtable, float* dw1, float* db1, float* dw2, float* db2, float* d_origins,
                             float* d_directions) {
  constexpr size_t smem = sizeof(float) * ((DF_H + 2 * L) * (DF_THREADS + 4) + DF_THREADS * (DF_CH * 2 * L + 1));
  if (d_origins != nullptr || d_directions != nullptr) {
    auto kernel = density_fused_bwd_kernel<L, MODE, true>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, DF_THREADS, smem, st>>>(gp, pp, net, rg, table, d_density, live, g_live_policy, dtable, dw1, db1, dw2, db2,
                                           d_origins, d_directions);
  } else {
    auto kernel = density_fused_bwd_kernel<L, MODE, false>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, DF_THREADS, smem, st>>>(gp, pp, net, rg, table, d_density, live, g_live_policy, dtable, dw1, db1, dw2, db2, nullptr,
                                           nullptr);
  }
}

#define DF_DISPATCH(KERNEL, ...)                                                          \
  do {                                                                                    \
    const bool torch_mode = gp.mode == B2N_GRID_TORCH;                                    \
    switch (gp.n_levels) {                                                                \
      case 1: if (torch_mode) KERNEL<1, B2N_GRID_TORCH> __VA_ARGS__; else KERNEL<1, B2N_GRID_TCNN> __VA_ARGS__; break; \
      case 2: if (torch_mode) KERNEL<2, B2N_GRID_TORCH> __VA_ARGS__; else KERNEL<2, B2N_GRID_TCNN> __VA_ARGS__; break; \
      case 3: if (torch_mode) KERNEL<3, B2N_GRID_TORCH> __VA_ARGS__; else KERNEL<3, B2N_GRID_TCNN> __VA_ARGS__; break; \
      case 4: if (torch_mode) KERNEL<4, B2N_GRID_TORCH> __VA_ARGS__; else KERNEL<4, B2N_GRID_TCNN> __VA_ARGS__; break; \
      case 5: if (torch_mode) KERNEL<5, B2N_GRID_TORCH> __VA_ARGS__; else KERNEL<5, B2N_GRID_TCNN> __VA_ARGS__; break; \
      case 6: if (torch_mode) KERNEL<6, B2N_GRID_TORCH> __VA_ARGS__; else KERNEL<6, B2N_GRID_TCNN> __VA_ARGS__; break; \
      case 7: if (torch_mode) KERNEL<7, B2N_GRID_TORCH> __VA_ARGS__; else KERNEL<7, B2N_GRID_TCNN> __VA_ARGS__; break; \
      default: if (torch_mode) KERNEL<8, B2N_GRID_TORCH> __VA_ARGS__; else KERNEL<8, B2N_GRID_TCNN> __VA_ARGS__; break; \
    }                                                                                     \
  } while (0)

extern "C" int b2n_density_field_fwd(const B2nGrid* grid_host, const B2nMlp* mlp_host, const float* table,
                                     const float* origins, const float* directions, const float* starts,
                                     const float* ends, int64_t bin_stride, int64_t n_rays, int32_t n_samples,
                                     int32_t contraction, const float* aabb_host6, float avg_init, float* density,
                                     void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(grid_host && mlp_host && table && origins && density, "null pointer");
  B2N_REQUIRE(directions == nullptr || (starts && ends), "ray form needs starts/ends");
  B2N_REQUIRE(contraction || aabb_host6, "aabb required without contraction");
  B2N_UNSUPPORTED(check_shape(grid_host, mlp_host) != 0,
                  "fused density field: F=2, <= 8 levels, MLP in->16->1 with ReLU (the nerfacto proposal networks)");
  GridParams gp;
  B2N_REQUIRE(fill_params(grid_host, gp) == 0, "bad grid description");
  PosParams pp;
  fill_pos_params(pp, contraction, aabb_host6);
  DensityNet net{mlp_host->w[0], mlp_host->b[0], mlp_host->w[1], mlp_host->b[1], avg_init, mlp_host->in_dim};
  RayGeom rg{origins, directions, starts, ends, bin_stride, n_rays, directions ? n_samples : 1, b2n_sm_count()};
  const int64_t n = rg.n_rays * rg.n_samples;
  const unsigned grid = (unsigned)div_up(n, DF_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  DF_DISPATCH(density_fused_fwd_kernel, <<<grid, DF_THREADS, 0, st>>>(gp, pp, net, rg, table, density));
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_density_field_bwd_rays(const B2nGrid* grid_host, const B2nMlp* mlp_host, const B2nMlpGrad* grad_host,
                                          const float* table, const float* origins, const float* directions,
                                          const float* starts, const float* ends, int64_t bin_stride, int64_t n_rays,
                                          int32_t n_samples, int32_t contraction, const float* aabb_host6, float avg_init,
                                          const float* d_density, float* dtable, int32_t* live_ws, float* d_origins,
                                          float* d_directions, void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(grid_host && mlp_host && grad_host && table && origins && d_density && dtable, "null pointer");
  B2N_REQUIRE(directions == nullptr || (starts && ends), "ray form needs starts/ends");
  B2N_REQUIRE(contraction || aabb_host6, "aabb required without contraction");
  B2N_UNSUPPORTED(check_shape(grid_host, mlp_host) != 0,
                  "fused density field: F=2, <= 8 levels, MLP in->16->1 with ReLU (the nerfacto proposal networks)");
  GridParams gp;
  B2N_REQUIRE(fill_params(grid_host, gp) == 0, "bad grid description");
  PosParams pp;
  fill_pos_params(pp, contraction, aabb_host6);
  DensityNet net{mlp_host->w[0], mlp_host->b[0], mlp_host->w[1], mlp_host->b[1], avg_init, mlp_host->in_dim};
  RayGeom rg{origins, directions, starts, ends, bin_stride, n_rays, directions ? n_samples : 1, b2n_sm_count()};
  const int64_t n = rg.n_rays * rg.n_samples;
  const int64_t tiles = div_up(n, (int64_t)DF_THREADS * DF_CH);
  const unsigned grid = (unsigned)min(tiles, (int64_t)b2n_sm_count() * 3);
  cudaStream_t st = (cudaStream_t)stream;
  if (live_ws != nullptr) {  // visit only the samples that carry a gradient
    B2N_REQUIRE(n < (int64_t)1 << 31 && (reinterpret_cast<uintptr_t>(d_density) & 15) == 0, "live compaction: n < 2^31, aligned d_density");
    cudaMemsetAsync(live_ws, 0, sizeof(int32_t), st);
    live_compact_kernel<<<(unsigned)div_up(n, 4096), 1024, 0, st>>>(d_density, n, live_ws);
  }
  DF_DISPATCH(launch_fused_bwd, (grid, st, gp, pp, net, rg, table, d_density, live_ws, dtable, grad_host->dw[0],
                                 grad_host->db[0], grad_host->dw[1], grad_host->db[1], d_origins, d_directions));
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_density_field_bwd_ws(const B2nGrid* grid_host, const B2nMlp* mlp_host, const B2nMlpGrad* grad_host,
                                        const float* table, const float* origins, const float* directions,
                                        const float* starts, const float* ends, int64_t bin_stride, int64_t n_rays,
                                        int32_t n_samples, int32_t contraction, const float* aabb_host6, float avg_init,
                                        const float* d_density, float* dtable, int32_t* live_ws, void* stream) {
  return b2n_density_field_bwd_rays(grid_host, mlp_host, grad_host, table, origins, directions, starts, ends, bin_stride, n_rays,
                                    n_samples, contraction, aabb_host6, avg_init, d_density, dtable, live_ws, nullptr, nullptr,
                                    stream);
}

extern "C" int b2n_density_field_bwd(const B2nGrid* grid_host, const B2nMlp* mlp_host, const B2nMlpGrad* grad_host,
                                     const float* table, const float* origins, const float* directions,
                                     const float* starts, const float* ends, int64_t bin_stride, int64_t n_rays,
                                     int32_t n_samples, int32_t contraction, const float* aabb_host6, float avg_init,
                                     const float* d_density, float* dtable, void* stream) {
  return b2n_density_field_bwd_ws(grid_host, mlp_host, grad_host, table, origins, directions, starts, ends, bin_stride, n_rays,
                                  n_samples, contraction, aabb_host6, avg_init, d_density, dtable, nullptr, stream);
}
