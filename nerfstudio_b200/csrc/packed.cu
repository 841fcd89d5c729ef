// K7-K12 — packed (instant-ngp) path: occupancy-grid marching, pack_info, packed transmittance weights,
// per-ray accumulation.  These replace the nerfacc 0.5.2 calls of the reference
// (nerfstudio/models/instant_ngp.py:120-198, model_components/ray_samplers.py:481-493,
// model_components/renderers.py:97-102,312-314,371-376).  nerfacc's source is not available: semantics follow
// its published behaviour and are restated in oracle/nerf_oracle.py ("parity unpinned").
#include <string.h>

#include "common.cuh"

#define PW 4  // warps (rays) per CTA for the per-ray scans

// ------------------------------------------------------------------------------------------------
__global__ void pack_info_kernel(const int64_t* __restrict__ ray_indices, int64_t m, int64_t n_rays,
                                 int64_t* __restrict__ packed_info) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  int64_t lo = 0, hi = m;  // lower_bound(r)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ray_indices[mid] < r) lo = mid + 1; else hi = mid;
  }
  const int64_t start = lo;
  hi = m;  // upper_bound(r)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ray_indices[mid] <= r) lo = mid + 1; else hi = mid;
  }
  packed_info[2 * r] = start;
  packed_info[2 * r + 1] = lo - start;
}

extern "C" int b2n_pack_info(const int64_t* ray_indices, int64_t m, int64_t n_rays, int64_t* packed_info, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(packed_info && (ray_indices || m == 0), "null pointer");
  if (n_rays == 0) return B2N_OK;
  pack_info_kernel<<<(unsigned)div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(ray_indices, m, n_rays, packed_info);
  B2N_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PW * 32) packed_weights_fwd_kernel(const float* __restrict__ ts, const float* __restrict__ te,
                                                                     const float* __restrict__ sig,
                                                                     const int64_t* __restrict__ info, int64_t n_rays,
                                                                     float* __restrict__ w, float* __restrict__ trans,
                                                                     float* __restrict__ alphas) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * PW + warp;
  if (r >= n_rays) return;
  const int64_t base = info[2 * r];
  const int cnt = (int)info[2 * r + 1];
  const int chunk = (cnt + 31) / 32, i0 = lane * chunk, i1 = min(cnt, i0 + chunk);
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += (double)mul_rn(__ldg(sig + base + i), sub_rn(__ldg(te + base + i), __ldg(ts + base + i)));
  double run = warp_scan_incl_d(local, lane) - local;
  for (int i = i0; i < i1; ++i) {
    const float sd = mul_rn(__ldg(sig + base + i), sub_rn(__ldg(te + base + i), __ldg(ts + base + i)));
    const float a = sub_rn(1.f, expf(-sd)), T = expf(-(float)run);
    w[base + i] = mul_rn(T, a);
    if (trans) trans[base + i] = T;
    if (alphas) alphas[base + i] = a;
    run += (double)sd;
  }
}

__global__ void __launch_bounds__(PW * 32) packed_weights_bwd_kernel(const float* __restrict__ ts, const float* __restrict__ te,
                                                                     const float* __restrict__ sig,
                                                                     const int64_t* __restrict__ info,
                                                                     const float* __restrict__ g, int64_t n_rays,
                                                                     float* __restrict__ dsig) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * PW + warp;
  if (r >= n_rays) return;
  const int64_t base = info[2 * r];
  const int cnt = (int)info[2 * r + 1];
  const int chunk = (cnt + 31) / 32, i0 = lane * chunk, i1 = min(cnt, i0 + chunk);
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += (double)mul_rn(__ldg(sig + base + i), sub_rn(__ldg(te + base + i), __ldg(ts + base + i)));
  const double excl = warp_scan_incl_d(local, lane) - local;
  double run = excl, gw = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float sd = mul_rn(__ldg(sig + base + i), sub_rn(__ldg(te + base + i), __ldg(ts + base + i)));
    gw += (double)(__ldg(g + base + i) * (1.f - expf(-sd)) * expf(-(float)run));
    run += (double)sd;
  }
  const double gw_incl = warp_scan_incl_d(gw, lane);
  double suffix = __shfl_sync(0xffffffffu, gw_incl, 31) - gw_incl;
  run = excl + local;
  for (int i = i1 - 1; i >= i0; --i) {
    const float dt = sub_rn(__ldg(te + base + i), __ldg(ts + base + i));
    const float sd = mul_rn(__ldg(sig + base + i), dt);
    run -= (double)sd;
    const float ea = expf(-sd), T = expf(-(float)run), gi = __ldg(g + base + i);
    dsig[base + i] = dt * (gi * T * ea - (float)suffix);
    suffix += (double)(gi * (1.f - ea) * T);
  }
}

extern "C" int b2n_packed_weights_fwd(const float* t_starts, const float* t_ends, const float* sigmas,
                                      const int64_t* packed_info, int64_t n_rays, float* weights, float* trans,
                                      float* alphas, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(t_starts && t_ends && sigmas && packed_info && weights, "null pointer");
  if (n_rays == 0) return B2N_OK;
  packed_weights_fwd_kernel<<<(unsigned)div_up(n_rays, PW), PW * 32, 0, (cudaStream_t)stream>>>(t_starts, t_ends, sigmas, packed_info, n_rays, weights, trans, alphas);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_packed_weights_bwd(const float* t_starts, const float* t_ends, const float* sigmas,
                                      const int64_t* packed_info, const float* dweights, int64_t n_rays, float* dsigmas,
                                      void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(t_starts && t_ends && sigmas && packed_info && dweights && dsigmas, "null pointer");
  if (n_rays == 0) return B2N_OK;
  packed_weights_bwd_kernel<<<(unsigned)div_up(n_rays, PW), PW * 32, 0, (cudaStream_t)stream>>>(t_starts, t_ends, sigmas, packed_info, dweights, n_rays, dsigmas);
  B2N_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PW * 32) packed_accum_fwd_kernel(const float* __restrict__ w, const float* __restrict__ v,
                                                                   int d, const int64_t* __restrict__ info, int64_t n_rays,
                                                                   float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * PW + warp;
  if (r >= n_rays) return;
  const int64_t base = info[2 * r];
  const int cnt = (int)info[2 * r + 1];
  for (int c = 0; c < d; ++c) {
    float s = 0.f;
    for (int i = lane; i < cnt; i += 32) s = fmaf(__ldg(w + base + i), v ? __ldg(v + (base + i) * d + c) : 1.f, s);
    s = warp_sum(s);
    if (lane == 0) out[r * d + c] = s;
  }
}

__global__ void __launch_bounds__(PW * 32) packed_accum_bwd_kernel(const float* __restrict__ w, const float* __restrict__ v,
                                                                   int d, const int64_t* __restrict__ info,
                                                                   const float* __restrict__ dout, int64_t n_rays,
                                                                   float* __restrict__ dw, float* __restrict__ dv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * PW + warp;
  if (r >= n_rays) return;
  const int64_t base = info[2 * r];
  const int cnt = (int)info[2 * r + 1];
  for (int i = lane; i < cnt; i += 32) {
    const float wi = __ldg(w + base + i);
    float acc = 0.f;
    for (int c = 0; c < d; ++c) {
      const float go = __ldg(dout + r * d + c);
      acc = fmaf(go, v ? __ldg(v + (base + i) * d + c) : 1.f, acc);
      if (dv) dv[(base + i) * d + c] = wi * go;
    }
    if (dw) dw[base + i] = acc;
  }
}

extern "C" int b2n_packed_accumulate_fwd(const float* weights, const float* values, int32_t d, const int64_t* packed_info,
                                         int64_t n_rays, float* out, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(weights && packed_info && out && d >= 1, "bad arguments");
  if (n_rays == 0) return B2N_OK;
  packed_accum_fwd_kernel<<<(unsigned)div_up(n_rays, PW), PW * 32, 0, (cudaStream_t)stream>>>(weights, values, d, packed_info, n_rays, out);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_packed_accumulate_bwd(const float* weights, const float* values, int32_t d, const int64_t* packed_info,
                                         const float* dout, int64_t n_rays, float* dweights, float* dvalues, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(weights && packed_info && dout && d >= 1, "bad arguments");
  B2N_REQUIRE(!(dvalues && !values), "dvalues without values");
  if (n_rays == 0) return B2N_OK;
  packed_accum_bwd_kernel<<<(unsigned)div_up(n_rays, PW), PW * 32, 0, (cudaStream_t)stream>>>(weights, values, d, packed_info, dout, n_rays, dweights, dvalues);
  B2N_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// occupancy-grid marching.  Same code path for the count and the fill pass, so counts/offsets/indices agree.
// ------------------------------------------------------------------------------------------------
struct March {
  int levels, res;
  float centre[3], half[3];
  float step, cone, near_plane, far_plane;
};

template <bool FILL>
__global__ void occgrid_march_kernel(const __grid_constant__ March mp, const float* __restrict__ origins,
                                     const float* __restrict__ directions, const float* __restrict__ t_min,
                                     const float* __restrict__ t_max, const uint8_t* __restrict__ binaries,
                                     const float* __restrict__ jitter, int64_t n_rays, int32_t* __restrict__ counts,
                                     const int64_t* __restrict__ offsets, int64_t* __restrict__ ray_indices,
                                     float* __restrict__ t_starts, float* __restrict__ t_ends) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  float o[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) o[a] = __ldg(origins + 3 * r + a), d[a] = __ldg(directions + 3 * r + a);
  const float big = (float)(1 << (mp.levels - 1));
  float tn = -INFINITY, tf = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lo = sub_rn(mp.centre[a], mul_rn(mp.half[a], big)), hi = add_rn(mp.centre[a], mul_rn(mp.half[a], big));
    const float inv = div_rn(1.f, d[a]);
    const float t1 = mul_rn(sub_rn(lo, o[a]), inv), t2 = mul_rn(sub_rn(hi, o[a]), inv);
    tn = fmaxf(tn, fminf(t1, t2)), tf = fminf(tf, fmaxf(t1, t2));
  }
  tn = fmaxf(tn, mp.near_plane), tf = fminf(tf, mp.far_plane);
  if (t_min) tn = fmaxf(tn, __ldg(t_min + r));
  if (t_max) tf = fminf(tf, __ldg(t_max + r));
  int n = 0;
  int64_t out = FILL ? offsets[r] : 0;
  if (tf > tn) {
    float t = tn;
    if (jitter) t = add_rn(t, mul_rn(__ldg(jitter + r), mp.step));
    const int res = mp.res;
    while (t < tf) {
      const float dt = fmaxf(mul_rn(t, mp.cone), mp.step);
      const float t1 = add_rn(t, dt);
      const float mid = mul_rn(add_rn(t, t1), 0.5f);
      float p[3], m = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        p[a] = add_rn(o[a], mul_rn(d[a], mid));
        m = fmaxf(m, div_rn(fabsf(sub_rn(p[a], mp.centre[a])), mp.half[a]));
      }
      int lvl = 0;
      while (lvl < mp.levels && m > (float)(1 << lvl)) ++lvl;
      if (lvl < mp.levels) {
        const float scale = (float)(1 << lvl);
        int cell[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float q = mul_rn(add_rn(div_rn(sub_rn(p[a], mp.centre[a]), mul_rn(mp.half[a], scale)), 1.f), 0.5f);
          cell[a] = min(max((int)floorf(mul_rn(q, (float)res)), 0), res - 1);
        }
        const size_t idx = (((size_t)lvl * res + cell[0]) * res + cell[1]) * res + cell[2];
        if (binaries[idx]) {
          if (FILL) {
            ray_indices[out] = r, t_starts[out] = t, t_ends[out] = t1;
            ++out;
          }
          ++n;
        }
      }
      t = t1;
    }
  }
  if (!FILL) counts[r] = n;
}

// Warp-per-ray variant (default).  One thread per ray gives 4096 threads for a 4096-ray batch — 128 warps on 148 SMs,
// each stepping ~1000 dependent iterations with a global load inside: pure latency (measured 0.69 + 0.57 ms for the two
// passes at BASELINE config 2, a third of the instant-ngp step).  Here a warp owns the ray and handles 32 consecutive
// candidate intervals per round: every lane re-runs the fp32 recurrence t <- t + max(t*cone, step) `lane` times from the
// round's start (the SAME separately rounded operations in the same order as the serial march, so t values are
// bit-identical), tests its own interval, and a ballot compacts the hits in order.
template <bool FILL>
__global__ void __launch_bounds__(128) occgrid_march_warp_kernel(
    const __grid_constant__ March mp, const float* __restrict__ origins, const float* __restrict__ directions,
    const float* __restrict__ t_min, const float* __restrict__ t_max, const uint8_t* __restrict__ binaries,
    const float* __restrict__ jitter, int64_t n_rays, int32_t* __restrict__ counts, const int64_t* __restrict__ offsets,
    int64_t* __restrict__ ray_indices, float* __restrict__ t_starts, float* __restrict__ t_ends) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rays) return;
  float o[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) o[a] = __ldg(origins + 3 * r + a), d[a] = __ldg(directions + 3 * r + a);
  const float big = (float)(1 << (mp.levels - 1));
  float tn = -INFINITY, tf = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lo = sub_rn(mp.centre[a], mul_rn(mp.half[a], big)), hi = add_rn(mp.centre[a], mul_rn(mp.half[a], big));
    const float inv = div_rn(1.f, d[a]);
    const float t1 = mul_rn(sub_rn(lo, o[a]), inv), t2 = mul_rn(sub_rn(hi, o[a]), inv);
    tn = fmaxf(tn, fminf(t1, t2)), tf = fminf(tf, fmaxf(t1, t2));
  }
  tn = fmaxf(tn, mp.near_plane), tf = fminf(tf, mp.far_plane);
  if (t_min) tn = fmaxf(tn, __ldg(t_min + r));
  if (t_max) tf = fminf(tf, __ldg(t_max + r));
  int n = 0;
  int64_t out = FILL ? offsets[r] : 0;
  if (tf > tn) {
    float base = tn;
    if (jitter) base = add_rn(base, mul_rn(__ldg(jitter + r), mp.step));
    const int res = mp.res;
    while (base < tf) {  // warp-uniform
      float t = base;
#pragma unroll 1
      for (int i = 0; i < 31; ++i)
        if (i < lane) t = add_rn(t, fmaxf(mul_rn(t, mp.cone), mp.step));
      const float dt = fmaxf(mul_rn(t, mp.cone), mp.step);
      const float t1 = add_rn(t, dt);
      bool hit = false;
      if (t < tf) {
        const float mid = mul_rn(add_rn(t, t1), 0.5f);
        float p[3], m = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          p[a] = add_rn(o[a], mul_rn(d[a], mid));
          m = fmaxf(m, div_rn(fabsf(sub_rn(p[a], mp.centre[a])), mp.half[a]));
        }
        int lvl = 0;
        while (lvl < mp.levels && m > (float)(1 << lvl)) ++lvl;
        if (lvl < mp.levels) {
          const float scale = (float)(1 << lvl);
          int cell[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float q = mul_rn(add_rn(div_rn(sub_rn(p[a], mp.centre[a]), mul_rn(mp.half[a], scale)), 1.f), 0.5f);
            cell[a] = min(max((int)floorf(mul_rn(q, (float)res)), 0), res - 1);
          }
          hit = binaries[(((size_t)lvl * res + cell[0]) * res + cell[1]) * res + cell[2]] != 0;
        }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (FILL && hit) {
        const int64_t dst = out + n + __popc(mask & ((1u << lane) - 1u));
        ray_indices[dst] = r, t_starts[dst] = t, t_ends[dst] = t1;
      }
      n += __popc(mask);
      base = __shfl_sync(0xffffffffu, t1, 31);  // the 33rd interval starts where lane 31's ended
    }
  }
  if (!FILL && lane == 0) counts[r] = n;
}

static int g_march_warp = 1;
int b2n_tune_packed(const char* key, int value) {
  if (!strcmp(key, "march_warp")) { g_march_warp = value; return 1; }
  return 0;
}

static int fill_march(March& mp, int levels, int res, const float* roi, float step, float cone, float near_plane, float far_plane) {
  if (levels < 1 || levels > 16 || res < 1 || !roi || !(step > 0.f)) return -1;
  mp.levels = levels, mp.res = res, mp.step = step, mp.cone = cone, mp.near_plane = near_plane, mp.far_plane = far_plane;
  for (int a = 0; a < 3; ++a) {
    mp.centre[a] = (roi[a] + roi[3 + a]) / 2.f;
    mp.half[a] = (roi[3 + a] - roi[a]) / 2.f;
  }
  return 0;
}

extern "C" int b2n_occgrid_count(const float* origins, const float* directions, const float* t_min, const float* t_max,
                                 const uint8_t* binaries, int32_t levels, int32_t res, const float* roi_host6, float step,
                                 float cone_angle, float near_plane, float far_plane, const float* jitter, int64_t n_rays,
                                 int32_t* counts, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(origins && directions && binaries && counts, "null pointer");
  March mp;
  B2N_REQUIRE(fill_march(mp, levels, res, roi_host6, step, cone_angle, near_plane, far_plane) == 0, "bad grid description");
  if (n_rays == 0) return B2N_OK;
  if (g_march_warp)
    occgrid_march_warp_kernel<false><<<(unsigned)div_up(n_rays, 4), 128, 0, (cudaStream_t)stream>>>(
        mp, origins, directions, t_min, t_max, binaries, jitter, n_rays, counts, nullptr, nullptr, nullptr, nullptr);
  else
    occgrid_march_kernel<false><<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
        mp, origins, directions, t_min, t_max, binaries, jitter, n_rays, counts, nullptr, nullptr, nullptr, nullptr);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_occgrid_fill(const float* origins, const float* directions, const float* t_min, const float* t_max,
                                const uint8_t* binaries, int32_t levels, int32_t res, const float* roi_host6, float step,
                                float cone_angle, float near_plane, float far_plane, const float* jitter, int64_t n_rays,
                                const int64_t* offsets, int64_t* ray_indices, float* t_starts, float* t_ends, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(origins && directions && binaries && offsets && ray_indices && t_starts && t_ends, "null pointer");
  March mp;
  B2N_REQUIRE(fill_march(mp, levels, res, roi_host6, step, cone_angle, near_plane, far_plane) == 0, "bad grid description");
  if (n_rays == 0) return B2N_OK;
  if (g_march_warp)
    occgrid_march_warp_kernel<true><<<(unsigned)div_up(n_rays, 4), 128, 0, (cudaStream_t)stream>>>(
        mp, origins, directions, t_min, t_max, binaries, jitter, n_rays, nullptr, offsets, ray_indices, t_starts, t_ends);
  else
    occgrid_march_kernel<true><<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
        mp, origins, directions, t_min, t_max, binaries, jitter, n_rays, nullptr, offsets, ray_indices, t_starts, t_ends);
  B2N_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of per-ray sample counts -> pack offsets + total (one CTA; R is a few thousand rays).
// Replaces the framework cumsum between the count and the fill pass of the march / the pruning.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t* __restrict__ counts, int64_t n,
                                                           int64_t* __restrict__ offsets, int64_t* __restrict__ total) {
  __shared__ long long warp_tot[32], warp_excl[32];
  __shared__ long long carry_s, chunk_tot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const long long v = i < n ? (long long)counts[i] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const long long w = warp_tot[lane];
      long long wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_excl[lane] = wi - w;
      if (lane == 31) chunk_tot = wi;
    }
    __syncthreads();
    if (i < n) offsets[i] = carry_s + warp_excl[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += chunk_tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

// Large inputs (1 M Gaussians' tile counts): three passes — per-block sums (1024 counts per CTA), a one-CTA scan of the
// block sums, then every CTA re-scans its 1024 counts starting from its block offset.
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(const int32_t* __restrict__ counts, int64_t n,
                                                               int32_t* __restrict__ block_sums) {
  __shared__ int warp_tot[32];
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  int v = i < n ? counts[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = warp_tot[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = w;
  }
}
__global__ void __launch_bounds__(1024) scan_apply_kernel(const int32_t* __restrict__ counts, int64_t n,
                                                          const int64_t* __restrict__ block_offsets,
                                                          int64_t* __restrict__ offsets) {
  __shared__ long long warp_tot[32], warp_excl[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  const long long v = i < n ? (long long)counts[i] : 0;
  long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const long long w = warp_tot[lane];
    long long wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_excl[lane] = wi - w;
  }
  __syncthreads();
  if (i < n) offsets[i] = block_offsets[blockIdx.x] + warp_excl[warp] + incl - v;
}

extern "C" int b2n_scan_counts_ws(const int32_t* counts, int64_t n, int64_t* offsets, int64_t* total, int32_t* scratch_sums,
                                  int64_t* scratch_offsets, void* stream) {
  B2N_REQUIRE(total != nullptr, "null pointer");
  B2N_REQUIRE(n == 0 || (counts && offsets), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t blocks = div_up(n, 1024);
  if (n <= 8192 || !scratch_sums || !scratch_offsets) {
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts, n, offsets, total);
    B2N_LAUNCH_CHECK();
  }
  scan_block_sums_kernel<<<(unsigned)blocks, 1024, 0, st>>>(counts, n, scratch_sums);
  scan_counts_kernel<<<1, 1024, 0, st>>>(scratch_sums, blocks, scratch_offsets, total);
  scan_apply_kernel<<<(unsigned)blocks, 1024, 0, st>>>(counts, n, scratch_offsets, offsets);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_scan_counts(const int32_t* counts, int64_t n, int64_t* offsets, int64_t* total, void* stream) {
  B2N_REQUIRE(total != nullptr, "null pointer");
  B2N_REQUIRE(n == 0 || (counts && offsets), "null pointer");
  scan_counts_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(counts, n, offsets, total);
  B2N_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// K8 — visibility / alpha pruning of marched samples (nerfacc render_visibility_from_density as called by
// OccGridEstimator.sampling; model_components/ray_samplers.py:481-493): keep = T >= early_stop_eps && alpha >= alpha_thre,
// order within a ray preserved.  Count and fill share the predicate, so counts/offsets/indices agree bit for bit.
// ------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(PW * 32) packed_prune_kernel(
    const float* __restrict__ trans, const float* __restrict__ alphas, const int64_t* __restrict__ info, int64_t n_rays,
    float eps, float alpha_thre_host, const float* __restrict__ alpha_cap, int32_t* __restrict__ counts,
    const int64_t* __restrict__ offsets,
    const float* __restrict__ ts, const float* __restrict__ te, int64_t* __restrict__ out_ri, float* __restrict__ out_ts,
    float* __restrict__ out_te) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * PW + warp;
  if (r >= n_rays) return;
  const float alpha_thre = alpha_cap ? fminf(alpha_thre_host, __ldg(alpha_cap)) : alpha_thre_host;
  const int64_t base = info[2 * r];
  const int cnt = (int)info[2 * r + 1];
  int64_t out = FILL ? offsets[r] : 0;
  int n = 0;
  for (int i0 = 0; i0 < cnt; i0 += 32) {
    const int i = i0 + lane;
    const bool keep = i < cnt && __ldg(trans + base + i) >= eps && __ldg(alphas + base + i) >= alpha_thre;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (FILL && keep) {
      const int64_t dst = out + n + __popc(m & ((1u << lane) - 1u));
      out_ri[dst] = r, out_ts[dst] = __ldg(ts + base + i), out_te[dst] = __ldg(te + base + i);
    }
    n += __popc(m);
  }
  if (!FILL && lane == 0) counts[r] = n;
}

extern "C" int b2n_packed_prune_count(const float* trans, const float* alphas, const int64_t* packed_info, int64_t n_rays,
                                      float early_stop_eps, float alpha_thre, const float* alpha_cap_dev, int32_t* counts,
                                      void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(trans && alphas && packed_info && counts, "null pointer");
  packed_prune_kernel<false><<<(unsigned)div_up(n_rays, PW), PW * 32, 0, (cudaStream_t)stream>>>(
      trans, alphas, packed_info, n_rays, early_stop_eps, alpha_thre, alpha_cap_dev, counts, nullptr, nullptr, nullptr, nullptr,
      nullptr, nullptr);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_packed_prune_fill(const float* trans, const float* alphas, const int64_t* packed_info, int64_t n_rays,
                                     float early_stop_eps, float alpha_thre, const float* alpha_cap_dev, const int64_t* offsets,
                                     const float* t_starts, const float* t_ends, int64_t* out_ray_indices, float* out_t_starts, float* out_t_ends,
                                     void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(trans && alphas && packed_info && offsets && t_starts && t_ends && out_ray_indices && out_t_starts && out_t_ends,
              "null pointer");
  packed_prune_kernel<true><<<(unsigned)div_up(n_rays, PW), PW * 32, 0, (cudaStream_t)stream>>>(
      trans, alphas, packed_info, n_rays, early_stop_eps, alpha_thre, alpha_cap_dev, nullptr, offsets, t_starts, t_ends, out_ray_indices,
      out_t_starts, out_t_ends);
  B2N_LAUNCH_CHECK();
}

// packed sample midpoints: x[i] = o[ray] + d[ray] * (ts+te)/2 — the positions VolumetricSampler's sigma_fn evaluates
// (model_components/ray_samplers.py:417-427), each op separately rounded like the reference's torch chain.
__global__ void packed_positions_kernel(const float* __restrict__ origins, const float* __restrict__ directions,
                                        const int64_t* __restrict__ ri, const float* __restrict__ ts,
                                        const float* __restrict__ te, int64_t m, float* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t r = ri[i];
  const float mid = div_rn(add_rn(__ldg(ts + i), __ldg(te + i)), 2.f);
#pragma unroll
  for (int a = 0; a < 3; ++a) x[3 * i + a] = add_rn(__ldg(origins + 3 * r + a), mul_rn(__ldg(directions + 3 * r + a), mid));
}

extern "C" int b2n_packed_positions(const float* origins, const float* directions, const int64_t* ray_indices,
                                    const float* t_starts, const float* t_ends, int64_t m, float* x, void* stream) {
  if (m == 0) return B2N_OK;
  B2N_REQUIRE(origins && directions && ray_indices && t_starts && t_ends && x, "null pointer");
  packed_positions_kernel<<<(unsigned)div_up(m, 256), 256, 0, (cudaStream_t)stream>>>(origins, directions, ray_indices, t_starts, t_ends, m, x);
  B2N_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// K9 — occupancy-grid update (nerfacc OccGridEstimator._update as called by
// models/instant_ngp.py:149-164): jittered cell centres -> occ_eval_fn -> EMA max -> threshold -> binaries.
// ------------------------------------------------------------------------------------------------
// x = lo + ((coord + jitter) / res) * (hi - lo), coord = (id / res^2, id / res % res, id % res); ids NULL = 0..n-1
__global__ void occgrid_points_kernel(const int64_t* __restrict__ ids, const float* __restrict__ jitter, int64_t n, int res,
                                      float lo0, float lo1, float lo2, float ex0, float ex1, float ex2, float* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ids ? ids[i] : i;
  const int c[3] = {(int)(id / ((int64_t)res * res)), (int)((id / res) % res), (int)(id % res)};
  const float lo[3] = {lo0, lo1, lo2}, ex[3] = {ex0, ex1, ex2};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float u = div_rn(add_rn((float)c[a], __ldg(jitter + 3 * i + a)), (float)res);
    x[3 * i + a] = add_rn(lo[a], mul_rn(u, ex[a]));
  }
}

extern "C" int b2n_occgrid_points(const int64_t* cell_ids, const float* jitter, int64_t n, int32_t res,
                                  const float* level_aabb_host6, float* x, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(jitter && x && level_aabb_host6 && res >= 1, "bad arguments");
  const float* b = level_aabb_host6;
  occgrid_points_kernel<<<(unsigned)div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
      cell_ids, jitter, n, res, b[0], b[1], b[2], b[3] - b[0], b[4] - b[1], b[5] - b[2], x);
  B2N_LAUNCH_CHECK();
}

// occs[off + id] <- max(occs[off + id] * decay, occ_new) for the sampled cells.  A cell may be sampled more than once
// (uniform + occupied draws overlap); the reference's indexed assignment then keeps an unspecified one of the
// candidates, all computed from the OLD value — here the largest (deterministic).  Three passes: candidates from the
// old values, reset, atomic max (non-negative floats order like their bit patterns).
__global__ void occ_cand_kernel(const float* __restrict__ occs, const int64_t* __restrict__ ids, const float* __restrict__ occ_new,
                                int64_t n, int64_t off, float decay, float* __restrict__ cand) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = off + (ids ? ids[i] : i);
  cand[i] = fmaxf(mul_rn(occs[id], decay), fmaxf(__ldg(occ_new + i), 0.f));
}
__global__ void occ_reset_kernel(float* __restrict__ occs, const int64_t* __restrict__ ids, int64_t n, int64_t off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) occs[off + (ids ? ids[i] : i)] = 0.f;
}
__global__ void occ_max_kernel(float* __restrict__ occs, const int64_t* __restrict__ ids, const float* __restrict__ cand,
                               int64_t n, int64_t off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMax(reinterpret_cast<int*>(occs + off + (ids ? ids[i] : i)), __float_as_int(cand[i]));
}

extern "C" int b2n_occgrid_ema(float* occs, const int64_t* cell_ids, const float* occ_new, int64_t n, int64_t level_offset,
                               float ema_decay, float* scratch_n, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(occs && occ_new && scratch_n, "null pointer");
  const unsigned g = (unsigned)div_up(n, 256);
  cudaStream_t s = (cudaStream_t)stream;
  occ_cand_kernel<<<g, 256, 0, s>>>(occs, cell_ids, occ_new, n, level_offset, ema_decay, scratch_n);
  occ_reset_kernel<<<g, 256, 0, s>>>(occs, cell_ids, n, level_offset);
  occ_max_kernel<<<g, 256, 0, s>>>(occs, cell_ids, scratch_n, n, level_offset);
  B2N_LAUNCH_CHECK();
}

// binaries = occs > min(mean(occs), occ_thre); the mean is a deterministic fp64 two-level reduction.
#define OCC_PARTS 512
__global__ void __launch_bounds__(256) occ_partial_kernel(const float* __restrict__ occs, int64_t n, double* __restrict__ parts) {
  __shared__ double sm[8];
  const int64_t per = div_up_dev(n, (int64_t)OCC_PARTS);
  const int64_t a = (int64_t)blockIdx.x * per, b = min(n, a + per);
  double s = 0.0;
  for (int64_t i = a + threadIdx.x; i < b; i += 256) s += (double)__ldg(occs + i);
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    parts[blockIdx.x] = t;
  }
}
__global__ void occ_threshold_kernel(const double* __restrict__ parts, int64_t n, float occ_thre, float* __restrict__ stats) {
  double t = 0.0;
  for (int i = 0; i < OCC_PARTS; ++i) t += parts[i];
  const float mean = (float)(t / (double)n);
  stats[0] = fminf(mean, occ_thre), stats[1] = mean;
}
__global__ void occ_binarize_kernel(const float* __restrict__ occs, int64_t n, const float* __restrict__ thre, uint8_t* __restrict__ bin) {
  const float t = __ldg(thre);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) bin[i] = occs[i] > t ? 1 : 0;
}

extern "C" int b2n_occgrid_binarize(const float* occs, int64_t n, float occ_thre, double* scratch_parts, float* stats_out,
                                    uint8_t* binaries, void* stream) {
  B2N_REQUIRE(occs && scratch_parts && stats_out && n >= 1, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  occ_partial_kernel<<<OCC_PARTS, 256, 0, s>>>(occs, n, scratch_parts);
  occ_threshold_kernel<<<1, 1, 0, s>>>(scratch_parts, n, occ_thre, stats_out);
  if (binaries)
    occ_binarize_kernel<<<(unsigned)min(div_up(n, 256), (int64_t)b2n_sm_count() * 16), 256, 0, s>>>(occs, n, stats_out, binaries);
  B2N_LAUNCH_CHECK();
}

// nerfacc.ray_aabb_intersect (the reference's proxy: utils/math.py:138-175 intersect_aabb, division by d, no epsilon):
// t_mins/t_maxs [n,K], hits uint8 [n,K]; misses get miss_value.
__global__ void ray_aabb_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ aabbs,
                                int64_t n, int k, float near_plane, float far_plane, float miss, float* __restrict__ tmin,
                                float* __restrict__ tmax, uint8_t* __restrict__ hits) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * k) return;
  const int64_t r = idx / k;
  const int b = (int)(idx - r * k);
  float tn = -INFINITY, tf = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float inv = div_rn(1.f, __ldg(d + 3 * r + a));
    const float t1 = mul_rn(sub_rn(__ldg(aabbs + 6 * b + a), __ldg(o + 3 * r + a)), inv);
    const float t2 = mul_rn(sub_rn(__ldg(aabbs + 6 * b + 3 + a), __ldg(o + 3 * r + a)), inv);
    tn = fmaxf(tn, fminf(t1, t2)), tf = fminf(tf, fmaxf(t1, t2));
  }
  tn = fmaxf(tn, near_plane), tf = fminf(tf, far_plane);
  const bool hit = tf > tn;
  tmin[idx] = hit ? tn : miss, tmax[idx] = hit ? tf : miss, hits[idx] = hit ? 1 : 0;
}

extern "C" int b2n_ray_aabb_intersect(const float* origins, const float* directions, const float* aabbs, int64_t n_rays,
                                      int32_t n_boxes, float near_plane, float far_plane, float miss_value, float* t_mins,
                                      float* t_maxs, uint8_t* hits, void* stream) {
  if (n_rays == 0 || n_boxes == 0) return B2N_OK;
  B2N_REQUIRE(origins && directions && aabbs && t_mins && t_maxs && hits, "null pointer");
  ray_aabb_kernel<<<(unsigned)div_up(n_rays * n_boxes, 256), 256, 0, (cudaStream_t)stream>>>(
      origins, directions, aabbs, n_rays, n_boxes, near_plane, far_plane, miss_value, t_mins, t_maxs, hits);
  B2N_LAUNCH_CHECK();
}
