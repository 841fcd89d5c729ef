// K7-K12 — packed (instant-ngp) path: occupancy-grid marching, pack_info, packed transmittance weights,
// per-ray accumulation.  These replace the nerfacc 0.5.2 calls of the reference
// (nerfstudio/models/instant_ngp.py:120-198, model_components/ray_samplers.py:481-493,
// model_components/renderers.py:97-102,312-314,371-376).  nerfacc's source is not available: semantics follow
// its published behaviour and are restated in oracle/nerf_oracle.py ("parity unpinned").
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
  occgrid_march_kernel<true><<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
      mp, origins, directions, t_min, t_max, binaries, jitter, n_rays, nullptr, offsets, ray_indices, t_starts, t_ends);
  B2N_LAUNCH_CHECK();
}
