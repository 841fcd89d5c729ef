// K5/K6 + a8-a10/a14 — SH / frequency encodings, sample positions -> unit cube, density activation.
#include "common.cuh"
#include "positions.cuh"

// ---------------------------------------------------------------------------------------------
// Spherical harmonics, positive-sign basis (nerfstudio/utils/spherical_harmonics.py:24-81)
// ---------------------------------------------------------------------------------------------
template <int LEVELS>
__global__ void sh_fwd_kernel(const float* __restrict__ dirs, int64_t n, int remap01, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = __ldg(dirs + 3 * i), y = __ldg(dirs + 3 * i + 1), z = __ldg(dirs + 3 * i + 2);
  if (remap01) {
    x = div_rn(add_rn(x, 1.f), 2.f), y = div_rn(add_rn(y, 1.f), 2.f), z = div_rn(add_rn(z, 1.f), 2.f);
  }
  const float xx = x * x, yy = y * y, zz = z * z;
  float c[LEVELS * LEVELS];
  c[0] = 0.28209479177387814f;
  if constexpr (LEVELS > 1) {
    c[1] = 0.4886025119029199f * y;
    c[2] = 0.4886025119029199f * z;
    c[3] = 0.4886025119029199f * x;
  }
  if constexpr (LEVELS > 2) {
    c[4] = 1.0925484305920792f * x * y;
    c[5] = 1.0925484305920792f * y * z;
    c[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
    c[7] = 1.0925484305920792f * x * z;
    c[8] = 0.5462742152960396f * (xx - yy);
  }
  if constexpr (LEVELS > 3) {
    c[9] = 0.5900435899266435f * y * (3.f * xx - yy);
    c[10] = 2.890611442640554f * x * y * z;
    c[11] = 0.4570457994644658f * y * (5.f * zz - 1.f);
    c[12] = 0.3731763325901154f * z * (5.f * zz - 3.f);
    c[13] = 0.4570457994644658f * x * (5.f * zz - 1.f);
    c[14] = 1.445305721320277f * z * (xx - yy);
    c[15] = 0.5900435899266435f * x * (xx - 3.f * yy);
  }
  if constexpr (LEVELS > 4) {
    c[16] = 2.5033429417967046f * x * y * (xx - yy);
    c[17] = 1.7701307697799304f * y * z * (3.f * xx - yy);
    c[18] = 0.9461746957575601f * x * y * (7.f * zz - 1.f);
    c[19] = 0.6690465435572892f * y * z * (7.f * zz - 3.f);
    c[20] = 0.10578554691520431f * (35.f * zz * zz - 30.f * zz + 3.f);
    c[21] = 0.6690465435572892f * x * z * (7.f * zz - 3.f);
    c[22] = 0.47308734787878004f * (xx - yy) * (7.f * zz - 1.f);
    c[23] = 1.7701307697799304f * x * z * (xx - 3.f * yy);
    c[24] = 0.6258357354491761f * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
  }
  float* o = out + i * (LEVELS * LEVELS);
  if constexpr ((LEVELS * LEVELS) % 4 == 0) {
#pragma unroll
    for (int j = 0; j < LEVELS * LEVELS; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(c[j], c[j + 1], c[j + 2], c[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < LEVELS * LEVELS; ++j) o[j] = c[j];
  }
}

extern "C" int b2n_sh_fwd(const float* dirs, int64_t n, int32_t levels, int32_t remap01, float* out, void* stream) {
  if (n == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(dirs && out, "null pointer");
  B2N_REQUIRE(levels >= 1 && levels <= 5, "levels must be in 1..5");
  if (n == 0) return B2N_OK;
  const unsigned grid = (unsigned)div_up(n, 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (levels) {
    case 1: sh_fwd_kernel<1><<<grid, 256, 0, st>>>(dirs, n, remap01, out); break;
    case 2: sh_fwd_kernel<2><<<grid, 256, 0, st>>>(dirs, n, remap01, out); break;
    case 3: sh_fwd_kernel<3><<<grid, 256, 0, st>>>(dirs, n, remap01, out); break;
    case 4: sh_fwd_kernel<4><<<grid, 256, 0, st>>>(dirs, n, remap01, out); break;
    default: sh_fwd_kernel<5><<<grid, 256, 0, st>>>(dirs, n, remap01, out); break;
  }
  B2N_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// NeRF frequency encoding (nerfstudio/field_components/encodings.py:148-186)
//   out[.., j] = sin(2*pi*x[d]*f),  j = d*F + fi ; second half = sin(. + pi/2) ; then x if include_input
// ---------------------------------------------------------------------------------------------
struct FreqParams {
  int d, n_freq, include_input;
  float freqs[32];
};

__global__ void freq_fwd_kernel(const __grid_constant__ FreqParams fp, const float* __restrict__ x, int64_t n,
                                float* __restrict__ out) {
  const int half = fp.d * fp.n_freq;
  const int width = 2 * half + (fp.include_input ? fp.d : 0);
  const int64_t total = n * width;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / width;
    const int j = (int)(idx - i * width);
    float v;
    if (j < 2 * half) {
      const int jj = j < half ? j : j - half;
      const int d = jj / fp.n_freq, f = jj - d * fp.n_freq;
      // torch: (2*pi*x) * freq, then + pi/2 for the second half, each separately rounded
      float s = mul_rn(mul_rn(6.283185307179586f, __ldg(x + i * fp.d + d)), fp.freqs[f]);
      if (j >= half) s = add_rn(s, 1.5707963267948966f);
      v = sinf(s);
    } else {
      v = __ldg(x + i * fp.d + (j - 2 * half));
    }
    out[idx] = v;
  }
}

__global__ void freq_bwd_kernel(const __grid_constant__ FreqParams fp, const float* __restrict__ x,
                                const float* __restrict__ dout, int64_t n, float* __restrict__ dx) {
  const int half = fp.d * fp.n_freq;
  const int width = 2 * half + (fp.include_input ? fp.d : 0);
  const int64_t total = n * fp.d;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / fp.d;
    const int d = (int)(idx - i * fp.d);
    const float xv = __ldg(x + idx);
    const float* g = dout + i * width;
    float acc = fp.include_input ? __ldg(g + 2 * half + d) : 0.f;
    for (int f = 0; f < fp.n_freq; ++f) {
      const float w = 6.283185307179586f * fp.freqs[f];
      const float s = mul_rn(mul_rn(6.283185307179586f, xv), fp.freqs[f]);
      acc += w * (__ldg(g + d * fp.n_freq + f) * cosf(s) + __ldg(g + half + d * fp.n_freq + f) * cosf(add_rn(s, 1.5707963267948966f)));
    }
    dx[idx] = acc;
  }
}

static int fill_freq(FreqParams& fp, int d, const float* freqs_host, int n_freq, int include_input) {
  if (d < 1 || n_freq < 1 || n_freq > 32 || !freqs_host) return -1;
  fp.d = d, fp.n_freq = n_freq, fp.include_input = include_input;
  for (int i = 0; i < n_freq; ++i) fp.freqs[i] = freqs_host[i];
  return 0;
}

extern "C" int b2n_freq_fwd(const float* x, int64_t n, int32_t d, const float* freqs_host, int32_t n_freq,
                            int32_t include_input, float* out, void* stream) {
  if (n == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(x && out, "null pointer");
  FreqParams fp;
  B2N_REQUIRE(fill_freq(fp, d, freqs_host, n_freq, include_input) == 0, "bad frequency table");
  if (n == 0) return B2N_OK;
  const int64_t total = n * (2 * d * n_freq + (include_input ? d : 0));
  const unsigned grid = (unsigned)min(div_up(total, 256), (int64_t)b2n_sm_count() * 16);
  freq_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(fp, x, n, out);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_freq_bwd(const float* x, const float* dout, int64_t n, int32_t d, const float* freqs_host,
                            int32_t n_freq, int32_t include_input, float* dx, void* stream) {
  if (n == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(x && dout && dx, "null pointer");
  FreqParams fp;
  B2N_REQUIRE(fill_freq(fp, d, freqs_host, n_freq, include_input) == 0, "bad frequency table");
  if (n == 0) return B2N_OK;
  const unsigned grid = (unsigned)min(div_up(n * d, 256), (int64_t)b2n_sm_count() * 16);
  freq_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(fp, x, dout, n, dx);
  B2N_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// positions: o + d*(s+e)/2 -> L-inf contraction -> (p+2)/4 | aabb normalise -> selector
// every op separately rounded like the reference's chain of torch kernels (the result feeds floor()/ceil()).
// ---------------------------------------------------------------------------------------------
__global__ void positions_fwd_kernel(const __grid_constant__ PosParams pp, const float* __restrict__ origins,
                                     const float* __restrict__ directions, const float* __restrict__ starts,
                                     const float* __restrict__ ends, int64_t bin_stride, int64_t n_rays, int n_samples,
                                     float* __restrict__ x_out, uint8_t* __restrict__ sel_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * n_samples) return;
  const int64_t r = idx / n_samples;
  const int s = (int)(idx - r * n_samples);
  float p[3];
  if (directions != nullptr) {
    frustum_centre(origins + 3 * r, directions + 3 * r, __ldg(starts + r * bin_stride + s), __ldg(ends + r * bin_stride + s), p);
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __ldg(origins + 3 * idx + a);
  }
  const bool sel = unit_cube_point(pp, p);
#pragma unroll
  for (int a = 0; a < 3; ++a) x_out[3 * idx + a] = p[a];
  if (sel_out) sel_out[idx] = sel ? 1 : 0;
}

extern "C" int b2n_positions_fwd(const float* origins, const float* directions, const float* starts,
                                 const float* ends, int64_t bin_stride, int64_t n_rays, int32_t n_samples,
                                 int32_t contraction, const float* aabb_host6, float* x_out, uint8_t* sel_out,
                                 void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(origins && x_out, "null pointer");
  B2N_REQUIRE(directions == nullptr || (starts && ends), "ray form needs starts/ends");
  B2N_REQUIRE(contraction || aabb_host6, "aabb required without contraction");
  B2N_REQUIRE(n_samples >= 1, "n_samples");
  PosParams pp;
  fill_pos_params(pp, contraction, aabb_host6);
  const int64_t total = n_rays * n_samples;
  if (total == 0) return B2N_OK;
  positions_fwd_kernel<<<(unsigned)div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
      pp, origins, directions, starts, ends, bin_stride, n_rays, n_samples, x_out, sel_out);
  B2N_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// density = avg_init * trunc_exp(h) * selector   (activations.py:28-41; nerfacto_field.py:226-232)
// ---------------------------------------------------------------------------------------------
__global__ void density_act_fwd_kernel(const float* __restrict__ h, int64_t h_stride, const uint8_t* __restrict__ sel,
                                       int64_t n, float avg_init, float* __restrict__ density) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = mul_rn(avg_init, expf(__ldg(h + i * h_stride)));
  density[i] = sel ? mul_rn(e, sel[i] ? 1.f : 0.f) : e;
}

__global__ void density_act_bwd_kernel(const float* __restrict__ h, int64_t h_stride, const uint8_t* __restrict__ sel,
                                       const float* __restrict__ g, int64_t n, float avg_init, float* __restrict__ dh,
                                       int64_t dh_stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float hv = fminf(fmaxf(__ldg(h + i * h_stride), -15.f), 15.f);
  const float m = (sel == nullptr || sel[i]) ? 1.f : 0.f;
  dh[i * dh_stride] = __ldg(g + i) * m * avg_init * expf(hv);
}

extern "C" int b2n_density_act_fwd(const float* h, int64_t h_stride, const uint8_t* sel, int64_t n, float avg_init,
                                   float* density, void* stream) {
  if (n == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(h && density, "null pointer");
  if (n == 0) return B2N_OK;
  density_act_fwd_kernel<<<(unsigned)div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(h, h_stride, sel, n, avg_init, density);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_density_act_bwd(const float* h, int64_t h_stride, const uint8_t* sel, const float* g, int64_t n,
                                   float avg_init, float* dh, int64_t dh_stride, void* stream) {
  if (n == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(h && g && dh, "null pointer");
  if (n == 0) return B2N_OK;
  density_act_bwd_kernel<<<(unsigned)div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(h, h_stride, sel, g, n, avg_init, dh, dh_stride);
  B2N_LAUNCH_CHECK();
}


// ---------------------------------------------------------------------------------------------
// positions backward: d loss / d x (unit-cube sample positions, [R*S,3]) -> d origins [R,3], d directions [R,3].
// x = sel * N(C(o + d t)) with t = (start + end)/2, C = L-inf scene contraction (identity inside the unit ball,
// (2 - 1/m) p/m outside, m = |p|_inf: spatial_distortions.py:66-69), N = (x + 2)/4 or the aabb normalisation.  This is the
// link that carries the photometric gradient back to the camera optimiser's pose corrections
// (camera_optimizers.py:148-153) and what the reference's autograd does through Frustums.get_positions (rays.py:50-59).
// One warp per ray: lanes stride over the samples, shuffle-reduce, one store per ray (no atomics).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) positions_bwd_kernel(const __grid_constant__ PosParams pp, const float* __restrict__ origins,
                                                            const float* __restrict__ directions,
                                                            const float* __restrict__ starts, const float* __restrict__ ends,
                                                            int64_t bin_stride, int64_t n_rays, int n_samples,
                                                            const float* __restrict__ dx, int accumulate,
                                                            float* __restrict__ d_origins,
                                                            float* __restrict__ d_directions) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rays) return;
  float o[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) o[a] = __ldg(origins + 3 * r + a), d[a] = __ldg(directions + 3 * r + a);
  float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
  for (int s = lane; s < n_samples; s += 32) {
    const float t = 0.5f * (__ldg(starts + r * bin_stride + s) + __ldg(ends + r * bin_stride + s));
    float p[3], g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = fmaf(d[a], t, o[a]), g[a] = __ldg(dx + 3 * (r * n_samples + s) + a);
    unit_cube_point_bwd(pp, p, g);  // zero outside (0,1)^3: x was multiplied by the selector
#pragma unroll
    for (int a = 0; a < 3; ++a) go[a] += g[a], gd[a] = fmaf(t, g[a], gd[a]);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    go[a] = warp_sum(go[a]), gd[a] = warp_sum(gd[a]);
    if (lane == 0) {  // this warp is the only writer of ray r in this launch
      if (accumulate == 2) {  // ... but another kernel may be adding to the same rays concurrently (forked streams)
        if (d_origins) atomicAdd(d_origins + 3 * r + a, go[a]);
        if (d_directions) atomicAdd(d_directions + 3 * r + a, gd[a]);
      } else {
        if (d_origins) d_origins[3 * r + a] = accumulate ? d_origins[3 * r + a] + go[a] : go[a];
        if (d_directions) d_directions[3 * r + a] = accumulate ? d_directions[3 * r + a] + gd[a] : gd[a];
      }
    }
  }
}

extern "C" int b2n_positions_bwd(const float* origins, const float* directions, const float* starts, const float* ends,
                                 int64_t bin_stride, int64_t n_rays, int32_t n_samples, int32_t contraction,
                                 const float* aabb_host6, const float* dx, int32_t accumulate, float* d_origins,
                                 float* d_directions, void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(origins && directions && starts && ends && dx && (d_origins || d_directions), "null pointer");
  B2N_REQUIRE(contraction || aabb_host6, "aabb required without contraction");
  PosParams pp;
  fill_pos_params(pp, contraction, aabb_host6);
  positions_bwd_kernel<<<(unsigned)div_up(n_rays, 4), 128, 0, (cudaStream_t)stream>>>(
      pp, origins, directions, starts, ends, bin_stride, n_rays, n_samples, dx, accumulate, d_origins, d_directions);
  B2N_LAUNCH_CHECK();
}
