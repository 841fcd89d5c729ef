// Sample position -> unit cube (frustum centre, L-inf contraction, normalisation, selector), shared device code.
// Every op is separately rounded like the reference's chain of torch kernels: the result feeds floor()/ceil().
// Reference: nerfstudio/cameras/rays.py:50-59, field_components/spatial_distortions.py:66-69,
// fields/nerfacto_field.py:205-213, data/scene_box.py:62-71.
#pragma once
#include "common.cuh"

struct PosParams {
  int contraction;
  float lo[3], len[3];
};

static inline void fill_pos_params(PosParams& pp, int contraction, const float* aabb_host6) {
  pp.contraction = contraction;
  for (int a = 0; a < 3; ++a) {
    pp.lo[a] = aabb_host6 ? aabb_host6[a] : 0.f;
    pp.len[a] = aabb_host6 ? aabb_host6[3 + a] - aabb_host6[a] : 1.f;
  }
}

// p: raw position in, unit-cube position (already multiplied by the selector) out; returns the selector
__device__ __forceinline__ bool unit_cube_point(const PosParams& pp, float (&p)[3]) {
  if (pp.contraction) {
    const float mag = fmaxf(fabsf(p[0]), fmaxf(fabsf(p[1]), fabsf(p[2])));
    if (!(mag < 1.f)) {
      const float k = sub_rn(2.f, div_rn(1.f, mag));
#pragma unroll
      for (int a = 0; a < 3; ++a) p[a] = mul_rn(k, div_rn(p[a], mag));
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = div_rn(add_rn(p[a], 2.f), 4.f);
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = div_rn(sub_rn(p[a], pp.lo[a]), pp.len[a]);
  }
  const bool sel = p[0] > 0.f && p[0] < 1.f && p[1] > 0.f && p[1] < 1.f && p[2] > 0.f && p[2] < 1.f;
  const float m = sel ? 1.f : 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) p[a] = mul_rn(p[a], m);  // NaN * 0 = NaN, as in the reference
  return sel;
}

// frustum centre o + d * (start + end) / 2
__device__ __forceinline__ void frustum_centre(const float* __restrict__ o, const float* __restrict__ d, float start,
                                               float end, float (&p)[3]) {
  const float t = add_rn(start, end);
#pragma unroll
  for (int a = 0; a < 3; ++a) p[a] = add_rn(__ldg(o + a), div_rn(mul_rn(__ldg(d + a), t), 2.f));
}

// Backward of unit_cube_point at raw position p: g (gradient w.r.t. the unit-cube position) -> gradient w.r.t. p, in place.
// Selector 0 -> zero; normalisation scale; L-inf contraction Jacobian  c(m) I + c'(m) sign(p_k) p e_k^T  with m = |p_k| the
// max-norm, c(m) = 2/m - 1/m^2 (spatial_distortions.py:66-69; torch.linalg.norm(ord=inf) back-propagates to the arg-max).
__device__ __forceinline__ void unit_cube_point_bwd(const PosParams& pp, const float (&p)[3], float (&g)[3]) {
  float q[3] = {p[0], p[1], p[2]};
  if (!unit_cube_point(pp, q)) {
    g[0] = g[1] = g[2] = 0.f;
    return;
  }
  if (pp.contraction) {
#pragma unroll
    for (int a = 0; a < 3; ++a) g[a] *= 0.25f;
    const float ax = fabsf(p[0]), ay = fabsf(p[1]), az = fabsf(p[2]);
    const float m = fmaxf(ax, fmaxf(ay, az));
    if (!(m < 1.f)) {
      const int k = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
      const float inv = 1.f / m, c = (2.f - inv) * inv, dc = 2.f * inv * inv * (inv - 1.f);
      const float dot = p[0] * g[0] + p[1] * g[1] + p[2] * g[2];
      const float sgn = p[k] < 0.f ? -1.f : 1.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) g[a] *= c;
      g[k] += sgn * dc * dot;
    }
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) g[a] /= pp.len[a];
  }
}
