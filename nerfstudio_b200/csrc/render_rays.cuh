// Per-ray bodies of the rendering / loss kernels (a21-a24), one warp per ray.  render.cu wraps each in its own kernel
// (the C-ABI operators); ray_tail.cu strings them together for the captured training step.  Template flag NC: true =
// the buffer was written by an EARLIER kernel (read-only loads), false = it was produced earlier in the SAME kernel by
// this warp (plain loads after a __syncwarp).  Same arithmetic either way: the fused step is bit-identical.
// Reference rows: cameras/rays.py:129-152, model_components/renderers.py:71-119,292-385, model_components/losses.py:53-155.
#pragma once
#include "common.cuh"

struct Bg {
  int mode, eval_mode;
  float c[3];
};

template <bool NC>
__device__ __forceinline__ float ldv(const float* p) {
  if constexpr (NC) return __ldg(p);
  else return *p;
}

// density = avg_init * trunc_exp(h) * selector   (activations.py:28-41; nerfacto_field.py:226-232)
__device__ __forceinline__ float density_act(float h, const uint8_t* sel, int64_t i, float avg_init) {
  const float e = mul_rn(avg_init, expf(h));
  return sel ? mul_rn(e, sel[i] ? 1.f : 0.f) : e;
}
__device__ __forceinline__ float density_act_grad(float h, const uint8_t* sel, int64_t i, float g, float avg_init) {
  const float hv = fminf(fmaxf(h, -15.f), 15.f);
  const float m = (sel == nullptr || sel[i]) ? 1.f : 0.f;
  return g * m * avg_init * expf(hv);
}

// get_weights for ray r: st/en/d/weights point at the ray's rows
template <bool NC>
__device__ __forceinline__ void weights_fwd_ray(const float* st, const float* en, const float* d, int S, float* weights, int lane) {
  const int chunk = (S + 31) / 32, i0 = lane * chunk, i1 = min(S, i0 + chunk);
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += (double)mul_rn(sub_rn(__ldg(en + i), __ldg(st + i)), ldv<NC>(d + i));
  double run = warp_scan_incl_d(local, lane) - local;  // exclusive prefix of this chunk
  for (int i = i0; i < i1; ++i) {
    const float dd = mul_rn(sub_rn(__ldg(en + i), __ldg(st + i)), ldv<NC>(d + i));
    const float alpha = sub_rn(1.f, expf(-dd));
    const float T = expf(-(float)run);
    weights[i] = nan_to_num(mul_rn(alpha, T));
    run += (double)dd;
  }
}

// dL/d(dd_k) = g_k T_k (1 - a_k) - sum_{i>k} g_i a_i T_i ;  dsigma_k = delta_k * that.   OUT(i, value) receives d density_i.
template <bool NC_D, bool NC_G, class Out>
__device__ __forceinline__ void weights_bwd_ray(const float* st, const float* en, const float* d, const float* g, int S, int lane,
                                                Out out) {
  const int chunk = (S + 31) / 32, i0 = lane * chunk, i1 = min(S, i0 + chunk);
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += (double)mul_rn(sub_rn(__ldg(en + i), __ldg(st + i)), ldv<NC_D>(d + i));
  const double excl = warp_scan_incl_d(local, lane) - local;
  // pass A: per-chunk sum of g_i * w_i (finite only), to build the suffix sums
  double run = excl, gw_local = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float dd = mul_rn(sub_rn(__ldg(en + i), __ldg(st + i)), ldv<NC_D>(d + i));
    const float w = mul_rn(sub_rn(1.f, expf(-dd)), expf(-(float)run));
    if (isfinite(w)) gw_local += (double)(ldv<NC_G>(g + i) * w);
    run += (double)dd;
  }
  const double gw_incl = warp_scan_incl_d(gw_local, lane);
  const double gw_total = __shfl_sync(0xffffffffu, gw_incl, 31);
  double suffix = gw_total - gw_incl;  // sum over chunks after this one
  // pass B: walk the chunk backwards
  run = excl + local;
  for (int i = i1 - 1; i >= i0; --i) {
    const float delta = sub_rn(__ldg(en + i), __ldg(st + i));
    const float dd = mul_rn(delta, ldv<NC_D>(d + i));
    run -= (double)dd;  // exclusive prefix at i (up to fp64 rounding)
    const float ea = expf(-dd), T = expf(-(float)run);
    const float w = mul_rn(sub_rn(1.f, ea), T);
    const float gi = isfinite(w) ? ldv<NC_G>(g + i) : 0.f;
    const float grad_dd = gi * T * ea - (float)suffix;
    out(i, delta * grad_dd);
    suffix += (double)(gi * w);
  }
}

struct CompositeOut {
  float r, g, b, acc, depth_exp, depth_med;  // valid on lane 0 (depth_med only if requested)
  int med_idx;
};

// rgb/w/st/en point at the ray's rows (rgb, st, en may be NULL)
template <bool NC_W>
__device__ __forceinline__ CompositeOut composite_fwd_ray(const Bg& bg, const float* rgb, const float* w, const float* st,
                                                          const float* en, int S, bool want_median, int lane) {
  const int chunk = (S + 31) / 32, i0 = lane * chunk, i1 = min(S, i0 + chunk);
  float cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f, wt = 0.f;
  double wl = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float wi = ldv<NC_W>(w + i);
    acc += wi;
    wl += (double)wi;
    if (rgb) {
      float a = __ldg(rgb + i * 3), b = __ldg(rgb + i * 3 + 1), c = __ldg(rgb + i * 3 + 2);
      if (bg.eval_mode) a = nan_to_num(a), b = nan_to_num(b), c = nan_to_num(c);
      cr = fmaf(wi, a, cr), cg = fmaf(wi, b, cg), cb = fmaf(wi, c, cb);
    }
    if (st) wt = fmaf(wi, div_rn(add_rn(__ldg(st + i), __ldg(en + i)), 2.f), wt);
  }
  CompositeOut o;
  o.med_idx = 0, o.depth_med = 0.f;
  // median: first index whose inclusive cumulative weight >= 0.5 (searchsorted left), clamped to S-1
  if (want_median) {
    double run = warp_scan_incl_d(wl, lane) - wl;
    int found = S;  // sentinel
    for (int i = i0; i < i1; ++i) {
      run += (double)ldv<NC_W>(w + i);
      if (found == S && (float)run >= 0.5f) found = i;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, s));
    o.med_idx = min(found, S - 1);
    if (lane == 0 && st) o.depth_med = div_rn(add_rn(__ldg(st + o.med_idx), __ldg(en + o.med_idx)), 2.f);
  }
  cr = warp_sum(cr), cg = warp_sum(cg), cb = warp_sum(cb), acc = warp_sum(acc), wt = warp_sum(wt);
  if (lane == 0 && rgb) {
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (bg.mode == B2N_BG_LAST_SAMPLE) {
      b0 = __ldg(rgb + (S - 1) * 3), b1 = __ldg(rgb + (S - 1) * 3 + 1), b2 = __ldg(rgb + (S - 1) * 3 + 2);
      if (bg.eval_mode) b0 = nan_to_num(b0), b1 = nan_to_num(b1), b2 = nan_to_num(b2);
    } else if (bg.mode == B2N_BG_CONSTANT) {
      b0 = bg.c[0], b1 = bg.c[1], b2 = bg.c[2];
    }
    if (bg.mode != B2N_BG_NONE) {
      const float k = 1.f - acc;
      cr += b0 * k, cg += b1 * k, cb += b2 * k;
    }
    if (bg.eval_mode) cr = fminf(fmaxf(cr, 0.f), 1.f), cg = fminf(fmaxf(cg, 0.f), 1.f), cb = fminf(fmaxf(cb, 0.f), 1.f);
  }
  o.r = cr, o.g = cg, o.b = cb, o.acc = acc, o.depth_exp = wt / (acc + 1e-10f);
  return o;
}

// g0..g2 = d loss / d rgb_out, ga = d / d accumulation, gd = d / d expected depth (st/en needed only if has_depth).
// d_rgb / d_w point at the ray's rows (either may be NULL); d_w_add (optional, same layout) is added to d_w.
template <bool NC_W>
__device__ __forceinline__ void composite_bwd_ray(const Bg& bg, const float* rgb, const float* w, const float* st, const float* en,
                                                  float g0, float g1, float g2, float ga, float gd, bool has_depth, int S,
                                                  float* d_rgb, float* d_w, const float* d_w_add, int lane) {
  float acc = 0.f, wt = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float wi = ldv<NC_W>(w + i);
    acc += wi;
    if (has_depth) wt = fmaf(wi, div_rn(add_rn(__ldg(st + i), __ldg(en + i)), 2.f), wt);
  }
  acc = warp_sum(acc), wt = warp_sum(wt);
  float bgdot = 0.f;
  if (bg.mode == B2N_BG_LAST_SAMPLE)
    bgdot = __ldg(rgb + (S - 1) * 3) * g0 + __ldg(rgb + (S - 1) * 3 + 1) * g1 + __ldg(rgb + (S - 1) * 3 + 2) * g2;
  else if (bg.mode == B2N_BG_CONSTANT)
    bgdot = bg.c[0] * g0 + bg.c[1] * g1 + bg.c[2] * g2;
  const float denom = acc + 1e-10f, D = wt / denom;
  for (int i = lane; i < S; i += 32) {
    const float wi = ldv<NC_W>(w + i);
    const float a = __ldg(rgb + i * 3), b = __ldg(rgb + i * 3 + 1), c = __ldg(rgb + i * 3 + 2);
    float k = wi;
    if (bg.mode == B2N_BG_LAST_SAMPLE && i == S - 1) k += 1.f - acc;
    if (d_rgb) d_rgb[i * 3] = k * g0, d_rgb[i * 3 + 1] = k * g1, d_rgb[i * 3 + 2] = k * g2;
    if (d_w) {
      float gw = a * g0 + b * g1 + c * g2 - bgdot + ga;
      if (has_depth) {
        const float t = div_rn(add_rn(__ldg(st + i), __ldg(en + i)), 2.f);
        gw += gd * (t - D) / denom;
      }
      if (d_w_add) gw = add_rn(gw, d_w_add[i]);  // same lane wrote d_w_add[i]
      d_w[i] = gw;
    }
  }
}

// searchsorted(a[0..n), v, right=True): number of entries <= v
__device__ __forceinline__ int upper_bound(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// interlevel (proposal) loss of one ray: c [Sc+1] / w [Sc] the main level's edges and weights, cp [Sp+1] / wp [Sp] the
// proposal's; scratch = 3 (Sp + 2) floats of this warp's shared memory.  Returns the ray's loss (all lanes); d_wp [Sp] optional.
template <bool NC_W>
__device__ __forceinline__ float interlevel_ray(const float* c, const float* w, const float* cp, const float* wp, int Sc, int Sp,
                                                float gscale, float* d_wp, float* scratch, int lane) {
  float* t1 = scratch;           // proposal edges [Sp+1]
  float* cy = t1 + (Sp + 2);     // [Sp+1] = [0, cumsum(wp)]
  float* diff = cy + (Sp + 2);   // [Sp+1] difference array for the gradient
  for (int i = lane; i <= Sp; i += 32) t1[i] = __ldg(cp + i), diff[i] = 0.f;
  const int chunk = (Sp + 31) / 32, i0 = lane * chunk, i1 = min(Sp, i0 + chunk);
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += (double)__ldg(wp + i);
  double run = warp_scan_incl_d(local, lane) - local;
  for (int i = i0; i < i1; ++i) {
    run += (double)__ldg(wp + i);
    cy[i + 1] = (float)run;
  }
  if (lane == 0) cy[0] = 0.f;
  __syncwarp();
  float loss = 0.f;
  for (int i = lane; i < Sc; i += 32) {
    const float t0s = __ldg(c + i), t0e = __ldg(c + i + 1), wi = ldv<NC_W>(w + i);
    int lo = upper_bound(t1, Sp, t0s) - 1;         // over t1_starts = cp[0..Sp)
    lo = min(max(lo, 0), Sp - 1);
    int hi = upper_bound(t1 + 1, Sp, t0e);         // over t1_ends = cp[1..Sp]
    hi = min(max(hi, 0), Sp - 1);
    const float w_outer = cy[hi + 1] - cy[lo];
    const float ex = fmaxf(wi - w_outer, 0.f);
    loss += ex * ex / (wi + 1.0e-7f);
    if (d_wp && ex > 0.f) {
      // d/dw_outer = -2 ex/(w+eps);  d w_outer / d wp_j = [j <= hi] - [j < lo]   (cy1[hi+1] - cy1[lo])
      const float g = -2.f * ex / (wi + 1.0e-7f) * gscale;
      atomicAdd(diff + hi, g);                 // suffix contribution: all j <= hi
      if (lo > 0) atomicAdd(diff + lo - 1, -g);  // remove j <= lo-1
    }
  }
  loss = warp_sum(loss);
  if (d_wp) {
    __syncwarp();
    // d_wp[j] = sum_{k >= j} diff[k]  (suffix sum)
    double sl = 0.0;
    for (int i = i0; i < i1; ++i) sl += (double)diff[i];
    const double incl = warp_scan_incl_d(sl, lane);
    const double total = __shfl_sync(0xffffffffu, incl, 31);
    double suffix = total - incl;  // chunks after mine
    for (int i = i1 - 1; i >= i0; --i) {
      suffix += (double)diff[i];
      d_wp[i] = (float)suffix;
    }
  }
  __syncwarp();  // scratch may be reused
  return loss;
}

// distortion loss of one ray: t [S+1] edges, w [S]; scratch = 2 S floats.  Returns the ray's loss; d_w [S] optional.
template <bool NC_W>
__device__ __forceinline__ float distortion_ray(const float* t, const float* w, int S, float gscale, float* d_w, float* scratch,
                                                int lane) {
  float* ut = scratch;
  float* ws = ut + S;
  for (int i = lane; i < S; i += 32) {
    ut[i] = div_rn(add_rn(__ldg(t + i + 1), __ldg(t + i)), 2.f);
    ws[i] = ldv<NC_W>(w + i);
  }
  __syncwarp();
  float loss = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float ui = ut[i], wi = ws[i];
    float inner = 0.f;
    for (int j = 0; j < S; ++j) inner = fmaf(ws[j], fabsf(ui - ut[j]), inner);
    const float delta = sub_rn(__ldg(t + i + 1), __ldg(t + i));
    loss += wi * inner + wi * wi * delta / 3.f;
    if (d_w) d_w[i] = gscale * (2.f * inner + 2.f * wi * delta / 3.f);
  }
  loss = warp_sum(loss);
  __syncwarp();  // scratch may be reused
  return loss;
}
