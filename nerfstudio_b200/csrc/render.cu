// a21-a24 — transmittance weights (warp-shuffle segmented scan), alpha compositing, depth, proposal losses.
//
// One warp owns one ray.  A lane holds a contiguous chunk of ceil(S/32) samples; running sums are fp64
// (torch's CPU cumsum accumulates in double and rounds per element — nerfstudio/cameras/rays.py:129-152),
// chunk totals are combined with a shuffle scan.  Reference rows: cameras/rays.py:129-152,
// model_components/renderers.py:71-119,292-385, model_components/losses.py:53-155.
#include "common.cuh"

#define RW 4  // warps (rays) per CTA

// ------------------------------------------------------------------------------------------------
// get_weights
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RW * 32) weights_fwd_kernel(const float* __restrict__ starts,
                                                              const float* __restrict__ ends, int64_t bs,
                                                              const float* __restrict__ density, int64_t n_rays, int S,
                                                              float* __restrict__ weights) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * RW + warp;
  if (r >= n_rays) return;
  const float* st = starts + r * bs;
  const float* en = ends + r * bs;
  const float* d = density + r * S;
  const int chunk = (S + 31) / 32, i0 = lane * chunk, i1 = min(S, i0 + chunk);
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += (double)mul_rn(sub_rn(__ldg(en + i), __ldg(st + i)), __ldg(d + i));
  double run = warp_scan_incl_d(local, lane) - local;  // exclusive prefix of this chunk
  for (int i = i0; i < i1; ++i) {
    const float dd = mul_rn(sub_rn(__ldg(en + i), __ldg(st + i)), __ldg(d + i));
    const float alpha = sub_rn(1.f, expf(-dd));
    const float T = expf(-(float)run);
    weights[r * S + i] = nan_to_num(mul_rn(alpha, T));
    run += (double)dd;
  }
}

// dL/d(dd_k) = g_k T_k (1 - a_k) - sum_{i>k} g_i a_i T_i ;  dsigma_k = delta_k * that.
__global__ void __launch_bounds__(RW * 32) weights_bwd_kernel(const float* __restrict__ starts,
                                                              const float* __restrict__ ends, int64_t bs,
                                                              const float* __restrict__ density,
                                                              const float* __restrict__ dweights, int64_t n_rays, int S,
                                                              float* __restrict__ ddensity) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * RW + warp;
  if (r >= n_rays) return;
  const float* st = starts + r * bs;
  const float* en = ends + r * bs;
  const float* d = density + r * S;
  const float* g = dweights + r * S;
  const int chunk = (S + 31) / 32, i0 = lane * chunk, i1 = min(S, i0 + chunk);
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += (double)mul_rn(sub_rn(__ldg(en + i), __ldg(st + i)), __ldg(d + i));
  const double excl = warp_scan_incl_d(local, lane) - local;
  // pass A: per-chunk sum of g_i * w_i (finite only), to build the suffix sums
  double run = excl, gw_local = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float dd = mul_rn(sub_rn(__ldg(en + i), __ldg(st + i)), __ldg(d + i));
    const float w = mul_rn(sub_rn(1.f, expf(-dd)), expf(-(float)run));
    if (isfinite(w)) gw_local += (double)(__ldg(g + i) * w);
    run += (double)dd;
  }
  const double gw_incl = warp_scan_incl_d(gw_local, lane);
  const double gw_total = __shfl_sync(0xffffffffu, gw_incl, 31);
  double suffix = gw_total - gw_incl;  // sum over chunks after this one
  // pass B: walk the chunk backwards
  run = excl + local;
  for (int i = i1 - 1; i >= i0; --i) {
    const float delta = sub_rn(__ldg(en + i), __ldg(st + i));
    const float dd = mul_rn(delta, __ldg(d + i));
    run -= (double)dd;  // exclusive prefix at i (up to fp64 rounding)
    const float ea = expf(-dd), T = expf(-(float)run);
    const float w = mul_rn(sub_rn(1.f, ea), T);
    const float gi = isfinite(w) ? __ldg(g + i) : 0.f;
    const float grad_dd = gi * T * ea - (float)suffix;
    ddensity[r * S + i] = delta * grad_dd;
    suffix += (double)(gi * w);
  }
}

extern "C" int b2n_weights_fwd(const float* starts, const float* ends, int64_t bin_stride, const float* density,
                               int64_t n_rays, int32_t n_samples, float* weights, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(starts && ends && density && weights, "null pointer");
  B2N_REQUIRE(n_samples >= 1, "n_samples");
  if (n_rays == 0) return B2N_OK;
  weights_fwd_kernel<<<(unsigned)div_up(n_rays, RW), RW * 32, 0, (cudaStream_t)stream>>>(starts, ends, bin_stride, density, n_rays, n_samples, weights);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_weights_bwd(const float* starts, const float* ends, int64_t bin_stride, const float* density,
                               const float* dweights, int64_t n_rays, int32_t n_samples, float* ddensity, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(starts && ends && density && dweights && ddensity, "null pointer");
  B2N_REQUIRE(n_samples >= 1, "n_samples");
  if (n_rays == 0) return B2N_OK;
  weights_bwd_kernel<<<(unsigned)div_up(n_rays, RW), RW * 32, 0, (cudaStream_t)stream>>>(starts, ends, bin_stride, density, dweights, n_rays, n_samples, ddensity);
  B2N_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// compositing: rgb, accumulation, expected depth, median depth
// ------------------------------------------------------------------------------------------------
struct Bg {
  int mode, eval_mode;
  float c[3];
};

__global__ void __launch_bounds__(RW * 32) composite_fwd_kernel(const __grid_constant__ Bg bg,
                                                                const float* __restrict__ rgb,
                                                                const float* __restrict__ weights,
                                                                const float* __restrict__ starts,
                                                                const float* __restrict__ ends, int64_t bs,
                                                                int64_t n_rays, int S,
                                                                float* __restrict__ rgb_out, float* __restrict__ acc_out,
                                                                float* __restrict__ depth_exp,
                                                                float* __restrict__ depth_med,
                                                                int64_t* __restrict__ med_idx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * RW + warp;
  if (r >= n_rays) return;
  const float* w = weights + r * S;
  const int chunk = (S + 31) / 32, i0 = lane * chunk, i1 = min(S, i0 + chunk);
  float cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f, wt = 0.f;
  double wl = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float wi = __ldg(w + i);
    acc += wi;
    wl += (double)wi;
    if (rgb) {
      float a = __ldg(rgb + (r * S + i) * 3), b = __ldg(rgb + (r * S + i) * 3 + 1), c = __ldg(rgb + (r * S + i) * 3 + 2);
      if (bg.eval_mode) a = nan_to_num(a), b = nan_to_num(b), c = nan_to_num(c);
      cr = fmaf(wi, a, cr), cg = fmaf(wi, b, cg), cb = fmaf(wi, c, cb);
    }
    if (starts) wt = fmaf(wi, div_rn(add_rn(__ldg(starts + r * bs + i), __ldg(ends + r * bs + i)), 2.f), wt);
  }
  // median: first index whose inclusive cumulative weight >= 0.5 (searchsorted left), clamped to S-1
  if (depth_med || med_idx) {
    double run = warp_scan_incl_d(wl, lane) - wl;
    int found = S;  // sentinel
    for (int i = i0; i < i1; ++i) {
      run += (double)__ldg(w + i);
      if (found == S && (float)run >= 0.5f) found = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
    const int idx = min(found, S - 1);
    if (lane == 0) {
      if (med_idx) med_idx[r] = idx;
      if (depth_med && starts)
        depth_med[r] = div_rn(add_rn(__ldg(starts + r * bs + idx), __ldg(ends + r * bs + idx)), 2.f);
    }
  }
  cr = warp_sum(cr), cg = warp_sum(cg), cb = warp_sum(cb), acc = warp_sum(acc), wt = warp_sum(wt);
  if (lane == 0) {
    if (rgb_out && rgb) {
      float b0 = 0.f, b1 = 0.f, b2 = 0.f;
      if (bg.mode == B2N_BG_LAST_SAMPLE) {
        b0 = __ldg(rgb + (r * S + S - 1) * 3), b1 = __ldg(rgb + (r * S + S - 1) * 3 + 1), b2 = __ldg(rgb + (r * S + S - 1) * 3 + 2);
        if (bg.eval_mode) b0 = nan_to_num(b0), b1 = nan_to_num(b1), b2 = nan_to_num(b2);
      } else if (bg.mode == B2N_BG_CONSTANT) {
        b0 = bg.c[0], b1 = bg.c[1], b2 = bg.c[2];
      }
      if (bg.mode != B2N_BG_NONE) {
        const float k = 1.f - acc;
        cr += b0 * k, cg += b1 * k, cb += b2 * k;
      }
      if (bg.eval_mode) cr = fminf(fmaxf(cr, 0.f), 1.f), cg = fminf(fmaxf(cg, 0.f), 1.f), cb = fminf(fmaxf(cb, 0.f), 1.f);
      rgb_out[3 * r] = cr, rgb_out[3 * r + 1] = cg, rgb_out[3 * r + 2] = cb;
    }
    if (acc_out) acc_out[r] = acc;
    if (depth_exp) depth_exp[r] = wt / (acc + 1e-10f);
  }
}

__global__ void __launch_bounds__(RW * 32) composite_bwd_kernel(const __grid_constant__ Bg bg,
                                                                const float* __restrict__ rgb,
                                                                const float* __restrict__ weights,
                                                                const float* __restrict__ starts,
                                                                const float* __restrict__ ends, int64_t bs,
                                                                const float* __restrict__ d_rgb_out,
                                                                const float* __restrict__ d_acc,
                                                                const float* __restrict__ d_depth, int64_t n_rays, int S,
                                                                float* __restrict__ d_rgb, float* __restrict__ d_w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * RW + warp;
  if (r >= n_rays) return;
  const float* w = weights + r * S;
  float acc = 0.f, wt = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float wi = __ldg(w + i);
    acc += wi;
    if (d_depth) wt = fmaf(wi, div_rn(add_rn(__ldg(starts + r * bs + i), __ldg(ends + r * bs + i)), 2.f), wt);
  }
  acc = warp_sum(acc), wt = warp_sum(wt);
  const float g0 = d_rgb_out ? __ldg(d_rgb_out + 3 * r) : 0.f, g1 = d_rgb_out ? __ldg(d_rgb_out + 3 * r + 1) : 0.f,
              g2 = d_rgb_out ? __ldg(d_rgb_out + 3 * r + 2) : 0.f;
  const float ga = d_acc ? __ldg(d_acc + r) : 0.f;
  const float gd = d_depth ? __ldg(d_depth + r) : 0.f;
  float bgdot = 0.f;
  if (bg.mode == B2N_BG_LAST_SAMPLE)
    bgdot = __ldg(rgb + (r * S + S - 1) * 3) * g0 + __ldg(rgb + (r * S + S - 1) * 3 + 1) * g1 + __ldg(rgb + (r * S + S - 1) * 3 + 2) * g2;
  else if (bg.mode == B2N_BG_CONSTANT)
    bgdot = bg.c[0] * g0 + bg.c[1] * g1 + bg.c[2] * g2;
  const float denom = acc + 1e-10f, D = wt / denom;
  for (int i = lane; i < S; i += 32) {
    const float wi = __ldg(w + i);
    const float a = __ldg(rgb + (r * S + i) * 3), b = __ldg(rgb + (r * S + i) * 3 + 1), c = __ldg(rgb + (r * S + i) * 3 + 2);
    float k = wi;
    if (bg.mode == B2N_BG_LAST_SAMPLE && i == S - 1) k += 1.f - acc;
    if (d_rgb) d_rgb[(r * S + i) * 3] = k * g0, d_rgb[(r * S + i) * 3 + 1] = k * g1, d_rgb[(r * S + i) * 3 + 2] = k * g2;
    if (d_w) {
      float gw = a * g0 + b * g1 + c * g2 - bgdot + ga;
      if (d_depth) {
        const float t = div_rn(add_rn(__ldg(starts + r * bs + i), __ldg(ends + r * bs + i)), 2.f);
        gw += gd * (t - D) / denom;
      }
      d_w[r * S + i] = gw;
    }
  }
}

static void fill_bg(Bg& bg, int mode, const float* c, int eval_mode) {
  bg.mode = mode, bg.eval_mode = eval_mode;
  for (int i = 0; i < 3; ++i) bg.c[i] = (mode == B2N_BG_CONSTANT && c) ? c[i] : 0.f;
}

extern "C" int b2n_composite_fwd(const float* rgb, const float* weights, const float* starts, const float* ends,
                                 int64_t bin_stride, int64_t n_rays, int32_t n_samples, int32_t bg_mode, const float* bg_host3, int32_t eval_mode,
                                 float* rgb_out, float* acc, float* depth_exp, float* depth_med, int64_t* med_idx,
                                 void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(weights, "null weights");
  B2N_REQUIRE(!(rgb_out && !rgb), "rgb_out needs rgb");
  B2N_REQUIRE(!((depth_exp || depth_med) && !(starts && ends)), "depth needs starts/ends");
  B2N_REQUIRE((starts == nullptr) == (ends == nullptr), "starts/ends must come together");
  B2N_REQUIRE(n_samples >= 1, "n_samples");
  B2N_REQUIRE(bg_mode != B2N_BG_CONSTANT || bg_host3, "constant background needs a colour");
  if (n_rays == 0) return B2N_OK;
  Bg bg;
  fill_bg(bg, bg_mode, bg_host3, eval_mode);
  composite_fwd_kernel<<<(unsigned)div_up(n_rays, RW), RW * 32, 0, (cudaStream_t)stream>>>(
      bg, rgb, weights, starts, ends, bin_stride, n_rays, n_samples, rgb_out, acc, depth_exp, depth_med, med_idx);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_composite_bwd(const float* rgb, const float* weights, const float* starts, const float* ends,
                                 int64_t bin_stride, const float* d_rgb_out,
                                 const float* d_acc, const float* d_depth_exp, int64_t n_rays, int32_t n_samples,
                                 int32_t bg_mode, const float* bg_host3, float* d_rgb, float* d_weights, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(rgb && weights, "null pointer");
  B2N_REQUIRE(!(d_depth_exp && !(starts && ends)), "depth grad needs starts/ends");
  B2N_REQUIRE(bg_mode != B2N_BG_CONSTANT || bg_host3, "constant background needs a colour");
  if (n_rays == 0) return B2N_OK;
  Bg bg;
  fill_bg(bg, bg_mode, bg_host3, 0);
  composite_bwd_kernel<<<(unsigned)div_up(n_rays, RW), RW * 32, 0, (cudaStream_t)stream>>>(
      bg, rgb, weights, starts, ends, bin_stride, d_rgb_out, d_acc, d_depth_exp, n_rays, n_samples, d_rgb, d_weights);
  B2N_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// proposal losses
// ------------------------------------------------------------------------------------------------
// searchsorted(a[0..n), v, right=True): number of entries <= v
__device__ __forceinline__ int upper_bound(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(RW * 32) interlevel_kernel(const float* __restrict__ c, const float* __restrict__ w,
                                                             const float* __restrict__ cp, const float* __restrict__ wp,
                                                             int64_t n_rays, int Sc, int Sp, float gscale,
                                                             float* __restrict__ loss_rows, float* __restrict__ d_wp) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * RW + warp;
  if (r >= n_rays) return;
  float* t1 = sm + (size_t)warp * (3 * (Sp + 2));  // proposal edges [Sp+1]
  float* cy = t1 + (Sp + 2);                       // [Sp+1] = [0, cumsum(wp)]
  float* diff = cy + (Sp + 2);                     // [Sp+1] difference array for the gradient
  for (int i = lane; i <= Sp; i += 32) t1[i] = __ldg(cp + r * (Sp + 1) + i), diff[i] = 0.f;
  const int chunk = (Sp + 31) / 32, i0 = lane * chunk, i1 = min(Sp, i0 + chunk);
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += (double)__ldg(wp + r * Sp + i);
  double run = warp_scan_incl_d(local, lane) - local;
  for (int i = i0; i < i1; ++i) {
    run += (double)__ldg(wp + r * Sp + i);
    cy[i + 1] = (float)run;
  }
  if (lane == 0) cy[0] = 0.f;
  __syncwarp();
  float loss = 0.f;
  for (int i = lane; i < Sc; i += 32) {
    const float t0s = __ldg(c + r * (Sc + 1) + i), t0e = __ldg(c + r * (Sc + 1) + i + 1), wi = __ldg(w + r * Sc + i);
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
  if (lane == 0 && loss_rows) loss_rows[r] = loss;
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
      d_wp[r * Sp + i] = (float)suffix;
    }
  }
}

extern "C" int b2n_interlevel_fwd_bwd(const float* c, const float* w, const float* cp, const float* wp, int64_t n_rays,
                                      int32_t sc, int32_t sp, float gscale, float* loss_rows, float* d_wp,
                                      void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(c && w && cp && wp, "null pointer");
  B2N_REQUIRE(sc >= 1 && sp >= 1 && sp <= 4096, "sample counts out of range");
  if (n_rays == 0) return B2N_OK;
  const size_t smem = sizeof(float) * RW * 3 * (sp + 2);
  if (smem > 48 * 1024) cudaFuncSetAttribute(interlevel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  interlevel_kernel<<<(unsigned)div_up(n_rays, RW), RW * 32, smem, (cudaStream_t)stream>>>(c, w, cp, wp, n_rays, sc, sp,
                                                                                          gscale, loss_rows, d_wp);
  B2N_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(RW * 32) distortion_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                                             int64_t n_rays, int S, float gscale,
                                                             float* __restrict__ loss_rows, float* __restrict__ d_w) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * RW + warp;
  if (r >= n_rays) return;
  float* ut = sm + (size_t)warp * 2 * S;
  float* ws = ut + S;
  for (int i = lane; i < S; i += 32) {
    ut[i] = div_rn(add_rn(__ldg(t + r * (S + 1) + i + 1), __ldg(t + r * (S + 1) + i)), 2.f);
    ws[i] = __ldg(w + r * S + i);
  }
  __syncwarp();
  float loss = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float ui = ut[i], wi = ws[i];
    float inner = 0.f;
    for (int j = 0; j < S; ++j) inner = fmaf(ws[j], fabsf(ui - ut[j]), inner);
    const float delta = sub_rn(__ldg(t + r * (S + 1) + i + 1), __ldg(t + r * (S + 1) + i));
    loss += wi * inner + wi * wi * delta / 3.f;
    if (d_w) d_w[r * S + i] = gscale * (2.f * inner + 2.f * wi * delta / 3.f);
  }
  loss = warp_sum(loss);
  if (lane == 0 && loss_rows) loss_rows[r] = loss;
}

extern "C" int b2n_distortion_fwd_bwd(const float* t, const float* w, int64_t n_rays, int32_t s, float gscale,
                                      float* loss_rows, float* d_w, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(t && w, "null pointer");
  B2N_REQUIRE(s >= 1 && s <= 4096, "sample count out of range");
  if (n_rays == 0) return B2N_OK;
  const size_t smem = sizeof(float) * RW * 2 * s;
  if (smem > 48 * 1024) cudaFuncSetAttribute(distortion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  distortion_kernel<<<(unsigned)div_up(n_rays, RW), RW * 32, smem, (cudaStream_t)stream>>>(t, w, n_rays, s, gscale, loss_rows, d_w);
  B2N_LAUNCH_CHECK();
}
