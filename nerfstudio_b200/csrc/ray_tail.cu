// The per-ray middle of the nerfacto training step in ONE launch: everything between the colour head's forward and its
// backward is local to a ray, so one warp per ray runs
//   density activation + get_weights (rays.py:129-152) -> RGB / accumulation / depth renderers (renderers.py:71-119,292-385)
//   -> MSE gradient (models/nerfacto.py:372) -> interlevel loss vs both proposal levels + distortion loss
//   (losses.py:53-155) -> compositing backward -> get_weights backward of all three levels -> density-activation backward
// with the per-ray bodies of render_rays.cuh — the same arithmetic as the separate operators (b2n_weights_fwd,
// b2n_composite_fwd, b2n_mse_fwd_bwd, b2n_interlevel_fwd_bwd, b2n_distortion_fwd_bwd, b2n_composite_bwd, b2n_add_inplace,
// b2n_weights_bwd, b2n_density_act_fwd/bwd), bit-identical gradients, 15 launches of 3-16 us each fewer per step.
// Loss terms leave as per-ray rows; b2n_loss_finalize reduces them in a fixed order (deterministic step losses).
#include "common.cuh"
#include "render_rays.cuh"

#define TW 4  // warps (rays) per CTA

struct RayTailParams {
  int64_t n_rays;
  int S2, Sp[2];
  int update_props;
  // main level
  const float* sb2;       // [R, S2+1] spacing-domain edges (losses)
  const float* eb2;       // [R, S2+1] euclidean edges (weights, depth)
  const float* h;         // [R*S2, h_stride] base MLP output, column 0 = density pre-activation
  int h_stride;
  const uint8_t* sel;     // [R*S2] selector
  float avg_init;
  const float* rgb;       // [R*S2, 3]
  const float* gt;        // [R, 3]
  float mse_gscale, il_gscale, dist_gscale;
  Bg bg;
  float *dens2, *w2, *rgb_out, *acc, *depth_exp, *depth_med, *d_rgb, *d_w2, *d_w_dist, *d_hpre;
  // proposal levels
  const float* sbp[2];    // [R, Sp+1] spacing-domain edges
  const float* ebp[2];    // [R, Sp+1] euclidean edges
  const float* wp[2];     // [R, Sp]
  const float* densp[2];  // [R, Sp]
  float* d_wp[2];
  float* d_densp[2];
  float* rows[4];         // per-ray loss terms: interlevel 0, interlevel 1, distortion, squared rgb error
};

__global__ void __launch_bounds__(TW * 32) ray_tail_kernel(const __grid_constant__ RayTailParams p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * TW + warp;
  if (r >= p.n_rays) return;
  const int S = p.S2;
  const int smax = max(3 * (max(p.Sp[0], p.Sp[1]) + 2), 2 * S);
  float* scratch = sm + (size_t)warp * smax;
  const float* st = p.eb2 + r * (S + 1);
  const float* en = st + 1;
  float* dens = p.dens2 + r * S;
  float* w = p.w2 + r * S;
  const float* rgb = p.rgb + r * S * 3;
  // ---- density activation (each lane the samples of its own chunk, which are the ones it reads back below)
  {
    const int chunk = (S + 31) / 32, i0 = lane * chunk, i1 = min(S, i0 + chunk);
    for (int i = i0; i < i1; ++i) dens[i] = density_act(__ldg(p.h + (r * S + i) * p.h_stride), p.sel, r * S + i, p.avg_init);
  }
  weights_fwd_ray<false>(st, en, dens, S, w, lane);
  __syncwarp();
  // ---- renderers + MSE
  const CompositeOut o = composite_fwd_ray<false>(p.bg, rgb, w, st, en, S, true, lane);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (lane == 0) {
    p.rgb_out[3 * r] = o.r, p.rgb_out[3 * r + 1] = o.g, p.rgb_out[3 * r + 2] = o.b;
    p.acc[r] = o.acc, p.depth_exp[r] = o.depth_exp, p.depth_med[r] = o.depth_med;
    const float inv = 1.f / (float)(3 * p.n_rays);
    const float d0 = o.r - __ldg(p.gt + 3 * r), d1 = o.g - __ldg(p.gt + 3 * r + 1), d2 = o.b - __ldg(p.gt + 3 * r + 2);
    float s = 0.f;
    s = fmaf(d0, d0, s), s = fmaf(d1, d1, s), s = fmaf(d2, d2, s);
    p.rows[3][r] = s;
    g0 = p.mse_gscale * 2.f * d0 * inv, g1 = p.mse_gscale * 2.f * d1 * inv, g2 = p.mse_gscale * 2.f * d2 * inv;
  }
  g0 = __shfl_sync(0xffffffffu, g0, 0), g1 = __shfl_sync(0xffffffffu, g1, 0), g2 = __shfl_sync(0xffffffffu, g2, 0);
  // ---- proposal losses (the main level's weights are detached: gradients only reach the proposal weights)
  const float* c2 = p.sb2 + r * (S + 1);
#pragma unroll
  for (int lvl = 0; lvl < 2; ++lvl) {
    const int Sp = p.Sp[lvl];
    const float loss = interlevel_ray<false>(c2, w, p.sbp[lvl] + r * (Sp + 1), p.wp[lvl] + r * Sp, S, Sp, p.il_gscale,
                                             p.update_props ? p.d_wp[lvl] + r * Sp : nullptr, scratch, lane);
    if (lane == 0) p.rows[lvl][r] = loss;
  }
  {
    const float loss = distortion_ray<false>(c2, w, S, p.dist_gscale, p.d_w_dist + r * S, scratch, lane);
    if (lane == 0) p.rows[2][r] = loss;
  }
  // ---- backward: compositing (+ the distortion term), get_weights, density activation
  composite_bwd_ray<false>(p.bg, rgb, w, st, en, g0, g1, g2, 0.f, 0.f, false, S, p.d_rgb + r * S * 3, p.d_w2 + r * S,
                           p.d_w_dist + r * S, lane);
  __syncwarp();
  {
    const float* hrow = p.h + r * S * p.h_stride;
    const uint8_t* sel = p.sel;
    float* dh = p.d_hpre + r * S;
    const int hs = p.h_stride;
    const float avg = p.avg_init;
    const int64_t base = r * S;
    weights_bwd_ray<false, false>(st, en, dens, p.d_w2 + r * S, S, lane, [=](int i, float v) {
      dh[i] = density_act_grad(__ldg(hrow + (int64_t)i * hs), sel, base + i, v, avg);
    });
  }
  if (p.update_props) {
#pragma unroll
    for (int lvl = 0; lvl < 2; ++lvl) {
      const int Sp = p.Sp[lvl];
      const float* eb = p.ebp[lvl] + r * (Sp + 1);
      float* out = p.d_densp[lvl] + r * Sp;
      weights_bwd_ray<true, false>(eb, eb + 1, p.densp[lvl] + r * Sp, p.d_wp[lvl] + r * Sp, Sp, lane,
                             [out](int i, float v) { out[i] = v; });
    }
  }
}

extern "C" int b2n_nerfacto_ray_tail(int64_t n_rays, int32_t s_main, int32_t s_prop0, int32_t s_prop1, const float* sbins_main,
                                     const float* ebins_main, const float* base_out, int32_t base_stride, const uint8_t* selector,
                                     float avg_init, const float* rgb, const float* gt, int32_t bg_mode, const float* bg_host3,
                                     float mse_gscale, float interlevel_gscale, float distortion_gscale,
                                     const float* const* prop_sbins2, const float* const* prop_ebins2,
                                     const float* const* prop_weights2, const float* const* prop_density2,
                                     float* const* d_prop_weights2, float* const* d_prop_density2, float* density, float* weights,
                                     float* rgb_out, float* accumulation, float* depth_expected, float* depth_median, float* d_rgb,
                                     float* d_weights, float* d_weights_distortion, float* d_density_pre, float* const* loss_rows4,
                                     void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(sbins_main && ebins_main && base_out && rgb && gt && prop_sbins2 && prop_ebins2 && prop_weights2 && loss_rows4,
              "null pointer");
  B2N_REQUIRE(density && weights && rgb_out && accumulation && depth_expected && depth_median && d_rgb && d_weights &&
                  d_weights_distortion && d_density_pre, "null output pointer");
  B2N_REQUIRE(s_main >= 1 && s_prop0 >= 1 && s_prop1 >= 1 && s_main <= 4096 && s_prop0 <= 4096 && s_prop1 <= 4096,
              "sample counts out of range");
  B2N_REQUIRE(bg_mode != B2N_BG_CONSTANT || bg_host3, "constant background needs a colour");
  const bool update = d_prop_weights2 != nullptr;
  B2N_REQUIRE(!update || (prop_density2 && d_prop_density2), "proposal gradients need densities and outputs");
  RayTailParams p;
  p.n_rays = n_rays, p.S2 = s_main, p.Sp[0] = s_prop0, p.Sp[1] = s_prop1, p.update_props = update ? 1 : 0;
  p.sb2 = sbins_main, p.eb2 = ebins_main, p.h = base_out, p.h_stride = base_stride, p.sel = selector, p.avg_init = avg_init;
  p.rgb = rgb, p.gt = gt, p.mse_gscale = mse_gscale, p.il_gscale = interlevel_gscale, p.dist_gscale = distortion_gscale;
  p.bg.mode = bg_mode, p.bg.eval_mode = 0;
  for (int i = 0; i < 3; ++i) p.bg.c[i] = (bg_mode == B2N_BG_CONSTANT) ? bg_host3[i] : 0.f;
  p.dens2 = density, p.w2 = weights, p.rgb_out = rgb_out, p.acc = accumulation, p.depth_exp = depth_expected;
  p.depth_med = depth_median, p.d_rgb = d_rgb, p.d_w2 = d_weights, p.d_w_dist = d_weights_distortion, p.d_hpre = d_density_pre;
  for (int l = 0; l < 2; ++l) {
    B2N_REQUIRE(prop_sbins2[l] && prop_ebins2[l] && prop_weights2[l], "null proposal pointer");
    p.sbp[l] = prop_sbins2[l], p.ebp[l] = prop_ebins2[l], p.wp[l] = prop_weights2[l];
    p.densp[l] = update ? prop_density2[l] : nullptr;
    p.d_wp[l] = update ? d_prop_weights2[l] : nullptr, p.d_densp[l] = update ? d_prop_density2[l] : nullptr;
    B2N_REQUIRE(!update || (p.densp[l] && p.d_wp[l] && p.d_densp[l]), "null proposal gradient pointer");
  }
  for (int k = 0; k < 4; ++k) {
    B2N_REQUIRE(loss_rows4[k], "null loss row pointer");
    p.rows[k] = loss_rows4[k];
  }
  const int smax = max(3 * (max(s_prop0, s_prop1) + 2), 2 * s_main);
  const size_t smem = sizeof(float) * TW * smax;
  if (smem > 48 * 1024) cudaFuncSetAttribute(ray_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ray_tail_kernel<<<(unsigned)div_up(n_rays, TW), TW * 32, smem, (cudaStream_t)stream>>>(p);
  B2N_LAUNCH_CHECK();
}

// out[k] = scale[k] * sum(rows_k[0..n)) for the step's loss terms in a fixed reduction order, then the total:
//   losses[0] = rgb, [1] = interlevel (both levels), [2] = distortion, [3] = ((l0 + l1) + l2) + l4, [4] untouched
__global__ void __launch_bounds__(1024) loss_finalize_kernel(const float* __restrict__ r_il0, const float* __restrict__ r_il1,
                                                             const float* __restrict__ r_dist, const float* __restrict__ r_rgb,
                                                             int64_t n, float s_il, float s_dist, float s_rgb,
                                                             float* __restrict__ losses) {
  __shared__ float part[4][32];
  const float* rows[4] = {r_il0, r_il1, r_dist, r_rgb};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float tot[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += __ldg(rows[k] + i);
    s = warp_sum(s);
    if (lane == 0) part[k][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tot[k] = warp_sum(part[k][lane]);
    if (lane == 0) {
      const float l_rgb = tot[3] * s_rgb, l_il = add_rn(tot[0] * s_il, tot[1] * s_il), l_dist = tot[2] * s_dist;
      losses[0] = l_rgb, losses[1] = l_il, losses[2] = l_dist;
      losses[3] = add_rn(add_rn(add_rn(l_rgb, l_il), l_dist), losses[4]);
    }
  }
}

extern "C" int b2n_loss_finalize(const float* const* loss_rows4, int64_t n_rays, float interlevel_scale, float distortion_scale,
                                 float rgb_scale, float* losses5, void* stream) {
  B2N_REQUIRE(loss_rows4 && losses5 && n_rays >= 1, "bad arguments");
  for (int k = 0; k < 4; ++k) B2N_REQUIRE(loss_rows4[k], "null loss row pointer");
  loss_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(loss_rows4[0], loss_rows4[1], loss_rows4[2], loss_rows4[3], n_rays,
                                                             interlevel_scale, distortion_scale, rgb_scale, losses5);
  B2N_LAUNCH_CHECK();
}
