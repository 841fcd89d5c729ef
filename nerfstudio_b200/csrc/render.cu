// a21-a24 — transmittance weights (warp-shuffle segmented scan), alpha compositing, depth, proposal losses.
//
// One warp owns one ray.  A lane holds a contiguous chunk of ceil(S/32) samples; running sums are fp64
// (torch's CPU cumsum accumulates in double and rounds per element — nerfstudio/cameras/rays.py:129-152),
// chunk totals are combined with a shuffle scan.  Reference rows: cameras/rays.py:129-152,
// model_components/renderers.py:71-119,292-385, model_components/losses.py:53-155.
#include "common.cuh"
#include "render_rays.cuh"

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
  weights_fwd_ray<true>(starts + r * bs, ends + r * bs, density + r * S, S, weights + r * S, lane);
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
  float* out = ddensity + r * S;
  weights_bwd_ray<true, true>(starts + r * bs, ends + r * bs, density + r * S, dweights + r * S, S, lane,
                        [out](int i, float v) { out[i] = v; });
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
  const CompositeOut o = composite_fwd_ray<true>(bg, rgb ? rgb + r * S * 3 : nullptr, weights + r * S,
                                                 starts ? starts + r * bs : nullptr, ends ? ends + r * bs : nullptr, S,
                                                 depth_med != nullptr || med_idx != nullptr, lane);
  if (lane == 0) {
    if (med_idx) med_idx[r] = o.med_idx;
    if (depth_med && starts) depth_med[r] = o.depth_med;
    if (rgb_out && rgb) rgb_out[3 * r] = o.r, rgb_out[3 * r + 1] = o.g, rgb_out[3 * r + 2] = o.b;
    if (acc_out) acc_out[r] = o.acc;
    if (depth_exp) depth_exp[r] = o.depth_exp;
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
  const float g0 = d_rgb_out ? __ldg(d_rgb_out + 3 * r) : 0.f, g1 = d_rgb_out ? __ldg(d_rgb_out + 3 * r + 1) : 0.f,
              g2 = d_rgb_out ? __ldg(d_rgb_out + 3 * r + 2) : 0.f;
  const float ga = d_acc ? __ldg(d_acc + r) : 0.f;
  const float gd = d_depth ? __ldg(d_depth + r) : 0.f;
  composite_bwd_ray<true>(bg, rgb + r * S * 3, weights + r * S, starts ? starts + r * bs : nullptr, ends ? ends + r * bs : nullptr,
                          g0, g1, g2, ga, gd, d_depth != nullptr, S, d_rgb ? d_rgb + r * S * 3 : nullptr,
                          d_w ? d_w + r * S : nullptr, nullptr, lane);
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

__global__ void __launch_bounds__(RW * 32) interlevel_kernel(const float* __restrict__ c, const float* __restrict__ w,
                                                             const float* __restrict__ cp, const float* __restrict__ wp,
                                                             int64_t n_rays, int Sc, int Sp, float gscale,
                                                             float* __restrict__ loss_rows, float* __restrict__ d_wp) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * RW + warp;
  if (r >= n_rays) return;
  const float loss = interlevel_ray<true>(c + r * (Sc + 1), w + r * Sc, cp + r * (Sp + 1), wp + r * Sp, Sc, Sp, gscale,
                                          d_wp ? d_wp + r * Sp : nullptr, sm + (size_t)warp * (3 * (Sp + 2)), lane);
  if (lane == 0 && loss_rows) loss_rows[r] = loss;
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
  const float loss = distortion_ray<true>(t + r * (S + 1), w + r * S, S, gscale, d_w ? d_w + r * S : nullptr,
                                          sm + (size_t)warp * 2 * S, lane);
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
