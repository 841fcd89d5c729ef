// a5/a6 — SpacedSampler bins and PDFSampler inverse-CDF resampling for sm_100a.
//
// Index work (searchsorted) must be bit-exact with the reference's torch path
// (nerfstudio/model_components/ray_samplers.py:78-128,276-372), so every elementwise step is a separately
// rounded fp32 op, the normaliser is summed in torch's CPU order (common.cuh:torch_cpu_row_sum), the running sums
// are accumulated in fp64 and rounded to fp32 per element (what torch's CPU cumsum does; the fp64 partial sums of
// pdf values >= 2^-15 are exact, so the scan order cannot change them), and linspace tables come from the host (torch.linspace) instead of being recomputed.
// One warp owns one ray; cdf and bin edges live in shared memory; the scan is a warp-shuffle scan over
// per-lane contiguous chunks.
#include "common.cuh"
#include "render_rays.cuh"

__device__ __forceinline__ float spacing_fn(int kind, float x) {
  switch (kind) {
    case B2N_SPACING_PIECEWISE: return x < 1.f ? div_rn(x, 2.f) : sub_rn(1.f, div_rn(1.f, mul_rn(2.f, x)));
    case B2N_SPACING_LINDISP: return div_rn(1.f, x);
    case B2N_SPACING_SQRT: return __fsqrt_rn(x);
    case B2N_SPACING_LOG: return logf(x);
    default: return x;
  }
}
__device__ __forceinline__ float spacing_inv(int kind, float x) {
  switch (kind) {
    case B2N_SPACING_PIECEWISE: return x < 0.5f ? mul_rn(2.f, x) : div_rn(1.f, sub_rn(2.f, mul_rn(2.f, x)));
    case B2N_SPACING_LINDISP: return div_rn(1.f, x);
    case B2N_SPACING_SQRT: return mul_rn(x, x);
    case B2N_SPACING_LOG: return expf(x);
    default: return x;
  }
}
__device__ __forceinline__ float to_euclid(int kind, float b, float s_near, float s_far) {
  return spacing_inv(kind, add_rn(mul_rn(b, s_far), mul_rn(sub_rn(1.f, b), s_near)));
}

__global__ void spaced_sample_kernel(const float* __restrict__ nears, const float* __restrict__ fars,
                                     const float* __restrict__ lin, const float* __restrict__ jitter,
                                     int jitter_per_bin, int64_t n_rays, int n_samples, int spacing,
                                     float* __restrict__ sbins, float* __restrict__ ebins) {
  const int nb = n_samples + 1;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * nb) return;
  const int64_t r = idx / nb;
  const int i = (int)(idx - r * nb);
  float b = __ldg(lin + i);
  if (jitter != nullptr) {
    const float u = jitter_per_bin ? __ldg(jitter + idx) : __ldg(jitter + r);
    const float upper = i < n_samples ? div_rn(add_rn(__ldg(lin + i + 1), b), 2.f) : b;
    const float lower = i > 0 ? div_rn(add_rn(b, __ldg(lin + i - 1)), 2.f) : b;
    b = add_rn(lower, mul_rn(sub_rn(upper, lower), u));
  }
  const float s_near = spacing_fn(spacing, __ldg(nears + r)), s_far = spacing_fn(spacing, __ldg(fars + r));
  sbins[idx] = b;
  ebins[idx] = to_euclid(spacing, b, s_near, s_far);
}

extern "C" int b2n_spaced_sample(const float* nears, const float* fars, const float* lin, const float* jitter,
                                 int32_t jitter_per_bin, int64_t n_rays, int32_t n_samples, int32_t spacing,
                                 float* sbins, float* ebins, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(nears && fars && lin && sbins && ebins, "null pointer");
  B2N_REQUIRE(n_samples >= 1, "n_samples");
  const int64_t total = n_rays * (n_samples + 1);
  if (total == 0) return B2N_OK;
  spaced_sample_kernel<<<(unsigned)div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
      nears, fars, lin, jitter, jitter_per_bin, n_rays, n_samples, spacing, sbins, ebins);
  B2N_LAUNCH_CHECK();
}

#define PDF_WARPS 4

__global__ void __launch_bounds__(PDF_WARPS * 32) pdf_sample_kernel(
    const float* __restrict__ bins, const float* weights, const float* __restrict__ u_base,
    const float* __restrict__ jitter, int jitter_per_bin, const float* __restrict__ nears,
    const float* __restrict__ fars, int64_t n_rays, int S, int n_out, float anneal_host,
    const float* __restrict__ anneal_dev, float pad_hist, float eps,
    int spacing, float* __restrict__ new_sbins, float* __restrict__ new_ebins, float* __restrict__ cdf_out,
    int64_t* __restrict__ inds_out, const float* __restrict__ ebins_in, const float* __restrict__ density,
    float* weights_out) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * PDF_WARPS + warp;
  if (r >= n_rays) return;
  if (density != nullptr) {
    // fused RaySamples.get_weights (cameras/rays.py:129-152) of the level being resampled: same per-ray body as
    // b2n_weights_fwd; the histogram below then reads what this warp just wrote
    weights_fwd_ray<true>(ebins_in + r * (S + 1), ebins_in + r * (S + 1) + 1, density + r * S, S, weights_out + r * S, lane);
    __syncwarp();
    weights = weights_out;
  }
  float* cdf = sm + (size_t)warp * 2 * (S + 1);
  float* eb = cdf + (S + 1);
  const int nb = n_out;  // number of new bin edges = num_samples + 1
  const float anneal = anneal_dev ? __ldg(anneal_dev) : anneal_host;
  const float* wrow = weights + r * S;
  const float* brow = bins + r * (S + 1);
  const int chunk = (S + 31) / 32;
  const int i0 = lane * chunk, i1 = min(S, i0 + chunk);

  // pass 1: padded weights into smem (eb reused as scratch for w); their sum in torch's CPU summation order
  for (int i = i0; i < i1; ++i) {
    float w = density != nullptr ? wrow[i] : __ldg(wrow + i);
    if (anneal != 1.f) w = powf(w, anneal);
    eb[i] = add_rn(w, pad_hist);
  }
  __syncwarp();
  float w_sum = torch_cpu_row_sum(eb, S, lane);
  const float padding = fmaxf(sub_rn(eps, w_sum), 0.f);
  const float pad_each = div_rn(padding, (float)S);
  w_sum = add_rn(w_sum, padding);
  // pass 2: pdf, chunk-local fp64 prefix
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += (double)div_rn(add_rn(eb[i], pad_each), w_sum);
  const double incl = warp_scan_incl_d(local, lane);
  double run = incl - local;
  __syncwarp();
  for (int i = i0; i < i1; ++i) {
    run += (double)div_rn(add_rn(eb[i], pad_each), w_sum);
    cdf[i + 1] = fminf(1.f, (float)run);
  }
  if (lane == 0) cdf[0] = 0.f;
  __syncwarp();
  for (int i = lane; i <= S; i += 32) eb[i] = __ldg(brow + i);
  __syncwarp();
  if (cdf_out)
    for (int i = lane; i <= S; i += 32) cdf_out[r * (S + 1) + i] = cdf[i];

  const float s_near = spacing_fn(spacing, __ldg(nears + r)), s_far = spacing_fn(spacing, __ldg(fars + r));
  const float half_step = (float)(1.0 / (2.0 * (double)nb));
  for (int j = lane; j < nb; j += 32) {
    float u = __ldg(u_base + j);
    if (jitter != nullptr) {
      const float jv = jitter_per_bin ? __ldg(jitter + r * nb + j) : __ldg(jitter + r);
      u = add_rn(u, div_rn(jv, (float)nb));
    } else {
      u = add_rn(u, half_step);
    }
    int lo = 0, hi = S + 1;  // searchsorted(cdf, u, right=True)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int below = min(max(lo - 1, 0), S), above = min(max(lo, 0), S);
    const float c0 = cdf[below], c1 = cdf[above], b0 = eb[below], b1 = eb[above];
    float t = div_rn(sub_rn(u, c0), sub_rn(c1, c0));
    if (isnan(t)) t = 0.f;  // nan_to_num(., 0); +-inf are clipped below anyway
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float b = add_rn(b0, mul_rn(t, sub_rn(b1, b0)));
    new_sbins[r * nb + j] = b;
    if (new_ebins) new_ebins[r * nb + j] = to_euclid(spacing, b, s_near, s_far);
    if (inds_out) inds_out[r * nb + j] = lo;
  }
}

extern "C" int b2n_pdf_sample(const float* bins, const float* weights, const float* u_base, const float* jitter,
                              int32_t jitter_per_bin, const float* nears, const float* fars, int64_t n_rays,
                              int32_t n_in, int32_t n_out, float anneal, const float* anneal_dev,
                              float histogram_padding, float eps, int32_t spacing, float* new_sbins, float* new_ebins, float* cdf_out, int64_t* inds_out,
                              void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(bins && weights && u_base && nears && fars && new_sbins, "null pointer");
  B2N_REQUIRE(n_in >= 1 && n_in <= 4096 && n_out >= 1, "sample counts out of range");
  if (n_rays == 0) return B2N_OK;
  const size_t smem = sizeof(float) * PDF_WARPS * 2 * (n_in + 1);
  if (smem > 48 * 1024) cudaFuncSetAttribute(pdf_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pdf_sample_kernel<<<(unsigned)div_up(n_rays, PDF_WARPS), PDF_WARPS * 32, smem, (cudaStream_t)stream>>>(
      bins, weights, u_base, jitter, jitter_per_bin, nears, fars, n_rays, n_in, n_out, anneal, anneal_dev,
      histogram_padding, eps,
      spacing, new_sbins, new_ebins, cdf_out, inds_out, nullptr, nullptr, nullptr);
  B2N_LAUNCH_CHECK();
}

// get_weights of the level being resampled + the PDF resampling in one launch (the proposal sampler's inner loop:
// model_components/ray_samplers.py:568-604 calls field density -> get_weights -> PDFSampler per level)
extern "C" int b2n_weights_pdf_sample(const float* sbins, const float* ebins, const float* density, const float* u_base,
                                      const float* jitter, int32_t jitter_per_bin, const float* nears, const float* fars,
                                      int64_t n_rays, int32_t n_in, int32_t n_out, float anneal, const float* anneal_dev,
                                      float histogram_padding, float eps, int32_t spacing, float* weights, float* new_sbins,
                                      float* new_ebins, void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(sbins && ebins && density && weights && u_base && nears && fars && new_sbins, "null pointer");
  B2N_REQUIRE(n_in >= 1 && n_in <= 4096 && n_out >= 1, "sample counts out of range");
  const size_t smem = sizeof(float) * PDF_WARPS * 2 * (n_in + 1);
  if (smem > 48 * 1024) cudaFuncSetAttribute(pdf_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pdf_sample_kernel<<<(unsigned)div_up(n_rays, PDF_WARPS), PDF_WARPS * 32, smem, (cudaStream_t)stream>>>(
      sbins, weights, u_base, jitter, jitter_per_bin, nears, fars, n_rays, n_in, n_out, anneal, anneal_dev,
      histogram_padding, eps, spacing, new_sbins, new_ebins, nullptr, nullptr, ebins, density, weights);
  B2N_LAUNCH_CHECK();
}

// diagnostic / parity pin: out[r] = torch.sum(x[r, :]) in torch's CPU summation order (common.cuh:torch_cpu_row_sum)
__global__ void torch_row_sum_kernel(const float* __restrict__ x, int64_t n_rows, int S, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= n_rows) return;
  const float v = torch_cpu_row_sum(x + r * S, S, lane);
  if (lane == 0) out[r] = v;
}

extern "C" int b2n_torch_row_sum(const float* x, int64_t n_rows, int32_t n_cols, float* out, void* stream) {
  if (n_rows == 0) return B2N_OK;
  B2N_REQUIRE(x && out && n_cols >= 1, "bad arguments");
  torch_row_sum_kernel<<<(unsigned)div_up(n_rows, 4), 128, 0, (cudaStream_t)stream>>>(x, n_rows, n_cols, out);
  B2N_LAUNCH_CHECK();
}
