// Glue kernels of the graph-captured nerfacto step: colour-head input assembly (SH ‖ geo features ‖ appearance
// embedding) forward/backward, and the rgb MSE loss with its gradient.  They replace the torch.cat / expand /
// embedding / index_put chain of nerfstudio/fields/nerfacto_field.py:234-310 and the MSELoss of
// models/nerfacto.py:372 so that one optimisation step is a fixed sequence of our own launches.
#include "common.cuh"

// head_in[n, :] = [ sh[ray(n), 0:n_sh] | base_out[n, 1:1+geo] | emb[cam[ray(n)], 0:n_emb] ]
__device__ __forceinline__ float head_input_elem(const float* __restrict__ sh, int n_sh, const float* __restrict__ base_out,
                                                 int base_w, int geo, const float* __restrict__ emb, int64_t cam_row,
                                                 int n_emb, int emb_mode, int64_t n, int64_t r, int c) {
  if (c < n_sh) return __ldg(sh + r * n_sh + c);
  if (c < n_sh + geo) return __ldg(base_out + n * base_w + 1 + (c - n_sh));
  if (c >= n_sh + geo + n_emb) return 0.f;                                    // padding inside the last 4-column group
  if (emb_mode == 1) return __ldg(emb + cam_row * n_emb + (c - n_sh - geo));  // training: per-image row
  if (emb_mode == 2) return __ldg(emb + (c - n_sh - geo));                    // eval: one (mean) row
  return 0.f;                                                                 // eval: zeros
}

// one thread per (sample, group of 4 output columns): one index division per 4 values and a 128-bit store; the element
// version below (unaligned output) spent its time in 64-bit divisions (58 us for 12.4 M values)
__global__ void head_input_fwd_vec_kernel(const float* __restrict__ sh, int n_sh, const float* __restrict__ base_out,
                                          int base_w, int geo, const float* __restrict__ emb,
                                          const int64_t* __restrict__ cam, int n_emb, int emb_mode, int64_t n_rays, int S,
                                          float* __restrict__ out, int out_stride, int part) {
  // part 0: every 4-column group; 1: the groups that do not read base_out (SH and embedding columns: per-ray constants,
  // available before the field runs); 2: the groups that do (geo features)
  const int width = n_sh + geo + n_emb, all_groups = (width + 3) >> 2;
  const int dyn0 = n_sh >> 2, dyn1 = min(all_groups, (n_sh + geo + 3) >> 2);
  const int groups = part == 0 ? all_groups : part == 2 ? dyn1 - dyn0 : all_groups - (dyn1 - dyn0);
  const uint32_t total = (uint32_t)(n_rays * S) * (uint32_t)groups;  // host guarantees < 2^31
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const uint32_t n = idx / (uint32_t)groups, qi = idx - n * (uint32_t)groups, r = n / (uint32_t)S;
    const uint32_t q = part == 0 ? qi : part == 2 ? qi + dyn0 : ((int)qi < dyn0 ? qi : qi + (dyn1 - dyn0));
    const int64_t cam_row = (emb_mode == 1 && (int)(4 * q + 3) >= n_sh + geo) ? __ldg(cam + r) : 0;
    float4 v;
    v.x = head_input_elem(sh, n_sh, base_out, base_w, geo, emb, cam_row, n_emb, emb_mode, n, r, 4 * q);
    v.y = head_input_elem(sh, n_sh, base_out, base_w, geo, emb, cam_row, n_emb, emb_mode, n, r, 4 * q + 1);
    v.z = head_input_elem(sh, n_sh, base_out, base_w, geo, emb, cam_row, n_emb, emb_mode, n, r, 4 * q + 2);
    v.w = head_input_elem(sh, n_sh, base_out, base_w, geo, emb, cam_row, n_emb, emb_mode, n, r, 4 * q + 3);
    *reinterpret_cast<float4*>(out + (size_t)n * out_stride + 4 * q) = v;
  }
}

__global__ void head_input_fwd_kernel(const float* __restrict__ sh, int n_sh, const float* __restrict__ base_out,
                                      int base_w, int geo, const float* __restrict__ emb,
                                      const int64_t* __restrict__ cam, int n_emb, int emb_mode, int64_t n_rays, int S,
                                      float* __restrict__ out, int out_stride) {
  const int width = n_sh + geo + n_emb;
  const int64_t total = n_rays * S * width;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / width;
    const int c = (int)(idx - n * width);
    const int64_t r = n / S;
    const int64_t cam_row = (emb_mode == 1 && c >= n_sh + geo) ? __ldg(cam + r) : 0;
    out[n * out_stride + c] = head_input_elem(sh, n_sh, base_out, base_w, geo, emb, cam_row, n_emb, emb_mode, n, r, c);
  }
}

// d_base_out[n, 0] = d_density_pre[n]; d_base_out[n, 1+g] = d_in[n, n_sh+g]; d_emb[cam[ray]] += sum_s d_in[n, n_sh+geo+e]
__global__ void head_input_bwd_kernel(const float* __restrict__ d_in, int in_stride, int n_sh, int geo, int n_emb,
                                      const float* __restrict__ d_dens_pre, const int64_t* __restrict__ cam,
                                      int64_t n_rays, int S, float* __restrict__ d_base_out, int base_w,
                                      float* __restrict__ d_emb, int part) {
  const int width = in_stride;
  // part 1: one thread per (sample, base column)          (launch part: 0 = both, 1 = this one, 2 = the embedding rows)
  const int64_t total1 = part == 2 ? 0 : n_rays * S * base_w;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = gtid; idx < total1; idx += gsz) {
    const int64_t n = idx / base_w;
    const int c = (int)(idx - n * base_w);
    float v = 0.f;
    if (c == 0) v = d_dens_pre ? __ldg(d_dens_pre + n) : 0.f;
    else if (c <= geo) v = __ldg(d_in + n * width + n_sh + c - 1);
    d_base_out[idx] = v;
  }
  // part 2: one thread per (ray, embedding column): reduce over the ray's samples, one atomic per ray
  if (d_emb != nullptr && n_emb > 0 && part != 1) {
    const int64_t total2 = n_rays * n_emb;
    for (int64_t idx = gtid; idx < total2; idx += gsz) {
      const int64_t r = idx / n_emb;
      const int e = (int)(idx - r * n_emb);
      float s = 0.f;
      const float* p = d_in + (r * S) * width + n_sh + geo + e;
      for (int i = 0; i < S; ++i) s += __ldg(p + (int64_t)i * width);
      atomicAdd(d_emb + __ldg(cam + r) * n_emb + e, s);
    }
  }
}

extern "C" int b2n_head_input_fwd_part(const float* sh, int32_t n_sh, const float* base_out, int32_t base_w, int32_t geo,
                                       const float* emb, const int64_t* cam, int32_t n_emb, int32_t emb_mode, int64_t n_rays,
                                       int32_t n_samples, float* out, int32_t out_stride, int32_t part, void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(sh && base_out && out, "null pointer");
  B2N_REQUIRE(part >= 0 && part <= 2, "part");
  B2N_REQUIRE(n_emb == 0 || emb_mode == 0 || emb, "embedding table missing");
  B2N_REQUIRE(emb_mode != 1 || cam, "camera indices missing");
  B2N_REQUIRE(1 + geo <= base_w, "geo features exceed base width");
  B2N_REQUIRE(out_stride >= n_sh + geo + n_emb, "out_stride too small");
  const int width = n_sh + geo + n_emb, groups = (width + 3) / 4;
  const int64_t total_vec = n_rays * n_samples * groups;
  // vector path: whole 4-column groups are written, i.e. columns [width, 4*groups) (< out_stride) are zero-filled
  if ((out_stride & 3) == 0 && ((uintptr_t)out & 15) == 0 && 4 * groups <= out_stride && total_vec < (1ll << 31)) {
    const unsigned grid = (unsigned)min(div_up(total_vec, 256), (int64_t)b2n_sm_count() * 16);
    head_input_fwd_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sh, n_sh, base_out, base_w, geo, emb, cam, n_emb,
                                                                      emb_mode, n_rays, n_samples, out, out_stride, part);
    B2N_LAUNCH_CHECK();
  }
  B2N_UNSUPPORTED(part != 0, "head input in parts needs 16-byte aligned rows (out_stride % 4 == 0)");
  const int64_t total = n_rays * n_samples * width;
  const unsigned grid = (unsigned)min(div_up(total, 256), (int64_t)b2n_sm_count() * 32);
  head_input_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sh, n_sh, base_out, base_w, geo, emb, cam, n_emb,
                                                                emb_mode, n_rays, n_samples, out, out_stride);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_head_input_fwd(const float* sh, int32_t n_sh, const float* base_out, int32_t base_w, int32_t geo,
                                  const float* emb, const int64_t* cam, int32_t n_emb, int32_t emb_mode, int64_t n_rays,
                                  int32_t n_samples, float* out, int32_t out_stride, void* stream) {
  return b2n_head_input_fwd_part(sh, n_sh, base_out, base_w, geo, emb, cam, n_emb, emb_mode, n_rays, n_samples, out, out_stride, 0,
                                 stream);
}

extern "C" int b2n_head_input_bwd_part(const float* d_in, int32_t in_stride, int32_t n_sh, int32_t geo, int32_t n_emb,
                                       const float* d_dens_pre, const int64_t* cam, int64_t n_rays, int32_t n_samples,
                                       float* d_base_out, int32_t base_w, float* d_emb, int32_t part, void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(d_in && (d_base_out || part == 2), "null pointer");
  B2N_REQUIRE(part >= 0 && part <= 2, "part");
  B2N_REQUIRE(d_emb == nullptr || cam, "camera indices missing");
  B2N_REQUIRE(in_stride >= n_sh + geo + n_emb, "in_stride too small");
  const int64_t total = part == 2 ? n_rays * n_emb : n_rays * n_samples * base_w;
  if (total == 0) return B2N_OK;
  const unsigned grid = (unsigned)min(div_up(total, 256), (int64_t)b2n_sm_count() * 32);
  head_input_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, in_stride, n_sh, geo, n_emb, d_dens_pre, cam, n_rays,
                                                                n_samples, d_base_out, base_w, d_emb, part);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_head_input_bwd(const float* d_in, int32_t in_stride, int32_t n_sh, int32_t geo, int32_t n_emb,
                                  const float* d_dens_pre, const int64_t* cam, int64_t n_rays, int32_t n_samples,
                                  float* d_base_out, int32_t base_w, float* d_emb, void* stream) {
  return b2n_head_input_bwd_part(d_in, in_stride, n_sh, geo, n_emb, d_dens_pre, cam, n_rays, n_samples, d_base_out, base_w, d_emb,
                                 0, stream);
}

// loss_out[0] += mean((pred - gt)^2) * 1 ; d_pred = gscale * 2 (pred - gt) / n     (nn.MSELoss, models/nerfacto.py:372)
__global__ void mse_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int64_t n, float gscale,
                           float* __restrict__ loss_out, float* __restrict__ d_pred) {
  float s = 0.f;
  const float inv = 1.f / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = __ldg(pred + i) - __ldg(gt + i);
    s = fmaf(d, d, s);
    if (d_pred) d_pred[i] = gscale * 2.f * d * inv;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0 && loss_out) atomicAdd(loss_out, s * inv);
}

extern "C" int b2n_mse_fwd_bwd(const float* pred, const float* gt, int64_t n, float gscale, float* loss_out,
                               float* d_pred, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(pred && gt, "null pointer");
  const unsigned grid = (unsigned)min(div_up(n, 256), (int64_t)b2n_sm_count() * 4);
  mse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, n, gscale, loss_out, d_pred);
  B2N_LAUNCH_CHECK();
}

// out[0] += scale * sum(rows[0..n))      (loss bookkeeping inside the captured step)
__global__ void sum_rows_kernel(const float* __restrict__ rows, int64_t n, float scale, float* __restrict__ out) {
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += __ldg(rows + i);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s * scale);
}

extern "C" int b2n_sum_rows(const float* rows, int64_t n, float scale, float* out, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(rows && out, "null pointer");
  sum_rows_kernel<<<(unsigned)min(div_up(n, 256), (int64_t)64), 256, 0, (cudaStream_t)stream>>>(rows, n, scale, out);
  B2N_LAUNCH_CHECK();
}
