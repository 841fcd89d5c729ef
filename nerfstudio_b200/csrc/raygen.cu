// a1-a3 — ray generation for perspective cameras (+ OpenCV undistortion) and the AABB collider.
// Reference: nerfstudio/cameras/cameras.py:599-929, cameras/camera_utils.py:318-330,375-478,
// model_components/ray_generators.py:41-56, model_components/scene_colliders.py:47-108.
// One thread per ray; the three coordinate variants (pixel, +1 in x, +1 in y) the reference stacks are
// carried in registers, the 10 Newton iterations of the undistortion are unrolled in-thread.
#include "common.cuh"

__device__ __forceinline__ void undistort(float& x, float& y, const float* k) {
  const float xd = x, yd = y;
  const float k1 = k[0], k2 = k[1], k3 = k[2], k4 = k[3], p1 = k[4], p2 = k[5];
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {
    const float r = x * x + y * y;
    const float d = 1.f + r * (k1 + r * (k2 + r * (k3 + r * k4)));
    const float fx = d * x + 2.f * p1 * x * y + p2 * (r + 2.f * x * x) - xd;
    const float fy = d * y + 2.f * p2 * x * y + p1 * (r + 2.f * y * y) - yd;
    const float d_r = k1 + r * (2.f * k2 + r * (3.f * k3 + r * 4.f * k4));
    const float d_x = 2.f * x * d_r, d_y = 2.f * y * d_r;
    const float fx_x = d + d_x * x + 2.f * p1 * y + 6.f * p2 * x;
    const float fx_y = d_y * x + 2.f * p1 * x + 2.f * p2 * y;
    const float fy_x = d_x * y + 2.f * p2 * y + 2.f * p1 * x;
    const float fy_y = d + d_y * y + 2.f * p2 * x + 6.f * p1 * y;
    const float den = fy_x * fx_y - fx_x * fy_y;
    const bool ok = fabsf(den) > 1e-3f;
    x += ok ? (fx * fy_y - fy * fx_y) / den : 0.f;
    y += ok ? (fy * fx_x - fx * fy_x) / den : 0.f;
  }
}

// One ray: pixel-centre coordinates (px, py) of camera `cam` -> origin, unit direction, pixel area, direction norm.
// cam_opt (optional, 12 floats): camera_opt_to_camera, composed as pose_utils.multiply(c2w, opt) (cameras.py:884-885).
__device__ __forceinline__ void ray_from_pixel(const float* __restrict__ c2w, const float* __restrict__ intr,
                                               const float* __restrict__ dist, const float* __restrict__ dist_delta,
                                               const float* __restrict__ cam_opt, int64_t cam, float px, float py, float* o,
                                               float* dir, float* area, float* nrm_out) {
  const float fx = __ldg(intr + 4 * cam), fy = __ldg(intr + 4 * cam + 1), cx = __ldg(intr + 4 * cam + 2), cy = __ldg(intr + 4 * cam + 3);
  float ux[3], uy[3];
  ux[0] = div_rn(sub_rn(px, cx), fx), uy[0] = div_rn(sub_rn(py, cy), fy);
  ux[1] = div_rn(add_rn(sub_rn(px, cx), 1.f), fx), uy[1] = uy[0];
  ux[2] = ux[0], uy[2] = div_rn(add_rn(sub_rn(py, cy), 1.f), fy);
  if (dist != nullptr || dist_delta != nullptr) {
    float k[6];
    bool any = false;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      k[j] = (dist ? __ldg(dist + 6 * cam + j) : 0.f) + (dist_delta ? __ldg(dist_delta + j) : 0.f);
      any |= k[j] != 0.f;
    }
    if (any) {
#pragma unroll
      for (int v = 0; v < 3; ++v) undistort(ux[v], uy[v], k);
    }
  }
  float R[9], t[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) R[a * 3 + b] = __ldg(c2w + cam * 12 + a * 4 + b);
    t[a] = __ldg(c2w + cam * 12 + a * 4 + 3);
  }
  if (cam_opt != nullptr) {  // R <- R * R_opt, t <- t + R * t_opt
    float R2[9], t2[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int b = 0; b < 3; ++b)
        R2[a * 3 + b] = R[a * 3] * __ldg(cam_opt + b) + R[a * 3 + 1] * __ldg(cam_opt + 4 + b) + R[a * 3 + 2] * __ldg(cam_opt + 8 + b);
      t2[a] = t[a] + (R[a * 3] * __ldg(cam_opt + 3) + R[a * 3 + 1] * __ldg(cam_opt + 7) + R[a * 3 + 2] * __ldg(cam_opt + 11));
    }
#pragma unroll
    for (int a = 0; a < 9; ++a) R[a] = R2[a];
#pragma unroll
    for (int a = 0; a < 3; ++a) t[a] = t2[a];
  }
  float d[3][3];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const float dx = ux[v], dy = -uy[v], dz = -1.f;
    float w[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) w[a] = add_rn(add_rn(mul_rn(dx, R[a * 3]), mul_rn(dy, R[a * 3 + 1])), mul_rn(dz, R[a * 3 + 2]));
    float nrm = __fsqrt_rn(add_rn(add_rn(mul_rn(w[0], w[0]), mul_rn(w[1], w[1])), mul_rn(w[2], w[2])));
    nrm = fmaxf(nrm, 8.881784197001252e-16f);
#pragma unroll
    for (int a = 0; a < 3; ++a) d[v][a] = div_rn(w[a], nrm);
    if (v == 0) *nrm_out = nrm;
  }
  float sx = 0.f, sy = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float ex = d[0][a] - d[1][a], ey = d[0][a] - d[2][a];
    sx += ex * ex, sy += ey * ey;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) o[a] = t[a], dir[a] = d[0][a];
  *area = mul_rn(__fsqrt_rn(sx), __fsqrt_rn(sy));
}

__global__ void raygen_kernel(const float* __restrict__ c2w, const float* __restrict__ intr,
                              const float* __restrict__ dist, int any_dist, const int64_t* __restrict__ ray_indices,
                              int64_t n_rays, float* __restrict__ origins, float* __restrict__ directions,
                              float* __restrict__ pixel_area, float* __restrict__ directions_norm,
                              int64_t* __restrict__ camera_indices) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  const int64_t cam = ray_indices[3 * i];
  const float py = (float)ray_indices[3 * i + 1] + 0.5f, px = (float)ray_indices[3 * i + 2] + 0.5f;
  float o[3], d[3], area, nrm;
  ray_from_pixel(c2w, intr, any_dist ? dist : nullptr, nullptr, nullptr, cam, px, py, o, d, &area, &nrm);
#pragma unroll
  for (int a = 0; a < 3; ++a) origins[3 * i + a] = o[a], directions[3 * i + a] = d[a];
  if (pixel_area) pixel_area[i] = area;
  if (directions_norm) directions_norm[i] = nrm;
  if (camera_indices) camera_indices[i] = cam;
}

extern "C" int b2n_raygen(const float* c2w, const float* intr, const float* dist, const int64_t* ray_indices,
                          int64_t n_rays, float* origins, float* directions, float* pixel_area, float* directions_norm,
                          int64_t* camera_indices, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(c2w && intr && ray_indices && origins && directions, "null pointer");
  if (n_rays == 0) return B2N_OK;
  raygen_kernel<<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
      c2w, intr, dist, dist != nullptr, ray_indices, n_rays, origins, directions, pixel_area, directions_norm, camera_indices);
  B2N_LAUNCH_CHECK();
}

// Cameras.generate_rays(camera_indices, coords, camera_opt_to_camera, distortion_params_delta, keep_shape, ...)
// (cameras/cameras.py:321-503).  coords != NULL: ray i uses camera cam_idx[i] and pixel-centre coords[i] = (y, x).
// coords == NULL: whole images — rays laid out (height, width, n_cams) with the camera index fastest, which is the
// reference's (h, w, num_rays) shape for coords=None; pixel centres row + 0.5 / col + 0.5 (get_image_coords).
// cam_opt / dist_delta: per ray (coords given) or per listed camera (whole images).
__global__ void raygen_coords_kernel(const float* __restrict__ c2w, const float* __restrict__ intr,
                                     const float* __restrict__ dist, const int64_t* __restrict__ cam_idx,
                                     const float* __restrict__ coords, int64_t n_rays, int n_cams, int width,
                                     const float* __restrict__ cam_opt, const float* __restrict__ dist_delta,
                                     float* __restrict__ origins, float* __restrict__ directions,
                                     float* __restrict__ pixel_area, float* __restrict__ directions_norm,
                                     int64_t* __restrict__ camera_indices) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  int64_t cam, sel;
  float px, py;
  if (coords != nullptr) {
    sel = i, cam = cam_idx[i];
    py = __ldg(coords + 2 * i), px = __ldg(coords + 2 * i + 1);
  } else {
    sel = i % n_cams, cam = cam_idx[sel];
    const int64_t pix = i / n_cams;
    py = (float)(pix / width) + 0.5f, px = (float)(pix % width) + 0.5f;
  }
  float o[3], d[3], area, nrm;
  ray_from_pixel(c2w, intr, dist, dist_delta ? dist_delta + 6 * sel : nullptr, cam_opt ? cam_opt + 12 * sel : nullptr, cam, px,
                 py, o, d, &area, &nrm);
#pragma unroll
  for (int a = 0; a < 3; ++a) origins[3 * i + a] = o[a], directions[3 * i + a] = d[a];
  if (pixel_area) pixel_area[i] = area;
  if (directions_norm) directions_norm[i] = nrm;
  if (camera_indices) camera_indices[i] = cam;
}

extern "C" int b2n_raygen_coords(const float* c2w, const float* intr, const float* dist, const int64_t* cam_idx,
                                 const float* coords, int64_t n_rays, int32_t n_cams, int32_t height, int32_t width,
                                 const float* cam_opt, const float* dist_delta, float* origins, float* directions,
                                 float* pixel_area, float* directions_norm, int64_t* camera_indices, void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(c2w && intr && cam_idx && origins && directions, "null pointer");
  B2N_REQUIRE(coords != nullptr || (n_cams >= 1 && height >= 1 && width >= 1 && n_rays == (int64_t)n_cams * height * width),
              "whole-image mode needs n_rays == n_cams * height * width");
  raygen_coords_kernel<<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
      c2w, intr, dist, cam_idx, coords, n_rays, n_cams, width, cam_opt, dist_delta, origins, directions, pixel_area,
      directions_norm, camera_indices);
  B2N_LAUNCH_CHECK();
}

// CameraOptimizer.get_loss_dict (cameras/camera_optimizers.py:155-162): loss += mean_c |t_c| * trans_pen + mean_c |w_c| *
// rot_pen, with its gradient accumulated into d_pose (the norm's sub-gradient at 0 is 0, as torch's).  One block.
__global__ void __launch_bounds__(256) pose_regularizer_kernel(const float* __restrict__ pose, int n_cams, float trans_pen,
                                                               float rot_pen, float gscale, float* __restrict__ loss,
                                                               float* __restrict__ d_pose) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int c = threadIdx.x; c < n_cams; c += blockDim.x) {
    const float* p = pose + 6 * c;
    const float nt = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]), nw = sqrtf(p[3] * p[3] + p[4] * p[4] + p[5] * p[5]);
    acc += (nt * trans_pen + nw * rot_pen) / (float)n_cams;
    if (d_pose) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (nt > 0.f) d_pose[6 * c + a] += gscale * trans_pen * p[a] / (nt * (float)n_cams);
        if (nw > 0.f) d_pose[6 * c + 3 + a] += gscale * rot_pen * p[3 + a] / (nw * (float)n_cams);
      }
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    *loss += t;
  }
}

extern "C" int b2n_pose_regularizer(const float* pose_adjustment, int32_t n_cams, float trans_l2_penalty, float rot_l2_penalty,
                                    float gscale, float* loss_out, float* d_pose_adjustment, void* stream) {
  if (n_cams == 0) return B2N_OK;
  B2N_REQUIRE(pose_adjustment && (loss_out || d_pose_adjustment), "null pointer");
  pose_regularizer_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pose_adjustment, n_cams, trans_l2_penalty, rot_l2_penalty, gscale,
                                                                loss_out, d_pose_adjustment);
  B2N_LAUNCH_CHECK();
}

// Device-side training-ray pipeline: PixelSampler.sample_method + collate (data/pixel_samplers.py:137-174,
// 265-318) + RayGenerator (model_components/ray_generators.py:41-56) in one launch over a uint8 image cache that lives
// in HBM.  u [R,3] are the reference's `torch.rand((R,3))` draws; indices = (u * [C,H,W]).long() — a separately rounded
// fp32 multiply, then truncation, as torch evaluates it.  Pixel colours are converted like `get_image_float32`
// (uint8 / 255).  Outputs are written where the captured step wants them (any of them may be NULL).
__global__ void pixel_sample_raygen_kernel(const float* __restrict__ c2w, const float* __restrict__ intr,
                                           const float* __restrict__ dist, const uint8_t* __restrict__ images,
                                           int n_images, int height, int width, int channels,
                                           const int64_t* __restrict__ image_idx, const float* __restrict__ u,
                                           int64_t n_rays, int64_t* __restrict__ ray_indices, float* __restrict__ origins,
                                           float* __restrict__ directions, float* __restrict__ pixel_area,
                                           float* __restrict__ directions_norm, int64_t* __restrict__ camera_indices,
                                           float* __restrict__ rgb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  const int c = min((int)mul_rn(__ldg(u + 3 * i), (float)n_images), n_images - 1);
  const int y = min((int)mul_rn(__ldg(u + 3 * i + 1), (float)height), height - 1);
  const int x = min((int)mul_rn(__ldg(u + 3 * i + 2), (float)width), width - 1);
  const int64_t cam = image_idx ? image_idx[c] : (int64_t)c;  // cached subset -> absolute camera index
  if (rgb) {
    const uint8_t* px = images + (((size_t)c * height + y) * width + x) * channels;
#pragma unroll
    for (int a = 0; a < 3; ++a) rgb[3 * i + a] = div_rn((float)px[a], 255.f);
  }
  if (ray_indices) ray_indices[3 * i] = cam, ray_indices[3 * i + 1] = y, ray_indices[3 * i + 2] = x;
  float o[3], d[3], area, nrm;
  ray_from_pixel(c2w, intr, dist, nullptr, nullptr, cam, (float)x + 0.5f, (float)y + 0.5f, o, d, &area, &nrm);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (origins) origins[3 * i + a] = o[a];
    if (directions) directions[3 * i + a] = d[a];
  }
  if (pixel_area) pixel_area[i] = area;
  if (directions_norm) directions_norm[i] = nrm;
  if (camera_indices) camera_indices[i] = cam;
}

extern "C" int b2n_pixel_sample_raygen(const float* c2w, const float* intr, const float* dist, const uint8_t* images,
                                       int32_t n_images, int32_t height, int32_t width, int32_t channels,
                                       const int64_t* image_idx, const float* u, int64_t n_rays, int64_t* ray_indices,
                                       float* origins, float* directions, float* pixel_area, float* directions_norm,
                                       int64_t* camera_indices, float* rgb, void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(c2w && intr && u && (images || !rgb), "null pointer");
  B2N_REQUIRE(n_images >= 1 && height >= 1 && width >= 1 && channels >= 3, "bad image cache shape");
  pixel_sample_raygen_kernel<<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
      c2w, intr, dist, images, n_images, height, width, channels, image_idx, u, n_rays, ray_indices, origins, directions,
      pixel_area, directions_norm, camera_indices, rgb);
  B2N_LAUNCH_CHECK();
}

struct Box {
  float lo[3], hi[3];
};

__global__ void aabb_collide_kernel(const __grid_constant__ Box box, const float* __restrict__ o,
                                    const float* __restrict__ d, float near_plane, int64_t n, float* __restrict__ nears,
                                    float* __restrict__ fars) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float tn = -INFINITY, tf = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float inv = div_rn(1.f, add_rn(__ldg(d + 3 * i + a), 1e-6f));
    const float t1 = mul_rn(sub_rn(box.lo[a], __ldg(o + 3 * i + a)), inv), t2 = mul_rn(sub_rn(box.hi[a], __ldg(o + 3 * i + a)), inv);
    tn = fmaxf(tn, fminf(t1, t2)), tf = fminf(tf, fmaxf(t1, t2));
  }
  tn = fmaxf(tn, near_plane);
  tf = fmaxf(tf, add_rn(tn, 1e-6f));
  nears[i] = tn, fars[i] = tf;
}

extern "C" int b2n_aabb_collide(const float* origins, const float* directions, const float* aabb_host6, float near_plane,
                                int64_t n_rays, float* nears, float* fars, void* stream) {
  if (n_rays == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(origins && directions && aabb_host6 && nears && fars, "null pointer");
  if (n_rays == 0) return B2N_OK;
  Box box;
  for (int a = 0; a < 3; ++a) box.lo[a] = aabb_host6[a], box.hi[a] = aabb_host6[3 + a];
  aabb_collide_kernel<<<(unsigned)div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(box, origins, directions, near_plane, n_rays, nears, fars);
  B2N_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// a4: CameraOptimizer.apply_to_raybundle, mode SO3xR3 (cameras/camera_optimizers.py:148-153 with the exponential
// map of cameras/lie_groups.py:25-58 evaluated per ray): origins += t[cam], directions = R(w[cam]) directions with
//   theta = sqrt(max(|w|^2, 1e-4)), R = I + sin(theta)/theta K + (1 - cos(theta))/theta^2 K^2, K = skew(w).
// The backward follows the same chain of elementary operations the reference's autograd differentiates
// (f1 = inv*sin, f2 = inv*inv*(1-cos), K^2 by matrix product, clamp passing the gradient where |w|^2 >= 1e-4).
// ------------------------------------------------------------------------------------------------
struct PoseTerms {
  float K[3][3], K2[3][3], f1, f2, inv, s, c, theta, n;
};

__device__ __forceinline__ void pose_terms(const float* __restrict__ w, PoseTerms& p) {
  const float w0 = w[0], w1 = w[1], w2 = w[2];
  p.n = w0 * w0 + w1 * w1 + w2 * w2;
  p.theta = sqrtf(fmaxf(p.n, 1e-4f));
  p.inv = 1.f / p.theta;
  p.s = sinf(p.theta), p.c = cosf(p.theta);
  p.f1 = p.inv * p.s;
  p.f2 = p.inv * p.inv * (1.f - p.c);
  p.K[0][0] = 0.f, p.K[0][1] = -w2, p.K[0][2] = w1;
  p.K[1][0] = w2, p.K[1][1] = 0.f, p.K[1][2] = -w0;
  p.K[2][0] = -w1, p.K[2][1] = w0, p.K[2][2] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) p.K2[i][j] = p.K[i][0] * p.K[0][j] + p.K[i][1] * p.K[1][j] + p.K[i][2] * p.K[2][j];
}

__global__ void pose_apply_fwd_kernel(const float* __restrict__ pose, const int64_t* __restrict__ cam,
                                      const uint8_t* __restrict__ frozen, const float* __restrict__ o,
                                      const float* __restrict__ d, int64_t n, float* __restrict__ out_o,
                                      float* __restrict__ out_d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = __ldg(cam + i);
  const float d0 = __ldg(d + 3 * i), d1 = __ldg(d + 3 * i + 1), d2 = __ldg(d + 3 * i + 2);
  if (frozen != nullptr && frozen[c]) {  // non-trainable camera: identity transform
    out_o[3 * i] = __ldg(o + 3 * i), out_o[3 * i + 1] = __ldg(o + 3 * i + 1), out_o[3 * i + 2] = __ldg(o + 3 * i + 2);
    out_d[3 * i] = d0, out_d[3 * i + 1] = d1, out_d[3 * i + 2] = d2;
    return;
  }
  float t[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) t[k] = __ldg(pose + 6 * c + k);
  PoseTerms p;
  pose_terms(t + 3, p);
  const float dv[3] = {d0, d1, d2};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    out_o[3 * i + a] = __ldg(o + 3 * i + a) + t[a];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) acc += (p.f1 * p.K[a][j] + p.f2 * p.K2[a][j] + (a == j ? 1.f : 0.f)) * dv[j];
    out_d[3 * i + a] = acc;
  }
}

__global__ void pose_apply_bwd_kernel(const float* __restrict__ pose, const int64_t* __restrict__ cam,
                                      const uint8_t* __restrict__ frozen, const float* __restrict__ d,
                                      const float* __restrict__ g_o, const float* __restrict__ g_d, int64_t n,
                                      float* __restrict__ d_pose) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = __ldg(cam + i);
  if (frozen != nullptr && frozen[c]) return;
  float w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = __ldg(pose + 6 * c + 3 + k);
  PoseTerms p;
  pose_terms(w, p);
  float dR[3][3];  // dL/dR = g_d d^T
  float df1 = 0.f, df2 = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      dR[a][j] = (g_d ? __ldg(g_d + 3 * i + a) : 0.f) * __ldg(d + 3 * i + j);
      df1 += dR[a][j] * p.K[a][j];
      df2 += dR[a][j] * p.K2[a][j];
    }
  // K2 = K K  =>  dK = f1 dR + f2 (dR K^T + K^T dR)
  float dK[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) t += dR[a][k] * p.K[j][k] + p.K[k][a] * dR[k][j];
      dK[a][j] = p.f1 * dR[a][j] + p.f2 * t;
    }
  float dw[3] = {dK[2][1] - dK[1][2], dK[0][2] - dK[2][0], dK[1][0] - dK[0][1]};
  // f1 = inv*s, f2 = inv*inv*(1-c), inv = 1/theta, theta = sqrt(clamp(n, 1e-4)), n = w.w
  const float dinv = df1 * p.s + df2 * 2.f * p.inv * (1.f - p.c);
  float dtheta = df1 * p.inv * p.c + df2 * p.inv * p.inv * p.s;
  dtheta -= dinv * p.inv * p.inv;
  const float dn = (p.n >= 1e-4f) ? dtheta * 0.5f / p.theta : 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) dw[k] += 2.f * w[k] * dn;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (g_o) atomicAdd(d_pose + 6 * c + k, __ldg(g_o + 3 * i + k));
    atomicAdd(d_pose + 6 * c + 3 + k, dw[k]);
  }
}

extern "C" int b2n_pose_apply_fwd(const float* pose_adjustment, const int64_t* camera_indices, const uint8_t* frozen,
                                  const float* origins, const float* directions, int64_t n_rays, float* out_origins,
                                  float* out_directions, void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(pose_adjustment && camera_indices && origins && directions && out_origins && out_directions, "null pointer");
  pose_apply_fwd_kernel<<<(unsigned)div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
      pose_adjustment, camera_indices, frozen, origins, directions, n_rays, out_origins, out_directions);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_pose_apply_bwd(const float* pose_adjustment, const int64_t* camera_indices, const uint8_t* frozen,
                                  const float* directions, const float* d_out_origins, const float* d_out_directions,
                                  int64_t n_rays, float* d_pose_adjustment, void* stream) {
  if (n_rays == 0) return B2N_OK;
  B2N_REQUIRE(pose_adjustment && camera_indices && directions && d_pose_adjustment, "null pointer");
  B2N_REQUIRE(d_out_origins || d_out_directions, "no upstream gradient");
  pose_apply_bwd_kernel<<<(unsigned)div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
      pose_adjustment, camera_indices, frozen, directions, d_out_origins, d_out_directions, n_rays, d_pose_adjustment);
  B2N_LAUNCH_CHECK();
}
