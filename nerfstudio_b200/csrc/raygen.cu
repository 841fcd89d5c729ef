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

__global__ void raygen_kernel(const float* __restrict__ c2w, const float* __restrict__ intr,
                              const float* __restrict__ dist, int any_dist, const int64_t* __restrict__ ray_indices,
                              int64_t n_rays, float* __restrict__ origins, float* __restrict__ directions,
                              float* __restrict__ pixel_area, float* __restrict__ directions_norm,
                              int64_t* __restrict__ camera_indices) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  const int64_t cam = ray_indices[3 * i];
  const float py = (float)ray_indices[3 * i + 1] + 0.5f, px = (float)ray_indices[3 * i + 2] + 0.5f;
  const float fx = __ldg(intr + 4 * cam), fy = __ldg(intr + 4 * cam + 1), cx = __ldg(intr + 4 * cam + 2), cy = __ldg(intr + 4 * cam + 3);
  float ux[3], uy[3];
  ux[0] = div_rn(sub_rn(px, cx), fx), uy[0] = div_rn(sub_rn(py, cy), fy);
  ux[1] = div_rn(add_rn(sub_rn(px, cx), 1.f), fx), uy[1] = uy[0];
  ux[2] = ux[0], uy[2] = div_rn(add_rn(sub_rn(py, cy), 1.f), fy);
  if (any_dist) {
    float k[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) k[j] = __ldg(dist + 6 * cam + j);
#pragma unroll
    for (int v = 0; v < 3; ++v) undistort(ux[v], uy[v], k);
  }
  float R[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) R[a * 3 + b] = __ldg(c2w + cam * 12 + a * 4 + b);
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
    if (v == 0 && directions_norm) directions_norm[i] = nrm;
  }
  float sx = 0.f, sy = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float ex = d[0][a] - d[1][a], ey = d[0][a] - d[2][a];
    sx += ex * ex, sy += ey * ey;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    origins[3 * i + a] = __ldg(c2w + cam * 12 + a * 4 + 3);
    directions[3 * i + a] = d[0][a];
  }
  if (pixel_area) pixel_area[i] = mul_rn(__fsqrt_rn(sx), __fsqrt_rn(sy));
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
