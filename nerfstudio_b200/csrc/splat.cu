// 3D Gaussian splatting rasteriser behind gsplat.rendering.rasterization (gsplat 1.4.0; reference call site
// nerfstudio/models/splatfacto.py:555-581, SURVEY 8f-3 / BASELINE configs[4]).  gsplat's sources are not available:
// the stages follow its published algorithm (restated in oracle/splat_oracle.py) — PARITY UNPINNED.
//
//   project    one thread per Gaussian: quat/scale -> Sigma, EWA projection (+ eps2d blur) -> conic, 3-sigma radius,
//              pixel mean, depth, tile box; view-dependent colour from SH (degree <= 3).  Culled Gaussians get radius 0.
//   emit       (tile id << 32 | depth bits) keys + Gaussian ids at the offsets of an exclusive scan of the tile counts
//              (b2n_scan_counts); the sort of the 64-bit keys is the caller's (a library radix sort).
//   ranges     first / one-past-last sorted entry of every tile.
//   rasterise  one CTA per 16x16 tile, one thread per pixel; the tile's Gaussians are staged through shared memory in
//              batches of 256; front-to-back alpha blending with the early stop at T <= 1e-4; backward walks the same
//              list back to front, reduces per-Gaussian gradients over the warp before the atomics.
// Memory-bound integer/float streaming work: no tensor cores.
#include "common.cuh"

#define GS_TILE 16
#define GS_BLOCK (GS_TILE * GS_TILE)

struct GsCam {
  float R[9], t[3];  // world -> camera
  float fx, fy, cx, cy;
  int width, height, tiles_x, tiles_y;
  float near_plane, far_plane, eps2d, radius_clip;
};

__constant__ float GS_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                               0.5462742152960396f};
__constant__ float GS_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

__device__ __forceinline__ void gs_sh_basis(int degree, float x, float y, float z, float* Y) {
  Y[0] = 0.28209479177387814f;
  if (degree >= 1) {
    Y[1] = -0.4886025119029199f * y, Y[2] = 0.4886025119029199f * z, Y[3] = -0.4886025119029199f * x;
  }
  if (degree >= 2) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[4] = GS_C2[0] * xy, Y[5] = GS_C2[1] * yz, Y[6] = GS_C2[2] * (2.f * zz - xx - yy), Y[7] = GS_C2[3] * xz,
    Y[8] = GS_C2[4] * (xx - yy);
    if (degree >= 3) {
      Y[9] = GS_C3[0] * y * (3.f * xx - yy), Y[10] = GS_C3[1] * xy * z, Y[11] = GS_C3[2] * y * (4.f * zz - xx - yy);
      Y[12] = GS_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy), Y[13] = GS_C3[4] * x * (4.f * zz - xx - yy);
      Y[14] = GS_C3[5] * z * (xx - yy), Y[15] = GS_C3[6] * x * (xx - 3.f * yy);
    }
  }
}

// quaternion (w,x,y,z), normalised here -> rotation matrix (row-major)
__device__ __forceinline__ void gs_quat_to_R(const float* q, float* R) {
  const float inv = rsqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const float w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
  R[0] = 1.f - 2.f * (y * y + z * z), R[1] = 2.f * (x * y - w * z), R[2] = 2.f * (x * z + w * y);
  R[3] = 2.f * (x * y + w * z), R[4] = 1.f - 2.f * (x * x + z * z), R[5] = 2.f * (y * z - w * x);
  R[6] = 2.f * (x * z - w * y), R[7] = 2.f * (y * z + w * x), R[8] = 1.f - 2.f * (x * x + y * y);
}

__global__ void gs_project_fwd_kernel(const __grid_constant__ GsCam cam, int64_t n, const float* __restrict__ means,
                                      const float* __restrict__ quats, const float* __restrict__ scales,
                                      const float* __restrict__ sh, int sh_k, int sh_degree, float* __restrict__ means2d,
                                      float* __restrict__ depths, float* __restrict__ conics, int32_t* __restrict__ radii,
                                      int32_t* __restrict__ tiles_touched, float* __restrict__ colors) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float mx = __ldg(means + 3 * i), my = __ldg(means + 3 * i + 1), mz = __ldg(means + 3 * i + 2);
  const float x = cam.R[0] * mx + cam.R[1] * my + cam.R[2] * mz + cam.t[0];
  const float y = cam.R[3] * mx + cam.R[4] * my + cam.R[5] * mz + cam.t[1];
  const float z = cam.R[6] * mx + cam.R[7] * my + cam.R[8] * mz + cam.t[2];
  int rad = 0, touched = 0;
  float m2x = 0.f, m2y = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
  if (z > cam.near_plane && z < cam.far_plane) {
    const float limx = 1.3f * (0.5f * cam.width / cam.fx), limy = 1.3f * (0.5f * cam.height / cam.fy);
    const float rz = 1.f / z;
    const float tx = z * fminf(limx, fmaxf(-limx, x * rz)), ty = z * fminf(limy, fmaxf(-limy, y * rz));
    // J = [[fx/z, 0, -fx tx/z^2], [0, fy/z, -fy ty/z^2]];  T = J * Rcam (2x3)
    const float j00 = cam.fx * rz, j02 = -cam.fx * tx * rz * rz, j11 = cam.fy * rz, j12 = -cam.fy * ty * rz * rz;
    float T[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) T[c] = j00 * cam.R[c] + j02 * cam.R[6 + c], T[3 + c] = j11 * cam.R[3 + c] + j12 * cam.R[6 + c];
    float Rq[9], q[4] = {__ldg(quats + 4 * i), __ldg(quats + 4 * i + 1), __ldg(quats + 4 * i + 2), __ldg(quats + 4 * i + 3)};
    gs_quat_to_R(q, Rq);
    const float s[3] = {__ldg(scales + 3 * i), __ldg(scales + 3 * i + 1), __ldg(scales + 3 * i + 2)};
    // A = T * Rq * S (2x3); cov2d = A A^T
    float A[6];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) A[3 * r + c] = (T[3 * r] * Rq[c] + T[3 * r + 1] * Rq[3 + c] + T[3 * r + 2] * Rq[6 + c]) * s[c];
    const float a = A[0] * A[0] + A[1] * A[1] + A[2] * A[2] + cam.eps2d;
    const float b = A[0] * A[3] + A[1] * A[4] + A[2] * A[5];
    const float c = A[3] * A[3] + A[4] * A[4] + A[5] * A[5] + cam.eps2d;
    const float det = a * c - b * b;
    if (det > 0.f) {
      const float idet = 1.f / det;
      c0 = c * idet, c1 = -b * idet, c2 = a * idet;
      const float mid = 0.5f * (a + c);
      const float lam = mid + sqrtf(fmaxf(mid * mid - det, 0.1f));
      const float rf = ceilf(3.f * sqrtf(lam));
      m2x = cam.fx * x * rz + cam.cx, m2y = cam.fy * y * rz + cam.cy;
      if (rf > cam.radius_clip) {
        const int x0 = min(max((int)floorf((m2x - rf) / GS_TILE), 0), cam.tiles_x), x1 = min(max((int)ceilf((m2x + rf) / GS_TILE), 0), cam.tiles_x);
        const int y0 = min(max((int)floorf((m2y - rf) / GS_TILE), 0), cam.tiles_y), y1 = min(max((int)ceilf((m2y + rf) / GS_TILE), 0), cam.tiles_y);
        touched = (x1 - x0) * (y1 - y0);
        if (touched > 0) rad = (int)rf;
        else touched = 0;
      }
    }
  }
  means2d[2 * i] = m2x, means2d[2 * i + 1] = m2y;
  depths[i] = z;
  conics[3 * i] = c0, conics[3 * i + 1] = c1, conics[3 * i + 2] = c2;
  radii[i] = rad;
  tiles_touched[i] = rad > 0 ? touched : 0;
  if (colors != nullptr && sh != nullptr) {  // view-dependent colour: max(SH(dir) + 0.5, 0), dir = normalize(mu - cam_pos)
    const float px = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
    const float py = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
    const float pz = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
    float dx = mx - px, dy = my - py, dz = mz - pz;
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
    dx *= inv, dy *= inv, dz *= inv;
    float Y[16];
    gs_sh_basis(sh_degree, dx, dy, dz, Y);
    const int nb = (sh_degree + 1) * (sh_degree + 1);
    float rgb[3] = {0.5f, 0.5f, 0.5f};
    for (int k = 0; k < nb; ++k) {
      const float* c3 = sh + ((size_t)i * sh_k + k) * 3;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) rgb[ch] = fmaf(Y[k], __ldg(c3 + ch), rgb[ch]);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) colors[3 * i + ch] = fmaxf(rgb[ch], 0.f);
  }
}

// backward of gs_project_fwd_kernel: (v_means2d, v_depths, v_conics, v_colors) -> (v_means, v_quats, v_scales, v_sh)
__global__ void gs_project_bwd_kernel(const __grid_constant__ GsCam cam, int64_t n, const float* __restrict__ means,
                                      const float* __restrict__ quats, const float* __restrict__ scales,
                                      const float* __restrict__ sh, int sh_k, int sh_degree, const int32_t* __restrict__ radii,
                                      const float* __restrict__ conics, const float* __restrict__ colors,
                                      const float* __restrict__ v_means2d, const float* __restrict__ v_depths,
                                      const float* __restrict__ v_conics, const float* __restrict__ v_colors,
                                      float* __restrict__ v_means, float* __restrict__ v_quats, float* __restrict__ v_scales,
                                      float* __restrict__ v_sh) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float vm[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f};
  const float mx = __ldg(means + 3 * i), my = __ldg(means + 3 * i + 1), mz = __ldg(means + 3 * i + 2);
  const bool live = radii[i] > 0;
  if (live) {
    const float x = cam.R[0] * mx + cam.R[1] * my + cam.R[2] * mz + cam.t[0];
    const float y = cam.R[3] * mx + cam.R[4] * my + cam.R[5] * mz + cam.t[1];
    const float z = cam.R[6] * mx + cam.R[7] * my + cam.R[8] * mz + cam.t[2];
    const float rz = 1.f / z, rz2 = rz * rz;
    const float limx = 1.3f * (0.5f * cam.width / cam.fx), limy = 1.3f * (0.5f * cam.height / cam.fy);
    const bool clx = fabsf(x * rz) >= limx, cly = fabsf(y * rz) >= limy;
    const float tx = z * fminf(limx, fmaxf(-limx, x * rz)), ty = z * fminf(limy, fmaxf(-limy, y * rz));
    const float j00 = cam.fx * rz, j02 = -cam.fx * tx * rz2, j11 = cam.fy * rz, j12 = -cam.fy * ty * rz2;
    float T[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) T[c] = j00 * cam.R[c] + j02 * cam.R[6 + c], T[3 + c] = j11 * cam.R[3 + c] + j12 * cam.R[6 + c];
    float q[4] = {__ldg(quats + 4 * i), __ldg(quats + 4 * i + 1), __ldg(quats + 4 * i + 2), __ldg(quats + 4 * i + 3)};
    float Rq[9];
    gs_quat_to_R(q, Rq);
    const float sc[3] = {__ldg(scales + 3 * i), __ldg(scales + 3 * i + 1), __ldg(scales + 3 * i + 2)};
    float M[9], A[6];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) M[3 * r + c] = Rq[3 * r + c] * sc[c];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) A[3 * r + c] = T[3 * r] * M[c] + T[3 * r + 1] * M[3 + c] + T[3 * r + 2] * M[6 + c];
    // conic = X^-1 (X = cov2d): dL/dX = -X^-1 G X^-1 with G = [[v0, v1/2], [v1/2, v2]]
    const float c0 = conics[3 * i], c1 = conics[3 * i + 1], c2 = conics[3 * i + 2];
    const float g0 = v_conics[3 * i], g1 = 0.5f * v_conics[3 * i + 1], g2 = v_conics[3 * i + 2];
    // P = conic * G
    const float p00 = c0 * g0 + c1 * g1, p01 = c0 * g1 + c1 * g2, p10 = c1 * g0 + c2 * g1, p11 = c1 * g1 + c2 * g2;
    const float dX00 = -(p00 * c0 + p01 * c1), dX01 = -(p00 * c1 + p01 * c2), dX10 = -(p10 * c0 + p11 * c1),
                dX11 = -(p10 * c1 + p11 * c2);
    // X = A A^T: dA = (dX + dX^T) A
    const float s00 = 2.f * dX00, s01 = dX01 + dX10, s11 = 2.f * dX11;
    float dA[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) dA[c] = s00 * A[c] + s01 * A[3 + c], dA[3 + c] = s01 * A[c] + s11 * A[3 + c];
    // A = T M: dT = dA M^T (2x3), dM = T^T dA (3x3)
    float dT[6], dM[9];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) dT[3 * r + c] = dA[3 * r] * M[3 * c] + dA[3 * r + 1] * M[3 * c + 1] + dA[3 * r + 2] * M[3 * c + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) dM[3 * r + c] = T[r] * dA[c] + T[3 + r] * dA[3 + c];
    // M = Rq S
    float dR[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      vs[c] = dM[c] * Rq[c] + dM[3 + c] * Rq[3 + c] + dM[6 + c] * Rq[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r) dR[3 * r + c] = dM[3 * r + c] * sc[c];
    }
    {  // rotation matrix -> normalised quaternion -> raw quaternion
      const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), inv = 1.f / nrm;
      const float w = q[0] * inv, qx = q[1] * inv, qy = q[2] * inv, qz = q[3] * inv;
      float vn[4];
      vn[0] = 2.f * (qx * (dR[7] - dR[5]) + qy * (dR[2] - dR[6]) + qz * (dR[3] - dR[1]));
      vn[1] = 2.f * (-2.f * qx * (dR[4] + dR[8]) + qy * (dR[1] + dR[3]) + qz * (dR[2] + dR[6]) + w * (dR[7] - dR[5]));
      vn[2] = 2.f * (qx * (dR[1] + dR[3]) - 2.f * qy * (dR[0] + dR[8]) + qz * (dR[5] + dR[7]) + w * (dR[2] - dR[6]));
      vn[3] = 2.f * (qx * (dR[2] + dR[6]) + qy * (dR[5] + dR[7]) - 2.f * qz * (dR[0] + dR[4]) + w * (dR[3] - dR[1]));
      const float dot = vn[0] * w + vn[1] * qx + vn[2] * qy + vn[3] * qz;
      const float qn[4] = {w, qx, qy, qz};
#pragma unroll
      for (int k = 0; k < 4; ++k) vq[k] = (vn[k] - dot * qn[k]) * inv;
    }
    // T = J Rcam: dJ = dT Rcam^T
    const float dJ00 = dT[0] * cam.R[0] + dT[1] * cam.R[1] + dT[2] * cam.R[2];
    const float dJ02 = dT[0] * cam.R[6] + dT[1] * cam.R[7] + dT[2] * cam.R[8];
    const float dJ11 = dT[3] * cam.R[3] + dT[4] * cam.R[4] + dT[5] * cam.R[5];
    const float dJ12 = dT[3] * cam.R[6] + dT[4] * cam.R[7] + dT[5] * cam.R[8];
    const float v2x = v_means2d[2 * i], v2y = v_means2d[2 * i + 1];
    float vt[3];
    vt[0] = cam.fx * rz * v2x + (clx ? 0.f : -cam.fx * rz2 * dJ02);
    vt[1] = cam.fy * rz * v2y + (cly ? 0.f : -cam.fy * rz2 * dJ12);
    vt[2] = -(cam.fx * x * v2x + cam.fy * y * v2y) * rz2 - cam.fx * rz2 * dJ00 - cam.fy * rz2 * dJ11 +
            (clx ? cam.fx * tx * rz2 * rz : 2.f * cam.fx * tx * rz2 * rz) * dJ02 +
            (cly ? cam.fy * ty * rz2 * rz : 2.f * cam.fy * ty * rz2 * rz) * dJ12 + (v_depths ? v_depths[i] : 0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) vm[c] = cam.R[c] * vt[0] + cam.R[3 + c] * vt[1] + cam.R[6 + c] * vt[2];
  }
  if (sh != nullptr && v_colors != nullptr) {  // colours are evaluated for every Gaussian (culled ones receive zero v_colors)
    const float px = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
    const float py = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
    const float pz = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
    const float ex = mx - px, ey = my - py, ez = mz - pz;
    const float inv = rsqrtf(ex * ex + ey * ey + ez * ez);
    const float x = ex * inv, y = ey * inv, z = ez * inv;
    float Y[16];
    gs_sh_basis(sh_degree, x, y, z, Y);
    const int nb = (sh_degree + 1) * (sh_degree + 1);
    float vr[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) vr[ch] = colors[3 * i + ch] > 0.f ? v_colors[3 * i + ch] : 0.f;  // clamp_min(0)
    float dots[16];  // sum over channels of coefficient * upstream gradient, per basis function
    for (int k = 0; k < nb; ++k) {
      const float* c3 = sh + ((size_t)i * sh_k + k) * 3;
      dots[k] = __ldg(c3) * vr[0] + __ldg(c3 + 1) * vr[1] + __ldg(c3 + 2) * vr[2];
      if (v_sh) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v_sh[((size_t)i * sh_k + k) * 3 + ch] = Y[k] * vr[ch];
      }
    }
    if (v_sh)
      for (int k = nb; k < sh_k; ++k)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v_sh[((size_t)i * sh_k + k) * 3 + ch] = 0.f;
    float dx = 0.f, dy = 0.f, dz = 0.f;  // d(sum_k Y_k dots_k) / d(dir)
    if (sh_degree >= 1) {
      dy += -0.4886025119029199f * dots[1], dz += 0.4886025119029199f * dots[2], dx += -0.4886025119029199f * dots[3];
    }
    if (sh_degree >= 2) {
      dx += GS_C2[0] * y * dots[4], dy += GS_C2[0] * x * dots[4];
      dy += GS_C2[1] * z * dots[5], dz += GS_C2[1] * y * dots[5];
      dx += GS_C2[2] * -2.f * x * dots[6], dy += GS_C2[2] * -2.f * y * dots[6], dz += GS_C2[2] * 4.f * z * dots[6];
      dx += GS_C2[3] * z * dots[7], dz += GS_C2[3] * x * dots[7];
      dx += GS_C2[4] * 2.f * x * dots[8], dy += GS_C2[4] * -2.f * y * dots[8];
    }
    if (sh_degree >= 3) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      dx += GS_C3[0] * 6.f * xy * dots[9], dy += GS_C3[0] * (3.f * xx - 3.f * yy) * dots[9];
      dx += GS_C3[1] * yz * dots[10], dy += GS_C3[1] * xz * dots[10], dz += GS_C3[1] * xy * dots[10];
      dx += GS_C3[2] * -2.f * xy * dots[11], dy += GS_C3[2] * (4.f * zz - xx - 3.f * yy) * dots[11], dz += GS_C3[2] * 8.f * yz * dots[11];
      dx += GS_C3[3] * -6.f * xz * dots[12], dy += GS_C3[3] * -6.f * yz * dots[12], dz += GS_C3[3] * (6.f * zz - 3.f * xx - 3.f * yy) * dots[12];
      dx += GS_C3[4] * (4.f * zz - 3.f * xx - yy) * dots[13], dy += GS_C3[4] * -2.f * xy * dots[13], dz += GS_C3[4] * 8.f * xz * dots[13];
      dx += GS_C3[5] * 2.f * xz * dots[14], dy += GS_C3[5] * -2.f * yz * dots[14], dz += GS_C3[5] * (xx - yy) * dots[14];
      dx += GS_C3[6] * (3.f * xx - 3.f * yy) * dots[15], dy += GS_C3[6] * -6.f * xy * dots[15];
    }
    const float dd = dx * x + dy * y + dz * z;  // through the normalisation: (I - d d^T) / |e|
    vm[0] += (dx - dd * x) * inv, vm[1] += (dy - dd * y) * inv, vm[2] += (dz - dd * z) * inv;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) v_means[3 * i + c] = vm[c], v_scales[3 * i + c] = vs[c];
#pragma unroll
  for (int c = 0; c < 4; ++c) v_quats[4 * i + c] = vq[c];
}

__global__ void gs_emit_kernel(int64_t n, int tiles_x, int tiles_y, const float* __restrict__ means2d,
                               const int32_t* __restrict__ radii, const float* __restrict__ depths,
                               const int64_t* __restrict__ offsets, int64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || radii[i] <= 0) return;
  const float rf = (float)radii[i], mx = means2d[2 * i], my = means2d[2 * i + 1];
  const int x0 = min(max((int)floorf((mx - rf) / GS_TILE), 0), tiles_x), x1 = min(max((int)ceilf((mx + rf) / GS_TILE), 0), tiles_x);
  const int y0 = min(max((int)floorf((my - rf) / GS_TILE), 0), tiles_y), y1 = min(max((int)ceilf((my + rf) / GS_TILE), 0), tiles_y);
  int64_t o = offsets[i];
  const uint32_t dbits = __float_as_uint(depths[i]);  // depth > 0: float order == unsigned order
  for (int ty = y0; ty < y1; ++ty)
    for (int tx = x0; tx < x1; ++tx) {
      keys[o] = ((int64_t)(ty * tiles_x + tx) << 32) | (int64_t)dbits;
      vals[o] = (int32_t)i;
      ++o;
    }
}

__global__ void gs_tile_ranges_kernel(int64_t m, const int64_t* __restrict__ keys, int32_t* __restrict__ tile_lo,
                                      int32_t* __restrict__ tile_hi) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int t = (int)(keys[i] >> 32);
  if (i == 0 || (int)(keys[i - 1] >> 32) != t) tile_lo[t] = (int32_t)i;
  if (i == m - 1 || (int)(keys[i + 1] >> 32) != t) tile_hi[t] = (int32_t)(i + 1);
}

// ------------------------------------------------------------------------------------------------
// tile rasteriser
// ------------------------------------------------------------------------------------------------
template <int CH>  // colour channels: 3 (RGB) or 4 (RGB + depth)
__global__ void __launch_bounds__(GS_BLOCK) gs_rasterize_fwd_kernel(
    int width, int height, int tiles_x, const int32_t* __restrict__ tile_lo, const int32_t* __restrict__ tile_hi,
    const int32_t* __restrict__ ids, const float* __restrict__ means2d, const float* __restrict__ conics,
    const float* __restrict__ opacities, const float* __restrict__ colors, const float* __restrict__ extra,
    float* __restrict__ out, float* __restrict__ out_alpha, int32_t* __restrict__ last_idx) {
  __shared__ int32_t s_id[GS_BLOCK];
  __shared__ float s_xy[GS_BLOCK][2], s_con[GS_BLOCK][3], s_op[GS_BLOCK], s_col[GS_BLOCK][CH];
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const int px = blockIdx.x * GS_TILE + (threadIdx.x % GS_TILE), py = blockIdx.y * GS_TILE + (threadIdx.x / GS_TILE);
  const bool inside = px < width && py < height;
  const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
  const int lo = tile_lo[tile], hi = tile_hi[tile];
  float T = 1.f, acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = 0.f;
  bool done = !inside;
  int last = lo - 1;
  for (int base = lo; base < hi; base += GS_BLOCK) {
    if (__syncthreads_count(done) == GS_BLOCK) break;
    const int j = base + threadIdx.x;
    if (j < hi) {
      const int g = ids[j];
      s_id[threadIdx.x] = g;
      s_xy[threadIdx.x][0] = means2d[2 * g], s_xy[threadIdx.x][1] = means2d[2 * g + 1];
      s_con[threadIdx.x][0] = conics[3 * g], s_con[threadIdx.x][1] = conics[3 * g + 1], s_con[threadIdx.x][2] = conics[3 * g + 2];
      s_op[threadIdx.x] = opacities[g];
#pragma unroll
      for (int c = 0; c < 3; ++c) s_col[threadIdx.x][c] = colors[3 * g + c];
      if (CH == 4) s_col[threadIdx.x][CH - 1] = extra[g];
    }
    __syncthreads();
    const int cnt = min(GS_BLOCK, hi - base);
    for (int k = 0; k < cnt && !done; ++k) {
      const float dx = s_xy[k][0] - fx, dy = s_xy[k][1] - fy;
      const float sigma = 0.5f * (s_con[k][0] * dx * dx + s_con[k][2] * dy * dy) + s_con[k][1] * dx * dy;
      if (sigma < 0.f) continue;
      const float alpha = fminf(0.999f, s_op[k] * __expf(-sigma));
      if (alpha < 1.f / 255.f) continue;
      const float nT = T * (1.f - alpha);
      if (nT <= 1e-4f) {
        done = true;
        break;
      }
      const float w = alpha * T;
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = fmaf(w, s_col[k][c], acc[c]);
      T = nT;
      last = base + k;
    }
    __syncthreads();
  }
  if (inside) {
    const int64_t p = (int64_t)py * width + px;
#pragma unroll
    for (int c = 0; c < CH; ++c) out[p * CH + c] = acc[c];
    out_alpha[p] = 1.f - T;
    if (last_idx) last_idx[p] = last;
  }
}

template <int CH>
__global__ void __launch_bounds__(GS_BLOCK) gs_rasterize_bwd_kernel(
    int width, int height, int tiles_x, const int32_t* __restrict__ tile_lo, const int32_t* __restrict__ tile_hi,
    const int32_t* __restrict__ ids, const float* __restrict__ means2d, const float* __restrict__ conics,
    const float* __restrict__ opacities, const float* __restrict__ colors, const float* __restrict__ extra,
    const float* __restrict__ out_alpha, const int32_t* __restrict__ last_idx, const float* __restrict__ v_out,
    const float* __restrict__ v_out_alpha, float* __restrict__ v_means2d, float* __restrict__ v_conics,
    float* __restrict__ v_opacities, float* __restrict__ v_colors, float* __restrict__ v_extra) {
  __shared__ int32_t s_id[GS_BLOCK];
  __shared__ float s_xy[GS_BLOCK][2], s_con[GS_BLOCK][3], s_op[GS_BLOCK], s_col[GS_BLOCK][CH];
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const int px = blockIdx.x * GS_TILE + (threadIdx.x % GS_TILE), py = blockIdx.y * GS_TILE + (threadIdx.x / GS_TILE);
  const bool inside = px < width && py < height;
  const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
  const int64_t p = (int64_t)min(py, height - 1) * width + min(px, width - 1);
  const int lo = tile_lo[tile], hi = tile_hi[tile];
  const float T_final = inside ? 1.f - out_alpha[p] : 1.f;
  float T = T_final, buf[CH], vo[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) buf[c] = 0.f, vo[c] = inside ? v_out[p * CH + c] : 0.f;
  const float va = inside && v_out_alpha ? v_out_alpha[p] : 0.f;
  const int last = inside ? last_idx[p] : lo - 1;
  const int lane = threadIdx.x & 31;
  // walk the tile's list back to front in batches; every thread of a warp visits the same Gaussian at the same time so
  // the per-Gaussian gradients can be summed with shuffles before one atomic per warp
  const int n_batches = (hi - lo + GS_BLOCK - 1) / GS_BLOCK;
  for (int b = n_batches - 1; b >= 0; --b) {
    const int base = lo + b * GS_BLOCK;
    const int j = base + threadIdx.x;
    __syncthreads();
    if (j < hi) {
      const int g = ids[j];
      s_id[threadIdx.x] = g;
      s_xy[threadIdx.x][0] = means2d[2 * g], s_xy[threadIdx.x][1] = means2d[2 * g + 1];
      s_con[threadIdx.x][0] = conics[3 * g], s_con[threadIdx.x][1] = conics[3 * g + 1], s_con[threadIdx.x][2] = conics[3 * g + 2];
      s_op[threadIdx.x] = opacities[g];
#pragma unroll
      for (int c = 0; c < 3; ++c) s_col[threadIdx.x][c] = colors[3 * g + c];
      if (CH == 4) s_col[threadIdx.x][CH - 1] = extra[g];
    }
    __syncthreads();
    const int cnt = min(GS_BLOCK, hi - base);
    for (int k = cnt - 1; k >= 0; --k) {
      const int idx = base + k;
      bool valid = inside && idx <= last;
      float dx = 0.f, dy = 0.f, vis = 0.f, alpha = 0.f;
      if (valid) {
        dx = s_xy[k][0] - fx, dy = s_xy[k][1] - fy;
        const float sigma = 0.5f * (s_con[k][0] * dx * dx + s_con[k][2] * dy * dy) + s_con[k][1] * dx * dy;
        vis = __expf(-sigma);
        alpha = fminf(0.999f, s_op[k] * vis);
        if (sigma < 0.f || alpha < 1.f / 255.f) valid = false;
      }
      if (!__any_sync(0xffffffffu, valid)) continue;
      float g_col[CH], g_con[3] = {0.f, 0.f, 0.f}, g_xy[2] = {0.f, 0.f}, g_op = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) g_col[c] = 0.f;
      if (valid) {
        const float ra = 1.f / (1.f - alpha);
        T *= ra;  // transmittance in front of this Gaussian
        const float fac = alpha * T;
        float v_alpha = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          g_col[c] = fac * vo[c];
          v_alpha += (s_col[k][c] * T - buf[c] * ra) * vo[c];
          buf[c] += s_col[k][c] * fac;
        }
        v_alpha += T_final * ra * va;
        if (s_op[k] * vis <= 0.999f) {
          const float v_sigma = -s_op[k] * vis * v_alpha;
          g_con[0] = 0.5f * v_sigma * dx * dx, g_con[1] = v_sigma * dx * dy, g_con[2] = 0.5f * v_sigma * dy * dy;
          g_xy[0] = v_sigma * (s_con[k][0] * dx + s_con[k][1] * dy), g_xy[1] = v_sigma * (s_con[k][1] * dx + s_con[k][2] * dy);
          g_op = vis * v_alpha;
        }
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) g_col[c] = warp_sum(g_col[c]);
#pragma unroll
      for (int c = 0; c < 3; ++c) g_con[c] = warp_sum(g_con[c]);
      g_xy[0] = warp_sum(g_xy[0]), g_xy[1] = warp_sum(g_xy[1]), g_op = warp_sum(g_op);
      if (lane == 0) {
        const int g = s_id[k];
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(v_colors + 3 * g + c, g_col[c]);
        if (CH == 4 && v_extra) atomicAdd(v_extra + g, g_col[CH - 1]);
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(v_conics + 3 * g + c, g_con[c]);
        atomicAdd(v_means2d + 2 * g, g_xy[0]), atomicAdd(v_means2d + 2 * g + 1, g_xy[1]);
        atomicAdd(v_opacities + g, g_op);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
static int fill_cam(GsCam& c, const float* viewmat_host16, const float* k_host9, int width, int height, float near_plane,
                    float far_plane, float eps2d, float radius_clip) {
  if (!viewmat_host16 || !k_host9 || width < 1 || height < 1) return -1;
  for (int r = 0; r < 3; ++r) {
    for (int q = 0; q < 3; ++q) c.R[3 * r + q] = viewmat_host16[4 * r + q];
    c.t[r] = viewmat_host16[4 * r + 3];
  }
  c.fx = k_host9[0], c.fy = k_host9[4], c.cx = k_host9[2], c.cy = k_host9[5];
  c.width = width, c.height = height, c.tiles_x = (width + GS_TILE - 1) / GS_TILE, c.tiles_y = (height + GS_TILE - 1) / GS_TILE;
  c.near_plane = near_plane, c.far_plane = far_plane, c.eps2d = eps2d, c.radius_clip = radius_clip;
  return 0;
}

extern "C" int b2n_gs_project_fwd(const float* means, const float* quats, const float* scales, const float* sh,
                                  int32_t sh_k, int32_t sh_degree, int64_t n, const float* viewmat_host16,
                                  const float* k_host9, int32_t width, int32_t height, float near_plane, float far_plane,
                                  float eps2d, float radius_clip, float* means2d, float* depths, float* conics,
                                  int32_t* radii, int32_t* tiles_touched, float* colors, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(means && quats && scales && means2d && depths && conics && radii && tiles_touched, "null pointer");
  B2N_REQUIRE(sh == nullptr || (colors && sh_degree >= 0 && sh_degree <= 3 && sh_k >= (sh_degree + 1) * (sh_degree + 1)),
              "SH: degree 0..3 and K >= (degree+1)^2 coefficients");
  GsCam cam;
  B2N_REQUIRE(fill_cam(cam, viewmat_host16, k_host9, width, height, near_plane, far_plane, eps2d, radius_clip) == 0, "bad camera");
  gs_project_fwd_kernel<<<(unsigned)div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(
      cam, n, means, quats, scales, sh, sh_k, sh_degree, means2d, depths, conics, radii, tiles_touched, colors);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_gs_project_bwd(const float* means, const float* quats, const float* scales, const float* sh,
                                  int32_t sh_k, int32_t sh_degree, int64_t n, const float* viewmat_host16,
                                  const float* k_host9, int32_t width, int32_t height, float near_plane, float far_plane,
                                  float eps2d, float radius_clip, const int32_t* radii, const float* conics,
                                  const float* colors, const float* v_means2d, const float* v_depths, const float* v_conics,
                                  const float* v_colors, float* v_means, float* v_quats, float* v_scales, float* v_sh,
                                  void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(means && quats && scales && radii && conics && v_means2d && v_conics && v_means && v_quats && v_scales,
              "null pointer");
  B2N_REQUIRE(sh == nullptr || (colors && sh_degree >= 0 && sh_degree <= 3 && sh_k >= (sh_degree + 1) * (sh_degree + 1)),
              "SH: degree 0..3 and K >= (degree+1)^2 coefficients");
  GsCam cam;
  B2N_REQUIRE(fill_cam(cam, viewmat_host16, k_host9, width, height, near_plane, far_plane, eps2d, radius_clip) == 0, "bad camera");
  gs_project_bwd_kernel<<<(unsigned)div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(
      cam, n, means, quats, scales, sh, sh_k, sh_degree, radii, conics, colors, v_means2d, v_depths, v_conics, v_colors, v_means,
      v_quats, v_scales, v_sh);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_gs_emit(const float* means2d, const int32_t* radii, const float* depths, const int64_t* offsets, int64_t n,
                           int32_t width, int32_t height, int64_t* keys, int32_t* gaussian_ids, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(means2d && radii && depths && offsets && keys && gaussian_ids, "null pointer");
  gs_emit_kernel<<<(unsigned)div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(
      n, (width + GS_TILE - 1) / GS_TILE, (height + GS_TILE - 1) / GS_TILE, means2d, radii, depths, offsets, keys, gaussian_ids);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_gs_tile_ranges(const int64_t* sorted_keys, int64_t m, int32_t* tile_lo, int32_t* tile_hi, void* stream) {
  if (m == 0) return B2N_OK;
  B2N_REQUIRE(sorted_keys && tile_lo && tile_hi, "null pointer");
  gs_tile_ranges_kernel<<<(unsigned)div_up(m, 256), 256, 0, (cudaStream_t)stream>>>(m, sorted_keys, tile_lo, tile_hi);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_gs_rasterize_fwd(int32_t width, int32_t height, const int32_t* tile_lo, const int32_t* tile_hi,
                                    const int32_t* sorted_ids, const float* means2d, const float* conics,
                                    const float* opacities, const float* colors, const float* extra, float* out,
                                    float* out_alpha, int32_t* last_idx, void* stream) {
  B2N_REQUIRE(width >= 1 && height >= 1 && tile_lo && tile_hi && means2d && conics && opacities && colors && out && out_alpha,
              "bad arguments");
  const dim3 grid((width + GS_TILE - 1) / GS_TILE, (height + GS_TILE - 1) / GS_TILE);
  if (extra)
    gs_rasterize_fwd_kernel<4><<<grid, GS_BLOCK, 0, (cudaStream_t)stream>>>(width, height, grid.x, tile_lo, tile_hi, sorted_ids,
                                                                             means2d, conics, opacities, colors, extra, out,
                                                                             out_alpha, last_idx);
  else
    gs_rasterize_fwd_kernel<3><<<grid, GS_BLOCK, 0, (cudaStream_t)stream>>>(width, height, grid.x, tile_lo, tile_hi, sorted_ids,
                                                                             means2d, conics, opacities, colors, nullptr, out,
                                                                             out_alpha, last_idx);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_gs_rasterize_bwd(int32_t width, int32_t height, const int32_t* tile_lo, const int32_t* tile_hi,
                                    const int32_t* sorted_ids, const float* means2d, const float* conics,
                                    const float* opacities, const float* colors, const float* extra, const float* out_alpha,
                                    const int32_t* last_idx, const float* v_out, const float* v_out_alpha, float* v_means2d,
                                    float* v_conics, float* v_opacities, float* v_colors, float* v_extra, void* stream) {
  B2N_REQUIRE(width >= 1 && height >= 1 && tile_lo && tile_hi && means2d && conics && opacities && colors && out_alpha &&
                  last_idx && v_out && v_means2d && v_conics && v_opacities && v_colors,
              "bad arguments");
  const dim3 grid((width + GS_TILE - 1) / GS_TILE, (height + GS_TILE - 1) / GS_TILE);
  if (extra)
    gs_rasterize_bwd_kernel<4><<<grid, GS_BLOCK, 0, (cudaStream_t)stream>>>(
        width, height, grid.x, tile_lo, tile_hi, sorted_ids, means2d, conics, opacities, colors, extra, out_alpha, last_idx, v_out,
        v_out_alpha, v_means2d, v_conics, v_opacities, v_colors, v_extra);
  else
    gs_rasterize_bwd_kernel<3><<<grid, GS_BLOCK, 0, (cudaStream_t)stream>>>(
        width, height, grid.x, tile_lo, tile_hi, sorted_ids, means2d, conics, opacities, colors, nullptr, out_alpha, last_idx, v_out,
        v_out_alpha, v_means2d, v_conics, v_opacities, v_colors, nullptr);
  B2N_LAUNCH_CHECK();
}
