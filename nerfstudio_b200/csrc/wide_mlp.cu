// Wide dense layers (vanilla-nerf: 8 x 256 with a skip, 128-wide colour branch; fields/vanilla_nerf_field.py:84-107,
// field_components/mlp.py:160-179) on our own kernels: networks whose weights do not fit the fused tiny-MLP kernels' shared
// memory run layer by layer through one register-tiled fp32 GEMM with a fused bias + activation epilogue.
//
//   linear_fwd   Y[n,out] = act(X[n,in] W[out,in]^T + b)                               (NT GEMM)
//   linear_bwd   dZ = dY * act'(Y);  dX[n,in] = dZ W (NN);  dW[out,in] += dZ^T X (TN);  db[out] += column sums of dZ
//
// Exact fp32 FFMA (the parity mode of BASELINE configs[0], the reference's CPU case).  64 x 64 output tile per CTA, 16-deep
// k-slabs staged in shared memory, 4 x 4 outputs per thread; all shapes guarded.  dW accumulates with atomics because the
// reduction dimension (rows) is split over blockIdx.z.
#include "common.cuh"

#define WM_BM 64
#define WM_BN 64
#define WM_BK 16

__device__ __forceinline__ float wm_act(int act, float v) {
  switch (act) {
    case B2N_ACT_RELU: return fmaxf(v, 0.f);
    case B2N_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case B2N_ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));  // torch.nn.Softplus (beta 1, threshold 20)
    case B2N_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// C[M,N] (+)= op(A) op(B), row-major.  TA: A is [K,M] (use A^T), else [M,K].  TB: B is [N,K] (use B^T), else [K,N].
// epilogue: + bias[N], activation.  split-K over blockIdx.z with atomics when gridDim.z > 1 (accumulate semantics).
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) wm_gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                      const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                                      const float* __restrict__ bias, int act, int accumulate) {
  __shared__ float As[WM_BK][WM_BM + 4], Bs[WM_BK][WM_BN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * WM_BM, n0 = blockIdx.x * WM_BN;
  const int kz = (K + gridDim.z - 1) / gridDim.z, k_lo = blockIdx.z * kz, k_hi = min(K, k_lo + kz);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = k_lo; k0 < k_hi; k0 += WM_BK) {
    for (int e = threadIdx.x; e < WM_BK * WM_BM; e += 256) {
      int kk, mm;
      if (TA) mm = e % WM_BM, kk = e / WM_BM;  // A^T stored [K,M]: consecutive threads along M
      else kk = e % WM_BK, mm = e / WM_BK;     // A stored [M,K]: consecutive threads along K
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < k_hi) ? (TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk]) : 0.f;
    }
    for (int e = threadIdx.x; e < WM_BK * WM_BN; e += 256) {
      int kk, nn;
      if (TB) kk = e % WM_BK, nn = e / WM_BK;  // B^T stored [N,K]
      else nn = e % WM_BN, kk = e / WM_BN;     // B stored [K,N]
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < k_hi) ? (TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WM_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i], b[i] = Bs[kk][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* dst = C + (size_t)gm * ldc + gn;
      if (gridDim.z > 1) {
        atomicAdd(dst, acc[i][j]);
      } else {
        float v = acc[i][j] + (bias ? __ldg(bias + gn) : 0.f);
        v = wm_act(act, v);
        *dst = accumulate ? *dst + v : v;
      }
    }
  }
}

// dZ = dY * act'(Y) (Y = post-activation output);  colsum[out] += sum over rows of dZ (optional)
__global__ void __launch_bounds__(256) wm_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int64_t n,
                                                         int width, int act, float* __restrict__ dz, float* __restrict__ colsum) {
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rlane = threadIdx.x >> 5;  // 8 row lanes
  float s = 0.f;
  if (col < width) {
    for (int64_t r = (int64_t)blockIdx.y * 8 + rlane; r < n; r += (int64_t)gridDim.y * 8) {
      const float g = dy[r * width + col], v = y[r * width + col];
      float d;
      switch (act) {
        case B2N_ACT_RELU: d = v > 0.f ? g : 0.f; break;
        case B2N_ACT_SIGMOID: d = g * v * (1.f - v); break;
        case B2N_ACT_TANH: d = g * (1.f - v * v); break;
        case B2N_ACT_SOFTPLUS: d = g * (1.f - expf(-v)); break;  // y = softplus(z): sigmoid(z) = 1 - exp(-y)
        default: d = g;
      }
      dz[r * width + col] = d;
      s += d;
    }
  }
  __shared__ float red[8][33];
  red[rlane][threadIdx.x & 31] = s;
  __syncthreads();
  if (rlane == 0 && col < width && colsum) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    atomicAdd(colsum + col, t);
  }
}

extern "C" int b2n_linear_fwd(const float* x, int64_t n, int32_t in_dim, int64_t x_stride, const float* w, const float* b,
                              int32_t out_dim, int32_t act, float* y, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(x && w && y && in_dim >= 1 && out_dim >= 1 && x_stride >= in_dim, "bad arguments");
  B2N_REQUIRE(n < ((int64_t)1 << 31), "n must fit int32");
  dim3 grid((out_dim + WM_BN - 1) / WM_BN, (unsigned)div_up(n, WM_BM), 1);
  wm_gemm_kernel<false, true><<<grid, 256, 0, (cudaStream_t)stream>>>((int)n, out_dim, in_dim, x, (int)x_stride, w, in_dim, y,
                                                                       out_dim, b, act, 0);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_linear_bwd(const float* x, int64_t n, int32_t in_dim, int64_t x_stride, const float* w, int32_t out_dim,
                              int32_t act, const float* y, const float* dy, float* dz_scratch, float* dx, int64_t dx_stride,
                              float* dw, float* db, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(x && w && y && dy && dz_scratch && in_dim >= 1 && out_dim >= 1, "bad arguments");
  B2N_REQUIRE(n < ((int64_t)1 << 31), "n must fit int32");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 ga((out_dim + 31) / 32, (unsigned)min((int64_t)512, div_up(n, 8)));
  wm_act_bwd_kernel<<<ga, 256, 0, st>>>(dy, y, n, out_dim, act, dz_scratch, db);
  if (dx) {  // dX = dZ W   [n,out] x [out,in]
    dim3 g1((in_dim + WM_BN - 1) / WM_BN, (unsigned)div_up(n, WM_BM), 1);
    wm_gemm_kernel<false, false><<<g1, 256, 0, st>>>((int)n, in_dim, out_dim, dz_scratch, out_dim, w, in_dim, dx, (int)dx_stride,
                                                      nullptr, B2N_ACT_NONE, 0);
  }
  if (dw) {  // dW += dZ^T X   [out,n] x [n,in], reduction over n split across blockIdx.z
    const int split = (int)min((int64_t)64, max((int64_t)2, div_up(n, 2048)));
    dim3 g2((in_dim + WM_BN - 1) / WM_BN, (out_dim + WM_BM - 1) / WM_BM, split);
    wm_gemm_kernel<true, false><<<g2, 256, 0, st>>>(out_dim, in_dim, (int)n, dz_scratch, out_dim, x, (int)x_stride, dw, in_dim,
                                                     nullptr, B2N_ACT_NONE, 1);
  }
  B2N_LAUNCH_CHECK();
}
