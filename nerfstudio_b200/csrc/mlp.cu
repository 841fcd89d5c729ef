// K3 — tiny fused MLP, fp32 SIMT (exact-parity path), forward and backward, for sm_100a.
//
// Replaces MLP.pytorch_fwd / tcnn FullyFusedMLP (nerfstudio/field_components/mlp.py:110-114,160-184) for the
// widths the hot path uses (base 32->64->16, head 63->64->64->3, proposal 10->16->1).  The whole network lives
// in one kernel: all layer weights are staged in shared memory once per CTA, a CTA owns a tile of 128 rows,
// activations stay in shared memory feature-major ([k][row], conflict-free for one-row-per-thread access) and
// every weight read is a warp-uniform 128-bit broadcast.  Accumulation is plain fp32 FFMA in k order, so the
// result agrees with the reference's nn.Linear to ~1e-6 relative — this is the mode the 1e-4 parity tests use.
//
// Backward is one persistent kernel (one CTA per SM): per tile it walks the layers in reverse, forming
//   dW += dZ^T A   (4x4 register patches over the 128-row tile, accumulated in shared memory across tiles)
//   dA  = dZ W     (one row per thread, weights broadcast)
// and flushes dW/db to global memory with one RED per entry per CTA at the end.
#include <string.h>

#include "common.cuh"

#define MLP_ROWS 128
#define MLP_LD (MLP_ROWS + 1)
#define MLP_THREADS 256

struct MlpParams {
  int n_layers, in_dim, hidden_act, out_act;
  int in[B2N_MAX_MLP_LAYERS];    // layer input width (incl. skip concat)
  int out[B2N_MAX_MLP_LAYERS];   // layer output width
  int skip[B2N_MAX_MLP_LAYERS];
  int outp[B2N_MAX_MLP_LAYERS];  // out rounded up to a multiple of 8 (fwd transposed copy)
  int inp[B2N_MAX_MLP_LAYERS];   // in rounded up to a multiple of 8 (bwd native copy)
  int w_s[B2N_MAX_MLP_LAYERS];   // smem float offsets
  int b_s[B2N_MAX_MLP_LAYERS];
  const float* w[B2N_MAX_MLP_LAYERS];   // device pointers, nn.Linear layout [out][in]
  const float* b[B2N_MAX_MLP_LAYERS];   // nullptr = no bias
  float* dw[B2N_MAX_MLP_LAYERS];        // gradient destinations (backward only)
  float* db[B2N_MAX_MLP_LAYERS];
  long long hid_off[B2N_MAX_MLP_LAYERS];  // feature offset of layer's saved activations
  int w_total;                            // floats of smem weights+biases
  int kmax_in, wmax_out;
  int any_skip;
};

__device__ __forceinline__ float act_apply(int act, float v) {
  switch (act) {
    case B2N_ACT_RELU: return fmaxf(v, 0.f);
    case B2N_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case B2N_ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));  // torch softplus threshold=20
    case B2N_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float act_grad_from_out(int act, float y) {
  switch (act) {
    case B2N_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case B2N_ACT_SIGMOID: return y * (1.f - y);
    case B2N_ACT_SOFTPLUS: return y > 20.f ? 1.f : 1.f - expf(-y);
    case B2N_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

// acc[0..CH) += sum_k a[k][r] * W[k][j0 + i]      (W row stride `ldw`, rows k0..k0+K)
template <int CH>
__device__ __forceinline__ void dot_chunk(float (&acc)[CH], const float* __restrict__ A, int K, int r,
                                          const float* __restrict__ W, int ldw) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float a = A[k * MLP_LD + r];
    const float4* w4 = reinterpret_cast<const float4*>(W + k * ldw);
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
      const float4 w = w4[i];
      acc[4 * i + 0] = fmaf(a, w.x, acc[4 * i + 0]);
      acc[4 * i + 1] = fmaf(a, w.y, acc[4 * i + 1]);
      acc[4 * i + 2] = fmaf(a, w.z, acc[4 * i + 2]);
      acc[4 * i + 3] = fmaf(a, w.w, acc[4 * i + 3]);
    }
  }
}

// One layer (or its transpose) for this thread's row and half of the outputs.
//   out[j][r] = epilogue( bias[j] + sum over segments seg: sum_k A_seg[k][r] * W[(k0_seg + k)][j] )
template <int CH, typename Epi>
__device__ __forceinline__ void layer_chunk(const float* A0, int K0, const float* A1, int K1, const float* W, int ldw,
                                            const float* bias, int j0, int r, Epi epi) {
  float acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = bias ? bias[j0 + i] : 0.f;
  dot_chunk<CH>(acc, A0, K0, r, W + j0, ldw);
  if (K1 > 0) dot_chunk<CH>(acc, A1, K1, r, W + (size_t)K0 * ldw + j0, ldw);
#pragma unroll
  for (int i = 0; i < CH; ++i) epi(j0 + i, acc[i]);
}

template <typename Epi>
__device__ __forceinline__ void layer_rows(const float* A0, int K0, const float* A1, int K1, const float* W, int ldw,
                                           const float* bias, int width_p /*multiple of 8*/, int r, int half,
                                           Epi epi) {
  const int wh = width_p >> 1;  // outputs per half, multiple of 4
  int j = half * wh;
  const int jend = j + wh;
  while (jend - j >= 32) { layer_chunk<32>(A0, K0, A1, K1, W, ldw, bias, j, r, epi); j += 32; }
  if (jend - j >= 16) { layer_chunk<16>(A0, K0, A1, K1, W, ldw, bias, j, r, epi); j += 16; }
  if (jend - j >= 8) { layer_chunk<8>(A0, K0, A1, K1, W, ldw, bias, j, r, epi); j += 8; }
  if (jend - j >= 4) { layer_chunk<4>(A0, K0, A1, K1, W, ldw, bias, j, r, epi); j += 4; }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MLP_THREADS) mlp_fwd_kernel(const __grid_constant__ MlpParams mp,
                                                              const float* __restrict__ x, int64_t n,
                                                              float* __restrict__ y, float* __restrict__ hidden) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                                   // transposed weights [in][outp] + bias per layer
  float* bufX = Ws + mp.w_total;                      // [in_dim][LD]
  float* H0 = bufX + (size_t)mp.in_dim * MLP_LD;      // [wmax_out][LD]
  float* H1 = H0 + (size_t)mp.wmax_out * MLP_LD;
  const int t = threadIdx.x;

  // stage weights: Wt[k][j] = W[j][k], zero padded
  for (int l = 0; l < mp.n_layers; ++l) {
    const int in = mp.in[l], out = mp.out[l], outp = mp.outp[l];
    float* wt = Ws + mp.w_s[l];
    const float* wg = mp.w[l];
    for (int idx = t; idx < in * outp; idx += MLP_THREADS) {
      const int k = idx / outp, j = idx - k * outp;
      wt[idx] = j < out ? __ldg(wg + (size_t)j * in + k) : 0.f;
    }
    float* bs = Ws + mp.b_s[l];
    for (int j = t; j < outp; j += MLP_THREADS) bs[j] = (j < out && mp.b[l] != nullptr) ? __ldg(mp.b[l] + j) : 0.f;
  }

  const int r = t & (MLP_ROWS - 1), half = t >> 7;
  const int64_t n_tiles = (n + MLP_ROWS - 1) / MLP_ROWS;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * MLP_ROWS;
    const int rows = (int)min((int64_t)MLP_ROWS, n - row0);
    __syncthreads();
    // x tile -> bufX[k][r]
    {
      const int in = mp.in_dim;
      const float* xg = x + row0 * in;
      for (int idx = t; idx < MLP_ROWS * in; idx += MLP_THREADS) {
        const int rr = idx / in, k = idx - rr * in;
        bufX[k * MLP_LD + rr] = rr < rows ? __ldg(xg + idx) : 0.f;
      }
    }
    __syncthreads();
    const float* prev = bufX;
    int prev_w = mp.in_dim;
    float* cur = H0;
    for (int l = 0; l < mp.n_layers; ++l) {
      const bool last = (l == mp.n_layers - 1);
      const int act = last ? mp.out_act : mp.hidden_act;
      const float* A0 = prev;
      int K0 = prev_w;
      const float* A1 = nullptr;
      int K1 = 0;
      if (mp.skip[l]) { A0 = bufX, K0 = mp.in_dim, A1 = prev, K1 = prev_w; }
      float* dst = cur;
      layer_rows(A0, K0, A1, K1, Ws + mp.w_s[l], mp.outp[l], Ws + mp.b_s[l], mp.outp[l], r, half,
                 [&](int j, float v) { dst[j * MLP_LD + r] = act_apply(act, v); });
      __syncthreads();
      const int out = mp.out[l];
      if (!last) {
        if (hidden != nullptr) {
          float* hg = hidden + mp.hid_off[l] * n + row0;
          for (int idx = t; idx < out * MLP_ROWS; idx += MLP_THREADS) {
            const int j = idx >> 7, rr = idx & (MLP_ROWS - 1);
            if (rr < rows) hg[(size_t)j * n + rr] = cur[j * MLP_LD + rr];
          }
        }
      } else {
        float* yg = y + row0 * out;
        for (int idx = t; idx < rows * out; idx += MLP_THREADS) {
          const int rr = idx / out, j = idx - rr * out;
          yg[idx] = cur[j * MLP_LD + rr];
        }
      }
      prev = cur, prev_w = out;
      cur = (cur == H0) ? H1 : H0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MLP_THREADS) mlp_bwd_kernel(const __grid_constant__ MlpParams mp,
                                                              const float* __restrict__ x,
                                                              const float* __restrict__ y,
                                                              const float* __restrict__ hidden,
                                                              const float* __restrict__ dy, int64_t n,
                                                              float* __restrict__ dx) {
  extern __shared__ __align__(16) float smem[];
  float* Wn = smem;                                   // native weights [out][inp] per layer (+ unused bias slot)
  float* dWs = Wn + mp.w_total;                       // same layout: dW [out][inp], db [out]
  float* dZ = dWs + mp.w_total;                       // [wmax_out][LD]
  float* A = dZ + (size_t)mp.wmax_out * MLP_LD;       // [kmax_in][LD] layer input
  float* dA = A + (size_t)mp.kmax_in * MLP_LD;        // [kmax_in][LD]
  float* dXs = dA + (size_t)mp.kmax_in * MLP_LD;      // [in_dim][LD], only when any_skip
  const int t = threadIdx.x;

  for (int l = 0; l < mp.n_layers; ++l) {
    const int in = mp.in[l], out = mp.out[l], inp = mp.inp[l];
    float* wn = Wn + mp.w_s[l];
    const float* wg = mp.w[l];
    for (int idx = t; idx < out * inp; idx += MLP_THREADS) {
      const int j = idx / inp, k = idx - j * inp;
      wn[idx] = k < in ? __ldg(wg + (size_t)j * in + k) : 0.f;
    }
  }
  for (int idx = t; idx < mp.w_total; idx += MLP_THREADS) dWs[idx] = 0.f;

  const int r = t & (MLP_ROWS - 1), half = t >> 7;
  const int64_t n_tiles = (n + MLP_ROWS - 1) / MLP_ROWS;
  const int L = mp.n_layers;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * MLP_ROWS;
    const int rows = (int)min((int64_t)MLP_ROWS, n - row0);
    __syncthreads();
    {  // dZ_L = dy * act'(y)
      const int out = mp.out[L - 1];
      const float* dyg = dy + row0 * out;
      const float* yg = y + row0 * out;
      for (int idx = t; idx < MLP_ROWS * out; idx += MLP_THREADS) {
        const int rr = idx / out, j = idx - rr * out;
        float v = 0.f;
        if (rr < rows) v = __ldg(dyg + idx) * act_grad_from_out(mp.out_act, __ldg(yg + idx));
        dZ[j * MLP_LD + rr] = v;
      }
      if (mp.any_skip)
        for (int idx = t; idx < mp.in_dim * MLP_LD; idx += MLP_THREADS) dXs[idx] = 0.f;
    }
    for (int l = L - 1; l >= 0; --l) {
      const int in = mp.in[l], out = mp.out[l], inp = mp.inp[l];
      const int hprev_w = l > 0 ? mp.out[l - 1] : 0;
      const int x_rows = (l == 0 || mp.skip[l]) ? mp.in_dim : 0;  // leading rows of A that are the network input
      // ---- layer input A = [x ; hidden_{l-1}] (either part may be absent)
      if (x_rows) {
        const float* xg = x + row0 * mp.in_dim;
        for (int idx = t; idx < MLP_ROWS * mp.in_dim; idx += MLP_THREADS) {
          const int rr = idx / mp.in_dim, k = idx - rr * mp.in_dim;
          A[k * MLP_LD + rr] = rr < rows ? __ldg(xg + idx) : 0.f;
        }
      }
      if (l > 0) {
        const float* hg = hidden + mp.hid_off[l - 1] * n + row0;
        for (int idx = t; idx < hprev_w * MLP_ROWS; idx += MLP_THREADS) {
          const int j = idx >> 7, rr = idx & (MLP_ROWS - 1);
          A[(x_rows + j) * MLP_LD + rr] = rr < rows ? __ldg(hg + (size_t)j * n + rr) : 0.f;
        }
      }
      __syncthreads();
      // ---- dW[j][k] += sum_r dZ[j][r] A[k][r]   (4x4 patches), db[j] += sum_r dZ[j][r]
      {
        float* dw = dWs + mp.w_s[l];
        const int pj = (out + 3) >> 2, pk = (in + 3) >> 2;
        for (int p = t; p < pj * pk; p += MLP_THREADS) {
          const int j0 = (p / pk) * 4, k0 = (p % pk) * 4;
          float acc[4][4] = {};
          const float* zr[4];
          const float* ar[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            zr[q] = dZ + min(j0 + q, out - 1) * MLP_LD;
            ar[q] = A + min(k0 + q, in - 1) * MLP_LD;
          }
#pragma unroll 4
          for (int rr = 0; rr < MLP_ROWS; ++rr) {
            float zv[4], av[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) zv[q] = zr[q][rr], av[q] = ar[q][rr];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(zv[a], av[b], acc[a][b]);
          }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (j0 + a < out && k0 + b < in) dw[(j0 + a) * inp + k0 + b] += acc[a][b];
        }
        if (mp.db[l] != nullptr) {
          float* db = dWs + mp.b_s[l];
          for (int j = t; j < out; j += MLP_THREADS) {
            float s = 0.f;
            for (int rr = 0; rr < MLP_ROWS; ++rr) s += dZ[j * MLP_LD + rr];
            db[j] += s;
          }
        }
      }
      // ---- dA[k][r] = sum_j dZ[j][r] Wn[j][k]
      const bool need_dA = (l > 0) || (dx != nullptr);
      if (need_dA) {
        layer_rows(dZ, out, nullptr, 0, Wn + mp.w_s[l], inp, nullptr, inp, r, half,
                   [&](int k, float v) { dA[k * MLP_LD + r] = v; });
      }
      __syncthreads();
      // ---- next dZ = dA[hidden part] * act'(hidden_{l-1}); input part accumulates into dXs / is dX
      if (l > 0) {
        if (mp.skip[l] && dx != nullptr)
          for (int idx = t; idx < mp.in_dim * MLP_ROWS; idx += MLP_THREADS) {
            const int k = idx >> 7, rr = idx & (MLP_ROWS - 1);
            dXs[k * MLP_LD + rr] += dA[k * MLP_LD + rr];
          }
        for (int idx = t; idx < hprev_w * MLP_ROWS; idx += MLP_THREADS) {
          const int j = idx >> 7, rr = idx & (MLP_ROWS - 1);
          dZ[j * MLP_LD + rr] =
              dA[(x_rows + j) * MLP_LD + rr] * act_grad_from_out(mp.hidden_act, A[(x_rows + j) * MLP_LD + rr]);
        }
        __syncthreads();
      }
    }
    if (dx != nullptr) {
      const int in = mp.in_dim;
      float* dxg = dx + row0 * in;
      for (int idx = t; idx < rows * in; idx += MLP_THREADS) {
        const int rr = idx / in, k = idx - rr * in;
        float v = dA[k * MLP_LD + rr];
        if (mp.any_skip) v += dXs[k * MLP_LD + rr];
        dxg[idx] = v;
      }
    }
  }
  __syncthreads();
  // ---- flush: one RED per entry per CTA
  for (int l = 0; l < L; ++l) {
    const int in = mp.in[l], out = mp.out[l], inp = mp.inp[l];
    const float* dw = dWs + mp.w_s[l];
    float* gw = mp.dw[l];
    if (gw != nullptr)
      for (int idx = t; idx < out * in; idx += MLP_THREADS) {
        const int j = idx / in, k = idx - j * in;
        atomicAdd(gw + idx, dw[j * inp + k]);
      }
    if (mp.db[l] != nullptr) {
      const float* db = dWs + mp.b_s[l];
      for (int j = t; j < out; j += MLP_THREADS) atomicAdd(mp.db[l] + j, db[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
static int build_params(const B2nMlp* m, const B2nMlpGrad* g, MlpParams& mp, bool transposed) {
  if (m->n_layers < 1 || m->n_layers > B2N_MAX_MLP_LAYERS || m->in_dim < 1) return -1;
  memset(&mp, 0, sizeof(mp));
  mp.n_layers = m->n_layers, mp.in_dim = m->in_dim, mp.hidden_act = m->hidden_act, mp.out_act = m->out_act;
  int prev = m->in_dim, off = 0;
  long long hid = 0;
  mp.kmax_in = m->in_dim, mp.wmax_out = 8;
  for (int l = 0; l < m->n_layers; ++l) {
    const int out = m->out_dims[l];
    if (out < 1) return -1;
    if (m->skip[l] && l == 0) return -1;
    const int in = m->skip[l] ? prev + m->in_dim : prev;
    mp.in[l] = in, mp.out[l] = out, mp.skip[l] = m->skip[l] ? 1 : 0;
    mp.outp[l] = (out + 7) & ~7, mp.inp[l] = (in + 7) & ~7;
    if (m->w[l] == nullptr) return -1;
    mp.w[l] = m->w[l], mp.b[l] = m->b[l];
    mp.dw[l] = g ? g->dw[l] : nullptr, mp.db[l] = g ? g->db[l] : nullptr;
    mp.w_s[l] = off;
    off += transposed ? in * mp.outp[l] : out * mp.inp[l];
    off = (off + 3) & ~3;
    mp.b_s[l] = off;
    off += mp.outp[l];
    mp.hid_off[l] = hid;
    hid += out;
    if (mp.skip[l]) mp.any_skip = 1;
    if (in > mp.kmax_in) mp.kmax_in = in;
    if (mp.outp[l] > mp.wmax_out) mp.wmax_out = mp.outp[l];
    prev = out;
  }
  mp.kmax_in = (mp.kmax_in + 7) & ~7;
  mp.w_total = off;
  return 0;
}

static size_t fwd_smem(const MlpParams& mp) {
  return sizeof(float) * ((size_t)mp.w_total + (size_t)(mp.in_dim + 2 * mp.wmax_out) * MLP_LD);
}
static size_t bwd_smem(const MlpParams& mp) {
  return sizeof(float) * (2 * (size_t)mp.w_total +
                          (size_t)(mp.wmax_out + 2 * mp.kmax_in + (mp.any_skip ? mp.in_dim : 0)) * MLP_LD);
}

static int smem_limit() {
  static int lim = 0;
  if (!lim) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || lim <= 0) lim = 232448;
  }
  return lim;
}

extern "C" int b2n_mlp_fwd(const B2nMlp* mlp_host, const float* x, int64_t n, float* y, float* hidden, void* stream) {
  if (n == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(mlp_host && x && y, "null pointer");
  MlpParams mp;
  B2N_REQUIRE(build_params(mlp_host, nullptr, mp, true) == 0, "bad mlp description");
  const size_t smem = fwd_smem(mp);
  B2N_UNSUPPORTED(smem > (size_t)smem_limit(), "network too wide for the fused kernel (shared memory)");
  if (n == 0) return B2N_OK;
  cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t tiles = div_up(n, MLP_ROWS);
  const int per_sm = (int)max((size_t)1, (size_t)(smem_limit() - 1024) / (smem + 1024));
  const int grid = (int)min(tiles, (int64_t)b2n_sm_count() * min(per_sm, 4));
  mlp_fwd_kernel<<<grid, MLP_THREADS, smem, (cudaStream_t)stream>>>(mp, x, n, y, hidden);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_mlp_bwd(const B2nMlp* mlp_host, const B2nMlpGrad* grad_host, const float* x, const float* y,
                           const float* hidden, const float* dy, int64_t n, float* dx, void* stream) {
  if (n == 0) return B2N_OK;  // empty batch: nothing to validate or launch
  B2N_REQUIRE(mlp_host && grad_host && x && y && dy, "null pointer");
  B2N_REQUIRE(mlp_host->n_layers == 1 || hidden != nullptr, "hidden activations required");
  MlpParams mp;
  B2N_REQUIRE(build_params(mlp_host, grad_host, mp, false) == 0, "bad mlp description");
  const size_t smem = bwd_smem(mp);
  B2N_UNSUPPORTED(smem > (size_t)smem_limit(), "network too wide for the fused kernel (shared memory)");
  if (n == 0) return B2N_OK;
  cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t tiles = div_up(n, MLP_ROWS);
  const int per_sm = (int)max((size_t)1, (size_t)(smem_limit() - 1024) / (smem + 1024));
  const int grid = (int)min(tiles, (int64_t)b2n_sm_count() * min(per_sm, 2));
  mlp_bwd_kernel<<<grid, MLP_THREADS, smem, (cudaStream_t)stream>>>(mp, x, y, hidden, dy, n, dx);
  B2N_LAUNCH_CHECK();
}
