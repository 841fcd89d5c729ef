// K3 on tensor cores — fused tiny MLP forward/backward with tcgen05.mma (kind::tf32), accumulators in TMEM.
//
// Replaces the same reference rows as mlp.cu (nerfstudio/field_components/mlp.py:110-114,160-184) for the networks
// of the nerfacto hot path (every width <= 64: base 32->64->16, head 63->64->64->3, proposal 10->16->1), keeping
// fp32-level accuracy: every product is formed as hi*hi + lo*hi + hi*lo with hi = top 19 bits of the fp32 value
// (exactly a tf32) and lo = a - hi ("3xTF32"), accumulated in fp32 in TMEM — measured 1e-6 relative against fp64,
// so the 1e-4 parity bar holds.  The tensor pipe is nowhere near its limit on K,N <= 64; what the kernel buys is
// that the 2.1 ms/step the SIMT kernels spend in FFMA disappears into operand staging.
//
// Structure (persistent CTAs, one per SM; a tile = 128 rows = 128 TMEM lanes):
//   forward  : default = the warp-specialised two-slot kernel (mlp_tc_fwd2_kernel, see its header); the serial kernel
//              (mlp_tc_fwd_kernel: x row -> hi/lo -> smem (K-major canonical, float4 stores) -> per layer: thread 0
//              issues the MMAs, tcgen05.commit -> mbarrier; everyone tcgen05.ld's their row, bias + activation, saves
//              the hidden row (row-major) for backward, writes the next layer's operand) is kept as the reference
//              variant.  Weights: hi and lo halves stacked along N -> 2 MMA streams per layer.
//   backward : 512 threads; threads t, t+128, t+256, t+384 own row t&127 and 16 of its columns (warps w, w+4, w+8,
//              w+12 share TMEM quarter w&3).  Per layer
//                         dA = dZ W      A = dZ tile (K-major), B = W^T copy (K-major, hi/lo stacked)  -> TMEM -> regs
//                         dW += dZ^T A   operands are the transposed tiles (feature x point), written by scalar
//                                        conflict-free stores; hi and lo of dZ^T are STACKED along M (rows 0-63 /
//                                        64-127) so one M=128 MMA yields both partial products; a row of ones
//                                        appended to A^T makes the same GEMM produce db.  dW/db stay resident in
//                                        TMEM across all tiles of the CTA and are flushed once with REDs.
// MN-major tf32 operands would need the SWIZZLE_128B_BASE32B layout (probed: the interleaved layout returns zeros),
// hence the explicit transposed copies.
#include <cuda.h>  // CUtensorMap and the cuTensorMapEncodeTiled prototype only: the entry point is fetched at run time
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

#define TC_MAXL 4
#define TC_ROWS 128
#define TC_CS_A 2048u                 // [128][64] K-major tile: column-group stride
#define TC_TILE_BYTES (16u * TC_CS_A)  // 64 cols
#define TC_CS_TZ (128u * 16u + 16u)    // stacked dZ^T [128 feature rows][64 points], padded against bank conflicts
#define TC_CS_TA (80u * 16u + 16u)     // A^T [<=80 rows][64 points]
#define TC_TZ_BYTES (16u * TC_CS_TZ)
#define TC_TA_BYTES (16u * TC_CS_TA)

struct TcParams {
  int n_layers, in_dim, in_pad, hidden_act, out_act;
  int K[TC_MAXL], N[TC_MAXL];          // padded input / output widths (K % 16 == 0, N % 16 == 0, <= 64)
  int kr[TC_MAXL], nr[TC_MAXL];        // real widths
  const float* w[TC_MAXL];
  const float* b[TC_MAXL];
  float* dw[TC_MAXL];
  float* db[TC_MAXL];
  uint32_t w_off[TC_MAXL];             // byte offset of the layer's hi weight tile in the weight region (lo follows)
  uint32_t w_bytes[TC_MAXL];           // bytes of one (hi or lo) weight tile
  uint32_t bias_off[TC_MAXL];          // float offset in the bias region
  long long hid_off[TC_MAXL];          // feature offset (sum of previous real widths) of the saved activations
  uint32_t w_total;                    // bytes of the whole weight region
};

// ---- TMA staging of the packed weight image -----------------------------------------------------------------------
// b2n_mlp_tc_pack writes, once per optimisation step, the exact shared-memory image the kernels want (hi/lo split,
// canonical core-matrix layout, hi and lo halves stacked, zero padding, biases) into a caller-provided workspace; every
// persistent CTA then pulls it with cp.async.bulk.tensor (a 2-D tensor map over [rows][64 floats], boxes of TC_WBOX_ROWS
// rows) completing on an mbarrier — instead of 148 CTAs each re-reading, re-splitting and scatter-storing the weights.
#define TC_WBOX_ROWS 16                       // rows of 256 B per TMA box (4 KB)
#define TC_WBOX_BYTES (TC_WBOX_ROWS * 256)

struct TcWeightsTma {
  CUtensorMap map;  // 64-byte aligned by its typedef
  int rows;         // rows of 256 B to copy (multiple of TC_WBOX_ROWS); 0 = stage with plain loads
};

__device__ __forceinline__ void tma_issue_weights(const TcWeightsTma& w, uint8_t* dst, uint64_t* bar) {
  const uint32_t b = tc::smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)w.rows * 256u) : "memory");
  for (int r0 = 0; r0 < w.rows; r0 += TC_WBOX_ROWS) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            tc::smem_u32(dst + (size_t)r0 * 256)),
        "l"(&w.map), "r"(0), "r"(r0), "r"(b)
        : "memory");
  }
}

// The activation id is warp-uniform: switch once per slice, not per element (the per-element form compiled to a
// branch ladder of ~12 instructions per value and made the epilogue the longest phase of the kernel).
__device__ __forceinline__ void bias_act_slice(int act, float (&v)[16], const float* __restrict__ bias16, int n_real) {
  float b[16];
#pragma unroll
  for (int c = 0; c < 16; c += 4) {
    const float4 q = *reinterpret_cast<const float4*>(bias16 + c);
    b[c] = q.x, b[c + 1] = q.y, b[c + 2] = q.z, b[c + 3] = q.w;
  }
  if (act == B2N_ACT_RELU) {
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = fmaxf(v[c] + b[c], 0.f);
  } else if (act == B2N_ACT_SIGMOID) {
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = (c < n_real) ? 1.f / (1.f + expf(-(v[c] + b[c]))) : 0.f;
  } else {
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] += b[c];
  }
}
// g[c] *= act'(y[c])
__device__ __forceinline__ void act_grad_slice(int act, float (&g)[16], const float (&y)[16]) {
  if (act == B2N_ACT_RELU) {
#pragma unroll
    for (int c = 0; c < 16; ++c) g[c] = y[c] > 0.f ? g[c] : 0.f;
  } else if (act == B2N_ACT_SIGMOID) {
#pragma unroll
    for (int c = 0; c < 16; ++c) g[c] *= y[c] * (1.f - y[c]);
  }
}

#ifdef B2N_TC_PROF
// clock64 phase profile (development builds only): per-thread register accumulators, flushed once at kernel end
__device__ unsigned long long tc_prof[16];
#define TCP_INIT                  \
  long long tcp_last = clock64(); \
  unsigned long long tcp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TCPT(i, tid)                                                  \
  do {                                                                \
    if (threadIdx.x == (tid) && blockIdx.x == 0) {                    \
      const long long now = clock64();                                \
      tcp_acc[(i) & 7] += (unsigned long long)(now - tcp_last);       \
      tcp_last = now;                                                 \
    }                                                                 \
  } while (0)
#define TCP(i) TCPT(i, 0)
#define TCP_FLUSH(tid)                                                \
  do {                                                                \
    if (threadIdx.x == (tid) && blockIdx.x == 0) {                    \
      _Pragma("unroll") for (int q_ = 0; q_ < 8; ++q_) if (tcp_acc[q_]) atomicAdd(&tc_prof[q_], tcp_acc[q_]); \
    }                                                                 \
  } while (0)
extern "C" int b2n_tc_prof_read(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, tc_prof, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(tc_prof, z, sizeof(z));
  }
  return 0;
}
#else
#define TCP_INIT
#define TCP(i)
#define TCPT(i, tid)
#define TCP_FLUSH(tid)
#endif

// Read-only global loads as volatile asm: the compiler otherwise sinks a "prefetch" (a load issued a layer or a tile
// before its first use) across the barriers down to the use, exposing the full DRAM latency (ncu: 13 % of the
// forward's stall samples sat on the first use of the prefetched input).
__device__ __forceinline__ float4 ldg4_early(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg1_early(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

static int g_tc_bwd_issuer = 1;  // backward: 1 = asynchronous MMA issue (mlp_tc_bwd2_kernel) for nets wider than 16, 2 = always, 0 = never
#define TC_THREADS 512  // 16 warps: warps w, w+4, w+8, w+12 share TMEM quarter w&3 and split the tile's 64 columns
#define TC_HALF 16     // columns per thread

// write this thread's half row (32 columns starting at c0; zero beyond width_pad) as hi/lo into a K-major tile pair
__device__ __forceinline__ void store_half_hilo(uint8_t* hi, uint8_t* lo, int r, int c0, const float (&v)[TC_HALF],
                                                int width_pad) {
#pragma unroll
  for (int c = 0; c < TC_HALF; c += 4) {
    if (c0 + c < width_pad) {
      float4 h, l;
      tc::split_tf32(v[c], h.x, l.x), tc::split_tf32(v[c + 1], h.y, l.y);
      tc::split_tf32(v[c + 2], h.z, l.z), tc::split_tf32(v[c + 3], h.w, l.w);
      const uint32_t off = tc::canon_off(r, c0 + c, TC_CS_A);
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    }
  }
}

// this thread's half of its TMEM row: columns [col + c0, col + c0 + 32) ∩ [col, col + width); warp-uniform control flow
__device__ __forceinline__ void load_half(uint32_t tmem, int quarter, int col, int c0, int width, float (&v)[TC_HALF]) {
#pragma unroll
  for (int c = 0; c < TC_HALF; c += 16) {
    if (c0 + c < width) {
      float t[16];
      tc::ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(col + c0 + c), t);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[c + i] = t[i];
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[c + i] = 0.f;
    }
  }
}

// this thread's half of a global row (row-major, `real` valid columns), zero padded
__device__ __forceinline__ void load_global_half(const float* __restrict__ src, bool live, bool vec, int c0, int real,
                                                 int pad, float (&v)[TC_HALF]) {
#pragma unroll
  for (int c = 0; c < TC_HALF; c += 4) {
    if (c0 + c < pad) {
      if (live && vec && c0 + c + 4 <= real) {
        const float4 q = ldg4_early(src + c0 + c);
        v[c] = q.x, v[c + 1] = q.y, v[c + 2] = q.z, v[c + 3] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[c + i] = (live && c0 + c + i < real) ? ldg1_early(src + c0 + c + i) : 0.f;
      }
    } else {
      v[c] = v[c + 1] = v[c + 2] = v[c + 3] = 0.f;
    }
  }
}


// ---- warp-cooperative row access ---------------------------------------------------------------------------------
// A warp owns rows row0..row0+31 of a row-major matrix and, per row, the 16 floats (64 B) starting at column c0.
// If lane L touched its own row, every access instruction would hit 32 different 128-byte lines (32 LSU cycles per
// instruction: measured 2-3k cycles per layer and tile).  Instead instruction k lets quad i = L/4 touch row 4i+k with
// lane j = L%4 on 16-byte chunk j — 8 rows x 64 contiguous bytes per instruction, whole sectors — and a 4x4
// transpose inside each quad (two shuffle stages) converts between "access order" and "row-owner order".
__device__ __forceinline__ float4 shfl_xor4(const float4 v, int m) {
  return make_float4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                     __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
// lane j of a quad holds q[k]; afterwards lane j's q[k] is what lane k held in q[j]
__device__ __forceinline__ void quad_transpose(float4 (&q)[4]) {
  const int lane = threadIdx.x & 31;
  {
    const bool odd = (lane & 1) != 0;
    const float4 r0 = shfl_xor4(odd ? q[0] : q[1], 1), r1 = shfl_xor4(odd ? q[2] : q[3], 1);
    if (odd) q[0] = r0, q[2] = r1;
    else q[1] = r0, q[3] = r1;
  }
  {
    const bool hi = (lane & 2) != 0;
    const float4 r0 = shfl_xor4(hi ? q[0] : q[2], 2), r1 = shfl_xor4(hi ? q[1] : q[3], 2);
    if (hi) q[0] = r0, q[1] = r1;
    else q[2] = r0, q[3] = r1;
  }
}
// v <- columns [c0, c0+16) of row (row0 + lane); rows >= n and columns >= real read as zero.  `stride` % 4 == 0 and a
// 16-byte aligned base are required (so a chunk that starts below `real` lies inside the row's storage).  Split in two
// so that a prefetch can stay in flight: `issue` only launches the loads (raw access-order chunks, 16 registers),
// `finish` — which needs the data — zero-fills the tail and transposes into row-owner order at the point of use.
__device__ __forceinline__ void load_rows_quad_issue(const float* __restrict__ base, int64_t stride, int64_t row0, int64_t n,
                                                     int c0, int real, float (&raw)[TC_HALF]) {
  const int lane = threadIdx.x & 31, col = c0 + 4 * (lane & 3);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t row = row0 + 4 * (lane >> 2) + k;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n && col < real) q = ldg4_early(base + row * stride + col);
    raw[4 * k] = q.x, raw[4 * k + 1] = q.y, raw[4 * k + 2] = q.z, raw[4 * k + 3] = q.w;
  }
}
__device__ __forceinline__ void load_rows_quad_finish(int c0, int real, const float (&raw)[TC_HALF], float (&v)[TC_HALF]) {
  const int lane = threadIdx.x & 31, col = c0 + 4 * (lane & 3);
  float4 q[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    q[k] = make_float4(raw[4 * k], raw[4 * k + 1], raw[4 * k + 2], raw[4 * k + 3]);
    if (col + 1 >= real) q[k].y = 0.f;
    if (col + 2 >= real) q[k].z = 0.f;
    if (col + 3 >= real) q[k].w = 0.f;
  }
  quad_transpose(q);
#pragma unroll
  for (int k = 0; k < 4; ++k) v[4 * k] = q[k].x, v[4 * k + 1] = q[k].y, v[4 * k + 2] = q[k].z, v[4 * k + 3] = q[k].w;
}
__device__ __forceinline__ void load_rows_quad(const float* __restrict__ base, int64_t stride, int64_t row0, int64_t n,
                                               int c0, int real, float (&v)[TC_HALF]) {
  float raw[TC_HALF];
  load_rows_quad_issue(base, stride, row0, n, c0, real, raw);
  load_rows_quad_finish(c0, real, raw, v);
}
// columns [c0, c0+16) ∩ [0, real) of row (row0 + lane) <- v   (rows >= n are skipped)
__device__ __forceinline__ void store_rows_quad(float* __restrict__ base, int64_t stride, int64_t row0, int64_t n, int c0,
                                                int real, const float (&v)[TC_HALF]) {
  const int lane = threadIdx.x & 31, col = c0 + 4 * (lane & 3);
  float4 q[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  quad_transpose(q);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t row = row0 + 4 * (lane >> 2) + k;
    if (row < n && col < real) {
      float* dst = base + row * stride + col;
      if (col + 4 <= real) {
        *reinterpret_cast<float4*>(dst) = q[k];
      } else {
        dst[0] = q[k].x;
        if (col + 1 < real) dst[1] = q[k].y;
        if (col + 2 < real) dst[2] = q[k].z;
      }
    }
  }
}
__device__ __forceinline__ bool quad_ok(const void* base, int64_t stride) {
  return ((stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_fwd_kernel(const __grid_constant__ TcParams p,
                                                                  const float* __restrict__ x, int64_t x_stride,
                                                                  int64_t n, float* __restrict__ y,
                                                                  float* __restrict__ hidden) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* Ah = smem;
  uint8_t* Al = smem + TC_TILE_BYTES;
  uint8_t* Wr = smem + 2 * TC_TILE_BYTES;                       // weight region
  float* bias = reinterpret_cast<float*>(Wr + p.w_total);        // [sum N]
  const int t = threadIdx.x, warp = t >> 5, quarter = warp & 3;
  const int r = t & (TC_ROWS - 1), c0 = (t >> 7) * TC_HALF;

  for (int l = 0; l < p.n_layers; ++l) {  // stage W hi/lo as K-major canonical [N rows][K cols]
    const int N = p.N[l], K = p.K[l], nr = p.nr[l], kr = p.kr[l];
    // hi and lo halves are STACKED along N (rows 0..N-1 = hi, N..2N-1 = lo): one N'=2N MMA stream then yields
    // [A_hi W_hi | A_hi W_lo] and a second N'=N stream adds A_lo W_hi onto the first half — 2 streams instead of 3
    uint8_t* wst = Wr + p.w_off[l];
    const uint32_t cs = (uint32_t)(2 * N) * 16u;
#pragma unroll 4
    for (int idx = t; idx < N * K; idx += TC_THREADS) {
      const int j = idx / K, k = idx - j * K;
      const float v = (j < nr && k < kr) ? __ldg(p.w[l] + (size_t)j * kr + k) : 0.f;
      float h, lo;
      tc::split_tf32(v, h, lo);
      *reinterpret_cast<float*>(wst + tc::canon_off(j, k, cs)) = h;
      *reinterpret_cast<float*>(wst + tc::canon_off(N + j, k, cs)) = lo;
    }
    for (int j = t; j < N; j += TC_THREADS) bias[p.bias_off[l] + j] = (j < nr && p.b[l]) ? __ldg(p.b[l] + j) : 0.f;
  }
  if (t == 0) tc::mbar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc<128>(&tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  uint32_t phase = 0;
  const bool xvec = ((x_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);

  const int64_t n_tiles = (n + TC_ROWS - 1) / TC_ROWS;
  float vnext[TC_HALF];  // next tile's input slice, loaded a whole tile ahead so its DRAM latency is never exposed
  {
    const int64_t row0 = (int64_t)blockIdx.x * TC_ROWS + r;
    load_global_half(x + row0 * x_stride, blockIdx.x < n_tiles && row0 < n, xvec, c0, p.in_dim, p.in_pad, vnext);
  }
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * TC_ROWS + r;
    const bool live = row < n;
    float v[TC_HALF];
#pragma unroll
    for (int c = 0; c < TC_HALF; ++c) v[c] = vnext[c];
    store_half_hilo(Ah, Al, r, c0, v, p.in_pad);
    {
      const int64_t nrow = (tile + gridDim.x) * TC_ROWS + r;
      load_global_half(x + nrow * x_stride, tile + gridDim.x < n_tiles && nrow < n, xvec, c0, p.in_dim, p.in_pad, vnext);
    }
    for (int l = 0; l < p.n_layers; ++l) {
      const int N = p.N[l], K = p.K[l];
      tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      if (t == 0) {
        const uint32_t cs = (uint32_t)(2 * N) * 16u;
        const uint32_t wst = tc::smem_u32(Wr + p.w_off[l]);
        // descriptors differ between k-steps only in the start-address field (bits 0-13, units of 16 B)
        const uint64_t da = (uint64_t)((2 * TC_CS_A) >> 4), db = (uint64_t)((2 * cs) >> 4);
        uint32_t acc = 0;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {  // pass 0: A_hi x [W_hi; W_lo] (N' = 2N), pass 1: A_lo x W_hi (N' = N)
          const uint32_t idesc = tc::make_idesc_tf32(TC_ROWS, pass ? N : 2 * N, false, false);
          uint64_t ad = tc::make_desc(tc::smem_u32(pass ? Al : Ah), TC_CS_A, 128), bd = tc::make_desc(wst, cs, 128);
#pragma unroll 2
          for (int s = 0; s < K / 8; ++s) {
            tc::mma_tf32(tmem, ad, bd, idesc, acc);
            ad += da, bd += db;
            acc = 1;
          }
        }
        tc::commit(&bar);
      }
      __syncwarp();  // lane 0 issued the MMAs alone: reconverge before the warp-aligned tcgen05 instructions
      tc::mbar_wait(&bar, phase);
      __syncwarp();
      phase ^= 1;
      tc::fence_after_sync();
      {
        float v2[TC_HALF];
        load_half(tmem, quarter, 0, c0, N, v);
        load_half(tmem, quarter, N, c0, N, v2);  // the A_hi W_lo partial product
#pragma unroll
        for (int c = 0; c < TC_HALF; ++c) v[c] += v2[c];
      }
      const bool last = (l == p.n_layers - 1);
      const int act = last ? p.out_act : p.hidden_act;
      const int nr = p.nr[l];
      // padded columns: zero weights and zero bias give act(0) = 0 for ReLU / identity (sigmoid only ends a network)
      if (c0 < N) bias_act_slice(act, v, bias + p.bias_off[l] + c0, nr - c0);  // warp-uniform
      if (!last) {
        if (hidden != nullptr && live) {
          float* h = hidden + p.hid_off[l] * n + row * nr;
#pragma unroll
          for (int c = 0; c < TC_HALF; c += 4)
            if (c0 + c < nr) *reinterpret_cast<float4*>(h + c0 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
        store_half_hilo(Ah, Al, r, c0, v, N);  // this layer's MMAs have completed: the operand tile is free
      } else if (live) {
        float* yr = y + row * nr;
#pragma unroll
        for (int c = 0; c < TC_HALF; ++c)
          if (c0 + c < nr) yr[c0 + c] = v[c];
      }
    }
    tc::fence_before_sync();
    __syncthreads();  // all TMEM reads of this tile done before the next tile's first MMA overwrites D
    tc::fence_after_sync();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

// ------------------------------------------------------------------------------------------------
// forward, warp-specialised variant: two tiles in flight per CTA (selected with b2n_tune("tc_fwd_slots", 2))
//
// ncu stall samples of the serial kernel above: 22 % waiting for the MMAs, 12 % at the CTA barrier, 8 % on the TMEM
// load — the tensor pipe and the CUDA cores take turns.  Here two SLOTS of 4 warps each own one 128-row tile (thread =
// one full row, TMEM quarter = warp & 3) with their own operand tiles and accumulator columns, and warp 8 only issues
// MMAs, alternating between the slots layer by layer:
//     worker:  write operand (hi/lo) -> fence.proxy.async -> arrive full[s]   ...   wait done[s] -> tcgen05.ld -> epilogue
//     issuer:  wait full[s] -> MMAs of that layer -> tcgen05.commit -> done[s]
// so slot 0's epilogue runs under slot 1's MMAs and vice versa.  (9 warps: a 17-warp version with 2 x 8 worker warps
// is capped at 96 registers by the 4-warp allocation granularity and spilled its prefetch registers.)
// ------------------------------------------------------------------------------------------------
#define TCF_SLOT_THREADS 128
#define TCF_WORKERS (2 * TCF_SLOT_THREADS)
#define TCF_THREADS (TCF_WORKERS + 32)

__global__ void __launch_bounds__(TCF_THREADS, 1) mlp_tc_fwd2_kernel(const __grid_constant__ TcParams p,
                                                                    const __grid_constant__ TcWeightsTma wt,
                                                                    const float* __restrict__ x, int64_t x_stride,
                                                                    int64_t n, float* __restrict__ y,
                                                                    float* __restrict__ hidden) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[2], bar_done[2], bar_w;
  __shared__ uint32_t tmem_slot;
  uint8_t* Wr = smem + 4 * TC_TILE_BYTES;                        // weight region behind the two slots' hi/lo tiles
  float* bias = reinterpret_cast<float*>(Wr + p.w_total);        // [sum N]
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  constexpr int ISSUER_WARP = TCF_WORKERS / 32;

  if (wt.rows > 0) {  // packed image (weights + biases) by TMA: one thread arms the barrier and issues the boxes
    if (t == 0) {
      tc::mbar_init(&bar_w, 1);
      tma_issue_weights(wt, Wr, &bar_w);
    }
  } else {
  for (int l = 0; l < p.n_layers; ++l) {  // stage W as K-major canonical [2N rows][K cols] (hi rows, then lo rows)
    const int N = p.N[l], K = p.K[l], nr = p.nr[l], kr = p.kr[l];
    uint8_t* wst = Wr + p.w_off[l];
    const uint32_t cs = (uint32_t)(2 * N) * 16u;
#pragma unroll 4
    for (int idx = t; idx < N * K; idx += TCF_THREADS) {
      const int j = idx / K, k = idx - j * K;
      const float v = (j < nr && k < kr) ? __ldg(p.w[l] + (size_t)j * kr + k) : 0.f;
      float h, lo;
      tc::split_tf32(v, h, lo);
      *reinterpret_cast<float*>(wst + tc::canon_off(j, k, cs)) = h;
      *reinterpret_cast<float*>(wst + tc::canon_off(N + j, k, cs)) = lo;
    }
    for (int j = t; j < N; j += TCF_THREADS) bias[p.bias_off[l] + j] = (j < nr && p.b[l]) ? __ldg(p.b[l] + j) : 0.f;
  }
  }
  if (t == 0) {
    tc::mbar_init(&bar_full[0], TCF_SLOT_THREADS), tc::mbar_init(&bar_full[1], TCF_SLOT_THREADS);
    tc::mbar_init(&bar_done[0], 1), tc::mbar_init(&bar_done[1], 1);
  }
  if (warp == ISSUER_WARP) tc::tmem_alloc<256>(&tmem_slot);
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (wt.rows > 0) tc::mbar_wait(&bar_w, 0);  // weight image has landed (TMA complete_tx)
  const uint32_t tmem = tmem_slot;
  const int64_t n_tiles = (n + TC_ROWS - 1) / TC_ROWS;
  const int L = p.n_layers;

  if (warp < ISSUER_WARP) {
    // ------------------------------------------------------------------ workers
    const int s = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    uint8_t* Ah = smem + (size_t)s * 2 * TC_TILE_BYTES;
    uint8_t* Al = Ah + TC_TILE_BYTES;
    const uint32_t tm = tmem + (uint32_t)(s * 128) + ((uint32_t)(quarter * 32) << 16);
    uint32_t ph_done = 0;
    const bool xvec = quad_ok(x, x_stride);
    const int in_chunks = p.in_pad >> 2;
    float4 xn[16];  // the next tile's input row, loaded a whole tile ahead so its DRAM latency is never exposed
    auto load_x = [&](int64_t tile) {
      const int64_t row = tile * TC_ROWS + r;
      const bool live = tile < n_tiles && row < n;
      const float* src = x + row * x_stride;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int c = 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < in_chunks && live && c < p.in_dim) {
          if (xvec) {  // a chunk that starts below in_dim lies inside the row's storage (stride % 4 == 0)
            v = ldg4_early(src + c);
          } else {
            v.x = ldg1_early(src + c);
            if (c + 1 < p.in_dim) v.y = ldg1_early(src + c + 1);
            if (c + 2 < p.in_dim) v.z = ldg1_early(src + c + 2);
            if (c + 3 < p.in_dim) v.w = ldg1_early(src + c + 3);
          }
        }
        xn[q] = v;
      }
    };
    // one 16-byte chunk (4 columns of this thread's row) as hi / lo into the slot's operand tiles
    auto put4 = [&](int c, const float4 v) {
      float4 h, lo;
      tc::split_tf32(v.x, h.x, lo.x), tc::split_tf32(v.y, h.y, lo.y);
      tc::split_tf32(v.z, h.z, lo.z), tc::split_tf32(v.w, h.w, lo.w);
      const uint32_t off = tc::canon_off(r, c, TC_CS_A);
      *reinterpret_cast<float4*>(Ah + off) = h;
      *reinterpret_cast<float4*>(Al + off) = lo;
    };
    load_x((int64_t)blockIdx.x + (int64_t)s * gridDim.x);
    for (int64_t k = s;; k += 2) {
      const int64_t tile = (int64_t)blockIdx.x + k * gridDim.x;
      if (tile >= n_tiles) break;
      const int64_t row = tile * TC_ROWS + r;
      const bool live = row < n;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (q < in_chunks) {
          float4 v = xn[q];
          const int c = 4 * q;  // zero the tail of a chunk that straddles in_dim
          if (c + 1 >= p.in_dim) v.y = 0.f;
          if (c + 2 >= p.in_dim) v.z = 0.f;
          if (c + 3 >= p.in_dim) v.w = 0.f;
          put4(c, v);
        }
      }
      load_x(tile + 2 * (int64_t)gridDim.x);
      tc::fence_smem_to_async();
      tc::fence_before_sync();  // orders this thread's TMEM reads of the previous tile before the issuer's next MMA
      tc::mbar_arrive(&bar_full[s]);
      for (int l = 0; l < L; ++l) {
        const int N = p.N[l], nr = p.nr[l];
        const bool last = (l == L - 1);
        const int act = last ? p.out_act : p.hidden_act;
        const float* bl = bias + p.bias_off[l];
        float* hrow = (!last && hidden != nullptr && live) ? hidden + p.hid_off[l] * n + row * nr : nullptr;
        tc::mbar_wait(&bar_done[s], ph_done);
        __syncwarp();
        ph_done ^= 1;
        tc::fence_after_sync();
#pragma unroll 1
        for (int cc = 0; cc < N; cc += 16) {
          float v[16], v2[16];
          tc::ld16(tm + (uint32_t)cc, v);
          tc::ld16(tm + (uint32_t)(N + cc), v2);  // the A_hi W_lo partial product
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] += v2[c];
          // padded columns: zero weights and zero bias give act(0) = 0 for ReLU / identity (sigmoid only ends a net)
          bias_act_slice(act, v, bl + cc, nr - cc);
          if (!last) {
            if (hrow != nullptr) {
              float* h = hrow + cc;
#pragma unroll
              for (int c = 0; c < 16; c += 4)
                if (cc + c < nr) *reinterpret_cast<float4*>(h + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
#pragma unroll
            for (int c = 0; c < 16; c += 4) put4(cc + c, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
          } else if (live) {
            float* yr = y + row * nr + cc;
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (cc + c < nr) yr[c] = v[c];
          }
        }
        if (!last) {
          tc::fence_smem_to_async();
          tc::fence_before_sync();
          tc::mbar_arrive(&bar_full[s]);
        }
      }
    }
  } else if (t == TCF_WORKERS) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    uint32_t ph_full0 = 0u, ph_full1 = 0u;
    for (int64_t k0 = 0; (int64_t)blockIdx.x + k0 * gridDim.x < n_tiles; k0 += 2) {
      for (int l = 0; l < L; ++l) {
        const int N = p.N[l], K = p.K[l];
        const uint32_t cs = (uint32_t)(2 * N) * 16u;
        const uint32_t wst = tc::smem_u32(Wr + p.w_off[l]);
        // descriptors differ between k-steps only in the start-address field (bits 0-13, units of 16 B)
        const uint64_t da = (uint64_t)((2 * TC_CS_A) >> 4), db = (uint64_t)((2 * cs) >> 4);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if ((int64_t)blockIdx.x + (k0 + s) * gridDim.x >= n_tiles) continue;
          if (s == 0) {
            tc::mbar_wait(&bar_full[0], ph_full0);
            ph_full0 ^= 1;
          } else {
            tc::mbar_wait(&bar_full[1], ph_full1);
            ph_full1 ^= 1;
          }
          tc::fence_after_sync();
          const uint32_t d = tmem + (uint32_t)(s * 128);
          const uint32_t ah = tc::smem_u32(smem + (size_t)s * 2 * TC_TILE_BYTES), al = ah + TC_TILE_BYTES;
          uint32_t acc = 0;
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {  // pass 0: A_hi x [W_hi; W_lo] (N' = 2N), pass 1: A_lo x W_hi (N' = N)
            const uint32_t idesc = tc::make_idesc_tf32(TC_ROWS, pass ? N : 2 * N, false, false);
            uint64_t ad = tc::make_desc(pass ? al : ah, TC_CS_A, 128), bd = tc::make_desc(wst, cs, 128);
#pragma unroll 2
            for (int ks = 0; ks < K / 8; ++ks) {
              tc::mma_tf32(d, ad, bd, idesc, acc);
              ad += da, bd += db;
              acc = 1;
            }
          }
          tc::commit(&bar_done[s]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == ISSUER_WARP) {
    __syncwarp();
    tc::tmem_dealloc<256>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_bwd_kernel(const __grid_constant__ TcParams p,
                                                                  const __grid_constant__ TcWeightsTma wt,
                                                                  const float* __restrict__ x, int64_t x_stride,
                                                                  const float* __restrict__ y,
                                                                  const float* __restrict__ hidden,
                                                                  const float* __restrict__ dy, int64_t n,
                                                                  float* __restrict__ dx, int64_t dx_stride) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* Zh = smem;                                   // dZ tile, K-major [128][64]
  uint8_t* Zl = Zh + TC_TILE_BYTES;
  uint8_t* TZ = Zl + TC_TILE_BYTES;                     // stacked dZ^T: rows 0-63 hi, 64-127 lo; 64 point columns
  uint8_t* TAh = TZ + TC_TZ_BYTES;                      // A^T hi (+ ones row)
  uint8_t* TAl = TAh + TC_TA_BYTES;                     // A^T lo
  uint8_t* Wr = TAl + TC_TA_BYTES;                      // W^T hi/lo per layer: K-major [K rows][N cols]
  const int t = threadIdx.x, warp = t >> 5, quarter = warp & 3;
  const int r = t & (TC_ROWS - 1), c0 = (t >> 7) * TC_HALF;
  const int L = p.n_layers;
  __shared__ uint64_t bar_w;

  if (wt.rows > 0) {  // packed W^T image by TMA
    if (t == 0) {
      tc::mbar_init(&bar_w, 1);
      tma_issue_weights(wt, Wr, &bar_w);
    }
  } else {
  for (int l = 0; l < L; ++l) {
    const int N = p.N[l], K = p.K[l], nr = p.nr[l], kr = p.kr[l];
    // transposed (row = input index k, col = output index j), hi rows 0..K-1 and lo rows K..2K-1 stacked along the
    // MMA's N: dA = [dZ_hi W_hi | dZ_hi W_lo] (N' = 2K) + dZ_lo W_hi (N' = K) in two streams
    uint8_t* wst = Wr + p.w_off[l];
    const uint32_t cs = (uint32_t)(2 * K) * 16u;
#pragma unroll 4
    for (int idx = t; idx < N * K; idx += TC_THREADS) {
      const int j = idx / K, k = idx - j * K;
      const float v = (j < nr && k < kr) ? __ldg(p.w[l] + (size_t)j * kr + k) : 0.f;
      float h, lo;
      tc::split_tf32(v, h, lo);
      *reinterpret_cast<float*>(wst + tc::canon_off(k, j, cs)) = h;
      *reinterpret_cast<float*>(wst + tc::canon_off(K + k, j, cs)) = lo;
    }
  }
  }
  if (t == 0) tc::mbar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (wt.rows > 0) tc::mbar_wait(&bar_w, 0);  // W^T image has landed (TMA complete_tx)
  const uint32_t tmem = tmem_slot;
  uint32_t phase = 0;
  uint32_t dw_started = 0;  // bit l: the layer's dW accumulator has been written once (thread 0 only)
  const int DW_COL0 = 128, DW_COLS = 80;  // D (dA, two partial products) occupies columns 0..127
  const bool xvec = quad_ok(x, x_stride);
  const bool dxvec = dx != nullptr && quad_ok(dx, dx_stride);
  const int wrow = 32 * quarter;  // first tile row of this warp
  // input rows of layer l (x or the saved hidden activations) for this thread's (row, column slice): `issue_a` starts
  // the loads one layer ahead (raw registers), `finish_a` turns them into row-owner values where they are consumed
  auto a_is_quad = [&](int l) -> bool {  // warp-uniform
    if (l == 0) return xvec;
    return (p.kr[l] & 3) == 0 && quad_ok(hidden + p.hid_off[l - 1] * n, p.kr[l]);
  };
  auto issue_a = [&](int l, int64_t tile, int64_t row, bool live, float (&raw)[TC_HALF]) {
    const int kr = p.kr[l];
    if (c0 >= p.K[l]) {  // warp-uniform
#pragma unroll
      for (int c = 0; c < TC_HALF; ++c) raw[c] = 0.f;
      return;
    }
    const float* src = (l == 0) ? x : hidden + p.hid_off[l - 1] * n;
    const int64_t stride = (l == 0) ? x_stride : (int64_t)kr;
    if (a_is_quad(l)) load_rows_quad_issue(src, stride, tile * TC_ROWS + wrow, n, c0, kr, raw);
    else load_global_half(src + row * stride, live, false, c0, kr, p.K[l], raw);
  };
  auto finish_a = [&](int l, const float (&raw)[TC_HALF], float (&dst)[TC_HALF]) {
    if (c0 < p.K[l] && a_is_quad(l)) {
      load_rows_quad_finish(c0, p.kr[l], raw, dst);
    } else {
#pragma unroll
      for (int c = 0; c < TC_HALF; ++c) dst[c] = raw[c];
    }
  };
  const int64_t n_tiles = (n + TC_ROWS - 1) / TC_ROWS;
  TCP_INIT;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * TC_ROWS + r;
    const bool live = row < n;
    {  // pull the NEXT tile's rows (dy, y, every layer's input slice) into L2: no registers, and the register
       // prefetches one layer ahead then see L2 latency instead of DRAM latency
      const int64_t nrow = row + (int64_t)gridDim.x * TC_ROWS;
      if (nrow < n) {
        if (c0 == 0) {
          tc::prefetch_l2(dy + nrow * p.nr[L - 1]);
          tc::prefetch_l2(y + nrow * p.nr[L - 1]);
        }
        if (c0 < p.K[0]) tc::prefetch_l2(x + nrow * x_stride + c0);
        for (int l = 1; l < L; ++l)
          if (c0 < p.K[l]) tc::prefetch_l2(hidden + p.hid_off[l - 1] * n + nrow * p.kr[l] + c0);
      }
    }
    float dz[TC_HALF], a[TC_HALF];
    {  // dZ of the last layer = dy * act'(y)   (this thread's column half)
      const int nr = p.nr[L - 1];
      if (c0 < nr && (nr & 3) == 0 && quad_ok(dy, nr) && quad_ok(y, nr)) {  // warp-uniform
        float yv[TC_HALF];
        load_rows_quad(dy, nr, tile * TC_ROWS + wrow, n, c0, nr, dz);
        load_rows_quad(y, nr, tile * TC_ROWS + wrow, n, c0, nr, yv);
        act_grad_slice(p.out_act, dz, yv);
      } else {
        float yv[TC_HALF];
#pragma unroll
        for (int c = 0; c < TC_HALF; ++c) {
          const bool on = c0 + c < nr && live;
          dz[c] = on ? __ldg(dy + row * nr + c0 + c) : 0.f;
          yv[c] = on ? __ldg(y + row * nr + c0 + c) : 0.f;
        }
        act_grad_slice(p.out_act, dz, yv);
      }
    }
    float a_next[TC_HALF];  // the input row slice of the NEXT layer to be processed (one layer of lookahead, raw)
    issue_a(L - 1, tile, row, live, a_next);
    TCP(0);
    for (int l = L - 1; l >= 0; --l) {
      const int N = p.N[l], K = p.K[l], kr = p.kr[l];
      finish_a(l, a_next, a);
      if (l > 0) issue_a(l - 1, tile, row, live, a_next);
      const bool need_da = (l > 0) || (dx != nullptr);
      if (need_da) store_half_hilo(Zh, Zl, r, c0, dz, N);
      TCP(1);
      // ---- dW_l += dZ^T A over the two 64-point halves of the tile
      for (int ph = 0; ph < 2; ++ph) {
        if ((r >> 6) == ph) {
          const int pc = r & 63;  // point column inside the half
#pragma unroll
          for (int j = 0; j < TC_HALF; ++j) {
            float h = 0.f, lo = 0.f;
            if (c0 + j < N) tc::split_tf32(dz[j], h, lo);
            *reinterpret_cast<float*>(TZ + tc::canon_off(c0 + j, pc, TC_CS_TZ)) = h;
            *reinterpret_cast<float*>(TZ + tc::canon_off(64 + c0 + j, pc, TC_CS_TZ)) = lo;
          }
#pragma unroll
          for (int k = 0; k < TC_HALF; ++k) {
            if (c0 + k < K) {
              float h, lo;
              tc::split_tf32(a[k], h, lo);
              *reinterpret_cast<float*>(TAh + tc::canon_off(c0 + k, pc, TC_CS_TA)) = h;
              *reinterpret_cast<float*>(TAl + tc::canon_off(c0 + k, pc, TC_CS_TA)) = lo;
            }
          }
          if (c0 == 0) {  // ones row (-> bias gradient) and zero padding rows K..K+15
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              *reinterpret_cast<float*>(TAh + tc::canon_off(K + k, pc, TC_CS_TA)) = (k == 0 && live) ? 1.f : 0.f;
              *reinterpret_cast<float*>(TAl + tc::canon_off(K + k, pc, TC_CS_TA)) = 0.f;
            }
          }
        }
        TCP(2);
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        TCP(3);
        if (t == 0) {
          const uint32_t idesc = tc::make_idesc_tf32(TC_ROWS, K + 16, false, false);
          const uint32_t tz = tc::smem_u32(TZ), tah = tc::smem_u32(TAh), tal = tc::smem_u32(TAl);
          const uint32_t dcol = tmem + (uint32_t)(DW_COL0 + DW_COLS * l);
          uint32_t acc = (dw_started >> l) & 1u;
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            uint64_t ad = tc::make_desc(tz, TC_CS_TZ, 128), bd = tc::make_desc(pass ? tal : tah, TC_CS_TA, 128);
#pragma unroll
            for (int s = 0; s < 8; ++s) {
              tc::mma_tf32(dcol, ad, bd, idesc, acc);
              ad += (uint64_t)((2 * TC_CS_TZ) >> 4), bd += (uint64_t)((2 * TC_CS_TA) >> 4);
              acc = 1;
            }
          }
          dw_started |= 1u << l;
          if (ph == 1 && need_da) {  // dA = dZ W : issued right behind, one commit covers both
            const uint32_t cs = (uint32_t)(2 * K) * 16u;
            const uint32_t wst = tc::smem_u32(Wr + p.w_off[l]);
            uint32_t acc2 = 0;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
              const uint32_t idx = tc::make_idesc_tf32(TC_ROWS, pass ? K : 2 * K, false, false);
              uint64_t ad = tc::make_desc(tc::smem_u32(pass ? Zl : Zh), TC_CS_A, 128), bd = tc::make_desc(wst, cs, 128);
#pragma unroll 2
              for (int s = 0; s < N / 8; ++s) {
                tc::mma_tf32(tmem, ad, bd, idx, acc2);
                ad += (uint64_t)((2 * TC_CS_A) >> 4), bd += (uint64_t)((2 * cs) >> 4);
                acc2 = 1;
              }
            }
          }
          tc::commit(&bar);
        }
        TCP(4);
        __syncwarp();  // lane 0 issued the MMAs alone: reconverge before the warp-aligned tcgen05 instructions
        tc::mbar_wait(&bar, phase);
        __syncwarp();
        phase ^= 1;
        tc::fence_after_sync();
        TCP(5);
      }
      // ---- dA half row -> next dZ (or dx)
      if (need_da) {
        float da[TC_HALF];
        {
          float da2[TC_HALF];
          load_half(tmem, quarter, 0, c0, K, da);
          load_half(tmem, quarter, K, c0, K, da2);  // the dZ_hi W_lo partial product
#pragma unroll
          for (int c = 0; c < TC_HALF; ++c) da[c] += da2[c];
        }
        if (l > 0) {
#pragma unroll
          for (int c = 0; c < TC_HALF; ++c) dz[c] = da[c];  // padded columns: zero weight rows give dA = 0
          act_grad_slice(p.hidden_act, dz, a);
        } else if (dxvec) {
          if (c0 < p.in_dim) store_rows_quad(dx, dx_stride, tile * TC_ROWS + wrow, n, c0, p.in_dim, da);  // warp-uniform
        } else if (live) {
          float* dr = dx + row * dx_stride;
#pragma unroll
          for (int c = 0; c < TC_HALF; ++c)
            if (c0 + c < p.in_dim) dr[c0 + c] = da[c];
        }
        TCP(6);
        tc::fence_before_sync();
        __syncthreads();  // D (columns 0..63) fully read before the next layer's dA MMA overwrites it
        tc::fence_after_sync();
        TCP(7);
      }
    }
  }
  TCP_FLUSH(0);
  // ---- flush dW / db: lanes 0-63 hold dZ_hi^T [A_hi + A_lo], lanes 64-127 the dZ_lo^T part
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  // The lo-part lanes hand their values to the hi-part lanes through shared memory (the dZ tile region is free now),
  // and rows go out as 128-bit REDs where the gradient row is 16-byte aligned: 8x fewer L2 atomics than one atomicAdd
  // per lane and element (148 CTAs all add into the same few KB).
  float* scratch = reinterpret_cast<float*>(Zh);  // [4 column groups][64 rows][20 floats (16 + pad)]
  for (int l = 0; l < L; ++l) {
    if (p.dw[l] == nullptr && p.db[l] == nullptr) continue;
    const int K = p.K[l], kr = p.kr[l], nr = p.nr[l];
    const int j = r & 63;
    const bool started = blockIdx.x < n_tiles;
    const bool vec = p.dw[l] != nullptr && (kr & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw[l]) & 15) == 0;
    for (int base = 0; base < K + 16; base += 64) {  // uniform trip count: the loop body synchronises the CTA
      const int cc = base + c0;
      const bool active = cc < K + 16;  // warp-uniform
      float v[16];
      if (active) tc::ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(DW_COL0 + DW_COLS * l + cc), v);
      float* sc = scratch + ((c0 >> 4) * 64 + j) * 20;
      if (active && r >= 64) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(sc + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      __syncthreads();
      if (active && r < 64 && started && j < nr) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 o = *reinterpret_cast<const float4*>(sc + i);
          const float g0 = v[i] + o.x, g1 = v[i + 1] + o.y, g2 = v[i + 2] + o.z, g3 = v[i + 3] + o.w;
          const int k = cc + i;
          if (vec && k + 4 <= kr) {
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dw[l] + (size_t)j * kr + k),
                         "f"(g0), "f"(g1), "f"(g2), "f"(g3)
                         : "memory");
          } else {
            const float g[4] = {g0, g1, g2, g3};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (k + e < kr && p.dw[l]) atomicAdd(p.dw[l] + (size_t)j * kr + k + e, g[e]);
              else if (k + e == K && p.db[l]) atomicAdd(p.db[l] + j, g[e]);
            }
          }
        }
      }
      __syncthreads();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}


// mbarrier wait that cannot hang the GPU: a protocol error traps (the launch fails with an error) instead of spinning
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 26)) asm volatile("trap;");
  }
}

// ------------------------------------------------------------------------------------------------
// backward, dedicated MMA-issuer warp (default; b2n_tune("tc_bwd_issuer", 0) selects the kernel above)
//
// Phase profile of the kernel above (profiles/r02_tc_phase.log): 25 % of a tile is thread 0's MMA issue loop — the tensor
// pipe's time — during which the other 15 warps sit at a CTA barrier, and the CUDA-core phases (input rows, Z store,
// transposed staging, tcgen05.ld + activation gradient) never overlap it.  Here the 16 warps
// never meet at a CTA barrier inside the tile loop: they signal 'operands staged' on an mbarrier (512 arrivals)
// and wait for the tcgen05.commit of exactly the MMA group whose result or whose operand buffer they need next:
//     worker, per layer:  [Z store] -> wait dW1(prev) -> [T half 0] -> arrive full -> wait dA -> tcgen05.ld dA
//                         -> wait dW0 -> [T half 1] -> arrive full -> dz = act'(a) * dA
//     issuer (lane 0 of warp 0, right after its own arrive):
//                         wait full -> dA MMAs, commit(barA); dW(half 0) MMAs, commit(barW)
//                         wait full -> dW(half 1) MMAs, commit(barW)
// so the dW MMAs run under the workers' tcgen05.ld / activation / next layer's input rows and Z store.  Arithmetic and
// operand layouts are those of the kernel above (dx, hidden-layer gradients bit-identical; dW up to atomics order).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_bwd2_kernel(const __grid_constant__ TcParams p,
                                                                  const __grid_constant__ TcWeightsTma wt,
                                                                  const float* __restrict__ x, int64_t x_stride,
                                                                  const float* __restrict__ y,
                                                                  const float* __restrict__ hidden,
                                                                  const float* __restrict__ dy, int64_t n,
                                                                  float* __restrict__ dx, int64_t dx_stride) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full, bar_a, bar_w;
  __shared__ uint32_t tmem_slot;
  uint8_t* Zh = smem;                                   // dZ tile, K-major [128][64]
  uint8_t* Zl = Zh + TC_TILE_BYTES;
  uint8_t* TZ = Zl + TC_TILE_BYTES;                     // stacked dZ^T: rows 0-63 hi, 64-127 lo; 64 point columns
  uint8_t* TAh = TZ + TC_TZ_BYTES;                      // A^T hi (+ ones row)
  uint8_t* TAl = TAh + TC_TA_BYTES;                     // A^T lo
  uint8_t* Wr = TAl + TC_TA_BYTES;                      // W^T hi/lo per layer: K-major [K rows][N cols]
  const int t = threadIdx.x, warp = t >> 5, quarter = warp & 3;
  const int r = t & (TC_ROWS - 1), c0 = ((t >> 7) & 3) * TC_HALF;
  const int L = p.n_layers;
  __shared__ uint64_t bar_wt;

  if (wt.rows > 0) {  // packed W^T image by TMA
    if (t == 0) {
      tc::mbar_init(&bar_wt, 1);
      tma_issue_weights(wt, Wr, &bar_wt);
    }
  } else {
  for (int l = 0; l < L; ++l) {
    const int N = p.N[l], K = p.K[l], nr = p.nr[l], kr = p.kr[l];
    // transposed (row = input index k, col = output index j), hi rows 0..K-1 and lo rows K..2K-1 stacked along the
    // MMA's N: dA = [dZ_hi W_hi | dZ_hi W_lo] (N' = 2K) + dZ_lo W_hi (N' = K) in two streams
    uint8_t* wst = Wr + p.w_off[l];
    const uint32_t cs = (uint32_t)(2 * K) * 16u;
#pragma unroll 4
    for (int idx = t; idx < N * K; idx += TC_THREADS) {
      const int j = idx / K, k = idx - j * K;
      const float v = (j < nr && k < kr) ? __ldg(p.w[l] + (size_t)j * kr + k) : 0.f;
      float h, lo;
      tc::split_tf32(v, h, lo);
      *reinterpret_cast<float*>(wst + tc::canon_off(k, j, cs)) = h;
      *reinterpret_cast<float*>(wst + tc::canon_off(K + k, j, cs)) = lo;
    }
  }
  }
  if (t == 0) {
    tc::mbar_init(&bar_full, TC_THREADS), tc::mbar_init(&bar_a, 1), tc::mbar_init(&bar_w, 1);
  }
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (wt.rows > 0) mbar_wait_guard(&bar_wt, 0);  // W^T image has landed (TMA complete_tx)
  const uint32_t tmem = tmem_slot;
  uint32_t dw_started = 0;  // bit l: the layer's dW accumulator has been written once (issuer only)
  const int DW_COL0 = 128, DW_COLS = 80;  // D (dA, two partial products) occupies columns 0..127
  const bool xvec = quad_ok(x, x_stride);
  const bool dxvec = dx != nullptr && quad_ok(dx, dx_stride);
  const int wrow = 32 * quarter;  // first tile row of this warp
  // input rows of layer l (x or the saved hidden activations) for this thread's (row, column slice): `issue_a` starts
  // the loads one layer ahead (raw registers), `finish_a` turns them into row-owner values where they are consumed
  auto a_is_quad = [&](int l) -> bool {  // warp-uniform
    if (l == 0) return xvec;
    return (p.kr[l] & 3) == 0 && quad_ok(hidden + p.hid_off[l - 1] * n, p.kr[l]);
  };
  auto issue_a = [&](int l, int64_t tile, int64_t row, bool live, float (&raw)[TC_HALF]) {
    const int kr = p.kr[l];
    if (c0 >= p.K[l]) {  // warp-uniform
#pragma unroll
      for (int c = 0; c < TC_HALF; ++c) raw[c] = 0.f;
      return;
    }
    const float* src = (l == 0) ? x : hidden + p.hid_off[l - 1] * n;
    const int64_t stride = (l == 0) ? x_stride : (int64_t)kr;
    if (a_is_quad(l)) load_rows_quad_issue(src, stride, tile * TC_ROWS + wrow, n, c0, kr, raw);
    else load_global_half(src + row * stride, live, false, c0, kr, p.K[l], raw);
  };
  auto finish_a = [&](int l, const float (&raw)[TC_HALF], float (&dst)[TC_HALF]) {
    if (c0 < p.K[l] && a_is_quad(l)) {
      load_rows_quad_finish(c0, p.kr[l], raw, dst);
    } else {
#pragma unroll
      for (int c = 0; c < TC_HALF; ++c) dst[c] = raw[c];
    }
  };
  const int64_t n_tiles = (n + TC_ROWS - 1) / TC_ROWS;
  uint32_t ph_a = 0, ph_w = 0, ph_f = 0;  // parities of bar_a / bar_w (workers) and bar_full (issuer)
  bool w_pending = false;                 // a dW group has been committed on bar_w and not yet waited for (uniform)
  // MMA groups of (layer l, tile half ph), issued by lane 0 of warp 0 once all 512 threads have staged their operands
  auto issue_group = [&](int l, int ph, bool need_da) {
    const int N = p.N[l], K = p.K[l];
    mbar_wait_guard(&bar_full, ph_f);
    ph_f ^= 1;
    tc::fence_after_sync();
    if ((t & 31) == 0) {
      if (ph == 0 && need_da) {  // dA = dZ W first: every thread's critical path waits for it
        const uint32_t cs = (uint32_t)(2 * K) * 16u;
        const uint32_t wst = tc::smem_u32(Wr + p.w_off[l]);
        uint32_t acc2 = 0;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const uint32_t idx = tc::make_idesc_tf32(TC_ROWS, pass ? K : 2 * K, false, false);
          uint64_t ad = tc::make_desc(tc::smem_u32(pass ? Zl : Zh), TC_CS_A, 128), bd = tc::make_desc(wst, cs, 128);
#pragma unroll 2
          for (int s = 0; s < N / 8; ++s) {
            tc::mma_tf32(tmem, ad, bd, idx, acc2);
            ad += (uint64_t)((2 * TC_CS_A) >> 4), bd += (uint64_t)((2 * cs) >> 4);
            acc2 = 1;
          }
        }
        tc::commit(&bar_a);
      }
      const uint32_t idesc_w = tc::make_idesc_tf32(TC_ROWS, K + 16, false, false);
      const uint32_t dcol = tmem + (uint32_t)(DW_COL0 + DW_COLS * l);
      uint32_t acc = (dw_started >> l) & 1u;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        uint64_t ad = tc::make_desc(tc::smem_u32(TZ), TC_CS_TZ, 128);
        uint64_t bd = tc::make_desc(tc::smem_u32(pass ? TAl : TAh), TC_CS_TA, 128);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          tc::mma_tf32(dcol, ad, bd, idesc_w, acc);
          ad += (uint64_t)((2 * TC_CS_TZ) >> 4), bd += (uint64_t)((2 * TC_CS_TA) >> 4);
          acc = 1;
        }
      }
      dw_started |= 1u << l;
      tc::commit(&bar_w);
    }
    __syncwarp();
  };
  {
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row = tile * TC_ROWS + r;
      const bool live = row < n;
      {  // pull the NEXT tile's rows into L2 (no registers)
        const int64_t nrow = row + (int64_t)gridDim.x * TC_ROWS;
        if (nrow < n) {
          if (c0 == 0) {
            tc::prefetch_l2(dy + nrow * p.nr[L - 1]);
            tc::prefetch_l2(y + nrow * p.nr[L - 1]);
          }
          if (c0 < p.K[0]) tc::prefetch_l2(x + nrow * x_stride + c0);
          for (int l = 1; l < L; ++l)
            if (c0 < p.K[l]) tc::prefetch_l2(hidden + p.hid_off[l - 1] * n + nrow * p.kr[l] + c0);
        }
      }
      float dz[TC_HALF], a[TC_HALF];
      {  // dZ of the last layer = dy * act'(y)   (this thread's column slice)
        const int nr = p.nr[L - 1];
        if (c0 < nr && (nr & 3) == 0 && quad_ok(dy, nr) && quad_ok(y, nr)) {  // warp-uniform
          float yv[TC_HALF];
          load_rows_quad(dy, nr, tile * TC_ROWS + wrow, n, c0, nr, dz);
          load_rows_quad(y, nr, tile * TC_ROWS + wrow, n, c0, nr, yv);
          act_grad_slice(p.out_act, dz, yv);
        } else {
          float yv[TC_HALF];
#pragma unroll
          for (int c = 0; c < TC_HALF; ++c) {
            const bool on = c0 + c < nr && live;
            dz[c] = on ? __ldg(dy + row * nr + c0 + c) : 0.f;
            yv[c] = on ? __ldg(y + row * nr + c0 + c) : 0.f;
          }
          act_grad_slice(p.out_act, dz, yv);
        }
      }
      float a_next[TC_HALF];
      issue_a(L - 1, tile, row, live, a_next);
      for (int l = L - 1; l >= 0; --l) {
        const int N = p.N[l], K = p.K[l];
        finish_a(l, a_next, a);
        if (l > 0) issue_a(l - 1, tile, row, live, a_next);
        const bool need_da = (l > 0) || (dx != nullptr);
        // Z is free: the dA MMAs of the previous layer (the only readers) completed before this thread's tcgen05.ld
        if (need_da) store_half_hilo(Zh, Zl, r, c0, dz, N);
        if (w_pending) {  // the transposed tiles are free once the previous dW(half 1) group has completed
          mbar_wait_guard(&bar_w, ph_w);
          ph_w ^= 1, w_pending = false;
        }
        auto stage_T = [&](int ph) {  // rows of tile half `ph`: dZ^T (hi/lo stacked), A^T (hi, lo), ones / padding rows
          if ((r >> 6) == ph) {
            const int pc = r & 63;  // point column inside the half
#pragma unroll
            for (int j = 0; j < TC_HALF; ++j) {
              float h = 0.f, lo = 0.f;
              if (c0 + j < N) tc::split_tf32(dz[j], h, lo);
              *reinterpret_cast<float*>(TZ + tc::canon_off(c0 + j, pc, TC_CS_TZ)) = h;
              *reinterpret_cast<float*>(TZ + tc::canon_off(64 + c0 + j, pc, TC_CS_TZ)) = lo;
            }
#pragma unroll
            for (int k = 0; k < TC_HALF; ++k) {
              if (c0 + k < K) {
                float h, lo;
                tc::split_tf32(a[k], h, lo);
                *reinterpret_cast<float*>(TAh + tc::canon_off(c0 + k, pc, TC_CS_TA)) = h;
                *reinterpret_cast<float*>(TAl + tc::canon_off(c0 + k, pc, TC_CS_TA)) = lo;
              }
            }
            if (c0 == 0) {  // ones row (-> bias gradient) and zero padding rows K..K+15
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                *reinterpret_cast<float*>(TAh + tc::canon_off(K + k, pc, TC_CS_TA)) = (k == 0 && live) ? 1.f : 0.f;
                *reinterpret_cast<float*>(TAl + tc::canon_off(K + k, pc, TC_CS_TA)) = 0.f;
              }
            }
          }
          tc::fence_smem_to_async();
          tc::fence_before_sync();  // orders this thread's earlier TMEM reads before the issuer's next MMAs
          tc::mbar_arrive(&bar_full);
          if (warp == 0) issue_group(l, ph, need_da);
        };
        stage_T(0);
        float da[TC_HALF];
        if (need_da) {  // dA of this layer is issued first: ready long before the dW(half 0) group finishes
          __syncwarp();
          mbar_wait_guard(&bar_a, ph_a);
          __syncwarp();
          ph_a ^= 1;
          tc::fence_after_sync();
          float da2[TC_HALF];
          load_half(tmem, quarter, 0, c0, K, da);
          load_half(tmem, quarter, K, c0, K, da2);  // the dZ_hi W_lo partial product
#pragma unroll
          for (int c = 0; c < TC_HALF; ++c) da[c] += da2[c];
        }
        mbar_wait_guard(&bar_w, ph_w);  // dW(half 0) done: the transposed tiles may be rewritten
        ph_w ^= 1;
        stage_T(1);  // still THIS layer's dz and a
        w_pending = true;
        if (need_da) {  // runs under the dW(half 1) MMAs
          if (l > 0) {
#pragma unroll
            for (int c = 0; c < TC_HALF; ++c) dz[c] = da[c];  // padded columns: zero weight rows give dA = 0
            act_grad_slice(p.hidden_act, dz, a);
          } else if (dxvec) {
            if (c0 < p.in_dim) store_rows_quad(dx, dx_stride, tile * TC_ROWS + wrow, n, c0, p.in_dim, da);  // warp-uniform
          } else if (live) {
            float* dr = dx + row * dx_stride;
#pragma unroll
            for (int c = 0; c < TC_HALF; ++c)
              if (c0 + c < p.in_dim) dr[c0 + c] = da[c];
          }
        }
      }
    }
    if (w_pending) {  // every MMA of this CTA has completed before the accumulators are read
      mbar_wait_guard(&bar_w, ph_w);
      ph_w ^= 1;
    }
    tc::fence_after_sync();
  }
  // ---- flush dW / db: lanes 0-63 hold dZ_hi^T [A_hi + A_lo], lanes 64-127 the dZ_lo^T part
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  // The lo-part lanes hand their values to the hi-part lanes through shared memory (the dZ tile region is free now),
  // and rows go out as 128-bit REDs where the gradient row is 16-byte aligned: 8x fewer L2 atomics than one atomicAdd
  // per lane and element (148 CTAs all add into the same few KB).
  float* scratch = reinterpret_cast<float*>(Zh);  // [4 column groups][64 rows][20 floats (16 + pad)]
  for (int l = 0; l < L; ++l) {
    if (p.dw[l] == nullptr && p.db[l] == nullptr) continue;
    const int K = p.K[l], kr = p.kr[l], nr = p.nr[l];
    const int j = r & 63;
    const bool started = blockIdx.x < n_tiles;
    const bool vec = p.dw[l] != nullptr && (kr & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw[l]) & 15) == 0;
    for (int base = 0; base < K + 16; base += 64) {  // uniform trip count: the loop body synchronises the CTA
      const int cc = base + c0;
      const bool active = cc < K + 16;  // warp-uniform
      float v[16];
      if (active) tc::ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(DW_COL0 + DW_COLS * l + cc), v);
      float* sc = scratch + ((c0 >> 4) * 64 + j) * 20;
      if (active && r >= 64) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(sc + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      __syncthreads();
      if (active && r < 64 && started && j < nr) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 o = *reinterpret_cast<const float4*>(sc + i);
          const float g0 = v[i] + o.x, g1 = v[i + 1] + o.y, g2 = v[i + 2] + o.z, g3 = v[i + 3] + o.w;
          const int k = cc + i;
          if (vec && k + 4 <= kr) {
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dw[l] + (size_t)j * kr + k),
                         "f"(g0), "f"(g1), "f"(g2), "f"(g3)
                         : "memory");
          } else {
            const float g[4] = {g0, g1, g2, g3};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (k + e < kr && p.dw[l]) atomicAdd(p.dw[l] + (size_t)j * kr + k + e, g[e]);
              else if (k + e == K && p.db[l]) atomicAdd(p.db[l] + j, g[e]);
            }
          }
        }
      }
      __syncthreads();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}


// ------------------------------------------------------------------------------------------------
static int tc_build(const B2nMlp* m, const B2nMlpGrad* g, TcParams& p, bool transposed) {
  memset(&p, 0, sizeof(p));
  if (m->n_layers < 1 || m->n_layers > TC_MAXL || m->in_dim < 1 || m->in_dim > 64) return -1;
  if (m->hidden_act != B2N_ACT_RELU && m->hidden_act != B2N_ACT_NONE) return -1;
  if (m->out_act != B2N_ACT_NONE && m->out_act != B2N_ACT_SIGMOID && m->out_act != B2N_ACT_RELU) return -1;
  p.n_layers = m->n_layers, p.in_dim = m->in_dim, p.in_pad = (m->in_dim + 15) & ~15;
  p.hidden_act = m->hidden_act, p.out_act = m->out_act;
  int prev = m->in_dim;
  uint32_t off = 0, boff = 0;
  long long hid = 0;
  for (int l = 0; l < m->n_layers; ++l) {
    const int out = m->out_dims[l];
    if (m->skip[l] || out < 1 || out > 64 || m->w[l] == nullptr) return -1;
    if (l < m->n_layers - 1 && (out & 3)) return -1;  // hidden rows are saved with float4 stores
    p.kr[l] = prev, p.nr[l] = out;
    p.K[l] = (prev + 15) & ~15, p.N[l] = (out + 15) & ~15;
    p.w[l] = m->w[l], p.b[l] = m->b[l];
    p.dw[l] = g ? g->dw[l] : nullptr, p.db[l] = g ? g->db[l] : nullptr;
    p.w_bytes[l] = (uint32_t)(p.K[l] / 4) * (uint32_t)((transposed ? p.K[l] : p.N[l]) * 16);
    if (transposed) p.w_bytes[l] = (uint32_t)(p.N[l] / 4) * (uint32_t)(p.K[l] * 16);
    p.w_off[l] = off;
    off += 2 * p.w_bytes[l];
    p.bias_off[l] = boff;
    boff += p.N[l];
    p.hid_off[l] = hid;
    hid += out;
    prev = out;
  }
  p.w_total = off;
  return 0;
}

// ---- weight image: layout, pack kernel, tensor map ------------------------------------------------------------------
static inline uint32_t pad_box(uint32_t bytes) { return (bytes + TC_WBOX_BYTES - 1) / TC_WBOX_BYTES * TC_WBOX_BYTES; }
static uint32_t fwd_image_bytes(const TcParams& pf) {
  uint32_t nb = 0;
  for (int l = 0; l < pf.n_layers; ++l) nb += (uint32_t)pf.N[l];
  return pad_box(pf.w_total + 4u * nb);
}
static uint32_t bwd_image_bytes(const TcParams& pb) { return pad_box(pb.w_total); }

__global__ void __launch_bounds__(256) mlp_tc_pack_kernel(const __grid_constant__ TcParams pf,
                                                          const __grid_constant__ TcParams pb, uint8_t* __restrict__ img_f,
                                                          uint32_t bytes_f, uint8_t* __restrict__ img_b, uint32_t bytes_b) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  uint32_t nb = 0;
  for (int l = 0; l < pf.n_layers; ++l) {
    const int N = pf.N[l], K = pf.K[l], nr = pf.nr[l], kr = pf.kr[l];
    uint8_t* wf = img_f + pf.w_off[l];
    uint8_t* wb = img_b + pb.w_off[l];
    const uint32_t csf = (uint32_t)(2 * N) * 16u, csb = (uint32_t)(2 * K) * 16u;
    for (int idx = t; idx < N * K; idx += nt) {
      const int j = idx / K, k = idx - j * K;
      const float v = (j < nr && k < kr) ? __ldg(pf.w[l] + (size_t)j * kr + k) : 0.f;
      float h, lo;
      tc::split_tf32(v, h, lo);
      *reinterpret_cast<float*>(wf + tc::canon_off(j, k, csf)) = h;      // forward: [2N rows][K cols], hi rows then lo rows
      *reinterpret_cast<float*>(wf + tc::canon_off(N + j, k, csf)) = lo;
      *reinterpret_cast<float*>(wb + tc::canon_off(k, j, csb)) = h;      // backward: W^T, [2K rows][N cols]
      *reinterpret_cast<float*>(wb + tc::canon_off(K + k, j, csb)) = lo;
    }
    float* bias = reinterpret_cast<float*>(img_f + pf.w_total);
    for (int j = t; j < N; j += nt) bias[pf.bias_off[l] + j] = (j < nr && pf.b[l]) ? __ldg(pf.b[l] + j) : 0.f;
    nb += (uint32_t)N;
  }
  for (uint32_t o = pf.w_total + 4u * nb + 4u * t; o < bytes_f; o += 4u * nt) *reinterpret_cast<float*>(img_f + o) = 0.f;
  for (uint32_t o = pb.w_total + 4u * t; o < bytes_b; o += 4u * nt) *reinterpret_cast<float*>(img_b + o) = 0.f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}
// tensor map over an image of `bytes` (multiple of TC_WBOX_BYTES): fp32 [bytes/256 rows][64], boxes of TC_WBOX_ROWS rows
static int make_weight_map(TcWeightsTma& w, const void* image, uint32_t bytes) {
  memset(&w, 0, sizeof(w));
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr || image == nullptr || (reinterpret_cast<uintptr_t>(image) & 127) != 0) return -1;
  const cuuint64_t dims[2] = {64, bytes / 256};
  const cuuint64_t strides[1] = {256};
  const cuuint32_t box[2] = {64, TC_WBOX_ROWS}, estr[2] = {1, 1};
  if (enc(&w.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(image), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return -1;
  w.rows = (int)(bytes / 256);
  return 0;
}

extern "C" int64_t b2n_mlp_tc_workspace_bytes(const B2nMlp* mlp_host) {
  TcParams pf, pb;
  if (!mlp_host || tc_build(mlp_host, nullptr, pf, false) != 0 || tc_build(mlp_host, nullptr, pb, true) != 0) return -1;
  return (int64_t)fwd_image_bytes(pf) + (int64_t)bwd_image_bytes(pb);
}

extern "C" int b2n_mlp_tc_pack(const B2nMlp* mlp_host, void* workspace, void* stream) {
  B2N_REQUIRE(mlp_host && workspace, "null pointer");
  B2N_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127) == 0, "workspace must be 128-byte aligned");
  TcParams pf, pb;
  B2N_UNSUPPORTED(tc_build(mlp_host, nullptr, pf, false) != 0 || tc_build(mlp_host, nullptr, pb, true) != 0,
                  "tensor-core MLP: needs <= 4 layers, widths <= 64 (hidden % 4 == 0), ReLU hidden, no skips");
  const uint32_t bf = fwd_image_bytes(pf), bb = bwd_image_bytes(pb);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  mlp_tc_pack_kernel<<<16, 256, 0, (cudaStream_t)stream>>>(pf, pb, ws, bf, ws + bf, bb);
  B2N_LAUNCH_CHECK();
}

static int g_tc_fwd_slots = 2;  // 2: warp-specialised two-slot kernel (default), 1: serial kernel
int b2n_tune_mlp_tc(const char* key, int value) {
  if (strcmp(key, "tc_bwd_issuer") == 0 && value >= 0 && value <= 2) {
    g_tc_bwd_issuer = value;
    return 1;
  }
  if (strcmp(key, "tc_fwd_slots") == 0 && (value == 1 || value == 2)) {
    g_tc_fwd_slots = value;
    return 1;
  }
  return 0;
}

static int tc_smem_limit() {
  static int lim = 0;
  if (!lim) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || lim <= 0) lim = 232448;
  }
  return lim;
}

extern "C" int b2n_mlp_tc_fwd_ws(const B2nMlp* mlp_host, const float* x, int64_t x_stride, int64_t n, float* y,
                                 float* hidden, const void* workspace, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(mlp_host && x && y, "null pointer");
  TcParams p;
  B2N_UNSUPPORTED(tc_build(mlp_host, nullptr, p, false) != 0,
                  "tensor-core MLP: needs <= 4 layers, widths <= 64 (hidden % 4 == 0), ReLU hidden, no skips");
  B2N_REQUIRE(x_stride >= mlp_host->in_dim, "x_stride smaller than in_dim");
  const int grid = (int)min(div_up(n, TC_ROWS), (int64_t)b2n_sm_count());
  const size_t smem2 = 4 * TC_TILE_BYTES + fwd_image_bytes(p) + 1024;
  if (g_tc_fwd_slots == 2 && smem2 <= (size_t)tc_smem_limit()) {  // two tiles in flight (needs room for 4 operand tiles)
    TcWeightsTma wt;
    if (workspace == nullptr || make_weight_map(wt, workspace, fwd_image_bytes(p)) != 0) wt.rows = 0;
    cudaFuncSetAttribute(mlp_tc_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    mlp_tc_fwd2_kernel<<<grid, TCF_THREADS, smem2, (cudaStream_t)stream>>>(p, wt, x, x_stride, n, y, hidden);
    B2N_LAUNCH_CHECK();
  }
  const size_t smem = 2 * TC_TILE_BYTES + p.w_total + 4 * 64 * TC_MAXL + 1024;
  B2N_UNSUPPORTED(smem > (size_t)tc_smem_limit(), "tensor-core MLP: shared memory");
  cudaFuncSetAttribute(mlp_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mlp_tc_fwd_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p, x, x_stride, n, y, hidden);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_mlp_tc_fwd(const B2nMlp* mlp_host, const float* x, int64_t x_stride, int64_t n, float* y,
                              float* hidden, void* stream) {
  return b2n_mlp_tc_fwd_ws(mlp_host, x, x_stride, n, y, hidden, nullptr, stream);
}

extern "C" int b2n_mlp_tc_bwd_ws(const B2nMlp* mlp_host, const B2nMlpGrad* grad_host, const float* x, int64_t x_stride,
                                 const float* y, const float* hidden, const float* dy, int64_t n, float* dx,
                                 int64_t dx_stride, const void* workspace, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(mlp_host && grad_host && x && y && dy, "null pointer");
  B2N_REQUIRE(mlp_host->n_layers == 1 || hidden != nullptr, "hidden activations required");
  TcParams p;
  B2N_UNSUPPORTED(tc_build(mlp_host, grad_host, p, true) != 0,
                  "tensor-core MLP: needs <= 4 layers, widths <= 64 (hidden % 4 == 0), ReLU hidden, no skips");
  B2N_REQUIRE(dx == nullptr || dx_stride >= mlp_host->in_dim, "dx_stride smaller than in_dim");
  const size_t smem = 2 * TC_TILE_BYTES + TC_TZ_BYTES + 2 * TC_TA_BYTES + bwd_image_bytes(p) + 1024;
  B2N_UNSUPPORTED(smem > (size_t)tc_smem_limit(), "tensor-core MLP: shared memory");
  TcWeightsTma wt;
  wt.rows = 0;
  if (workspace != nullptr) {
    TcParams pf;
    if (tc_build(mlp_host, nullptr, pf, false) != 0 ||
        make_weight_map(wt, static_cast<const uint8_t*>(workspace) + fwd_image_bytes(pf), bwd_image_bytes(p)) != 0)
      wt.rows = 0;
  }
  const int grid = (int)min(div_up(n, TC_ROWS), (int64_t)b2n_sm_count());
  int widest = 0;
  for (int l = 0; l < p.n_layers; ++l) widest = max(widest, max(p.N[l], p.K[l]));
  // measured (scripts/tc_bench.py): head 64-64-3 0.279 -> 0.259 ms, base 32-64-16 unchanged, a 10-16-1 net 5 % slower
  if (g_tc_bwd_issuer == 2 || (g_tc_bwd_issuer == 1 && widest > 16)) {
    cudaFuncSetAttribute(mlp_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mlp_tc_bwd2_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p, wt, x, x_stride, y, hidden, dy, n, dx, dx_stride);
    B2N_LAUNCH_CHECK();
  }
  cudaFuncSetAttribute(mlp_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mlp_tc_bwd_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p, wt, x, x_stride, y, hidden, dy, n, dx, dx_stride);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_mlp_tc_bwd(const B2nMlp* mlp_host, const B2nMlpGrad* grad_host, const float* x, int64_t x_stride,
                              const float* y, const float* hidden, const float* dy, int64_t n, float* dx,
                              int64_t dx_stride, void* stream) {
  return b2n_mlp_tc_bwd_ws(mlp_host, grad_host, x, x_stride, y, hidden, dy, n, dx, dx_stride, nullptr, stream);
}
