// Self-test of the tcgen05 building blocks (descriptors, operand majors, TMEM read-back, 3xTF32 split).
// One CTA, one 128-row tile.  Exposed through the C-ABI as b2n_tc_selftest so that tests/ can pin the hardware
// semantics the tensor-core MLP relies on against plain fp32 matmuls.
//
//   mode 0: D[128][N]   = A[128][K]   * B[N][K]^T     A K-major,  B K-major     (forward layer)
//   mode 1: D[128][N]   = A[128][K]   * W[K][N]       A K-major,  B MN-major    (dX = dZ W; W stored as in mode 0: rows=K)
//   mode 2: D[Ma][N]    = A[128][Ma]^T * B[128][N]    A MN-major, B MN-major    (dW = dZ^T X; reduction over the 128 rows)
// three_pass != 0 runs hi*hi + lo*hi + hi*lo (3xTF32); out receives the raw TMEM image [128 lanes][N columns].
#include "common.cuh"
#include "tc_common.cuh"

#define TST_MAXC 64  // max columns of any operand buffer

__global__ void __launch_bounds__(128) tc_selftest_kernel(int mode, int three_pass, const float* __restrict__ Ag,
                                                          int a_rows, int a_cols, const float* __restrict__ Bg, int b_rows,
                                                          int b_cols, int M, int N, int K, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  const int flags = three_pass >> 1;  // diagnostics: 1 = B buffer holds (word index & 1023), 2 = (word index >> 10),
  three_pass &= 1;                    //              4 = swap LBO/SBO of MN-major operands
  // four operand buffers (A hi/lo, B hi/lo), each 128 rows x TST_MAXC cols, column-group stride 2048 B
  const uint32_t CS = 2048, BUF = (TST_MAXC / 4) * CS;
  uint8_t* Ah = smem; uint8_t* Al = smem + BUF; uint8_t* Bh = smem + 2 * BUF; uint8_t* Bl = smem + 3 * BUF;
  for (int i = t; i < 4 * (int)BUF / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  for (int i = t; i < a_rows * a_cols; i += 128) {
    const int r = i / a_cols, c = i % a_cols;
    float hi, lo;
    tc::split_tf32(Ag[i], hi, lo);
    if (!three_pass) hi = Ag[i], lo = 0.f;
    *reinterpret_cast<float*>(Ah + tc::canon_off(r, c, CS)) = hi;
    *reinterpret_cast<float*>(Al + tc::canon_off(r, c, CS)) = lo;
  }
  for (int i = t; i < b_rows * b_cols; i += 128) {
    const int r = i / b_cols, c = i % b_cols;
    float hi, lo;
    tc::split_tf32(Bg[i], hi, lo);
    if (!three_pass) hi = Bg[i], lo = 0.f;
    *reinterpret_cast<float*>(Bh + tc::canon_off(r, c, CS)) = hi;
    *reinterpret_cast<float*>(Bl + tc::canon_off(r, c, CS)) = lo;
  }
  if (flags & 3) {
    for (int i = t; i < (int)BUF / 4; i += 128) {
      reinterpret_cast<float*>(Bh)[i] = (flags & 1) ? (float)(i & 1023) : (float)(i >> 10);
      reinterpret_cast<float*>(Bl)[i] = 0.f;
    }
  }
  if (t == 0) tc::mbar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc<64>(&tmem_slot);
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    const bool a_mn = (mode == 2), b_mn = (mode != 0);
    const uint32_t idesc = tc::make_idesc_tf32(M, N, a_mn, b_mn);
    const int ksteps = K / 8;
    uint32_t acc = 0;
    for (int pass = 0; pass < (three_pass ? 3 : 1); ++pass) {
      const uint8_t* A = (pass == 1) ? Al : Ah;
      const uint8_t* B = (pass == 2) ? Bl : Bh;
      for (int j = 0; j < ksteps; ++j) {
        const uint32_t mn_lbo = (flags & 4) ? CS : 128, mn_sbo = (flags & 4) ? 128 : CS;
        const uint64_t ad = a_mn ? tc::make_desc(tc::smem_u32(A) + j * 128, mn_lbo, mn_sbo)
                                 : tc::make_desc(tc::smem_u32(A) + j * 2 * CS, CS, 128);
        const uint64_t bd = b_mn ? tc::make_desc(tc::smem_u32(B) + j * 128, mn_lbo, mn_sbo)
                                 : tc::make_desc(tc::smem_u32(B) + j * 2 * CS, CS, 128);
        tc::mma_tf32(tmem, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    tc::commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < N) out[t * N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<64>(tmem);
}

extern "C" int b2n_tc_selftest(int32_t mode, int32_t three_pass, const float* a, int32_t a_rows, int32_t a_cols,
                               const float* b, int32_t b_rows, int32_t b_cols, int32_t m, int32_t n, int32_t k,
                               float* out128xn, void* stream) {
  B2N_REQUIRE(a && b && out128xn, "null pointer");
  B2N_REQUIRE(mode >= 0 && mode <= 2, "mode");
  B2N_REQUIRE(a_rows <= 128 && b_rows <= 128 && a_cols <= TST_MAXC && b_cols <= TST_MAXC, "operand too large");
  B2N_REQUIRE((m == 64 || m == 128) && n % 8 == 0 && n >= 8 && n <= 64 && k % 8 == 0 && k >= 8, "bad MMA shape");
  const size_t smem = 4 * (TST_MAXC / 4) * 2048;
  cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, three_pass, a, a_rows, a_cols, b, b_rows, b_cols, m, n, k,
                                                             out128xn);
  B2N_LAUNCH_CHECK();
}

// Timing probe: cycles from the first tcgen05.mma issue to completion for `n_mma` M=128 x N x K=8 tf32 MMAs split
// over `n_issuers` threads (lane 0 of warps 0..n_issuers-1), each with its own accumulator and mbarrier
// (development aid for the MLP kernels' pipeline model).
__global__ void __launch_bounds__(128) tc_timing_kernel(int n_mma, int N, int reps, int n_issuers, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 2 * 16 * 2048 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (t == 0)
    for (int i = 0; i < 4; ++i) tc::mbar_init(&bar[i], 1);
  if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  uint32_t phase = 0;
  long long issue = 0, total = 0;
  const int per = n_mma / n_issuers;
  for (int rep = 0; rep < reps; ++rep) {
    __syncthreads();
    const long long t0 = clock64();
    long long t1 = t0;
    if ((t & 31) == 0 && warp < n_issuers) {
      const uint32_t idesc = tc::make_idesc_tf32(128, N, false, false);
      const uint32_t a = tc::smem_u32(smem), b = a + 16 * 2048;
      uint64_t ad = tc::make_desc(a, 2048, 128), bd = tc::make_desc(b, 1024, 128);
#pragma unroll 2
      for (int j = 0; j < per; ++j) {
        tc::mma_tf32(tmem + warp * 64, ad, bd, idesc, j > 0);
        ad += ((j & 7) == 7) ? (uint64_t)0 - 7 * ((2 * 2048) >> 4) : (uint64_t)((2 * 2048) >> 4);
        bd += ((j & 7) == 7) ? (uint64_t)0 - 7 * ((2 * 1024) >> 4) : (uint64_t)((2 * 1024) >> 4);
      }
      tc::commit(&bar[warp]);
      t1 = clock64();
    }
    for (int i = 0; i < n_issuers; ++i) tc::mbar_wait(&bar[i], phase);
    phase ^= 1;
    tc::fence_after_sync();
    const long long t2 = clock64();
    if (rep > 0) issue += t1 - t0, total += t2 - t0;
  }
  if (t == 0) out[0] = issue / (reps - 1), out[1] = total / (reps - 1);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

extern "C" int b2n_tc_timing(int32_t n_mma, int32_t n, int32_t reps, long long* out2, void* stream) {
  // n_mma's upper byte selects the number of issuing threads (0 -> 1)
  const int issuers = (n_mma >> 24) ? (n_mma >> 24) : 1;
  n_mma &= 0xFFFFFF;
  B2N_REQUIRE(out2 && n_mma >= 1 && reps >= 2 && n % 16 == 0 && n >= 16 && n <= 64 && issuers <= 4, "bad arguments");
  const size_t smem = 2 * 16 * 2048;
  cudaFuncSetAttribute(tc_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_timing_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(n_mma, n, reps, issuers, out2);
  B2N_LAUNCH_CHECK();
}
