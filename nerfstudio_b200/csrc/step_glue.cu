// Small kernels that keep the captured training step free of framework launches: the stratified-sampling draws
// (counter-based Philox-4x32-10, state on the device so that every graph replay gets a fresh stream), the zeroing of the
// step's small accumulators, an in-place add and the total of the loss terms.
// Reference call sites: the stratified draws are torch.rand in model_components/ray_samplers.py:99-105 (spaced sampler)
// and :335-340 (PDF sampler) — any uniform [0,1) stream is a valid draw; the sum of the loss dict is
// engine/trainer.py:511 (functools.reduce(torch.add, loss_dict.values())).
#include "common.cuh"

namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0, c[1] = lo1, c[2] = n2, c[3] = lo0;
}

// Philox-4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0, k1) -> 4 x 32 random bits
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
}

// state: [0] seed, [1] draw counter (one per launch), [2] block tickets of the running launch
__global__ void __launch_bounds__(256) step_begin_kernel(float* __restrict__ u, int64_t n_u, unsigned long long* __restrict__ state,
                                                         float* __restrict__ z0, int64_t n0, float* __restrict__ z1, int64_t n1) {
  const unsigned long long seed = state[0];
  const unsigned long long draw = *reinterpret_cast<volatile unsigned long long*>(state + 1);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = tid; q * 4 < n_u; q += nth) {  // element 4q+j = word j of block (q, draw) under key seed
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)draw, (uint32_t)(draw >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n_u) u[q * 4 + j] = (float)(c[j] >> 8) * (1.0f / 16777216.0f);  // 24 bits: [0, 1) exactly
  }
  for (int64_t i = tid; i < n0; i += nth) z0[i] = 0.f;
  for (int64_t i = tid; i < n1; i += nth) z1[i] = 0.f;
  __syncthreads();  // every thread of the block has read `draw`
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(state + 2, 1ull);
    if (ticket == gridDim.x - 1) {  // last block: all blocks have read the counter
      state[1] = draw + 1;
      state[2] = 0;
    }
  }
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = add_rn(dst[i], src[i]);
}

__global__ void loss_total_kernel(const float* __restrict__ terms, int n_terms, const float* __restrict__ extra, float* __restrict__ out) {
  float s = terms[0];
  for (int i = 1; i < n_terms; ++i) s = add_rn(s, terms[i]);  // left to right like reduce(torch.add, ...)
  if (extra) s = add_rn(s, extra[0]);
  out[0] = s;
}

}  // namespace

extern "C" int b2n_step_begin(float* uniforms, int64_t n_uniforms, uint64_t* rng_state3, float* zero0, int64_t n_zero0,
                              float* zero1, int64_t n_zero1, void* stream) {
  B2N_REQUIRE(rng_state3 != nullptr, "rng state");
  B2N_REQUIRE((uniforms || n_uniforms == 0) && (zero0 || n_zero0 == 0) && (zero1 || n_zero1 == 0), "null pointer");
  const int64_t work = max(max(div_up(n_uniforms, 4), n_zero0), max(n_zero1, (int64_t)1));
  const int grid = (int)min(div_up(work, 256), (int64_t)(4 * b2n_sm_count()));
  step_begin_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(uniforms, n_uniforms, reinterpret_cast<unsigned long long*>(rng_state3),
                                                             zero0, n_zero0, zero1, n_zero1);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_add_inplace(float* dst, const float* src, int64_t n, void* stream) {
  if (n == 0) return B2N_OK;
  B2N_REQUIRE(dst && src, "null pointer");
  add_inplace_kernel<<<(unsigned)div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(dst, src, n);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_loss_total(const float* terms, int32_t n_terms, const float* extra, float* out, void* stream) {
  B2N_REQUIRE(terms && out && n_terms >= 1, "bad arguments");
  loss_total_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(terms, n_terms, extra, out);
  B2N_LAUNCH_CHECK();
}

extern "C" int b2n_zero_async(void* ptr, int64_t bytes, void* stream) {
  if (bytes == 0) return B2N_OK;
  B2N_REQUIRE(ptr && bytes > 0, "bad arguments");
  const cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    b2n_set_error("%s: cudaMemsetAsync failed: %s", __func__, cudaGetErrorString(e));
    return (int)e;
  }
  return B2N_OK;
}
