/* CPU restatement (plain C) of the integer part of the multiresolution hash grid — TEST INFRASTRUCTURE ONLY.
 *
 * Follows nerfstudio/field_components/encodings.py:398-435 (HashEncoding.hash_fn / pytorch_fwd):
 *   scaled = x * scalings[l]                (float32 multiply)
 *   c = (int32) ceil(scaled), f = (int32) floor(scaled)
 *   h = (X * 1) ^ (Y * 2654435761) ^ (Z * 805459861)      (int64 arithmetic)
 *   row = h mod T + l * T
 * for the eight corner selections hashed_0..7 of pytorch_fwd (encodings.py:427-434).
 * Pinned by tests/test_oracle_golden.py against indices recorded from the reference (tests/golden/hash_encoding.npz).
 * Build: make -C oracle   ->  oracle/libhash_index.so  (never linked into the product library).
 */
#include <math.h>
#include <stdint.h>

static int64_t pymod(int64_t a, int64_t m) { /* Python / torch remainder: result has the sign of m */
  int64_t r = a % m;
  return r < 0 ? r + m : r;
}

/* x [n,3] float32, scalings [L] float32, out [n, L, 8] int64 */
void hash_corner_indices(const float* x, int64_t n, const float* scalings, int32_t n_levels, int32_t log2_T,
                         int64_t* out) {
  const int64_t T = (int64_t)1 << log2_T;
  static const int pick[8][3] = {/* 1 = ceil, 0 = floor, per (x, y, z) */
                                 {1, 1, 1}, {1, 0, 1}, {0, 0, 1}, {0, 1, 1}, {1, 1, 0}, {1, 0, 0}, {0, 0, 0}, {0, 1, 0}};
  for (int64_t i = 0; i < n; ++i)
    for (int l = 0; l < n_levels; ++l) {
      int64_t c[3], f[3];
      for (int a = 0; a < 3; ++a) {
        volatile float s = x[3 * i + a] * scalings[l]; /* volatile: keep the separately rounded float32 product */
        c[a] = (int64_t)(int32_t)ceilf(s);
        f[a] = (int64_t)(int32_t)floorf(s);
      }
      for (int k = 0; k < 8; ++k) {
        const int64_t X = pick[k][0] ? c[0] : f[0], Y = pick[k][1] ? c[1] : f[1], Z = pick[k][2] ? c[2] : f[2];
        const int64_t h = (X * 1) ^ (Y * 2654435761LL) ^ (Z * 805459861LL);
        out[(i * n_levels + l) * 8 + k] = pymod(h, T) + (int64_t)l * T;
      }
    }
}
