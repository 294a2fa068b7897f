/* TEST INFRASTRUCTURE ONLY — plain-C oracle for the engine's random stream.
 *
 * Never linked into or loaded by optionslab_b200; built by __graft_entry__.build() into
 * oracle/_build/libphilox_oracle.so and loaded by tests/ (ctypes) as the checker.
 *
 * (1) philox4x32_10(): restatement of the published Philox4x32-10 algorithm (Salmon et al., SC'11;
 *     Random123 v1.14 include/Random123/philox.h).  The reference (OptionsLab) uses NumPy PCG64 /
 *     MT19937 and pins nothing about Philox, so this is pinned by the Random123 known-answer
 *     vectors only (tests/test_philox_oracle.py): "parity unpinned" w.r.t. the reference.
 * (2) b200mc_oracle_normals(): this repo's documented uniform->normal mapping (DESIGN.md "RNG stream
 *     contract"), evaluated in double precision with libm, as the ground truth the device's
 *     MUFU-approximated normals are compared to (abs tol ~1e-5).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double word_to_unit(uint32_t x) { /* float in [1,2) built from the top 23 bits */
  uint32_t bits = (x >> 9) | 0x3f800000u;
  float f;
  memcpy(&f, &bits, sizeof f);
  return (double)f;
}

/* One Box-Muller pair from two words: radius from xa, angle from xb. */
static void pair_to_normals(uint32_t xa, uint32_t xb, double* z_cos, double* z_sin) {
  double u = 2.0 - word_to_unit(xa);       /* (0, 1], grid 2^-23 */
  double turn = word_to_unit(xb) - 1.5;    /* [-0.5, 0.5) revolutions */
  double radius = sqrt(-2.0 * log(u));
  *z_cos = radius * cos(6.283185307179586476925 * turn);
  *z_sin = radius * sin(6.283185307179586476925 * turn);
}

/* out[(p - path_begin) * n_steps + s] = normal for step s of global path p. */
void b200mc_oracle_normals(uint64_t seed, uint32_t stream, uint64_t path_begin, uint64_t n_paths,
                           uint32_t n_steps, double* out) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (uint64_t i = 0; i < n_paths; ++i) {
    uint64_t p = path_begin + i;
    for (uint32_t blk = 0; blk * 4u < n_steps; ++blk) {
      uint32_t ctr[4] = {(uint32_t)p, (uint32_t)(p >> 32), blk, stream};
      uint32_t x[4];
      double z[4];
      philox4x32_10(ctr, key, x);
      pair_to_normals(x[0], x[1], &z[0], &z[1]);
      pair_to_normals(x[2], x[3], &z[2], &z[3]);
      for (uint32_t j = 0; j < 4 && blk * 4u + j < n_steps; ++j) out[i * n_steps + blk * 4u + j] = z[j];
    }
  }
}
