/* TEST INFRASTRUCTURE ONLY — plain-C oracle for the engine's random stream.
 *
 * Never linked into or loaded by optionslab_b200; built by __graft_entry__.build() into
 * oracle/_build/libphilox_oracle.so and loaded by tests/ (ctypes) as the checker.
 *
 * (1) philox4x32_10(): restatement of the published Philox4x32-10 algorithm (Salmon et al., SC'11;
 *     Random123 v1.14 include/Random123/philox.h).  The reference (OptionsLab) uses NumPy PCG64 /
 *     MT19937 and pins nothing about Philox, so this is pinned by the Random123 known-answer
 *     vectors only (tests/test_philox_oracle.py): "parity unpinned" w.r.t. the reference.
 * (2) b200mc_oracle_normals(): this repo's documented word->normal mapping (normal.cuh / DESIGN.md
 *     "RNG stream contract"), evaluated in double precision with libm, as the ground truth the device's
 *     MUFU-approximated normals are compared to (abs tol ~1e-5).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
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

/* One Box-Muller pair: radius from a full word, angle from a 16-bit integer h.
 * theta = (2^23 + h) * step + bias with the device's FP32 constants (normal.cuh kAngleStep /
 * kAngleBias), evaluated here in double: 65536 equally spaced angles covering [-pi, pi). */
static void pair_to_normals(uint32_t radius_word, uint32_t h, double* z_cos, double* z_sin) {
  const float step_f = 9.58737992428525768573e-5f;
  const float bias_f = -807.38931197248091f;
  double u = 2.0 - word_to_unit(radius_word);       /* (0, 1], grid 2^-23 */
  double theta = (8388608.0 + (double)h) * (double)step_f + (double)bias_f;
  double radius = sqrt(-2.0 * log(u));
  *z_cos = radius * cos(theta);
  *z_sin = radius * sin(theta);
}

/* out[(p - path_begin) * n_steps + s] = normal for step s of global path p.
 * Word stream of a path: Philox outputs of counters (path_lo, j, path_hi, stream), j = 0,1,2,...
 * Triple t = words 3t, 3t+1, 3t+2 -> steps 4t..4t+3 (see optionslab_b200/csrc/normal.cuh). */
void b200mc_oracle_normals(uint64_t seed, uint32_t stream, uint64_t path_begin, uint64_t n_paths,
                           uint32_t n_steps, double* out) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t n_triples = (n_steps + 3u) / 4u;
  uint32_t n_calls = (3u * n_triples + 3u) / 4u;
  uint32_t* w = (uint32_t*)malloc(sizeof(uint32_t) * 4u * (n_calls ? n_calls : 1u));
  for (uint64_t i = 0; i < n_paths; ++i) {
    uint64_t p = path_begin + i;
    for (uint32_t j = 0; j < n_calls; ++j) {
      uint32_t ctr[4] = {(uint32_t)p, j, (uint32_t)(p >> 32), stream};
      philox4x32_10(ctr, key, w + 4u * j);
    }
    for (uint32_t t = 0; t < n_triples; ++t) {
      double z[4];
      pair_to_normals(w[3u * t], w[3u * t + 2u] & 0xffffu, &z[0], &z[1]);
      pair_to_normals(w[3u * t + 1u], w[3u * t + 2u] >> 16, &z[2], &z[3]);
      for (uint32_t j = 0; j < 4u && 4u * t + j < n_steps; ++j) out[i * n_steps + 4u * t + j] = z[j];
    }
  }
  free(w);
}
