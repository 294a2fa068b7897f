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

static double mantissa_to_1_2(uint32_t mantissa23) { /* float in [1,2) with the given 23-bit mantissa */
  uint32_t bits = (mantissa23 & 0x007fffffu) | 0x3f800000u;
  float f;
  memcpy(&f, &bits, sizeof f);
  return (double)f;
}

static uint32_t byte_reverse(uint32_t w) {
  return (w >> 24) | ((w >> 8) & 0x0000ff00u) | ((w << 8) & 0x00ff0000u) | (w << 24);
}

/* One Box-Muller pair from one word (normal.cuh): radius mantissa = top 23 bits of the word, angle
 * mantissa = top 23 bits of the byte-reversed word (angle in turns, [1,2)); theta = turns * 2pi - 3pi
 * with the device's FP32 constants, evaluated here in double.  The radius carries the grid
 * normalisation kRadNormD (E[z^2] = 1). */
static void pair_to_normals(uint32_t w, double* z_cos, double* z_sin) {
  const float two_pi_f = 6.28318530717958647692f;
  const float minus_three_pi_f = -9.42477796076937971538f;
  const double rad_norm = 1.000000529893528569531;
  double u = 2.0 - mantissa_to_1_2(w >> 9);           /* (0, 1], grid 2^-23 */
  double theta = mantissa_to_1_2(byte_reverse(w) >> 9) * (double)two_pi_f + (double)minus_three_pi_f;
  double radius = rad_norm * sqrt(-2.0 * log(u));
  *z_cos = radius * cos(theta);
  *z_sin = radius * sin(theta);
}

/* The ONE normal of a single-step path (normal.cuh, box_muller_single): radius from the whole first word,
 * u = (w0 + 1) * 2^-32 evaluated as the device does (FP32 conversion of w0, one FP32 fma), angle from the top 23 bits of
 * the second word; no grid normalisation. */
static double single_step_normal(uint32_t w0, uint32_t w1) {
  const float two_pi_f = 6.28318530717958647692f;
  const float minus_three_pi_f = -9.42477796076937971538f;
  float fu = fmaf((float)w0, 0x1p-32f, 0x1p-32f);
  double theta = mantissa_to_1_2(w1 >> 9) * (double)two_pi_f + (double)minus_three_pi_f;
  return sqrt(-2.0 * log((double)fu)) * cos(theta);
}

/* out[(p - path_begin) * n_steps + s] = normal for step s of global path p.
 * Word stream of a path: Philox outputs of counters (path_lo, j, path_hi, stream), j = 0,1,2,...
 * Word n = 4j + i -> steps 2n (cos) and 2n+1 (sin)  (see optionslab_b200/csrc/normal.cuh); n_steps == 1: words 0 and 1 make
 * the path's one normal. */
void b200mc_oracle_normals(uint64_t seed, uint32_t stream, uint64_t path_begin, uint64_t n_paths,
                           uint32_t n_steps, double* out) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t n_pairs = (n_steps + 1u) / 2u;
  uint32_t n_calls = (n_pairs + 3u) / 4u;
  uint32_t* w = (uint32_t*)malloc(sizeof(uint32_t) * 4u * (n_calls ? n_calls : 1u));
  for (uint64_t i = 0; i < n_paths; ++i) {
    uint64_t p = path_begin + i;
    for (uint32_t j = 0; j < n_calls; ++j) {
      uint32_t ctr[4] = {(uint32_t)p, j, (uint32_t)(p >> 32), stream};
      philox4x32_10(ctr, key, w + 4u * j);
    }
    if (n_steps == 1u) {
      out[i] = single_step_normal(w[0], w[1]);
      continue;
    }
    for (uint32_t n = 0; n < n_pairs; ++n) {
      double z[2];
      pair_to_normals(w[n], &z[0], &z[1]);
      out[i * n_steps + 2u * n] = z[0];
      if (2u * n + 1u < n_steps) out[i * n_steps + 2u * n + 1u] = z[1];
    }
  }
  free(w);
}
