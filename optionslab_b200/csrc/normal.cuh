// Philox words -> standard normals, entirely in registers (device only).
//
// RNG stream contract (restated in double precision in oracle/philox_oracle.c, tests only):
//   The word stream of (seed, stream, path) is the concatenation of the Philox4x32-10 outputs for
//   counters (path_lo, j, path_hi, stream), j = 0, 1, 2, ...   (key = seed).  Word n of the stream
//   (n = 4j + i) is ONE Box-Muller pair and feeds steps 2n (cosine branch) and 2n+1 (sine branch):
//   32 random bits per 2 normals, one Philox call per 8 path-steps.
//
// Box-Muller on the XU pipe: 4 MUFU per pair (LG2, SQRT, SIN, COS), no I2F anywhere.
//   word w = bytes [b3 b2 b1 b0] (b3 most significant).
//   radius: f = float in [1,2) whose mantissa is the word's TOP 23 bits [b3 b2 b1>>1] (one LEA.HI);
//           u = 2 - f in [2^-23, 1]; rad = sqrt(-log2 u).  The true radius is sqrt(-2 ln u) =
//           sqrt(2 ln 2) * rad; kernels fold kRadScale into their per-scenario diffusion coefficient
//           instead of multiplying every draw.  kRadScale also carries the factor 1 + 5.3e-7 that
//           makes the second moment of the 2^23-point radius grid exactly 2 (E[-2 ln u] over
//           u = j/2^23 is 2 * (1 - 1.0598e-6)), so E[z^2] = 1 exactly; the radius is capped at 5.65.
//   angle:  g = float in [1,2) whose mantissa is the top 23 bits of the BYTE-REVERSED word
//           [b0 b1 b2>>1] (one PRMT + one LEA.HI) = the angle in turns; theta = fma(g, 2pi, -3pi) in
//           [-pi, pi).  The radius reads the word's bits 31..9 (b3, b2, and b1 without its lowest bit); the angle's
//           leading 8 bits are b0 - bits the radius never sees - so every one of the 2^23 radius values is paired
//           with exactly 256 equally spaced angles: every product moment E[r^a cos^b sin^c] with b + c < 256
//           equals that of a continuous uniform angle.  From its 9th bit on the angle re-uses bits of the radius,
//           the radius' LEAST significant ones first (bit 0 of b1 is the angle's 16th bit and outside the radius;
//           b1>>1 moves r by < 2^-16 relative), which shifts those 256-angle combs by a quasi-independent offset:
//           the 2^32 (u, theta) points fill the square evenly instead of stacking on 256 lines (a fixed angle grid
//           puts an atom at z = 0; taking the offset from the radius' leading bits skews z near 0 and in the tails
//           -- both fail a Kolmogorov-Smirnov test at 1e7 draws; this layout passes at 3e7,
//           tests/test_philox_oracle.py, and a 1e10-draw binned chi-square on the device, tests/test_gpu_rng.py).
//   z_cos = kRadScale*rad*cos(theta), z_sin = kRadScale*rad*sin(theta)
//
// Consumers that only ever ADD the two normals of a pair (the terminal log-price of the European kernels) take them as ONE
// term, z_cos + z_sin = sqrt(2) * kRadScale*rad*sin(theta + pi/4) (box_muller_pair_sum): the same value from 3 MUFU.
//
// Single-step paths (n_steps == 1: the reference's DEFAULT, monte_carlo.py:59, gbm_numpy.py:56-83) are the one place where an
// option's value is a direct functional of ONE draw's tail, and the one place where a draw is not on the hot loop.  Their
// single normal therefore spends 64 bits (box_muller_single): the radius takes the WHOLE first word of the path's stream,
// u = (w0 + 1) * 2^-32 in (0, 1] (FP32 conversion: 24 significant bits, i.e. full relative resolution where u is small - the
// tail - and a 2^-24 grid near 1), cap sqrt(64 ln 2) = 6.66 sigma; the angle takes the top 23 bits of the second word.
// P(z > 5) is then exact to the sampling error at 2^32 draws instead of 3.7% thin.  Every other step count keeps the
// 32-bit-per-pair layout above (a sum of >= 2 draws no longer sees a single draw's far tail).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200mc {

constexpr double kRadNormD = 1.000000529893528569531;    // sqrt(2 / E[-2 ln u]) on the 2^23-point grid u = j / 2^23
constexpr double kRadScaleD = 1.177410646417426094868;   // sqrt(2 ln 2) * kRadNormD
constexpr float kRadScale = (float)kRadScaleD;
constexpr double kCoefScaleD = 1.698644500676289376733;  // sqrt(2 / ln 2) * kRadNormD: sigma*sqrt(dt) -> log2 units per rad
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kMinusThreePi = -9.42477796076937971538f;

__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sin(float x) {
  float y;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_cos(float x) {
  float y;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One Box-Muller pair from one 32-bit word.  Normals are kRadScale*rad*cs and kRadScale*rad*sn.
struct NormalPair {
  float rad, cs, sn;
};

// SQUARED = true leaves rad = -log2(u), the squared radius, for callers that fold another factor under the same
// square root (Heston: sqrt(v) * rad = sqrt(v * rad^2) is one MUFU.SQRT instead of two).
template <bool SQUARED = false>
__device__ __forceinline__ NormalPair box_muller(uint32_t w) {
  NormalPair p;
  const float u = 2.0f - __uint_as_float((w >> 9) | 0x3f800000u);
  p.rad = SQUARED ? -mufu_lg2(u) : mufu_sqrt(-mufu_lg2(u));
  const float turns = __uint_as_float((__byte_perm(w, 0u, 0x0123) >> 9) | 0x3f800000u);
  const float theta = fmaf(turns, kTwoPi, kMinusThreePi);
  p.cs = mufu_cos(theta);
  p.sn = mufu_sin(theta);
  return p;
}

// The SUM of a pair's two normals, for consumers that only ever add them (the terminal log-price of the European kernels):
//   rad cos(theta) + rad sin(theta) = sqrt(2) rad sin(theta + pi/4)
// - the same number the two branches add up to, from ONE MUFU.SIN instead of COS + SIN and one FFMA instead of two.  The
// pair comes back with cs = sin(theta + pi/4) (the caller owns the factor sqrt(2)) and sn = 0.
constexpr float kMinusElevenQuarterPi = -8.63937979737193138115f;  // -3 pi + pi / 4
template <bool SQUARED = false>
__device__ __forceinline__ NormalPair box_muller_pair_sum(uint32_t w) {
  NormalPair p;
  const float u = 2.0f - __uint_as_float((w >> 9) | 0x3f800000u);
  p.rad = SQUARED ? -mufu_lg2(u) : mufu_sqrt(-mufu_lg2(u));
  const float turns = __uint_as_float((__byte_perm(w, 0u, 0x0123) >> 9) | 0x3f800000u);
  p.cs = mufu_sin(fmaf(turns, kTwoPi, kMinusElevenQuarterPi));
  p.sn = 0.0f;
  return p;
}

// The one normal of a single-step path from the first two words of its stream (see the header).  rad is divided by the
// 2^23-grid normalisation the kernels fold into their diffusion coefficient (this grid needs none: E[-2 ln u] = 2 to 1e-8).
template <bool SQUARED = false>
__device__ __forceinline__ NormalPair box_muller_single(uint32_t w0, uint32_t w1) {
  NormalPair p;
  const float u = fmaf(__uint2float_rn(w0), 2.3283064365386963e-10f, 2.3283064365386963e-10f);  // (w0 + 1) / 2^32, <= 1
  const float r2 = -mufu_lg2(u);
  constexpr float kInvNorm = (float)(1.0 / kRadNormD), kInvNorm2 = (float)(1.0 / (kRadNormD * kRadNormD));
  p.rad = SQUARED ? r2 * kInvNorm2 : mufu_sqrt(r2) * kInvNorm;
  const float turns = __uint_as_float((w1 >> 9) | 0x3f800000u);
  const float theta = fmaf(turns, kTwoPi, kMinusThreePi);
  p.cs = mufu_cos(theta);
  p.sn = mufu_sin(theta);
  return p;
}

}  // namespace b200mc
