// Philox words -> standard normals, entirely in registers (device only).
//
// RNG stream contract (restated in double precision in oracle/philox_oracle.c, tests only):
//   The word stream of (seed, stream, path) is the concatenation of the Philox4x32-10 outputs for
//   counters (path_lo, j, path_hi, stream), j = 0, 1, 2, ...   (key = seed).  Word triple t =
//   (w[3t], w[3t+1], w[3t+2]) yields the four normals of steps 4t .. 4t+3 by two Box-Muller pairs:
//     pair A: radius from w[3t],   angle from the LOW  16 bits of w[3t+2]  -> steps 4t   (cos), 4t+1 (sin)
//     pair B: radius from w[3t+1], angle from the HIGH 16 bits of w[3t+2]  -> steps 4t+2 (cos), 4t+3 (sin)
//   96 random bits per 4 normals: 3 Philox calls feed 16 path-steps.
//
// Box-Muller on the XU pipe: 4 MUFU per pair (LG2, SQRT, SIN, COS), no I2F anywhere.
//   radius: f = float in [1,2) from the word's top 23 bits (one LEA.HI); u = 2 - f in [2^-23, 1];
//           rad = sqrt(-log2 u).  The true radius is sqrt(-2 ln u) = kRadScale * rad with
//           kRadScale = sqrt(2 ln 2); kernels fold kRadScale into their per-scenario diffusion
//           coefficient instead of multiplying every draw.
//   angle:  h = 16-bit integer; g = float 2^23 + h (bit pattern 0x4b000000 | h, exact);
//           theta = fma(g, 2pi/65536, -(2^23 * 2pi/65536 + pi))  in [-pi, pi): one FFMA.  The grid of
//           65536 equally spaced angles integrates every trigonometric polynomial of degree < 65536
//           exactly, so all mixed moments of (z_cos, z_sin) up to that degree are those of the
//           continuous angle (the float32 sin/cos resolve ~2^-22 of a turn anyway).
//   z_cos = kRadScale*rad*cos(theta), z_sin = kRadScale*rad*sin(theta)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200mc {

constexpr float kRadScale = 1.17741002251547469101f;    // sqrt(2 ln 2)
constexpr double kRadScaleD = 1.17741002251547469101;
constexpr float kAngleStep = 9.58737992428525768573e-5f;   // 2*pi / 65536
constexpr float kAngleBias = -807.38931197248091f;          // -(2^23 * 2*pi/65536 + pi) = -(256 + 1) * pi

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

__device__ __forceinline__ float word_to_unit_1_2(uint32_t x) {
  return __uint_as_float((x >> 9) | 0x3f800000u);
}

// One Box-Muller pair from a radius word and a 16-bit angle (passed already OR-ed under 0x4b000000).
// Normals are kRadScale*rad*cs and kRadScale*rad*sn.
struct NormalPair {
  float rad, cs, sn;
};

__device__ __forceinline__ NormalPair box_muller(uint32_t radius_word, uint32_t angle_bits_4b) {
  NormalPair p;
  const float u = 2.0f - word_to_unit_1_2(radius_word);
  p.rad = mufu_sqrt(-mufu_lg2(u));
  const float theta = fmaf(__uint_as_float(angle_bits_4b), kAngleStep, kAngleBias);
  p.cs = mufu_cos(theta);
  p.sn = mufu_sin(theta);
  return p;
}

// Word triple -> two pairs (four consecutive steps).
__device__ __forceinline__ void box_muller_quad(uint32_t wa, uint32_t wb, uint32_t wc, NormalPair& A, NormalPair& B) {
  A = box_muller(wa, (wc & 0xffffu) | 0x4b000000u);
  B = box_muller(wb, (wc >> 16) | 0x4b000000u);
}

}  // namespace b200mc
