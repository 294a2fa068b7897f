// Uniform words -> standard normals, entirely in registers (device only).
//
// Box-Muller on the XU pipe: 4 MUFU per pair (LG2, SQRT, SIN, COS), no I2F.
//   word -> float in [1,2) by OR-ing the top 23 bits under exponent 0x3f8   (one LEA.HI)
//   u     = 2 - f            in [2^-23, 1]                                   (one FADD)
//   rad   = sqrt(-log2 u)    "radius in log2 units": the true Box-Muller radius is
//                            sqrt(-2 ln u) = kRadScale * rad, kRadScale = sqrt(2 ln 2).
//                            Kernels fold kRadScale into their per-scenario diffusion
//                            coefficient instead of multiplying every draw.
//   theta = 2*pi*g - 3*pi    g from the second word, theta in [-pi, pi)      (one FFMA)
//   z0 = kRadScale*rad*cos(theta), z1 = kRadScale*rad*sin(theta)
// The same mapping, in double precision with libm, is oracle/philox_oracle.c (tests only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200mc {

constexpr float kRadScale = 1.17741002251547469101f;    // sqrt(2 ln 2)
constexpr double kRadScaleD = 1.17741002251547469101;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kThreePi = 9.42477796076937971538f;

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

// One Box-Muller pair: rad (log2-unit radius), cs, sn.  Normals are kRadScale*rad*cs, kRadScale*rad*sn.
__device__ __forceinline__ void box_muller_pair(uint32_t xa, uint32_t xb, float& rad, float& cs, float& sn) {
  const float u = 2.0f - word_to_unit_1_2(xa);
  rad = mufu_sqrt(-mufu_lg2(u));
  const float theta = fmaf(word_to_unit_1_2(xb), kTwoPi, -kThreePi);
  cs = mufu_cos(theta);
  sn = mufu_sin(theta);
}

}  // namespace b200mc
