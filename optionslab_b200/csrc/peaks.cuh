// Pipe-rate probes: what the B200 in front of us can issue per second on the pipes the simulation
// kernels live on (FMA, integer multiply-wide, ALU logic, XU/MUFU, warp-instruction issue), plus
// two composite probes (Philox only; Philox + Box-Muller only).  They give the MEASURED roofline
// denominators that bench.py reports achieved fractions against — the kernels are bound by
// instruction issue and the XU pipe, not by HBM or tensor cores.
//
// Every probe keeps kChains independent dependency chains per thread (inline PTX so the optimiser
// cannot fold them), runs a full grid (SMs x 8 CTAs x 256 threads), and is timed with CUDA events.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_kernels.cuh"
#include "normal.cuh"
#include "philox.cuh"

namespace b200mc {
namespace probe {

constexpr int kChains = 8;

__global__ void ffma(uint32_t iters, float a, float b, float* out) {
  float x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = (float)(threadIdx.x + i);
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < kChains; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void imad_wide(uint32_t iters, uint32_t m, uint64_t* out) {
  uint64_t x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = threadIdx.x * 2654435761u + i;
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        uint32_t lo = (uint32_t)x[i];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(lo), "r"(m));
      }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void lop3(uint32_t iters, uint32_t a, uint32_t b, uint32_t* out) {
  uint32_t x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = threadIdx.x + i;
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < kChains; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a + u), "r"(b));
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The generator's MUFU mix: lg2, sqrt, sin, cos in equal parts (4 chains each kind would need
// range control; instead each chain cycles  x -> cos(x) -> sqrt -> lg2(1+.) -> sin, all bounded).
__global__ void mufu_mix(uint32_t iters, float* out) {
  float x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = 0.1f * (float)(threadIdx.x + i + 1);
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      asm volatile("abs.f32 %0, %0; sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void mufu_ex2(uint32_t iters, float* out) {
  float x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = -0.01f * (float)(threadIdx.x + i + 1);
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < kChains; ++i) asm volatile("neg.f32 %0, %0; ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Issue ceiling: 2 FFMA : 1 LOP3, all independent chains -> neither the FMA pipe (128 lanes/clk/SM)
// nor the ALU pipe (64) saturates before the 4 warp-instructions/clk/SM issue limit does.
__global__ void issue_mix(uint32_t iters, float a, float b, uint32_t c, float* out) {
  float x[kChains];
  uint32_t y[kChains / 2];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = (float)(threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < kChains / 2; ++i) y[i] = threadIdx.x + i;
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < kChains / 2; ++i) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[2 * i]) : "f"(a), "f"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(c + u), "r"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[2 * i + 1]) : "f"(a), "f"(b));
      }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < kChains / 2; ++i) s += (float)y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void philox_only(uint32_t iters, uint32_t k0, uint32_t k1, uint32_t* out) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t s = 0;
  for (uint32_t it = 0; it < iters; ++it) {
    const u32x4 x = philox4x32<10>(gid, it, 0u, 7u, k0, k1);
    s ^= x.x ^ x.y ^ x.z ^ x.w;
  }
  out[gid] = s;
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// clocks[2*cta] = SM cycles, clocks[2*cta+1] = nanoseconds spent by that CTA: their ratio is the SM
// clock actually sustained under this (integer + XU) load.
__global__ void normals_only(uint32_t iters, const PhiloxKeys rk, float* out, long long* clocks) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long n0 = global_ns();
  const long long t0 = clock64();
  // iters Philox calls -> iters * 8 normals, through the production word layout (mc_kernels.cuh)
  float W = 0.f;
  for_each_pair((uint64_t)gid, iters * 8, 7u, rk, [&](const NormalPair& p, int n_use) {
    W = fmaf(p.rad, p.cs, W);
    if (n_use > 1) W = fmaf(p.rad, p.sn, W);
  });
  out[gid] = W;
  if (threadIdx.x == 0) {
    clocks[2 * blockIdx.x] = clock64() - t0;
    clocks[2 * blockIdx.x + 1] = (long long)(global_ns() - n0);
  }
}

}  // namespace probe
}  // namespace b200mc
