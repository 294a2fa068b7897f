// Device-side statistics of the normal stream at GPU scale (tests / diagnostics; b200mc_rng_statistics).
//
// A B200 produces 2e12 of these normals per second, so the evidence for the stream contract of normal.cuh (32 random
// bits per Box-Muller pair, 23-bit radius grid, angle bits shared with the radius' low bits) can be collected at 1e10
// draws in milliseconds instead of the 3e7 the CPU oracle manages:
//   * a 256-bin histogram of z over [-6, 6)                         -> binned chi-square against the normal law
//   * a 64 x 64 histogram of the two normals of ONE word over [-4, 4)^2 -> joint law of a pair (cos / sin branch)
//   * power sums and cross moments: same-word  E[z1 z2], E[z1^2 z2^2], E[z1 z2^3], E[z1^3 z2]
//                                   lag 1      E[a b], E[a^2 b^2], E[a b^3]  (sine branch of word n, cosine of word n+1)
//   * exact tail counts beyond +-4 sigma and +-5 sigma.
// The draws are exactly the ones the simulation kernels consume (for_each_pair over the same (seed, stream, path)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_kernels.cuh"

namespace b200mc {

constexpr int kStatsBinsZ = 256;
constexpr int kStatsBinsJoint = 64;
constexpr int kStatsMoments = 16;

struct RngStatsArgs {
  PhiloxKeys rk;
  uint32_t stream;
  uint64_t path_begin, n_paths;
  uint32_t n_steps;
  unsigned long long* hist_z;      // [256]
  unsigned long long* hist_joint;  // [64 * 64], row = cosine-branch normal, column = sine-branch normal
  double* moments;                 // [16], see b200mc_rng_stats_t
  unsigned long long* tails;       // [4]: z > 4, z > 5, z < -4, z < -5
};

__global__ void __launch_bounds__(kBlock) rng_stats_kernel(const RngStatsArgs a) {
  __shared__ unsigned int hz[kStatsBinsZ];
  __shared__ unsigned int hj[kStatsBinsJoint * kStatsBinsJoint];
  for (int i = threadIdx.x; i < kStatsBinsZ; i += kBlock) hz[i] = 0u;
  for (int i = threadIdx.x; i < kStatsBinsJoint * kStatsBinsJoint; i += kBlock) hj[i] = 0u;
  __syncthreads();
  double m[kStatsMoments];
#pragma unroll
  for (int i = 0; i < kStatsMoments; ++i) m[i] = 0.0;
  unsigned int t4 = 0, t5 = 0, tm4 = 0, tm5 = 0;
  // a CTA's shared counters are flushed every kFlush paths per thread so that no 32-bit bin can overflow
  constexpr uint32_t kFlush = 4096;
  uint32_t since_flush = 0;
  auto flush = [&]() {
    __syncthreads();
    for (int i = threadIdx.x; i < kStatsBinsZ; i += kBlock)
      if (hz[i]) atomicAdd(a.hist_z + i, (unsigned long long)hz[i]), hz[i] = 0u;
    for (int i = threadIdx.x; i < kStatsBinsJoint * kStatsBinsJoint; i += kBlock)
      if (hj[i]) atomicAdd(a.hist_joint + i, (unsigned long long)hj[i]), hj[i] = 0u;
    __syncthreads();
  };
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  const uint64_t rounds = (a.n_paths + stride - 1) / stride;  // every thread makes the same number of rounds (barriers in flush)
  for (uint64_t round = 0; round < rounds; ++round) {
    const uint64_t local = round * stride + (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (local < a.n_paths) {
      float prev = 0.0f;
      bool have_prev = false;
      float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, c11 = 0.f, c22 = 0.f, c13 = 0.f, c31 = 0.f, l11 = 0.f, l22 = 0.f, l13 = 0.f;
      uint32_t n_pairs = 0, n_lag = 0, n_z = 0;
      for_each_pair(a.path_begin + local, a.n_steps, a.stream, a.rk, [&](const NormalPair& p, int n_use) {
        const float z0 = kRadScale * p.rad * p.cs;
        const float z1 = kRadScale * p.rad * p.sn;
        auto one = [&](float z) {
          const int b = min(max(__float2int_rd((z + 6.0f) * (kStatsBinsZ / 12.0f)), 0), kStatsBinsZ - 1);
          atomicAdd(&hz[b], 1u);
          const float z2 = z * z;
          s1 += z, s2 += z2, s3 = fmaf(z2, z, s3), s4 = fmaf(z2, z2, s4);
          t4 += z > 4.0f, t5 += z > 5.0f, tm4 += z < -4.0f, tm5 += z < -5.0f;
          n_z += 1;
        };
        one(z0);
        if (have_prev) {  // lag 1 across words: sine branch of the previous word, cosine branch of this one
          l11 = fmaf(prev, z0, l11), l22 = fmaf(prev * prev, z0 * z0, l22), l13 = fmaf(prev, z0 * z0 * z0, l13);
          n_lag += 1;
        }
        if (n_use > 1) {
          one(z1);
          const int r = min(max(__float2int_rd((z0 + 4.0f) * (kStatsBinsJoint / 8.0f)), 0), kStatsBinsJoint - 1);
          const int c = min(max(__float2int_rd((z1 + 4.0f) * (kStatsBinsJoint / 8.0f)), 0), kStatsBinsJoint - 1);
          atomicAdd(&hj[r * kStatsBinsJoint + c], 1u);
          c11 = fmaf(z0, z1, c11), c22 = fmaf(z0 * z0, z1 * z1, c22), c13 = fmaf(z0, z1 * z1 * z1, c13), c31 = fmaf(z0 * z0 * z0, z1, c31);
          n_pairs += 1;
          prev = z1, have_prev = true;
        }
      });
      // per-path FP32 sums (<= a few thousand terms) -> FP64 per thread
      m[0] += s1, m[1] += s2, m[2] += s3, m[3] += s4, m[4] += c11, m[5] += c22, m[6] += c13, m[7] += c31;
      m[8] += l11, m[9] += l22, m[10] += l13, m[11] += (double)n_z, m[12] += (double)n_pairs, m[13] += (double)n_lag;
    }
    if (++since_flush == kFlush / 16 || round + 1 == rounds) {
      flush();
      since_flush = 0;
    }
  }
  // moments: FP64 warp reduction, one atomic per warp and moment (diagnostic sums: the order is not fixed)
#pragma unroll
  for (int i = 0; i < kStatsMoments; ++i) {
    double x = m[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(a.moments + i, x);
  }
  unsigned int t[4] = {t4, t5, tm4, tm5};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unsigned int x = t[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(a.tails + i, (unsigned long long)x);
  }
}

}  // namespace b200mc
