// FP64 parity mode: price from caller-supplied normal draws Z[n_paths][n_steps] (row-major FP64),
// mirroring the reference's expression shapes so that, fed the reference's own NumPy draws, the
// per-path payoffs agree to ~1e-14 (bar: 1e-12 relative, tests/test_parity_f64.py).
//
// HBM-bound by design: 8 bytes of Z per path-step, read exactly once.  Each warp owns 32 consecutive
// paths; a 32-step chunk of their rows is loaded coalesced (one 256-byte row segment per warp load,
// 32 independent loads in flight per lane), parked in a padded warp-private shared-memory tile, and
// then every lane walks its own row sequentially — the reference's cumsum order
// (exotic_options.py:62-65, monte_carlo_unified.py:333-337).  No FMA contraction: adds and
// multiplies are issued as separate correctly-rounded ops, as NumPy evaluates them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200mc.h"

namespace b200mc {

constexpr int kF64Warps = 4;
constexpr int kF64Block = kF64Warps * 32;
constexpr int kF64Chunk = 32;

struct F64Args {
  const double* Z;    // [n_paths][n_steps]
  double* payoffs;    // [n_paths] or [2*n_paths]; may be null
  double* partials;   // [gridDim.x][2]
  uint64_t n_paths;
  uint32_t n_steps;
  int32_t accumulate, antithetic, is_put, barrier_down, barrier_in, lookback_fixed;
  double S, K, T, r, sigma, q, barrier;
};

__device__ __forceinline__ double vanilla64(double s, double K, bool is_put) {
  return is_put ? fmax(__dsub_rn(K, s), 0.0) : fmax(__dsub_rn(s, K), 0.0);
}

template <int KIND>
__global__ void __launch_bounds__(kF64Block) from_normals_kernel(const F64Args a) {
  __shared__ double tile[kF64Warps][32][kF64Chunk + 1];
  __shared__ double red[kF64Warps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warp_first = ((uint64_t)blockIdx.x * kF64Warps + warp) * 32;
  const uint64_t path = warp_first + lane;
  const bool live = path < a.n_paths;
  const bool is_put = a.is_put != 0;
  const uint32_t n = a.n_steps;

  // constants exactly as gbm_numpy.py:35-39 / exotic_options.py:54-56 compute them
  const double dt = __ddiv_rn(a.T, (double)n);
  const double drift = __dmul_rn(__dsub_rn(__dsub_rn(a.r, a.q), __dmul_rn(__dmul_rn(0.5, a.sigma), a.sigma)), dt);
  const double vol = __dmul_rn(a.sigma, sqrt(dt));
  const double log_S0 = log(a.S);

  double acc_pos = 0.0, acc_neg = 0.0;  // running sums (W when accumulate == 0)
  double aux = 0.0;                       // Asian: running sum of S_t or log S_t
  double s_first = exp(log_S0);           // column 0 of the reference's path array (exotic_options.py:64,67)
  double ext_max = s_first, ext_min = s_first;
  bool crossed = false;
  if (KIND == B200MC_BARRIER) crossed = a.barrier_down ? (s_first <= a.barrier) : (s_first >= a.barrier);

  for (uint32_t c0 = 0; c0 < n; c0 += kF64Chunk) {
    const uint32_t width = min((uint32_t)kF64Chunk, n - c0);
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const uint64_t row = warp_first + i;
      if (row < a.n_paths && (uint32_t)lane < width) tile[warp][i][lane] = __ldg(a.Z + row * n + c0 + lane);
    }
    __syncwarp();
    if (live) {
      for (uint32_t j = 0; j < width; ++j) {
        const double z = tile[warp][lane][j];
        if (KIND == B200MC_EUROPEAN && !a.accumulate) {
          acc_pos = __dadd_rn(acc_pos, z);
        } else {
          acc_pos = __dadd_rn(acc_pos, __dadd_rn(drift, __dmul_rn(vol, z)));
          if (KIND == B200MC_EUROPEAN) {
            if (a.antithetic) acc_neg = __dadd_rn(acc_neg, __dsub_rn(drift, __dmul_rn(vol, z)));
          } else {
            const double s_t = exp(__dadd_rn(log_S0, acc_pos));
            if (KIND == B200MC_ASIAN_ARITH) aux = __dadd_rn(aux, s_t);
            if (KIND == B200MC_ASIAN_GEOM) aux = __dadd_rn(aux, log(s_t));
            if (KIND == B200MC_BARRIER) crossed = crossed || (a.barrier_down ? (s_t <= a.barrier) : (s_t >= a.barrier));
            if (KIND == B200MC_LOOKBACK) ext_max = fmax(ext_max, s_t), ext_min = fmin(ext_min, s_t);
          }
        }
      }
    }
    __syncwarp();
  }

  double p0 = 0.0, p1 = 0.0;
  if (live) {
    if (KIND == B200MC_EUROPEAN) {
      double up, down;
      if (!a.accumulate) {  // gbm_numpy.py:46,50
        const double base = __dadd_rn(log_S0, __dmul_rn(drift, (double)n));
        up = __dadd_rn(base, __dmul_rn(vol, acc_pos));
        down = __dsub_rn(base, __dmul_rn(vol, acc_pos));
      } else {              // monte_carlo_unified.py:333-341
        up = __dadd_rn(log_S0, acc_pos);
        down = __dadd_rn(log_S0, acc_neg);
      }
      p0 = vanilla64(exp(up), a.K, is_put);
      if (a.antithetic) p1 = vanilla64(exp(down), a.K, is_put);
    } else {
      const double s_T = exp(__dadd_rn(log_S0, acc_pos));
      if (KIND == B200MC_ASIAN_ARITH) p0 = vanilla64(__ddiv_rn(aux, (double)n), a.K, is_put);
      if (KIND == B200MC_ASIAN_GEOM) p0 = vanilla64(exp(__ddiv_rn(aux, (double)n)), a.K, is_put);
      if (KIND == B200MC_BARRIER) {
        const bool active = crossed == (a.barrier_in != 0);
        p0 = active ? vanilla64(s_T, a.K, is_put) : 0.0;
      }
      if (KIND == B200MC_LOOKBACK) {
        if (!a.lookback_fixed) p0 = is_put ? __dsub_rn(ext_max, s_T) : __dsub_rn(s_T, ext_min);
        else p0 = is_put ? fmax(__dsub_rn(a.K, ext_min), 0.0) : fmax(__dsub_rn(ext_max, a.K), 0.0);
      }
    }
    if (a.payoffs) {
      a.payoffs[path] = p0;
      if (KIND == B200MC_EUROPEAN && a.antithetic) a.payoffs[a.n_paths + path] = p1;
    }
  }
  double s1 = p0 + p1, s2 = p0 * p0 + p1 * p1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  }
  if (lane == 0) red[warp][0] = s1, red[warp][1] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int w = 0; w < kF64Warps; ++w) t1 += red[w][0], t2 += red[w][1];
    a.partials[2 * (size_t)blockIdx.x] = t1;
    a.partials[2 * (size_t)blockIdx.x + 1] = t2;
  }
}

}  // namespace b200mc
