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


// ---- European parity mode with bulk-async (TMA engine) staging ---------------------------------------------------
// The kernel above keeps at most 8 x 256 B per warp in flight and alternates load and compute phases; this one keeps
// the copy engine busy all the time.  Each warp owns 32 consecutive rows of Z and a private ring of kTmaStages
// shared-memory tiles; every lane issues ONE cp.async.bulk per chunk for its own row segment (kF64Chunk steps =
// 256 B, global -> shared, completion counted in bytes on the stage's mbarrier), so nothing passes through
// registers and kTmaStages x 8 KB per warp are in flight while the warp adds up the previous chunk.  Rows are parked at a
// pitch of 17 x 16 B, which makes the per-lane LDS.128 row walk bank-conflict free.  Requirements (checked by the
// host, which otherwise launches from_normals_kernel): n_steps even (16-byte row pitch in HBM) and Z 16-byte aligned.
// The arithmetic is the same statement sequence as from_normals_kernel<B200MC_EUROPEAN>.
constexpr int kTmaStages = 3;
constexpr int kTmaPitch = kF64Chunk * 8 + 16;                                   // bytes per parked row segment
constexpr int kTmaWarpBytes = kTmaStages * 32 * kTmaPitch;                      // 26112 B per warp
constexpr int kTmaSmemBytes = kF64Warps * kTmaWarpBytes;                        // 104448 B per CTA -> 2 CTAs per SM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(kF64Block) european_from_normals_tma_kernel(const F64Args a) {
  extern __shared__ __align__(128) unsigned char tma_tiles[];
  __shared__ __align__(8) unsigned long long bars[kF64Warps][kTmaStages];
  __shared__ double red[kF64Warps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warp_first = ((uint64_t)blockIdx.x * kF64Warps + warp) * 32;
  const uint64_t path = warp_first + lane;
  const bool live = path < a.n_paths;
  const uint32_t live_rows = warp_first < a.n_paths ? (uint32_t)min((uint64_t)32, a.n_paths - warp_first) : 0u;
  const bool is_put = a.is_put != 0;
  const uint32_t n = a.n_steps;
  const uint32_t n_chunks = (n + kF64Chunk - 1) / kF64Chunk;

  const double dt = __ddiv_rn(a.T, (double)n);
  const double drift = __dmul_rn(__dsub_rn(__dsub_rn(a.r, a.q), __dmul_rn(__dmul_rn(0.5, a.sigma), a.sigma)), dt);
  const double vol = __dmul_rn(a.sigma, sqrt(dt));
  const double log_S0 = log(a.S);

  unsigned char* my_tiles = tma_tiles + (size_t)warp * kTmaWarpBytes;
  const uint32_t tiles_u32 = smem_u32(my_tiles);
  const uint32_t bar0 = smem_u32(&bars[warp][0]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) mbar_init(bar0 + 8u * s, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const double* my_row = a.Z + path * n;  // only dereferenced (by the copy engine) when live
  auto issue = [&](uint32_t c) {
    const uint32_t stage = c % kTmaStages;
    const uint32_t bytes = min((uint32_t)kF64Chunk, n - c * kF64Chunk) * 8u;
    if (lane == 0) mbar_expect_tx(bar0 + 8u * stage, live_rows * bytes);
    __syncwarp();
    if (live) bulk_g2s(tiles_u32 + (stage * 32u + (uint32_t)lane) * kTmaPitch, my_row + (size_t)c * kF64Chunk, bytes, bar0 + 8u * stage);
  };

  double acc_pos = 0.0, acc_neg = 0.0;
  if (live_rows > 0) {
    for (uint32_t c = 0; c < (uint32_t)kTmaStages && c < n_chunks; ++c) issue(c);
    for (uint32_t c = 0; c < n_chunks; ++c) {
      const uint32_t stage = c % kTmaStages;
      const uint32_t width = min((uint32_t)kF64Chunk, n - c * kF64Chunk);
      mbar_wait(bar0 + 8u * stage, (c / kTmaStages) & 1u);
      if (live) {
        const double2* row = reinterpret_cast<const double2*>(my_tiles + (size_t)(stage * 32 + lane) * kTmaPitch);
        for (uint32_t j = 0; j < width / 2; ++j) {
          const double2 zz = row[j];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double z = h ? zz.y : zz.x;
            if (!a.accumulate) {
              acc_pos = __dadd_rn(acc_pos, z);
            } else {
              acc_pos = __dadd_rn(acc_pos, __dadd_rn(drift, __dmul_rn(vol, z)));
              if (a.antithetic) acc_neg = __dadd_rn(acc_neg, __dsub_rn(drift, __dmul_rn(vol, z)));
            }
          }
        }
      }
      __syncwarp();  // every lane has consumed this stage (in-order issue: its LDS results were used) before it is refilled
      if (c + kTmaStages < n_chunks) issue(c + kTmaStages);
    }
  }

  double p0 = 0.0, p1 = 0.0;
  if (live) {
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
    if (a.payoffs) {
      a.payoffs[path] = p0;
      if (a.antithetic) a.payoffs[a.n_paths + path] = p1;
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
