// Fused simulation kernels: Philox -> normals -> log-Euler GBM -> payoff -> (sum, sum^2).
//
// Nothing per path or per step ever touches HBM: the only global traffic is the per-option
// parameter block read once per CTA (64 B x n_scen) and one (sum, sum^2) FP64 pair per
// (tile, scenario) written at the end.  A "tile" is one CTA's share of one option's paths:
// kBlock threads x paths_per_thread paths.  A second, tiny kernel folds the tile partials in a
// fixed order (deterministic: same seed => bit-identical moments, as the reference guarantees,
// tests/test_monte_carlo.py:153-158).
//
// All per-path arithmetic is FP32 on the quantity  l_t = log2(S_t / S_0)  (small magnitude, so
// FP32 rounding is ~3e-8 per step), payoffs are normalised by S_0 and re-scaled in FP64 by the
// fold kernel; cross-path accumulation is FP64 from the warp level up.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200mc.h"
#include "normal.cuh"
#include "philox.cuh"

namespace b200mc {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr int kMaxPathsPerThread = 32;

struct SimArgs {
  const b200mc_params_t* params;  // [n_opt][n_scen]
  double* partials;               // [n_opt * tiles][2 * NS]
  uint64_t path_begin;
  uint64_t n_paths;
  uint32_t n_opt, n_scen;
  uint32_t tiles;                 // tiles per option
  uint32_t paths_per_thread;
  uint32_t n_steps;
  uint32_t seed_lo, seed_hi;
  uint32_t stream_base;
  int32_t is_put, barrier_in, lookback_fixed, sgn_negative;
};

// Per-scenario FP32 coefficients, computed in FP64 once per CTA (the reference's constants:
// dt, drift, vol of gbm_numpy.py:35-39 / exotic_options.py:54-56, moved to log2 units).
struct Coef {
  float c;      // sgn * sigma*sqrt(dt) * sqrt(2/ln2): log2-diffusion per unit of log2-radius draw
  float d;      // sgn * (r - q - sigma^2/2)*dt / ln2: log2-drift per step
  float a;      // n_steps * (unsigned d): terminal log2-drift (European)
  float kappa;  // K / S
  float beta;   // sgn * log2(B / S)  (barrier)
  float inv_n;  // 1 / n_steps
};

__device__ __forceinline__ Coef make_coef(const b200mc_params_t& p, uint32_t n_steps, float sgn) {
  const double inv_ln2 = 1.44269504088896340736;
  const double dt = p.T / (double)n_steps;
  const double d = (p.r - p.q - 0.5 * p.sigma * p.sigma) * dt * inv_ln2;
  const double c = p.sigma * sqrt(dt) * 1.69864364966231197413;  // sqrt(2/ln 2)
  Coef k;
  k.c = sgn * (float)c;
  k.d = sgn * (float)d;
  k.a = (float)(d * (double)n_steps);
  k.kappa = (float)(p.K / p.S);
  k.beta = sgn * (float)(log2(p.barrier / p.S));
  k.inv_n = (float)(1.0 / (double)n_steps);
  return k;
}

__device__ __forceinline__ float vanilla(float e, float kappa, bool is_put) {
  return is_put ? fmaxf(kappa - e, 0.0f) : fmaxf(e - kappa, 0.0f);
}

// ---- block-level FP64 reduction of per-thread FP32 partial sums ------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce_store(const float (&v)[NV], double* dst) {
  __shared__ double warp_sums[kWarps][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = (double)v[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if (lane == 0) warp_sums[warp][i] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) x += warp_sums[w][threadIdx.x];
    dst[threadIdx.x] = x;
  }
}

// ---- Philox call for (path, step block), counter layout documented in philox.cuh ---------------
// The fast-varying word (step block) sits in c1, which round 1 only XORs: with path/stream/seed
// loop-invariant the compiler hoists both round-1 multiplies and one round-2 multiply out of the
// step loop (17 IMAD.WIDE + 18 LOP3 per call instead of 20 + 20).
__device__ __forceinline__ u32x4 draw4(uint64_t path, uint32_t blk, uint32_t stream, uint32_t k0, uint32_t k1) {
  return philox4x32<10>((uint32_t)path, blk, (uint32_t)(path >> 32), stream, k0, k1);
}

// ================================ European (terminal payoff) ====================================
// W' = sum over steps of rad*cos / rad*sin (log2-radius units); everything else happens once per path.
template <int ILP>
__device__ __forceinline__ void terminal_sums(const uint64_t (&path)[ILP], uint32_t n_steps, uint32_t stream,
                                              uint32_t k0, uint32_t k1, float (&W)[ILP]) {
#pragma unroll
  for (int i = 0; i < ILP; ++i) W[i] = 0.0f;
  const uint32_t full = n_steps >> 2;
  for (uint32_t blk = 0; blk < full; ++blk) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const u32x4 x = draw4(path[i], blk, stream, k0, k1);
      float r0, c0, s0, r1, c1, s1;
      box_muller_pair(x.x, x.y, r0, c0, s0);
      box_muller_pair(x.z, x.w, r1, c1, s1);
      W[i] = fmaf(r0, c0, W[i]);
      W[i] = fmaf(r0, s0, W[i]);
      W[i] = fmaf(r1, c1, W[i]);
      W[i] = fmaf(r1, s1, W[i]);
    }
  }
  const uint32_t rem = n_steps & 3u;
  if (rem) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const u32x4 x = draw4(path[i], full, stream, k0, k1);
      float r0, c0, s0, r1, c1, s1;
      box_muller_pair(x.x, x.y, r0, c0, s0);
      box_muller_pair(x.z, x.w, r1, c1, s1);
      W[i] = fmaf(r0, c0, W[i]);
      if (rem > 1) W[i] = fmaf(r0, s0, W[i]);
      if (rem > 2) W[i] = fmaf(r1, c1, W[i]);
    }
  }
}

template <int NS, bool ANTI, int ILP>
__global__ void __launch_bounds__(kBlock) european_kernel(const SimArgs a) {
  __shared__ Coef coef[NS];
  const uint32_t opt = blockIdx.x / a.tiles;
  const uint32_t tile = blockIdx.x - opt * a.tiles;
  if (threadIdx.x < NS) {
    const uint32_t k = threadIdx.x < a.n_scen ? threadIdx.x : a.n_scen - 1;
    coef[threadIdx.x] = make_coef(a.params[(size_t)opt * a.n_scen + k], a.n_steps, 1.0f);
  }
  __syncthreads();

  float acc[2 * NS];
#pragma unroll
  for (int i = 0; i < 2 * NS; ++i) acc[i] = 0.0f;

  const uint32_t stream = a.stream_base + opt;
  const bool is_put = a.is_put != 0;
  const uint64_t tile_first = (uint64_t)tile * (uint64_t)(kBlock * a.paths_per_thread);
  for (uint32_t j = 0; j < a.paths_per_thread; j += ILP) {
    uint64_t path[ILP];
    bool live[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const uint64_t local = tile_first + (uint64_t)(j + i) * kBlock + threadIdx.x;
      live[i] = (j + i) < a.paths_per_thread && local < a.n_paths;
      path[i] = a.path_begin + local;
    }
    if (!live[0]) break;  // paths are assigned in increasing order: nothing further for this thread
    float W[ILP];
    terminal_sums<ILP>(path, a.n_steps, stream, a.seed_lo, a.seed_hi, W);
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (!live[i]) continue;
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        const Coef q = coef[k];
        float p = vanilla(mufu_ex2(fmaf(q.c, W[i], q.a)), q.kappa, is_put);
        acc[2 * k] += p;
        acc[2 * k + 1] = fmaf(p, p, acc[2 * k + 1]);
        if (ANTI) {
          p = vanilla(mufu_ex2(fmaf(-q.c, W[i], q.a)), q.kappa, is_put);
          acc[2 * k] += p;
          acc[2 * k + 1] = fmaf(p, p, acc[2 * k + 1]);
        }
      }
    }
  }
  block_reduce_store<2 * NS>(acc, a.partials + (size_t)blockIdx.x * (2 * NS));
}

// ============================ path-dependent kinds (register state) =============================
template <int KIND>
__device__ __forceinline__ void step_update(float l, float& aux) {
  if (KIND == B200MC_ASIAN_ARITH) aux += mufu_ex2(l);
  else if (KIND == B200MC_ASIAN_GEOM) aux += l;
  else aux = fmaxf(aux, l);  // BARRIER / LOOKBACK: running max of sgn*l, seeded with l_0 = 0
}

template <int KIND, int NS>
__device__ __forceinline__ void advance_pair(float rad, float cs, float sn, int n_use, const Coef (&q)[NS],
                                             float (&l)[NS], float (&aux)[NS]) {
  if (NS == 1) {
    const float rc = rad * q[0].c;
    l[0] = fmaf(rc, cs, l[0] + q[0].d);
    step_update<KIND>(l[0], aux[0]);
    if (n_use > 1) {
      l[0] = fmaf(rc, sn, l[0] + q[0].d);
      step_update<KIND>(l[0], aux[0]);
    }
  } else {
    const float z0 = rad * cs, z1 = rad * sn;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      l[k] = fmaf(q[k].c, z0, l[k] + q[k].d);
      step_update<KIND>(l[k], aux[k]);
    }
    if (n_use > 1) {
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        l[k] = fmaf(q[k].c, z1, l[k] + q[k].d);
        step_update<KIND>(l[k], aux[k]);
      }
    }
  }
}

template <int KIND>
__device__ __forceinline__ float path_payoff(float l, float aux, const Coef& q, const SimArgs& a) {
  const bool is_put = a.is_put != 0;
  if (KIND == B200MC_ASIAN_ARITH) return vanilla(aux * q.inv_n, q.kappa, is_put);
  if (KIND == B200MC_ASIAN_GEOM) return vanilla(mufu_ex2(aux * q.inv_n), q.kappa, is_put);
  const float sgn = a.sgn_negative ? -1.0f : 1.0f;
  const float e_T = mufu_ex2(sgn * l);
  if (KIND == B200MC_BARRIER) {
    const bool crossed = aux >= q.beta;
    const bool active = crossed == (a.barrier_in != 0);
    return active ? vanilla(e_T, q.kappa, is_put) : 0.0f;
  }
  // LOOKBACK: aux tracks max(l) (sgn=+1) or max(-l) = -min(l) (sgn=-1)
  const float e_ext = mufu_ex2(sgn * aux);
  if (a.lookback_fixed) return vanilla(e_ext, q.kappa, is_put);
  return is_put ? e_ext - e_T : e_T - e_ext;
}

template <int KIND, int NS>
__global__ void __launch_bounds__(kBlock) pathdep_kernel(const SimArgs a) {
  __shared__ Coef coef_s[NS];
  const uint32_t opt = blockIdx.x / a.tiles;
  const uint32_t tile = blockIdx.x - opt * a.tiles;
  if (threadIdx.x < NS) {
    const uint32_t k = threadIdx.x < a.n_scen ? threadIdx.x : a.n_scen - 1;
    coef_s[threadIdx.x] = make_coef(a.params[(size_t)opt * a.n_scen + k], a.n_steps, a.sgn_negative ? -1.0f : 1.0f);
  }
  __syncthreads();
  Coef q[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) q[k] = coef_s[k];

  float acc[2 * NS];
#pragma unroll
  for (int i = 0; i < 2 * NS; ++i) acc[i] = 0.0f;

  const uint32_t stream = a.stream_base + opt;
  const uint64_t tile_first = (uint64_t)tile * (uint64_t)(kBlock * a.paths_per_thread);
  const uint32_t full = a.n_steps >> 2, rem = a.n_steps & 3u;
  for (uint32_t j = 0; j < a.paths_per_thread; ++j) {
    const uint64_t local = tile_first + (uint64_t)j * kBlock + threadIdx.x;
    if (local >= a.n_paths) break;
    const uint64_t path = a.path_begin + local;
    float l[NS], aux[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) l[k] = 0.0f, aux[k] = 0.0f;
    for (uint32_t blk = 0; blk < full; ++blk) {
      const u32x4 x = draw4(path, blk, stream, a.seed_lo, a.seed_hi);
      float r0, c0, s0, r1, c1, s1;
      box_muller_pair(x.x, x.y, r0, c0, s0);
      box_muller_pair(x.z, x.w, r1, c1, s1);
      advance_pair<KIND, NS>(r0, c0, s0, 2, q, l, aux);
      advance_pair<KIND, NS>(r1, c1, s1, 2, q, l, aux);
    }
    if (rem) {
      const u32x4 x = draw4(path, full, stream, a.seed_lo, a.seed_hi);
      float r0, c0, s0, r1, c1, s1;
      box_muller_pair(x.x, x.y, r0, c0, s0);
      box_muller_pair(x.z, x.w, r1, c1, s1);
      advance_pair<KIND, NS>(r0, c0, s0, rem > 1 ? 2 : 1, q, l, aux);
      if (rem > 2) advance_pair<KIND, NS>(r1, c1, s1, 1, q, l, aux);
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const float p = path_payoff<KIND>(l[k], aux[k], q[k], a);
      acc[2 * k] += p;
      acc[2 * k + 1] = fmaf(p, p, acc[2 * k + 1]);
    }
  }
  block_reduce_store<2 * NS>(acc, a.partials + (size_t)blockIdx.x * (2 * NS));
}

// ================================ fold tile partials -> moments ================================
// One warp per (option, scenario): lanes stride over the tiles in a fixed order, then a fixed
// shuffle tree.  Re-scales the S_0-normalised sums to currency units in FP64.
__global__ void __launch_bounds__(32) fold_kernel(const double* __restrict__ partials, const b200mc_params_t* __restrict__ params,
                                                  b200mc_moments_t* __restrict__ out, uint32_t n_scen, uint32_t ns_pad,
                                                  uint32_t tiles, double samples) {
  const uint32_t opt = blockIdx.x / n_scen, k = blockIdx.x - opt * n_scen;
  double s1 = 0.0, s2 = 0.0;
  for (uint32_t t = threadIdx.x; t < tiles; t += 32) {
    const double* p = partials + ((size_t)opt * tiles + t) * (2 * ns_pad) + 2 * k;
    s1 += p[0];
    s2 += p[1];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  }
  if (threadIdx.x == 0) {
    const double S = params[(size_t)opt * n_scen + k].S;
    out[blockIdx.x].sum = s1 * S;
    out[blockIdx.x].sum_sq = s2 * S * S;
    out[blockIdx.x].n = samples;
  }
}

// ================================ stream inspection kernels ====================================
__global__ void normals_kernel(uint32_t k0, uint32_t k1, uint32_t stream, uint64_t path_begin, uint64_t n_paths,
                               uint32_t n_steps, float* __restrict__ out) {
  const uint32_t nblk = (n_steps + 3) >> 2;
  const uint64_t total = n_paths * nblk;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t local = i / nblk;
    const uint32_t blk = (uint32_t)(i - local * nblk);
    const u32x4 x = draw4(path_begin + local, blk, stream, k0, k1);
    float r0, c0, s0, r1, c1, s1;
    box_muller_pair(x.x, x.y, r0, c0, s0);
    box_muller_pair(x.z, x.w, r1, c1, s1);
    const float z[4] = {kRadScale * r0 * c0, kRadScale * r0 * s0, kRadScale * r1 * c1, kRadScale * r1 * s1};
    for (uint32_t s = 0; s < 4 && blk * 4 + s < n_steps; ++s) out[local * n_steps + blk * 4 + s] = z[s];
  }
}

__global__ void philox_raw_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32x4 x = philox4x32<10>(in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3], in[6 * i + 4], in[6 * i + 5]);
  out[4 * i] = x.x, out[4 * i + 1] = x.y, out[4 * i + 2] = x.z, out[4 * i + 3] = x.w;
}

}  // namespace b200mc
