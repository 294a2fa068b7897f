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
  PhiloxKeys rk;                  // expanded from the 64-bit seed on the host
  uint32_t stream_base;
  int32_t is_put, barrier_in, lookback_fixed, sgn_negative;
  int32_t force_mufu_ex2;         // arithmetic Asian: always take the MUFU.EX2 form (tests / A-B timing)
};

// Per-scenario FP32 coefficients, computed in FP64 once per CTA (the reference's constants:
// dt, drift, vol of gbm_numpy.py:35-39 / exotic_options.py:54-56, moved to log2 units).
struct Coef {
  float c;      // sgn * sigma*sqrt(dt) * kCoefScaleD: log2-diffusion per unit of log2-radius draw
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
  const double c = p.sigma * sqrt(dt) * kCoefScaleD;
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

// ---- Philox call j of a path's word stream (layout documented in normal.cuh) -------------------
// The fast-varying word (call index) sits in c1, which round 1 only XORs: with path/stream/seed
// loop-invariant the compiler hoists both round-1 multiplies and one round-2 multiply out of the
// step loop (16 IMAD.WIDE + 18 LOP3 per call instead of 20 + 20).
__device__ __forceinline__ u32x4 draw4(uint64_t path, uint32_t call, uint32_t stream, const PhiloxKeys& rk) {
  return philox4x32_10((uint32_t)path, call, (uint32_t)(path >> 32), stream, rk);
}

// Visit the Box-Muller pairs of one path in step order: f(pair, n_use) with n_use = 2 except for a
// trailing odd step.  One Philox call = 4 words = 4 pairs = 8 steps (layout documented in normal.cuh).
// UNROLL calls are drawn before any is consumed, so their multiply chains interleave on the fmaheavy
// pipe; which UNROLL wins depends on the consumer's register appetite (profiles/r01_variants.txt).
template <bool SQUARED, class F>
__device__ __forceinline__ void consume_call(const u32x4& x, F&& f) {
  f(box_muller<SQUARED>(x.x), 2);
  f(box_muller<SQUARED>(x.y), 2);
  f(box_muller<SQUARED>(x.z), 2);
  f(box_muller<SQUARED>(x.w), 2);
}

// SQUARED: the pairs carry rad^2 = -log2(u) instead of rad (see box_muller).
template <int UNROLL = 1, bool SQUARED = false, class F>
__device__ __forceinline__ void for_each_pair(uint64_t path, uint32_t n_steps, uint32_t stream, const PhiloxKeys& rk, F&& f) {
  const uint32_t full = n_steps >> 3;
  uint32_t j = 0;
  if (UNROLL > 1) {
    for (; j + UNROLL <= full; j += UNROLL) {
      u32x4 x[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) x[u] = draw4(path, j + u, stream, rk);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) consume_call<SQUARED>(x[u], f);
    }
  }
  // (plain #pragma unroll 2 / 4 of this loop measures 2.5% / 2% slower on the European kernel: profiles/r01_variants16_ffma2.txt)
  for (; j < full; ++j) consume_call<SQUARED>(draw4(path, j, stream, rk), f);
  const int rem = (int)(n_steps & 7u);
  if (rem) {  // 1..7 trailing steps: same word layout, only the pairs that are needed
    const u32x4 x = draw4(path, full, stream, rk);
    f(box_muller<SQUARED>(x.x), rem >= 2 ? 2 : 1);
    if (rem > 2) f(box_muller<SQUARED>(x.y), rem >= 4 ? 2 : 1);
    if (rem > 4) f(box_muller<SQUARED>(x.z), rem >= 6 ? 2 : 1);
    if (rem > 6) f(box_muller<SQUARED>(x.w), 1);
  }
}

// Packed FP32 pairs (Blackwell fma.rn.f32x2 / mul.rn.f32x2 -> FFMA2 / FMUL2): two FP32 FMAs per issue slot, with the
// roundings of the scalar fmaf sequence.  Operands whose halves are equal compile to FFMA2 immediates / broadcasts.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ================================ European (terminal payoff) ====================================
// W' = sum over steps of rad*cos / rad*sin (log2-radius units); everything else happens once per path.
// (Accumulating the two branches in the halves of one packed FFMA2 register saves 4 issue slots per 8 steps and
// measures 0.6% slower: the European loop is not issue-bound.  profiles/r01_variants16_ffma2.txt)
template <int UNROLL = 1>
__device__ __forceinline__ float terminal_sum(uint64_t path, uint32_t n_steps, uint32_t stream, const PhiloxKeys& rk) {
  float W = 0.0f;
  for_each_pair<UNROLL>(path, n_steps, stream, rk, [&](const NormalPair& p, int n_use) {
    W = fmaf(p.rad, p.cs, W);
    if (n_use > 1) W = fmaf(p.rad, p.sn, W);
  });
  return W;
}

// CV = true additionally accumulates sum S_T, sum S_T^2 and sum payoff*S_T per scenario (the
// terminal-spot control variate of monte_carlo.py:154-186): 5 sums instead of 2.
template <int NS, bool ANTI, int MINB, bool CV = false, int UNROLL = 1>
__global__ void __launch_bounds__(kBlock, MINB) european_kernel(const SimArgs a) {
  constexpr int NM = CV ? 5 : 2;
  __shared__ Coef coef[NS];
  const uint32_t opt = blockIdx.x / a.tiles;
  const uint32_t tile = blockIdx.x - opt * a.tiles;
  if (threadIdx.x < NS) {
    const uint32_t k = threadIdx.x < a.n_scen ? threadIdx.x : a.n_scen - 1;
    coef[threadIdx.x] = make_coef(a.params[(size_t)opt * a.n_scen + k], a.n_steps, 1.0f);
  }
  __syncthreads();

  float acc[NM * NS];
#pragma unroll
  for (int i = 0; i < NM * NS; ++i) acc[i] = 0.0f;

  const uint32_t stream = a.stream_base + opt;
  const bool is_put = a.is_put != 0;
  const uint64_t tile_first = (uint64_t)tile * (uint64_t)(kBlock * a.paths_per_thread);
  for (uint32_t j = 0; j < a.paths_per_thread; ++j) {
    const uint64_t local = tile_first + (uint64_t)j * kBlock + threadIdx.x;
    if (local >= a.n_paths) break;  // paths are assigned in increasing order: nothing further for this thread
    const float W = terminal_sum<UNROLL>(a.path_begin + local, a.n_steps, stream, a.rk);
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const Coef q = coef[k];
#pragma unroll
      for (int mirror = 0; mirror < (ANTI ? 2 : 1); ++mirror) {
        const float e = mufu_ex2(fmaf(mirror ? -q.c : q.c, W, q.a));  // S_T / S_0
        const float p = vanilla(e, q.kappa, is_put);
        acc[NM * k] += p;
        acc[NM * k + 1] = fmaf(p, p, acc[NM * k + 1]);
        if (CV) {
          acc[NM * k + 2] += e;
          acc[NM * k + 3] = fmaf(e, e, acc[NM * k + 3]);
          acc[NM * k + 4] = fmaf(p, e, acc[NM * k + 4]);
        }
      }
    }
  }
  block_reduce_store<NM * NS>(acc, a.partials + (size_t)blockIdx.x * (NM * NS));
}

// ============================ path-dependent kinds (register state) =============================
template <int KIND>
__device__ __forceinline__ void step_update(float l, float& aux) {
  if (KIND == B200MC_ASIAN_ARITH) aux += mufu_ex2(l);
  else if (KIND == B200MC_ASIAN_GEOM) aux += l;
  else aux = fmaxf(aux, l);  // BARRIER / LOOKBACK: running max of sgn*l, seeded with l_0 = 0
}

// Two consecutive steps from one pair.  Both increments come from ONE packed FFMA2, (rc*cos + d, rc*sin + d), and are
// then added to the running log2-price in step order.  The same expression shape for every NS, so a scenario's
// result does not depend on how many other scenarios share the launch: fused Greeks == separate re-pricings, bit for bit.
template <int KIND, int NS>
__device__ __forceinline__ void advance_pair(const NormalPair& p, int n_use, const Coef (&q)[NS], float (&l)[NS], float (&aux)[NS]) {
  const f32x2 cssn = pack2(p.cs, p.sn);
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const float rc = p.rad * q[k].c;
    float inc0, inc1;
    unpack2(fma2(pack2(rc, rc), cssn, pack2(q[k].d, q[k].d)), inc0, inc1);
    l[k] += inc0;
    step_update<KIND>(l[k], aux[k]);
    if (n_use > 1) {
      l[k] += inc1;
      step_update<KIND>(l[k], aux[k]);
    }
  }
}

// ---- arithmetic Asian, small per-step moves: multiplicative update without MUFU.EX2 -------------------
// The arithmetic average needs S_t itself every step, i.e. a third MUFU per path-step on a kernel the XU pipe
// already bounds.  When every possible log2-increment x = d + c*rad*cos of an option is small
// (|d| + |c|*kRadMax <= kSmallMove; daily steps up to sigma ~ 0.48) the kernel tracks s_t = S_t/S_0 directly:
//   s_t = s_{t-1} + s_{t-1} * (2^x - 1),   2^x - 1 = x*ln2 * (1 + x*ln2/2 * (1 + ...)) to degree 4 on the FMA pipe,
// the two steps of a Box-Muller pair evaluated together with packed FFMA2 (119 instead of 147 issue slots per 8 steps).
// Truncation: the dropped (x ln2)^5/120 term is odd in the draw (mean zero; 1.3e-6 relative for a single 5.65-sigma draw
// at the bound, ~1e-11 for a typical one), the first even - biased - term (x ln2)^6/720 is <= 3.8e-8 at the bound and
// ~1e-14 typically; rounding is half an ulp of s per step, the same order as the additive form's 3e-8 on l_t.
// Measured on identical draws at 16M paths: prices of the two forms differ by <= 3e-8 relative (tools/asian_forms_diff.py).
// The choice is made per CTA (= per option, over all its scenarios) from the coefficients alone.
constexpr float kRadMax = 4.79583152331271954f;  // sqrt(23): u >= 2^-23 (normal.cuh)
constexpr float kSmallMove = 0.25f;

template <int DEG = 5>
__device__ __forceinline__ float exp2m1_small(float x) {
  float t = DEG >= 5 ? fmaf(x, 1.3333558146e-3f, 9.6181291076e-3f) : 9.6181291076e-3f;  // ln2^5/120, ln2^4/24
  if (DEG >= 4) t = fmaf(x, t, 5.5504108665e-2f);                                         // ln2^3/6
  else t = 5.5504108665e-2f;
  t = fmaf(x, t, 2.4022650696e-1f);                                                       // ln2^2/2
  t = fmaf(x, t, 6.9314718056e-1f);                                                       // ln2
  return x * t;
}

// DEG = kSmallPacked4 (shipped) / kSmallPacked5: packed degree 4 / 5.  DEG = 3..5: scalar Horner forms (4+ scenarios, scratch/variants14.cu).
constexpr int kSmallPacked5 = -5, kSmallPacked4 = -4;

template <int NS, int DEG = kSmallPacked4>
__device__ __forceinline__ void advance_pair_small(const NormalPair& p, int n_use, const Coef (&q)[NS], float (&s)[NS], float (&aux)[NS]) {
  if (DEG < 0) {  // packed forms: degree -DEG
    const f32x2 cssn = pack2(p.cs, p.sn);
    const f32x2 c5 = pack2(1.3333558146e-3f, 1.3333558146e-3f), c4 = pack2(9.6181291076e-3f, 9.6181291076e-3f),
                c3 = pack2(5.5504108665e-2f, 5.5504108665e-2f), c2 = pack2(2.4022650696e-1f, 2.4022650696e-1f),
                c1 = pack2(6.9314718056e-1f, 6.9314718056e-1f);
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const float rc = p.rad * q[k].c;
      const f32x2 x = fma2(pack2(rc, rc), cssn, pack2(q[k].d, q[k].d));
      f32x2 t = DEG <= -5 ? fma2(x, c5, c4) : c4;  // degree 4: drop the x^5 term
      t = fma2(x, t, c3);
      t = fma2(x, t, c2);
      t = fma2(x, t, c1);
      float y0, y1;
      unpack2(mul2(x, t), y0, y1);
      s[k] = fmaf(s[k], y0, s[k]);
      aux[k] += s[k];
      if (n_use > 1) {
        s[k] = fmaf(s[k], y1, s[k]);
        aux[k] += s[k];
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const float rc = p.rad * q[k].c;
      s[k] = fmaf(s[k], exp2m1_small<DEG>(fmaf(rc, p.cs, q[k].d)), s[k]);
      aux[k] += s[k];
      if (n_use > 1) {
        s[k] = fmaf(s[k], exp2m1_small<DEG>(fmaf(rc, p.sn, q[k].d)), s[k]);
        aux[k] += s[k];
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

// One CTA's share of one option: paths_per_thread paths per thread, payoffs accumulated in FP32 per thread.
// SMALL selects the multiplicative arithmetic-Asian update (state = S_t/S_0 instead of log2 of it).
template <int KIND, int NS, bool SMALL, int UNROLL, int DEG = kSmallPacked4>
__device__ __forceinline__ void simulate_tile(const SimArgs& a, const Coef (&q)[NS], uint32_t stream, uint64_t tile_first, float (&acc)[2 * NS]) {
  for (uint32_t j = 0; j < a.paths_per_thread; ++j) {
    const uint64_t local = tile_first + (uint64_t)j * kBlock + threadIdx.x;
    if (local >= a.n_paths) break;
    float l[NS], aux[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) l[k] = SMALL ? 1.0f : 0.0f, aux[k] = 0.0f;
    for_each_pair<UNROLL>(a.path_begin + local, a.n_steps, stream, a.rk, [&](const NormalPair& p, int n_use) {
      if (SMALL) advance_pair_small<NS, DEG>(p, n_use, q, l, aux);
      else advance_pair<KIND, NS>(p, n_use, q, l, aux);
    });
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const float p = path_payoff<KIND>(l[k], aux[k], q[k], a);
      acc[2 * k] += p;
      acc[2 * k + 1] = fmaf(p, p, acc[2 * k + 1]);
    }
  }
}

template <int KIND, int NS, int MINB, int UNROLL = 1, int DEG = kSmallPacked4>
__global__ void __launch_bounds__(kBlock, MINB) pathdep_kernel(const SimArgs a) {
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
  bool small = KIND == B200MC_ASIAN_ARITH && !a.force_mufu_ex2;
  if (KIND == B200MC_ASIAN_ARITH) {
#pragma unroll
    for (int k = 0; k < NS; ++k) small = small && (fabsf(q[k].d) + fabsf(q[k].c) * kRadMax <= kSmallMove);  // NaN -> false
  }
  if (small) simulate_tile<B200MC_ASIAN_ARITH, NS, true, UNROLL, DEG>(a, q, stream, tile_first, acc);  // CTA-uniform branch
  else simulate_tile<KIND, NS, false, UNROLL>(a, q, stream, tile_first, acc);
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
    const double S = params ? params[(size_t)opt * n_scen + k].S : 1.0;  // parity mode folds currency sums
    out[blockIdx.x].sum = s1 * S;
    out[blockIdx.x].sum_sq = s2 * S * S;
    out[blockIdx.x].n = samples;
  }
}

// Control-variate flavour: 5 sums per (option, scenario), all quadratic ones scale with S^2.
__global__ void __launch_bounds__(32) fold_cv_kernel(const double* __restrict__ partials, const b200mc_params_t* __restrict__ params,
                                                     b200mc_cv_moments_t* __restrict__ out, uint32_t n_scen, uint32_t ns_pad,
                                                     uint32_t tiles, double samples) {
  const uint32_t opt = blockIdx.x / n_scen, k = blockIdx.x - opt * n_scen;
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (uint32_t t = threadIdx.x; t < tiles; t += 32) {
    const double* p = partials + ((size_t)opt * tiles + t) * (5 * ns_pad) + 5 * k;
#pragma unroll
    for (int i = 0; i < 5; ++i) s[i] += p[i];
  }
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], off);
  if (threadIdx.x == 0) {
    const double S = params[(size_t)opt * n_scen + k].S;
    b200mc_cv_moments_t m;
    m.sum_payoff = s[0] * S;
    m.sum_payoff_sq = s[1] * S * S;
    m.sum_terminal = s[2] * S;
    m.sum_terminal_sq = s[3] * S * S;
    m.sum_payoff_terminal = s[4] * S * S;
    m.n = samples;
    out[blockIdx.x] = m;
  }
}

// ================================ terminal price arrays (simulation layer) =======================
// out[i] = S_T of path i, out[n_paths + i] = S_T of its mirror (-Z) when anti: what simulate_gbm_numpy returns
// (gbm_numpy.py:46-51).  One parameter set; grid-stride over the paths.
__global__ void __launch_bounds__(kBlock) terminal_prices_kernel(const SimArgs a, int anti, double* __restrict__ out) {
  __shared__ Coef coef;
  __shared__ double spot;
  if (threadIdx.x == 0) {
    coef = make_coef(a.params[0], a.n_steps, 1.0f);
    spot = a.params[0].S;
  }
  __syncthreads();
  const Coef q = coef;
  const double S = spot;
  for (uint64_t local = (uint64_t)blockIdx.x * kBlock + threadIdx.x; local < a.n_paths; local += (uint64_t)gridDim.x * kBlock) {
    const float W = terminal_sum(a.path_begin + local, a.n_steps, a.stream_base, a.rk);
    out[local] = S * (double)mufu_ex2(fmaf(q.c, W, q.a));
    if (anti) out[a.n_paths + local] = S * (double)mufu_ex2(fmaf(-q.c, W, q.a));
  }
}

// ================================ stream inspection kernels ====================================
__global__ void normals_kernel(const PhiloxKeys rk, uint32_t stream, uint64_t path_begin, uint64_t n_paths,
                               uint32_t n_steps, float* __restrict__ out) {
  for (uint64_t local = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; local < n_paths; local += (uint64_t)gridDim.x * blockDim.x) {
    float* row = out + local * n_steps;
    uint32_t s = 0;
    for_each_pair(path_begin + local, n_steps, stream, rk, [&](const NormalPair& p, int n_use) {
      row[s++] = kRadScale * p.rad * p.cs;
      if (n_use > 1) row[s++] = kRadScale * p.rad * p.sn;
    });
  }
}

__global__ void philox_raw_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // the keyed form every simulation kernel calls (round keys expanded once), so the known-answer test checks production code
  const u32x4 x = philox4x32_10(in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3], philox_expand_key(in[6 * i + 4], in[6 * i + 5]));
  out[4 * i] = x.x, out[4 * i + 1] = x.y, out[4 * i + 2] = x.z, out[4 * i + 3] = x.w;
}

}  // namespace b200mc
