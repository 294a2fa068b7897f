// Structured products of src/pricing_models/exotic_options.py:404-552 on the fused path: cliquet (sum of locally
// clipped period returns) and autocallable (early redemption on observation dates, coupon, knock-in put).
//
// Same skeleton as pathdep_kernel (mc_kernels.cuh): one thread = one path at a time, Philox normals in registers,
// l_t = log2(S_t/S_0) per scenario in FP32, scenarios on common random numbers, FP64 block reduction.  The period /
// observation schedule is a count-down register shared by all scenarios of the thread; it is the same for every
// thread of the launch, so the event branch is warp-uniform.
//
// Output conventions (include/b200mc.h): CLIQUET accumulates the undiscounted payoff per unit spot (the fold
// multiplies by S); AUTOCALLABLE accumulates the DISCOUNTED payoff per unit notional (folded with S = 1).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "mc_kernels.cuh"

namespace b200mc {

struct StructuredArgs {
  SimArgs sim;          // params, partials, path range, tiles, n_steps (TOTAL steps: dt = T / n_steps), keys
  double a, b, c, d;    // b200mc_product_t
  uint32_t period;      // CLIQUET: steps per period (n_steps / n_periods);  AUTOCALLABLE: observation_freq
  uint32_t n_events;    // CLIQUET: n_periods;  AUTOCALLABLE: number of observations = n_steps / observation_freq
  uint32_t sim_steps;   // steps that can influence the payoff (CLIQUET: period * n_events, AUTOCALLABLE: n_steps)
};

// Per-scenario FP32 constants beyond Coef.
struct StructCoef {
  float c, d;           // log2-diffusion per unit rad, log2-drift per step (as Coef)
  float rT_log2;        // r * T / ln2: exp(-r t_i) = 2^(-rT_log2 * steps_i / n_steps)
  float disc_T;         // exp(-r T)
  float coupon_T;       // coupon_rate * T
};

template <int KIND, int NS, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) structured_kernel(const StructuredArgs g) {
  const SimArgs& a = g.sim;
  __shared__ StructCoef coef_s[NS];
  const uint32_t opt = blockIdx.x / a.tiles;
  const uint32_t tile = blockIdx.x - opt * a.tiles;
  if (threadIdx.x < NS) {
    const uint32_t k = threadIdx.x < a.n_scen ? threadIdx.x : a.n_scen - 1;
    const b200mc_params_t p = a.params[(size_t)opt * a.n_scen + k];
    const Coef base = make_coef(p, a.n_steps, 1.0f);
    StructCoef s;
    s.c = base.c, s.d = base.d;
    s.rT_log2 = (float)(p.r * p.T * 1.44269504088896340736);
    s.disc_T = (float)exp(-p.r * p.T);
    s.coupon_T = (float)(g.c * p.T);
    coef_s[threadIdx.x] = s;
  }
  __syncthreads();
  StructCoef q[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) q[k] = coef_s[k];

  // product terms in the units of the state
  const float local_cap = (float)g.a, local_floor = (float)g.b, global_cap = (float)g.c, global_floor = (float)g.d;
  const float l_autocall = g.a > 0.0 ? (float)log2(g.a) : -CUDART_INF_F;  // S_t/S >= barrier  <=>  l_t >= log2(barrier)
  const float l_coupon = g.b > 0.0 ? (float)log2(g.b) : -CUDART_INF_F;
  const float l_knock_in = g.d > 0.0 ? (float)log2(g.d) : -CUDART_INF_F;  // min_t l_t <= log2(ki); never for ki <= 0
  const float inv_events = g.n_events ? 1.0f / (float)g.n_events : 0.0f;
  const float step_frac = (float)g.period / (float)a.n_steps;             // t_i / T per observation

  float acc[2 * NS];
  uint32_t paid[(NS + 3) / 4];
#pragma unroll
  for (int i = 0; i < 2 * NS; ++i) acc[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < (NS + 3) / 4; ++i) paid[i] = 0u;

  const uint32_t stream = a.stream_base + opt;
  const uint64_t tile_first = (uint64_t)tile * (uint64_t)(kBlock * a.paths_per_thread);
  for (uint32_t j = 0; j < a.paths_per_thread; ++j) {
    const uint64_t local = tile_first + (uint64_t)j * kBlock + threadIdx.x;
    if (local >= a.n_paths) break;
    float l[NS], aux[NS], pay[NS];  // CLIQUET: aux = l at the period start, pay = running total.  AUTOCALLABLE: aux = min l, pay = redemption
    bool done[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) l[k] = 0.0f, aux[k] = 0.0f, pay[k] = 0.0f, done[k] = false;
    uint32_t left = g.period, event = 0;

    auto after_step = [&]() {
      if (KIND == B200MC_AUTOCALLABLE) {
#pragma unroll
        for (int k = 0; k < NS; ++k) aux[k] = fminf(aux[k], l[k]);
      }
      if (--left == 0) {  // same schedule for every thread: warp-uniform
        left = g.period;
        event += 1;
        if (event <= g.n_events) {
#pragma unroll
          for (int k = 0; k < NS; ++k) {
            if (KIND == B200MC_CLIQUET) {  // exotic_options.py:538-546
              const float ret = mufu_ex2(l[k] - aux[k]) - 1.0f;
              pay[k] += fminf(fmaxf(ret, local_floor), local_cap);
              aux[k] = l[k];
            } else if (!done[k] && l[k] >= l_autocall) {  // exotic_options.py:454-467
              const float coupon = q[k].coupon_T * ((float)event * inv_events);
              pay[k] = (1.0f + coupon) * mufu_ex2(-q[k].rT_log2 * ((float)event * step_frac));
              done[k] = true;
            }
          }
        }
      }
    };

    for_each_pair(a.path_begin + local, g.sim_steps, stream, a.rk, [&](const NormalPair& p, int n_use) {
#pragma unroll
      for (int k = 0; k < NS; ++k) l[k] += fmaf(p.rad * q[k].c, p.cs, q[k].d);
      after_step();
      if (n_use > 1) {
#pragma unroll
        for (int k = 0; k < NS; ++k) l[k] += fmaf(p.rad * q[k].c, p.sn, q[k].d);
        after_step();
      }
    });

#pragma unroll
    for (int k = 0; k < NS; ++k) {
      float p;
      if (KIND == B200MC_CLIQUET) {  // exotic_options.py:548-552
        p = fmaxf(fminf(fmaxf(pay[k], global_floor), global_cap), 0.0f);
      } else if (done[k]) {
        p = pay[k];
      } else {                        // exotic_options.py:469-486
        float f = 1.0f;
        if (l[k] >= l_coupon) f += q[k].coupon_T;
        if (aux[k] <= l_knock_in && l[k] < 0.0f) f = mufu_ex2(l[k]);
        p = f * q[k].disc_T;
      }
      add_sample<NS>(acc + 2 * k, paid, k, p, p);
    }
  }
  // neither product has a strike term; CLIQUET payoffs are per unit spot, AUTOCALLABLE per unit notional
  finish_tile<2, NS>(acc, paid, a.fold, opt, tile, a.tiles, a.n_scen, [&](uint32_t k) {
    ScenScale sc;
    sc.spot = KIND == B200MC_CLIQUET ? a.params[(size_t)opt * a.n_scen + k].S : 1.0;
    sc.kappa = 0.0, sc.kappa32 = 0.0f, sc.has_strike = false, sc.is_put = false;
    return sc;
  });
}

// ---- FP64 on caller-supplied draws: the reference's statements, one thread per path -----------------------------
struct StructuredF64Args {
  const double* Z;    // [n_paths][n_steps]
  double* payoffs;    // [n_paths]; may be null
  double* partials;   // [gridDim.x][2]
  uint64_t n_paths;
  uint32_t n_steps, period, n_events;
  double S, T, r, sigma, q;
  double a, b, c, d;
};

template <int KIND>
__global__ void __launch_bounds__(128) structured_from_normals_kernel(const StructuredF64Args g) {
  __shared__ double red[4][2];
  const uint64_t path = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = g.n_steps;
  // exotic_options.py:54-56
  const double dt = __ddiv_rn(g.T, (double)n);
  const double drift = __dmul_rn(__dsub_rn(__dsub_rn(g.r, g.q), __dmul_rn(__dmul_rn(0.5, g.sigma), g.sigma)), dt);
  const double diffusion = __dmul_rn(g.sigma, sqrt(dt));
  const double log_S0 = log(g.S);
  double p = 0.0;
  if (path < g.n_paths) {
    const double* z = g.Z + path * n;
    double cum = 0.0;
    const double s0 = exp(log_S0);  // column 0 of the path array (exotic_options.py:64,67)
    double s_t = s0;
    uint32_t left = g.period, event = 0;
    if (KIND == B200MC_CLIQUET) {
      double s_start = s0, total = 0.0;
      for (uint32_t t = 0; t < n; ++t) {
        cum = __dadd_rn(cum, __dadd_rn(drift, __dmul_rn(diffusion, z[t])));
        if (--left == 0) {
          left = g.period;
          event += 1;
          if (event <= g.n_events) {
            s_t = exp(__dadd_rn(log_S0, cum));
            const double ret = __ddiv_rn(__dsub_rn(s_t, s_start), s_start);  // exotic_options.py:542
            total = __dadd_rn(total, fmin(fmax(ret, g.b), g.a));            // np.clip(x, floor, cap)
            s_start = s_t;
          }
        }
      }
      total = fmin(fmax(total, g.d), g.c);
      p = __dmul_rn(fmax(total, 0.0), g.S);
    } else {
      double min_rel = __ddiv_rn(s0, g.S);
      bool redeemed = false;
      for (uint32_t t = 0; t < n; ++t) {
        cum = __dadd_rn(cum, __dadd_rn(drift, __dmul_rn(diffusion, z[t])));
        s_t = exp(__dadd_rn(log_S0, cum));
        const double rel = __ddiv_rn(s_t, g.S);
        min_rel = fmin(min_rel, rel);
        if (--left == 0) {
          left = g.period;
          event += 1;
          if (event <= g.n_events && !redeemed && rel >= g.a) {
            const double coupon = __dmul_rn(__dmul_rn(g.c, __ddiv_rn((double)event, (double)g.n_events)), g.T);  // :463-464
            const double disc = exp(__dmul_rn(__dmul_rn(-g.r, (double)(event * g.period)), dt));                 // :466
            p = __dmul_rn(__dadd_rn(1.0, coupon), disc);
            redeemed = true;
          }
        }
      }
      if (!redeemed) {
        const double rel = __ddiv_rn(s_t, g.S);
        double f = 1.0;
        if (rel >= g.b) f = __dadd_rn(f, __dmul_rn(g.c, g.T));
        if (min_rel <= g.d && rel < 1.0) f = rel;
        p = __dmul_rn(f, exp(__dmul_rn(-g.r, g.T)));
      }
    }
    if (g.payoffs) g.payoffs[path] = p;
  }
  double s1 = p, s2 = p * p;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp][0] = s1, red[warp][1] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) t1 += red[w][0], t2 += red[w][1];
    g.partials[2 * (size_t)blockIdx.x] = t1;
    g.partials[2 * (size_t)blockIdx.x + 1] = t2;
  }
}

}  // namespace b200mc
