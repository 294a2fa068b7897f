// Path-parallel kernels for the reference's other Euler Monte Carlo models: same skeleton
// as mc_kernels.cuh -- Philox words -> Box-Muller pairs in registers, per-path state in registers, payoff fused
// with the (sum, sum^2) reduction, nothing per path or per step in HBM.
//
//   heston_kernel : HestonPricer.price_monte_carlo, full-truncation Euler (src/pricing_models/heston.py:184-255)
//   jump_kernel   : MertonJumpDiffusion.price_monte_carlo (src/pricing_models/jump_diffusion.py:160-225) and
//                   KouJumpDiffusion.price_monte_carlo (:325-377)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_kernels.cuh"

namespace b200mc {

// ============================================ Heston ============================================================
// One Box-Muller pair per step: the cosine branch is Z1 (spot), the sine branch the independent normal that is
// mixed into Z2 = rho*Z1 + sqrt(1-rho^2)*Z (heston.py:228-229).  State: l = log2(S_t/S_0) and the variance v.
// The reference clamps v at the end of every step (heston.py:240), so its max(v, 0) at the start of the next
// step (:232) is the identity for v0 > 0 and is not repeated here.
struct HestonArgs {
  const b200mc_heston_params_t* params;  // [n_opt]
  FoldArgs fold;                         // tile partials -> out[n_opt]
  uint64_t path_begin, n_paths;
  uint32_t n_opt, tiles, paths_per_thread, n_steps;
  PhiloxKeys rk;
  uint32_t stream_base;
  int32_t is_put;  // B200MC_MODEL_* bits
};

// The kernel tracks the SCALED variance u = A^2 * v, A = sqrt(dt) * kRadScale / ln2 (log2-spot diffusion per unit of
// sqrt(v) * rad), so that the step's diffusion magnitude is g = sqrt(rad^2 * u) with no further multiply:
//   l' = l + g*cos - (dt / (2 ln2 A^2)) * u                              (drift (r-q) dt/ln2 added once, at the end)
//   u' = max(u * (1 - kappa dt) + A^2 kappa theta dt + g * (B*rho*cos + B*rho_bar*sin), 0),   B = A * sigma_v * ln2 ... see below
// which is heston.py:229-240 multiplied through by constants: 7 FP32 + 1 FMNMX per step instead of 13.
struct HestonCoef {
  float mu_total;       // (r - q) T / ln2
  float neg_half;       // -(0.5 dt / ln2) / A^2
  float b_rho, b_rho_bar;  // A^2 * (sigma_v * ln2) * {rho, sqrt(1 - rho^2)}: u-diffusion per unit of g, along cos / sin
  float one_minus_kdt, ktheta_dt;  // 1 - kappa dt,  A^2 * kappa * theta * dt
  float u0, kappa_strike;          // A^2 * v0,  K / S
};

__global__ void __launch_bounds__(kBlock, 4) heston_kernel(const HestonArgs a) {
  __shared__ HestonCoef coef_s;
  const uint32_t opt = blockIdx.x / a.tiles;
  const uint32_t tile = blockIdx.x - opt * a.tiles;
  if (threadIdx.x == 0) {
    const b200mc_heston_params_t p = a.params[opt];
    const double inv_ln2 = 1.44269504088896340736;
    const double dt = p.T / (double)a.n_steps;
    const double A = sqrt(dt) * kRadScaleD * inv_ln2, A2 = A * A;
    HestonCoef c;
    c.mu_total = (float)((p.r - p.q) * p.T * inv_ln2);
    c.neg_half = (float)(-0.5 * dt * inv_ln2 / A2);
    c.b_rho = (float)(A2 * (p.sigma_v / inv_ln2) * p.rho);
    c.b_rho_bar = (float)(A2 * (p.sigma_v / inv_ln2) * sqrt(1.0 - p.rho * p.rho));
    c.one_minus_kdt = (float)(1.0 - p.kappa * dt);
    c.ktheta_dt = (float)(A2 * p.kappa * p.theta * dt);
    c.u0 = (float)(A2 * p.v0);
    c.kappa_strike = (float)(p.K / p.S);
    coef_s = c;
  }
  __syncthreads();
  const HestonCoef c = coef_s;
  float acc[2] = {0.0f, 0.0f};
  uint32_t paid[1] = {0u};
  const uint32_t stream = a.stream_base + ((a.is_put & B200MC_MODEL_SHARED_STREAM) ? 0u : opt);  // shared: CRN across the option axis
  const bool is_put = (a.is_put & B200MC_MODEL_PUT) != 0;
  const uint64_t tile_first = (uint64_t)tile * (uint64_t)(kBlock * a.paths_per_thread);
  for (uint32_t j = 0; j < a.paths_per_thread; ++j) {
    const uint64_t local = tile_first + (uint64_t)j * kBlock + threadIdx.x;
    if (local >= a.n_paths) break;
    float l = 0.0f, u = c.u0;
    // n_steps pairs = 2*n_steps draws of the path's stream.  p.rad carries rad^2: sqrt(v) * rad = sqrt(v * rad^2) is
    // ONE MUFU.SQRT (4 MUFU per step instead of 5).
    for_each_pair<1, true>(a.path_begin + local, 2u * a.n_steps, stream, a.rk, [&](const NormalPair& p, int) {
      const float g = mufu_sqrt(p.rad * u);                              // A * sqrt(v) * |draw|, log2 units
      l = fmaf(g, p.cs, fmaf(c.neg_half, u, l));                         // heston.py:236
      const float w = fmaf(c.b_rho, p.cs, c.b_rho_bar * p.sn);           // heston.py:229 (direction of Z2), pre-scaled
      u = fmaxf(fmaf(g, w, fmaf(u, c.one_minus_kdt, c.ktheta_dt)), 0.0f);  // heston.py:239-240
    });
    l += c.mu_total;
    const float e = mufu_ex2(l);
    add_sample<1>(acc, paid, 0, e, vanilla(e, c.kappa_strike, is_put));
  }
  finish_tile<2, 1>(acc, paid, a.fold, opt, tile, a.tiles, 1u, [&](uint32_t) {
    const b200mc_heston_params_t p = a.params[opt];
    ScenScale sc;
    sc.spot = p.S, sc.kappa = p.K / p.S, sc.kappa32 = (float)sc.kappa, sc.has_strike = true, sc.is_put = is_put;
    return sc;
  });
}

// ======================================= Merton / Kou jump diffusion =============================================
// The diffusion is stepped n_steps times exactly as the European kernel does (drift compensated by lambda*kappa,
// jump_diffusion.py:198-199, :346-347).  The jumps of the reference are a compound Poisson process independent of
// the diffusion, added to log S step by step (:209-216, :355-367); only S_T is priced, so their sum over the path
// is drawn once per path with the exact law: N ~ Poisson(lambda*T) by CDF inversion of one uniform, then
//   Merton : sum of N iid N(mu_j, sigma_j^2)  =  N*mu_j + sqrt(N)*sigma_j*Z                      (one normal)
//   Kou    : N draws of  +Exp(eta1) w.p. p,  -Exp(eta2) w.p. 1-p   (jump_diffusion.py:310-323)   (2 uniforms each)
// Jump draws come from the same (seed, stream, path) Philox stream at call indices >= 2^31, which the diffusion
// (call indices < n_steps/8) never reaches.
struct JumpArgs {
  const b200mc_params_t* params;       // [n_opt] (one scenario)
  const b200mc_jump_params_t* jumps;   // [n_opt]
  FoldArgs fold;                       // tile partials -> out[n_opt]
  uint64_t path_begin, n_paths;
  uint32_t n_opt, tiles, paths_per_thread, n_steps;
  PhiloxKeys rk;
  uint32_t stream_base;
  int32_t is_put;  // B200MC_MODEL_* bits
};

struct JumpCoef {
  float c, a, kappa_strike;  // as Coef (log2 units), drift compensated
  double lam_T, p0;          // Poisson mean over the path and exp(-lam_T) (the count is inverted in FP64: a
                             // 32-bit uniform resolves 2^-33, an FP32 running CDF would saturate below it)
  float j1, j2, j3;          // Merton: mu_j/ln2, sigma_j*kRadScale/ln2, -;  Kou: p, 1/(eta1*ln2)... see below
  int32_t model;
};

constexpr uint32_t kJumpCallBase = 0x80000000u;

__device__ __forceinline__ float uniform_open_closed(uint32_t w) {  // (0, 1], 24 significant bits
  return ((float)(w >> 8) + 1.0f) * 5.9604644775390625e-08f;
}

__global__ void __launch_bounds__(kBlock, 4) jump_kernel(const JumpArgs a) {
  __shared__ JumpCoef coef_s;
  const uint32_t opt = blockIdx.x / a.tiles;
  const uint32_t tile = blockIdx.x - opt * a.tiles;
  if (threadIdx.x == 0) {
    const b200mc_params_t p = a.params[opt];
    const b200mc_jump_params_t jp = a.jumps[opt];
    const double inv_ln2 = 1.44269504088896340736;
    double kappa_j;  // E[e^Y - 1]
    if (jp.model == B200MC_JUMP_MERTON) kappa_j = exp(jp.a + 0.5 * jp.b * jp.b) - 1.0;            // jump_diffusion.py:65-67
    else kappa_j = jp.a * jp.b / (jp.b - 1.0) + (1.0 - jp.a) * jp.c / (jp.c + 1.0) - 1.0;         // jump_diffusion.py:302-308
    const double dt = p.T / (double)a.n_steps;
    JumpCoef c;
    c.c = (float)(p.sigma * sqrt(dt) * kCoefScaleD);
    c.a = (float)((p.r - p.q - jp.lambda_j * kappa_j - 0.5 * p.sigma * p.sigma) * dt * inv_ln2 * (double)a.n_steps);
    c.kappa_strike = (float)(p.K / p.S);
    c.lam_T = jp.lambda_j * p.T;
    c.p0 = exp(-jp.lambda_j * p.T);
    c.model = jp.model;
    if (jp.model == B200MC_JUMP_MERTON) {
      c.j1 = (float)(jp.a * inv_ln2);                 // mu_j in log2 units
      c.j2 = (float)(jp.b * kRadScaleD * inv_ln2);    // sigma_j per unit of log2-radius draw
      c.j3 = 0.0f;
    } else {
      c.j1 = (float)jp.a;                             // p (up probability)
      c.j2 = (float)(1.0 / jp.b);                     // mean up jump 1/eta1, applied to -log2(U) * ln2 / ln2 = -log2(U)/eta1 in log2 units
      c.j3 = (float)(1.0 / jp.c);                     // mean down jump 1/eta2
    }
    coef_s = c;
  }
  __syncthreads();
  const JumpCoef c = coef_s;
  float acc[2] = {0.0f, 0.0f};
  uint32_t paid[1] = {0u};
  const uint32_t stream = a.stream_base + ((a.is_put & B200MC_MODEL_SHARED_STREAM) ? 0u : opt);  // shared: CRN across the option axis
  const bool is_put = (a.is_put & B200MC_MODEL_PUT) != 0;
  const uint64_t tile_first = (uint64_t)tile * (uint64_t)(kBlock * a.paths_per_thread);
  for (uint32_t j = 0; j < a.paths_per_thread; ++j) {
    const uint64_t local = tile_first + (uint64_t)j * kBlock + threadIdx.x;
    if (local >= a.n_paths) break;
    const uint64_t path = a.path_begin + local;
    const float W = terminal_sum(path, a.n_steps, stream, a.rk);
    float l = fmaf(c.c, W, c.a);
    // --- compound Poisson part: one call for (count uniform, Merton normal) -------------------------------------
    const u32x4 x = draw4(path, kJumpCallBase, stream, a.rk);
    const double u = ((double)x.x + 0.5) * 2.3283064365386963e-10;  // (0, 1), 32 bits
    uint32_t n_jumps = 0;
    if (u > c.p0) {  // CDF inversion of Poisson(lam_T); not taken by exp(-lam_T) of the paths
      double pk = c.p0, cdf = c.p0;
      do {
        n_jumps += 1;
        pk *= c.lam_T / (double)n_jumps;
        cdf += pk;
      } while (u > cdf && n_jumps < 4096u);
    }
    if (n_jumps > 0) {
      if (c.model == B200MC_JUMP_MERTON) {
        const NormalPair z = box_muller(x.y);
        l += fmaf(c.j2 * sqrtf((float)n_jumps), z.rad * z.cs, c.j1 * (float)n_jumps);
      } else {
        float jump_log2 = 0.0f;
        for (uint32_t k = 0; k < n_jumps; k += 2) {  // two jumps per Philox call: (direction, magnitude) x 2
          const u32x4 y = draw4(path, kJumpCallBase + 1u + (k >> 1), stream, a.rk);
          const float m0 = -mufu_lg2(uniform_open_closed(y.y));  // Exp(1) / ln2
          jump_log2 += (uniform_open_closed(y.x) <= c.j1) ? m0 * c.j2 : -m0 * c.j3;
          if (k + 1 < n_jumps) {
            const float m1 = -mufu_lg2(uniform_open_closed(y.w));
            jump_log2 += (uniform_open_closed(y.z) <= c.j1) ? m1 * c.j2 : -m1 * c.j3;
          }
        }
        l += jump_log2;  // -log2(U)/eta is the exponential jump already expressed in log2 units of S
      }
    }
    const float e = mufu_ex2(l);
    add_sample<1>(acc, paid, 0, e, vanilla(e, c.kappa_strike, is_put));
  }
  finish_tile<2, 1>(acc, paid, a.fold, opt, tile, a.tiles, 1u, [&](uint32_t) { return vanilla_scale(a.params[opt], is_put); });
}

// ================================ FP64 parity kernels (caller-supplied draws) =====================================
// Heston: Z is STEP-major [n_steps][2][n_paths] -- the order in which the reference consumes its generator
// (heston.py:228-229: standard_normal(n_paths) twice per step) -- so thread = path reads are coalesced.
struct HestonF64Args {
  const double* Z;
  double* payoffs;   // [n_paths], may be null
  double* partials;  // [gridDim.x][2]
  uint64_t n_paths;
  uint32_t n_steps;
  int32_t is_put;
  double S, K, T, r, q, kappa, theta, sigma_v, rho, v0;
};

__device__ __forceinline__ void block_reduce_pair_f64(double s1, double s2, double* dst) {
  __shared__ double red[kBlock / 32][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  }
  if (lane == 0) red[warp][0] = s1, red[warp][1] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t1 += red[w][0], t2 += red[w][1];
    dst[0] = t1, dst[1] = t2;
  }
}

__global__ void __launch_bounds__(kBlock) heston_from_normals_kernel(const HestonF64Args a) {
  const uint64_t path = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
  double pay = 0.0;
  if (path < a.n_paths) {
    // constants and update order exactly as heston.py:213-240 evaluates them (no FMA contraction)
    const double dt = __ddiv_rn(a.T, (double)a.n_steps);
    const double sqrt_dt = sqrt(dt);
    const double rho_sqrt = sqrt(__dsub_rn(1.0, __dmul_rn(a.rho, a.rho)));
    double log_S = log(a.S), v = a.v0;
    for (uint32_t t = 0; t < a.n_steps; ++t) {
      const double Z1 = __ldg(a.Z + ((size_t)t * 2) * a.n_paths + path);
      const double Zx = __ldg(a.Z + ((size_t)t * 2 + 1) * a.n_paths + path);
      const double Z2 = __dadd_rn(__dmul_rn(a.rho, Z1), __dmul_rn(rho_sqrt, Zx));
      const double v_pos = fmax(v, 0.0);
      const double sqrt_v = sqrt(v_pos);
      // log_S += (r - q - 0.5 * v_pos) * dt + sqrt_v * sqrt_dt * Z1
      const double drift = __dmul_rn(__dsub_rn(__dsub_rn(a.r, a.q), __dmul_rn(0.5, v_pos)), dt);
      log_S = __dadd_rn(log_S, __dadd_rn(drift, __dmul_rn(__dmul_rn(sqrt_v, sqrt_dt), Z1)));
      // v += kappa * (theta - v_pos) * dt + sigma_v * sqrt_v * sqrt_dt * Z2 ; v = max(v, 0)
      const double mean_rev = __dmul_rn(__dmul_rn(a.kappa, __dsub_rn(a.theta, v_pos)), dt);
      v = __dadd_rn(v, __dadd_rn(mean_rev, __dmul_rn(__dmul_rn(__dmul_rn(a.sigma_v, sqrt_v), sqrt_dt), Z2)));
      v = fmax(v, 0.0);
    }
    const double s_T = exp(log_S);
    pay = a.is_put ? fmax(__dsub_rn(a.K, s_T), 0.0) : fmax(__dsub_rn(s_T, a.K), 0.0);
    if (a.payoffs) a.payoffs[path] = pay;
  }
  block_reduce_pair_f64(pay, pay * pay, a.partials + 2 * (size_t)blockIdx.x);
}

// Jump diffusion: dW is step-major [n_steps][n_paths]; J (may be null) holds the sum of the jump sizes that hit
// path i during step t, [n_steps][n_paths] -- what the reference adds at jump_diffusion.py:216 / :366.
struct JumpF64Args {
  const double* dW;
  const double* J;
  double* payoffs;
  double* partials;
  uint64_t n_paths;
  uint32_t n_steps;
  int32_t is_put;
  double S, K, T, r, sigma, q, lambda_kappa;  // lambda_j * kappa, the compensator (jump_diffusion.py:198,346)
};

__global__ void __launch_bounds__(kBlock) jump_from_draws_kernel(const JumpF64Args a) {
  const uint64_t path = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
  double pay = 0.0;
  if (path < a.n_paths) {
    const double dt = __ddiv_rn(a.T, (double)a.n_steps);
    // drift = (r - q - lambda_j * kappa - 0.5 * sigma**2) * dt ; vol = sigma * sqrt(dt)
    const double drift = __dmul_rn(__dsub_rn(__dsub_rn(__dsub_rn(a.r, a.q), a.lambda_kappa), __dmul_rn(0.5, __dmul_rn(a.sigma, a.sigma))), dt);
    const double vol = __dmul_rn(a.sigma, sqrt(dt));
    double log_S = log(a.S);
    for (uint32_t t = 0; t < a.n_steps; ++t) {
      const double z = __ldg(a.dW + (size_t)t * a.n_paths + path);
      log_S = __dadd_rn(log_S, __dadd_rn(drift, __dmul_rn(vol, z)));
      if (a.J) {
        const double jump = __ldg(a.J + (size_t)t * a.n_paths + path);
        if (jump != 0.0) log_S = __dadd_rn(log_S, jump);
      }
    }
    const double s_T = exp(log_S);
    pay = a.is_put ? fmax(__dsub_rn(a.K, s_T), 0.0) : fmax(__dsub_rn(s_T, a.K), 0.0);
    if (a.payoffs) a.payoffs[path] = pay;
  }
  block_reduce_pair_f64(pay, pay * pay, a.partials + 2 * (size_t)blockIdx.x);
}

}  // namespace b200mc
