// Fused simulation kernels: Philox -> normals -> log-Euler GBM -> payoff -> (sum, sum^2) -> moments, ONE launch.
//
// Nothing per path or per step ever touches HBM: the only global traffic is the per-option
// parameter block read once per CTA (64 B x n_scen; single-option launches carry it in the kernel
// arguments instead) and one FP64 partial per (tile, scenario, moment) written at the end.  A "tile"
// is one CTA's share of one option's paths: kBlock threads x paths_per_thread paths.  The CTA that
// arrives LAST at an option's ticket counter folds that option's tile partials in a fixed order
// (finish_tile below) - no second launch, no atomics on the sums, same seed => bit-identical moments,
// as the reference guarantees (tests/test_monte_carlo.py:153-158).
//
// All per-path arithmetic is FP32 on the quantity  l_t = log2(S_t / S_0)  (small magnitude, so
// FP32 rounding is ~3e-8 per step), payoffs are normalised by S_0 and re-scaled in FP64 by the
// fold; cross-path accumulation is FP64 from the warp level up.
//
// Strike precision.  Scenarios that differ only in the spot (the S +- h re-pricings of delta_gamma,
// monte_carlo_unified.py:513-560 with its default h = 1e-4) share every draw AND every normalised terminal
// value e = S_T/S_0; the bump reaches the payoff max(e - K/S, 0) only through the strike ratio.  An FP32
// K/S resolves a 2e-6 bump to ~17 ulps, and the FP32 difference e - K/S itself rounds in a K/S-dependent way
// for large e - systematic errors no number of paths averages out.  So the strike never enters the FP32
// first-moment arithmetic at all: per scenario a thread accumulates  sum of e over the samples that PAID
// (payoff > 0, decided against the FP32 ratio) and counts them; the fold forms
//     sum(payoff) = +-(sum_paid(e) - (K/S in FP64) * paid)
// which is the exact piecewise-linear function of K/S for the draws at hand (the paid set is off only for paths
// with e between the FP32 and FP64 ratio: measure ~1e-8, error < 6e-8 each).  sum_paid(e) is the SAME number for
// every spot-bumped scenario, so its FP32 rounding cancels in finite differences.  The second moment keeps the
// well-conditioned FP32 sum of payoff^2 and is moved to the FP64 strike to first order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200mc.h"
#include "normal.cuh"
#include "philox.cuh"

namespace b200mc {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr int kMaxPathsPerThread = 32;  // <= 64 samples per thread with mirrors: the paid-counts fit 8-bit lanes
constexpr int kMaxSplit = 8;            // lanes that may share one European path (see european_kernel)

// Fused all-reduce of a launch's moment records across the GPUs of one NVSwitch domain (world == 0: off).  Every rank
// owns an exchange block in its HBM that its peers map (CUDA IPC between processes, peer access inside one):
//   [2 slots] x { uint64 epoch flag | pad to 128 B | records[n_opt * n_scen] }
// The CTA that finishes a rank's LAST option publishes the rank's records (slot = epoch & 1) with a system-scope release,
// waits for every peer's flag of the same epoch, adds the peers' records over NVLink in rank order - the same order on
// every rank, so all ranks end with identical bits - and writes the totals to `out`.  No second kernel, no host hop:
// the exchange is the tail of the simulation kernel.  Two slots suffice: a rank can start epoch e+1 only after every peer
// has published e, i.e. after every peer finished reading slot (e-1) & 1.
constexpr int kMaxRanks = 8;
constexpr size_t kXchgHeaderBytes = 128;
struct XchgArgs {
  uint32_t world, rank;
  unsigned long long epoch;
  char* peer[kMaxRanks];      // exchange blocks by rank; peer[rank] is this rank's own
  size_t slot_bytes;
  uint32_t* launch_ticket;    // options of this launch whose records are complete; zero between launches
  unsigned int* timed_out;    // mapped host word, set to 1 if a peer's flag did not arrive within timeout_ns
  unsigned long long timeout_ns;
};
constexpr unsigned long long kXchgTimeoutNs = 20ull * 1000 * 1000 * 1000;  // default bound on the wait for a peer

// Where a launch's results go.
struct FoldArgs {
  double* partials;          // [n_opt * tiles][values per tile]: per-CTA FP64 partial sums (stay in L2)
  uint32_t* tickets;         // [n_opt]: CTAs of the option that have delivered; zero between launches
  void* out;                 // [n_opt][n_scen] moment records - device memory, or the device alias of mapped pinned host memory
  unsigned long long* done;  // non-null: mapped host word that receives `seq` once `out` is complete (single-option launches)
  unsigned long long seq;
  double samples;            // the n of every record
  XchgArgs x;                // x.world > 1: `out` receives the sum over all ranks
  uint32_t n_opt;            // options of this launch (the exchange starts when all of them are folded)
};

struct SimArgs {
  const b200mc_params_t* params;  // [n_opt][n_scen]; null: single-option launch, coefficients precomputed by the host (inl)
  FoldArgs fold;
  uint64_t path_begin;
  uint64_t n_paths;
  uint32_t n_opt, n_scen;
  uint32_t tiles;                 // tiles per option
  uint32_t paths_per_thread;
  uint32_t n_steps;
  uint32_t split_shift;           // European: 2^split_shift adjacent lanes share one path's Philox calls
  PhiloxKeys rk;                  // expanded from the 64-bit seed on the host
  uint32_t stream_base;
  int32_t is_put, barrier_in, lookback_fixed, sgn_negative;
  int32_t force_mufu_ex2;         // arithmetic Asian: always take the MUFU.EX2 form (tests / A-B timing)
  // params == null (latency path of the host entry points, n_opt == 1): the host evaluated make_coef / the fold scale
  // with the same FP64 expressions and passes the results in the kernel arguments - no H2D copy, no FP64 prologue
  struct Inline {
    float c[B200MC_MAX_SCENARIOS], d[B200MC_MAX_SCENARIOS], a[B200MC_MAX_SCENARIOS], kappa[B200MC_MAX_SCENARIOS],
        beta[B200MC_MAX_SCENARIOS], inv_n[B200MC_MAX_SCENARIOS];
    double spot[B200MC_MAX_SCENARIOS], kappa64[B200MC_MAX_SCENARIOS];
  } inl;
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

// Written with explicitly rounded operations so that host (engine.cu, single-option launches) and device
// (batched launches) produce the same bits: nvcc would otherwise contract a*b+c into an FMA on the device only.
#if defined(__CUDA_ARCH__)
#define B200MC_DMUL(x, y) __dmul_rn((x), (y))
#define B200MC_DSUB(x, y) __dsub_rn((x), (y))
#define B200MC_DDIV(x, y) __ddiv_rn((x), (y))
#else
#define B200MC_DMUL(x, y) ((x) * (y))
#define B200MC_DSUB(x, y) ((x) - (y))
#define B200MC_DDIV(x, y) ((x) / (y))
#endif

__host__ __device__ __forceinline__ Coef make_coef(const b200mc_params_t& p, uint32_t n_steps, float sgn) {
  const double inv_ln2 = 1.44269504088896340736;
  const double dt = B200MC_DDIV(p.T, (double)n_steps);
  const double mu = B200MC_DSUB(B200MC_DSUB(p.r, p.q), B200MC_DMUL(B200MC_DMUL(0.5, p.sigma), p.sigma));
  const double d = B200MC_DMUL(B200MC_DMUL(mu, dt), inv_ln2);
  const double c = B200MC_DMUL(B200MC_DMUL(p.sigma, sqrt(dt)), kCoefScaleD);
  Coef k;
  k.c = sgn * (float)c;
  k.d = sgn * (float)d;
  k.a = (float)B200MC_DMUL(d, (double)n_steps);
  k.kappa = (float)B200MC_DDIV(p.K, p.S);
  k.beta = sgn * (float)(log2(B200MC_DDIV(p.barrier, p.S)));
  k.inv_n = (float)B200MC_DDIV(1.0, (double)n_steps);
  return k;
}

// Scenario k of option opt: from the parameter block in HBM, or from the host-evaluated inline block.
__device__ __forceinline__ Coef scen_coef(const SimArgs& a, uint32_t opt, uint32_t k, float sgn) {
  if (a.params) return make_coef(a.params[(size_t)opt * a.n_scen + k], a.n_steps, sgn);
  Coef q;
  q.c = a.inl.c[k], q.d = a.inl.d[k], q.a = a.inl.a[k], q.kappa = a.inl.kappa[k], q.beta = a.inl.beta[k], q.inv_n = a.inl.inv_n[k];
  return q;
}

// max(+-(e - kappa), 0) with the option type as DATA (a sign), not as a branch: is_put is a launch argument, and written as
// is_put ? max(kappa - e, 0) : max(e - kappa, 0) the compiler evaluates both sides and selects - 2 FADD + 2 FMNMX + a select
// per sample, a quarter of the multi-scenario epilogue.  fma(+-1, e, -+kappa) rounds once, exactly like the subtraction.
__device__ __forceinline__ float vanilla(float e, float kappa, bool is_put) {
  const float s = is_put ? -1.0f : 1.0f;
  return fmaxf(fmaf(s, e, -s * kappa), 0.0f);
}

// ---- per-sample accumulation ------------------------------------------------------------------------
// x: the quantity the strike ratio is compared with (S_T/S_0, the average, the extremum ...); p: the FP32 payoff
// max(+-(x - k32), 0) (0 when a barrier switched the path off).  acc[0] += x over the paid samples, acc[1] += p^2,
// and the paid-count of scenario k sits in an 8-bit lane (a thread sees <= 64 samples: kMaxPathsPerThread x mirrors).
// Payoffs without a strike (floating lookback, cliquet, autocallable) pass x = p: acc[0] is then sum(p) itself.
template <int NS>
__device__ __forceinline__ void add_sample(float* acc2, uint32_t (&cnt)[(NS + 3) / 4], int k, float x, float p) {
  const bool paid = p > 0.0f;
  acc2[0] += paid ? x : 0.0f;
  acc2[1] = fmaf(p, p, acc2[1]);
  cnt[k >> 2] += paid ? (1u << (8 * (k & 3))) : 0u;
}

// ---- CTA partials -> fixed-order two-level fold -> moment records ------------------------------------
// What the fold needs to know about scenario k, evaluated in FP64 by the finishing CTA only.
struct ScenScale {
  double spot;     // currency units per normalised payoff unit (1 for per-notional payoffs)
  double kappa;    // FP64 strike ratio K/S; has_strike = false: the payoff has no strike term
  float kappa32;   // the FP32 value the paid / unpaid decision was made against
  bool has_strike, is_put;
};

constexpr int kFoldGroup = 1024;  // tiles per first-level fold: four per thread of the folding CTA

// NM sums per scenario: 2 = (sum_paid x, sum p^2); 5 adds (sum S_T, sum S_T^2, sum_paid x^2) for the control variate.
// Every CTA: FP32 thread sums -> FP64 warp shuffles -> FP64 CTA partial (fixed order), written to the option's scratch.
// Tiles of an option form groups of 1024; whichever CTA takes a group's last ticket folds the group with every load in
// flight at once (one L2 round trip), and - only for options of more than 1024 tiles - whichever group finishes last
// folds the group totals; that CTA also writes the option's moment records.  Every level reads ALL its inputs in index
// order with a fixed association, so the result does not depend on which CTA did the work: no atomics on the sums, same
// launch => bit-identical moments, no second kernel.
// Scratch per option (doubles): [tiles][NV] tile partials, then [n_groups][NV] group totals.
// Tickets per option (uint32): [0] groups folded, [1 + g] tiles delivered in group g; all zero between launches.
__host__ __device__ inline uint32_t fold_groups(uint32_t tiles) { return (tiles + kFoldGroup - 1) / kFoldGroup; }
__host__ __device__ inline size_t fold_scratch_doubles(uint32_t tiles, uint32_t nv) { return ((size_t)tiles + fold_groups(tiles)) * nv; }
__host__ __device__ inline size_t fold_ticket_words(uint32_t tiles) { return (size_t)fold_groups(tiles) + 1; }

// Sum x[v] over the CTA in a fixed order: a butterfly per warp, then the warps in index order.  Threads v < NV return
// the total of value v (others: garbage).  `red` is CTA-shared scratch.
// Few values: one xor-butterfly per value (5 double shuffles each).  Many values (4-16 scenarios: 12..96 values): the
// butterfly runs on ALL values at once by recursive halving - at offset 16 a lane keeps one half of the values and hands
// the other half to its partner, at offset 8 a quarter, ... - so a warp spends ~P double shuffles for P values instead
// of 5P (the 14-scenario Greeks launch was SHFL-bound in its reduction: 480 -> 124 SHFL per warp).  After the five
// rounds lane L holds the warp totals of values L*P/32 .. L*P/32 + P/32 - 1.  Same association for every value.
template <int NV>
__device__ __forceinline__ double cta_sum(const double (&x)[NV], double (*red)[NV]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if constexpr (NV <= 8) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double y = x[v];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) y += __shfl_xor_sync(0xffffffffu, y, off);
      if (lane == 0) red[warp][v] = y;
    }
  } else {
    constexpr int P = NV <= 32 ? 32 : NV <= 64 ? 64 : 128;
    double y[P];
#pragma unroll
    for (int i = 0; i < P; ++i) y[i] = i < NV ? x[i] : 0.0;
#pragma unroll
    for (int round = 0; round < 5; ++round) {
      const int off = 16 >> round;
      const int h = P >> (round + 1);
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < h; ++i) {
        const double keep = upper ? y[i + h] : y[i];
        const double send = upper ? y[i] : y[i + h];
        y[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
#pragma unroll
    for (int j = 0; j < P / 32; ++j) {
      const int v = lane * (P / 32) + j;
      if (v < NV) red[warp][v] = y[j];
    }
  }
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < NV) {
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += red[w][threadIdx.x];
  }
  return t;
}

// Fold `count` (<= 1024) rows of NV doubles starting at `rows` into dst[NV] (CTA-shared or global), fixed association.
template <int NV, class Store>
__device__ __forceinline__ void fold_rows(const double* rows, uint32_t count, double (*red)[NV], Store&& store) {
  if constexpr (NV <= 8) {
    // few values per row: thread t takes rows t, t+256, t+512, t+768 - up to 4*NV independent loads - then a CTA sum
    double x[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double part[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t r = threadIdx.x + (uint32_t)kBlock * i;
        part[i] = r < count ? __ldcg(rows + (size_t)r * NV + v) : 0.0;
      }
      x[v] = (part[0] + part[1]) + (part[2] + part[3]);
    }
    __syncthreads();  // `red` may still be read by the caller's previous cta_sum
    const double t = cta_sum<NV>(x, red);
    if (threadIdx.x < NV) store(threadIdx.x, t);
  } else {
    // many values per row (4-16 scenarios): eight values per pass, thread t again takes rows t, t+256, t+512, t+768
    // (32 independent loads), then one CTA sum per pass
    static_assert(NV % 4 == 0, "rows of 3*NS doubles with NS >= 4");
    constexpr int kPass = NV % 8 == 0 ? 8 : 4;
    for (int v0 = 0; v0 < NV; v0 += kPass) {
      double part[4][kPass];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t r = threadIdx.x + (uint32_t)kBlock * i;
#pragma unroll
        for (int v = 0; v < kPass; ++v) part[i][v] = r < count ? __ldcg(rows + (size_t)r * NV + v0 + v) : 0.0;
      }
      double x[kPass];
#pragma unroll
      for (int v = 0; v < kPass; ++v) x[v] = (part[0][v] + part[1][v]) + (part[2][v] + part[3][v]);
      __syncthreads();
      double (*red8)[kPass] = reinterpret_cast<double (*)[kPass]>(&red[0][0]);  // NV >= kPass: the scratch is large enough
      const double t = cta_sum<kPass>(x, red8);
      if (threadIdx.x < kPass) store(v0 + threadIdx.x, t);
    }
  }
}

template <int NM, int NS, class ScaleFn>
__device__ __forceinline__ void finish_tile(const float (&acc)[NM * NS], const uint32_t (&cnt)[(NS + 3) / 4], const FoldArgs& f,
                                            uint32_t opt, uint32_t tile, uint32_t tiles, uint32_t n_scen, ScaleFn&& scale) {
  constexpr int NA = NM * NS;  // accumulated sums
  constexpr int NV = NA + NS;  // + one paid-count per scenario
  __shared__ double red[kWarps][NV];
  __shared__ double tot[NV];
  __shared__ uint32_t is_last;
  double mine[NV];
#pragma unroll
  for (int i = 0; i < NA; ++i) mine[i] = (double)acc[i];
#pragma unroll
  for (int k = 0; k < NS; ++k) mine[NA + k] = (double)((cnt[k >> 2] >> (8 * (k & 3))) & 0xffu);
  const double cta_total = cta_sum<NV>(mine, red);

  const uint32_t n_groups = fold_groups(tiles);
  const uint32_t group = tile / kFoldGroup;
  double* const scratch = f.partials + (size_t)opt * fold_scratch_doubles(tiles, NV);
  double* const group_tot = scratch + (size_t)tiles * NV;  // [n_groups][NV]
  uint32_t* const tickets = f.tickets + (size_t)opt * fold_ticket_words(tiles);
  if (threadIdx.x < NV) {
    if (tiles == 1) {
      tot[threadIdx.x] = cta_total;  // the only tile of its option: nothing to fold
    } else {
      __stcg(scratch + (size_t)tile * NV + threadIdx.x, cta_total);
      __threadfence();
    }
  }
  if (tiles > 1) {
    // ---- level 1: the group's last CTA folds its <= 1024 tiles ---------------------------------------------------
    const uint32_t group_size = min((uint32_t)kFoldGroup, tiles - group * kFoldGroup);
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(tickets + 1 + group, 1u) == group_size - 1 ? 1u : 0u;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (n_groups == 1) {
      fold_rows<NV>(scratch, group_size, red, [&](int v, double x) { tot[v] = x; });
    } else {
      fold_rows<NV>(scratch + (size_t)group * kFoldGroup * NV, group_size, red, [&](int v, double x) {
        __stcg(group_tot + (size_t)group * NV + v, x);
        __threadfence();
      });
    }
    if (threadIdx.x == 0) tickets[1 + group] = 0u;  // ready for the next launch
    if (n_groups > 1) {
      // ---- level 2 (options of more than 1024 tiles): the last group folds the group totals ----------------------
      __syncthreads();
      if (threadIdx.x == 0) is_last = atomicAdd(tickets, 1u) == n_groups - 1 ? 1u : 0u;
      __syncthreads();
      if (!is_last) return;
      __threadfence();
      for (uint32_t g0 = 0; g0 < n_groups; g0 += kFoldGroup) {  // > 2^20 tiles: chunks of 1024 groups, in order
        const uint32_t n = min((uint32_t)kFoldGroup, n_groups - g0);
        __syncthreads();
        fold_rows<NV>(group_tot + (size_t)g0 * NV, n, red, [&](int v, double x) { tot[v] = g0 ? tot[v] + x : x; });
      }
      if (threadIdx.x == 0) tickets[0] = 0u;
    }
  }
  __syncthreads();
  constexpr int kRecDoubles = NM == 2 ? 3 : 6;
  const bool exchange = f.x.world > 1;
  // with the exchange on, this rank's records go to its exchange slot first
  double* const records = exchange ? reinterpret_cast<double*>(f.x.peer[f.x.rank] + (f.x.epoch & 1ull) * f.x.slot_bytes + kXchgHeaderBytes)
                                   : static_cast<double*>(f.out);
  if (threadIdx.x < n_scen) {
    const uint32_t k = threadIdx.x;
    const ScenScale sc = scale(k);
    const double paid = tot[NA + k];
    const double a1 = tot[NM * k], p2 = tot[NM * k + 1];  // sum_paid(x), FP32-strike sum(p^2)
    const double sgn = sc.is_put ? -1.0 : 1.0;
    const double S = sc.spot;
    double sum = a1, sum_sq = p2;
    if (sc.has_strike) {
      sum = sgn * (a1 - sc.kappa * paid);
      const double sum32 = sgn * (a1 - (double)sc.kappa32 * paid);  // what the FP32 payoffs add up to
      const double d = sgn * ((double)sc.kappa32 - sc.kappa);       // FP64-strike payoff = FP32-strike payoff + d on every paid sample
      sum_sq = p2 + 2.0 * d * sum32 + d * d * paid;
    }
    double* const rec = records + ((size_t)opt * n_scen + k) * kRecDoubles;
    if (NM == 2) {  // b200mc_moments_t
      rec[0] = sum * S;
      rec[1] = sum_sq * S * S;
      rec[2] = f.samples;
    } else {        // b200mc_cv_moments_t
      const double e1 = tot[NM * k + 2], e2 = tot[NM * k + 3], a2 = tot[NM * k + 4];  // sum S_T, sum S_T^2, sum_paid(x^2)
      rec[0] = sum * S;
      rec[1] = sum_sq * S * S;
      rec[2] = e1 * S;
      rec[3] = e2 * S * S;
      rec[4] = sgn * (a2 - sc.kappa * a1) * S * S;  // payoff * S_T = +-(x - k) * x over the paid samples
      rec[5] = f.samples;
    }
    if (exchange) __threadfence();
    else if (f.done) __threadfence_system();
  }
  if (exchange) {
    // ---- the launch's last option triggers the all-reduce over peer memory ------------------------------------------
    __syncthreads();
    if (threadIdx.x == 0) {
      is_last = atomicAdd(f.x.launch_ticket, 1u) == f.n_opt - 1 ? 1u : 0u;
      if (is_last) *f.x.launch_ticket = 0u;
    }
    __syncthreads();
    if (!is_last) return;
    const size_t slot_off = (f.x.epoch & 1ull) * f.x.slot_bytes;
    __threadfence_system();  // every option's records (observed through the ticket) before the flag, at system scope
    __syncthreads();
    if (threadIdx.x == 0)
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.x.peer[f.x.rank] + slot_off), "l"(f.x.epoch) : "memory");
    if (threadIdx.x < f.x.world && threadIdx.x != f.x.rank) {
      const char* flag = f.x.peer[threadIdx.x] + slot_off;
      unsigned long long seen, t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      } while (seen != f.x.epoch && t1 - t0 < f.x.timeout_ns);
      if (seen != f.x.epoch) *reinterpret_cast<volatile unsigned int*>(f.x.timed_out) = 1u;
    }
    __syncthreads();
    const size_t n_doubles = (size_t)f.n_opt * n_scen * kRecDoubles;
    double* const out = static_cast<double*>(f.out);
    for (size_t i = threadIdx.x; i < n_doubles; i += kBlock) {
      double total = 0.0;
      for (uint32_t r = 0; r < f.x.world; ++r) {  // rank order: identical bits on every rank
        const double* src = reinterpret_cast<const double*>(f.x.peer[r] + slot_off + kXchgHeaderBytes) + i;
        double v;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(src) : "memory");
        total += v;
      }
      out[i] = total;
    }
    if (f.done) __threadfence_system();
  }
  if (f.done) {  // host is polling: publish after every record is visible to it
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(f.done) = f.seq;
  }
}

// The strike-bearing GBM payoffs of b200mc_params_t.
__device__ __forceinline__ ScenScale vanilla_scale(const b200mc_params_t& p, bool is_put, bool has_strike = true) {
  ScenScale sc;
  sc.spot = p.S;
  sc.kappa = B200MC_DDIV(p.K, p.S);
  sc.kappa32 = (float)sc.kappa;
  sc.has_strike = has_strike;
  sc.is_put = is_put;
  return sc;
}

__device__ __forceinline__ ScenScale scen_scale(const SimArgs& a, uint32_t opt, uint32_t k, bool is_put, bool has_strike) {
  if (a.params) return vanilla_scale(a.params[(size_t)opt * a.n_scen + k], is_put, has_strike);
  ScenScale sc;
  sc.spot = a.inl.spot[k];
  sc.kappa = a.inl.kappa64[k];
  sc.kappa32 = a.inl.kappa[k];
  sc.has_strike = has_strike;
  sc.is_put = is_put;
  return sc;
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
// PAIRSUM: the consumer only adds a pair's two normals; full pairs arrive as box_muller_pair_sum (cs = sin(theta + pi/4),
// n_use = 2), a trailing odd step as the plain cosine branch (n_use = 1).
template <bool SQUARED, bool PAIRSUM>
__device__ __forceinline__ NormalPair full_pair(uint32_t w) {
  if (PAIRSUM) return box_muller_pair_sum<SQUARED>(w);
  return box_muller<SQUARED>(w);
}
template <bool SQUARED, bool PAIRSUM, class F>
__device__ __forceinline__ void consume_call(const u32x4& x, F&& f) {
  f(full_pair<SQUARED, PAIRSUM>(x.x), 2);
  f(full_pair<SQUARED, PAIRSUM>(x.y), 2);
  f(full_pair<SQUARED, PAIRSUM>(x.z), 2);
  f(full_pair<SQUARED, PAIRSUM>(x.w), 2);
}

// SQUARED: the pairs carry rad^2 = -log2(u) instead of rad (see box_muller).
// [j0, j1): the Philox calls of the path this thread visits (default: all of them; the lane-split European kernel
// hands each lane of a group its own range).  The trailing 1..7 steps belong to call index n_steps / 8.
template <int UNROLL = 1, bool SQUARED = false, bool PAIRSUM = false, class F>
__device__ __forceinline__ void for_each_pair(uint64_t path, uint32_t n_steps, uint32_t stream, const PhiloxKeys& rk, F&& f,
                                              uint32_t j0 = 0u, uint32_t j1 = 0xffffffffu) {
  if (n_steps == 1u) {  // single-step paths: one 64-bit draw (normal.cuh, box_muller_single)
    if (j0 == 0u) {
      const u32x4 x = draw4(path, 0u, stream, rk);
      f(box_muller_single<SQUARED>(x.x, x.y), 1);
    }
    return;
  }
  const uint32_t full = n_steps >> 3;
  const uint32_t stop = j1 < full ? j1 : full;
  uint32_t j = j0;
  if (UNROLL > 1) {
    for (; j + UNROLL <= stop; j += UNROLL) {
      u32x4 x[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) x[u] = draw4(path, j + u, stream, rk);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) consume_call<SQUARED, PAIRSUM>(x[u], f);
    }
  }
  // (plain #pragma unroll 2 / 4 of this loop measures 2.5% / 2% slower on the European kernel: profiles/r01_variants16_ffma2.txt)
  for (; j < stop; ++j) consume_call<SQUARED, PAIRSUM>(draw4(path, j, stream, rk), f);
  const int rem = (int)(n_steps & 7u);
  if (rem && j0 <= full && j1 > full) {  // 1..7 trailing steps: same word layout, only the pairs that are needed
    const u32x4 x = draw4(path, full, stream, rk);
    // (a pair whose second step is not needed is always the plain cosine branch)
    if (rem >= 2) f(full_pair<SQUARED, PAIRSUM>(x.x), 2);
    else f(box_muller<SQUARED>(x.x), 1);
    if (rem >= 4) f(full_pair<SQUARED, PAIRSUM>(x.y), 2);
    else if (rem > 2) f(box_muller<SQUARED>(x.y), 1);
    if (rem >= 6) f(full_pair<SQUARED, PAIRSUM>(x.z), 2);
    else if (rem > 4) f(box_muller<SQUARED>(x.z), 1);
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
// W = sum over steps of rad*cos / rad*sin (log2-radius units); everything else happens once per path.
// The terminal price only ever sees the SUM of a pair's two normals, rad (cos + sin) = sqrt(2) rad sin(theta + pi/4): one
// MUFU.SIN and one FFMA per pair instead of COS + SIN and two (box_muller_pair_sum: 3 MUFU per two path-steps instead of 4,
// 78 instructions per Philox call instead of 86).  The loop adds up rad * sin(theta + pi/4), a trailing odd step enters with
// the weight 1 / sqrt(2), and sqrt(2) is applied once per path.  The draws - and the value of W up to FP32 rounding - are
// those of the step-by-step kernels (normals_kernel, the path-dependent kinds, the FP64 oracle).
template <int UNROLL = 1>
__device__ __forceinline__ float terminal_sum(uint64_t path, uint32_t n_steps, uint32_t stream, const PhiloxKeys& rk,
                                              uint32_t j0 = 0u, uint32_t j1 = 0xffffffffu) {
  float W = 0.0f;
  for_each_pair<UNROLL, false, true>(path, n_steps, stream, rk, [&](const NormalPair& p, int n_use) {
    W = fmaf(n_use > 1 ? p.rad : p.rad * 0.70710678118654752440f, p.cs, W);
  }, j0, j1);
  return W * 1.41421356237309504880f;
}

// CV = true additionally accumulates sum S_T, sum S_T^2 and sum payoff*S_T per scenario (the
// terminal-spot control variate of monte_carlo.py:154-186): 5 sums instead of 2.
//
// SPLIT: launches too small to fill the chip with one thread per path (the reference's own sizes: 1e4..1e5 paths)
// give each path to 2^split_shift ADJACENT lanes.  W is a plain sum over the path's draws, so every lane adds up a
// contiguous range of the path's Philox calls and a butterfly of shuffles completes it; lane 0 of the group prices the
// payoff.  Same draws, same stream contract - only the FP32 association of W changes (by ~1e-7 relative).
template <int NS, bool ANTI, int MINB, bool CV = false, int UNROLL = 1, bool SPLIT = false>
__global__ void __launch_bounds__(kBlock, MINB) european_kernel(const SimArgs a) {
  constexpr int NM = CV ? 5 : 2;
  __shared__ Coef coef[NS];
  const uint32_t opt = blockIdx.x / a.tiles;
  const uint32_t tile = blockIdx.x - opt * a.tiles;
  if (threadIdx.x < NS) {
    const uint32_t k = threadIdx.x < a.n_scen ? threadIdx.x : a.n_scen - 1;
    coef[threadIdx.x] = scen_coef(a, opt, k, 1.0f);
  }
  __syncthreads();

  float acc[NM * NS];
  uint32_t paid[(NS + 3) / 4];
#pragma unroll
  for (int i = 0; i < NM * NS; ++i) acc[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < (NS + 3) / 4; ++i) paid[i] = 0u;

  const uint32_t stream = a.stream_base + opt;
  const bool is_put = a.is_put != 0;
  const uint32_t shift = SPLIT ? a.split_shift : 0u;
  const uint32_t lanes = 1u << shift;                      // lanes per path
  const uint32_t sub = threadIdx.x & (lanes - 1u);         // this lane's share of the path
  const uint32_t slot = threadIdx.x >> shift;              // the path this lane works on, within one pass of the CTA
  const uint32_t per_pass = (uint32_t)kBlock >> shift;     // paths per pass
  uint32_t j0 = 0u, j1 = 0xffffffffu;
  if (SPLIT) {
    const uint32_t calls = (a.n_steps + 7u) >> 3;
    const uint32_t per = (calls + lanes - 1u) >> shift;
    j0 = min(sub * per, calls);
    j1 = min(j0 + per, calls);
  }
  const uint64_t tile_first = (uint64_t)tile * (uint64_t)(per_pass * a.paths_per_thread);
  for (uint32_t j = 0; j < a.paths_per_thread; ++j) {
    const uint64_t local = tile_first + (uint64_t)j * per_pass + slot;
    const bool live = local < a.n_paths;
    if (!SPLIT && !live) break;  // paths are assigned in increasing order: nothing further for this thread
    float W = 0.0f;
    if (live) W = terminal_sum<UNROLL>(a.path_begin + local, a.n_steps, stream, a.rk, j0, j1);
    if (SPLIT) {
      if (__all_sync(0xffffffffu, !live)) break;
      for (uint32_t off = 1u; off < lanes; off <<= 1) W += __shfl_xor_sync(0xffffffffu, W, off);
      if (!live || sub != 0u) continue;
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const Coef q = coef[k];
#pragma unroll
      for (int mirror = 0; mirror < (ANTI ? 2 : 1); ++mirror) {
        const float e = mufu_ex2(fmaf(mirror ? -q.c : q.c, W, q.a));  // S_T / S_0
        const float p = vanilla(e, q.kappa, is_put);
        add_sample<NS>(acc + NM * k, paid, k, e, p);
        if (CV) {
          acc[NM * k + 2] += e;
          acc[NM * k + 3] = fmaf(e, e, acc[NM * k + 3]);
          acc[NM * k + 4] += p > 0.0f ? e * e : 0.0f;
        }
      }
    }
  }
  finish_tile<NM, NS>(acc, paid, a.fold, opt, tile, a.tiles, a.n_scen,
                      [&](uint32_t k) { return scen_scale(a, opt, k, is_put, true); });
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

// -> 2^x - 1 + plus (plus = 0 or 2, see advance_pair_small)
template <int DEG = 5>
__device__ __forceinline__ float exp2m1_small(float x, float plus = 0.0f) {
  float t = DEG >= 5 ? fmaf(x, 1.3333558146e-3f, 9.6181291076e-3f) : 9.6181291076e-3f;  // ln2^5/120, ln2^4/24
  if (DEG >= 4) t = fmaf(x, t, 5.5504108665e-2f);                                         // ln2^3/6
  else t = 5.5504108665e-2f;
  t = fmaf(x, t, 2.4022650696e-1f);                                                       // ln2^2/2
  t = fmaf(x, t, 6.9314718056e-1f);                                                       // ln2
  return fmaf(x, t, plus);
}

// DEG = kSmallPacked4 (shipped) / kSmallPacked5: packed degree 4 / 5.  DEG = 3..5: scalar Horner forms (4+ scenarios, scratch/variants14.cu).
constexpr int kSmallPacked5 = -5, kSmallPacked4 = -4;

// One pair of the multiplicative update: y0 = 2^x0 - 1, y1p2 = 2^x1 + 1 (the same three roundings in the packed and scalar forms).
__device__ __forceinline__ void pair_update(float& s, float& sum, float y0, float y1p2, int n_use) {
  const float s1 = fmaf(s, y0, s);
  if (n_use > 1) {
    sum = fmaf(s1, y1p2, sum);
    s = fmaf(s1, y1p2, -s1);
  } else {
    sum += s1;
    s = s1;
  }
}

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
      // (y0, y1 + 2) = (2^x0 - 1, 2^x1 + 1): with s1 = s * 2^x0 the pair adds s1 + s2 = s1 * (2^x1 + 1) to the running sum and
      // leaves s2 = s1 * (2^x1 + 1) - s1 - three FMAs instead of two FMAs and two adds
      float y0, y1p2;
      unpack2(fma2(x, t, pack2(0.0f, 2.0f)), y0, y1p2);
      pair_update(s[k], aux[k], y0, y1p2, n_use);
    }
  } else {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const float rc = p.rad * q[k].c;
      pair_update(s[k], aux[k], exp2m1_small<DEG>(fmaf(rc, p.cs, q[k].d)), exp2m1_small<DEG>(fmaf(rc, p.sn, q[k].d), 2.0f), n_use);
    }
  }
}

// -> FP32 payoff p; x = the quantity compared with the strike ratio (p itself for the strike-less floating lookback).
template <int KIND>
__device__ __forceinline__ float path_payoff(float l, float aux, const Coef& q, const SimArgs& a, float& x) {
  const bool is_put = a.is_put != 0;
  if (KIND == B200MC_ASIAN_ARITH) return vanilla(x = aux * q.inv_n, q.kappa, is_put);
  if (KIND == B200MC_ASIAN_GEOM) return vanilla(x = mufu_ex2(aux * q.inv_n), q.kappa, is_put);
  const float sgn = a.sgn_negative ? -1.0f : 1.0f;
  const float e_T = mufu_ex2(sgn * l);
  if (KIND == B200MC_BARRIER) {
    const bool crossed = aux >= q.beta;
    const bool active = crossed == (a.barrier_in != 0);
    x = e_T;
    return active ? vanilla(e_T, q.kappa, is_put) : 0.0f;
  }
  // LOOKBACK: aux tracks max(l) (sgn=+1) or max(-l) = -min(l) (sgn=-1)
  const float e_ext = mufu_ex2(sgn * aux);
  if (a.lookback_fixed) return vanilla(x = e_ext, q.kappa, is_put);
  return x = (is_put ? e_ext - e_T : e_T - e_ext);
}

// One CTA's share of one option: paths_per_thread paths per thread, payoffs accumulated in FP32 per thread.
// SMALL selects the multiplicative arithmetic-Asian update (state = S_t/S_0 instead of log2 of it).
template <int KIND, int NS, bool SMALL, int UNROLL, int DEG = kSmallPacked4>
__device__ __forceinline__ void simulate_tile(const SimArgs& a, const Coef (&q)[NS], uint32_t stream, uint64_t tile_first, float (&acc)[2 * NS],
                                              uint32_t (&paid)[(NS + 3) / 4]) {
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
      float x;
      const float p = path_payoff<KIND>(l[k], aux[k], q[k], a, x);
      add_sample<NS>(acc + 2 * k, paid, k, x, p);
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
    coef_s[threadIdx.x] = scen_coef(a, opt, k, a.sgn_negative ? -1.0f : 1.0f);
  }
  __syncthreads();
  Coef q[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) q[k] = coef_s[k];

  float acc[2 * NS];
  uint32_t paid[(NS + 3) / 4];
#pragma unroll
  for (int i = 0; i < 2 * NS; ++i) acc[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < (NS + 3) / 4; ++i) paid[i] = 0u;

  const uint32_t stream = a.stream_base + opt;
  const uint64_t tile_first = (uint64_t)tile * (uint64_t)(kBlock * a.paths_per_thread);
  bool small = KIND == B200MC_ASIAN_ARITH && !a.force_mufu_ex2;
  if (KIND == B200MC_ASIAN_ARITH) {
#pragma unroll
    for (int k = 0; k < NS; ++k) small = small && (fabsf(q[k].d) + fabsf(q[k].c) * kRadMax <= kSmallMove);  // NaN -> false
  }
  if (small) simulate_tile<B200MC_ASIAN_ARITH, NS, true, UNROLL, DEG>(a, q, stream, tile_first, acc, paid);  // CTA-uniform branch
  else simulate_tile<KIND, NS, false, UNROLL>(a, q, stream, tile_first, acc, paid);
  // the floating-strike lookback pays S_T - min S / max S - S_T: no strike term to refine
  const bool has_strike = !(KIND == B200MC_LOOKBACK && !a.lookback_fixed);
  finish_tile<2, NS>(acc, paid, a.fold, opt, tile, a.tiles, a.n_scen,
                     [&](uint32_t k) { return scen_scale(a, opt, k, a.is_put != 0, has_strike); });
}

// ================================ fold of the FP64 parity kernels ===============================
// The from-normals kernels (f64_kernels.cuh, models.cuh, structured.cuh: test infrastructure fed the reference's own
// draws) write one currency-unit (sum, sum^2) pair per CTA; one warp folds them in a fixed order.
__global__ void __launch_bounds__(32) fold_kernel(const double* __restrict__ partials, b200mc_moments_t* __restrict__ out, uint32_t tiles,
                                                  double samples) {
  double s1 = 0.0, s2 = 0.0;
  for (uint32_t t = threadIdx.x; t < tiles; t += 32) {
    s1 += partials[2 * (size_t)t];
    s2 += partials[2 * (size_t)t + 1];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  }
  if (threadIdx.x == 0) {
    out->sum = s1;
    out->sum_sq = s2;
    out->n = samples;
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
