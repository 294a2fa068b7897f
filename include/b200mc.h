/* b200mc — C ABI of the B200-native Monte Carlo pricing engine (libb200mc.so).
 *
 * This is the drop-in boundary for OptionsLab's Monte Carlo hot path.  The reference is pure
 * Python and has no FFI of its own; its boundary is the duck-typed pricer interface
 *   PricerProtocol.price(S, K, T, r, sigma, option_type, q=0.0, **kw)   src/greeks/unified_greeks.py:45-66
 * plus the concrete classes behind it.  The Python classes in optionslab_b200/ mirror those classes
 * and call ONLY the functions declared here (ctypes; see INTEGRATION.md for the binding stub a
 * reference maintainer would add).  Each entry point names the reference computation it replaces.
 *
 * Conventions: plain pointers and sizes, no C++ / torch types.  Every function returns 0 on
 * success or a negative b200mc_status; the message is available from b200mc_last_error().  The
 * library owns all device scratch memory; the caller owns every buffer it passes in.  "host"
 * entry points block until results are in the caller's host buffers; "_device" entry points take
 * device pointers plus a cudaStream_t (as void*) and are asynchronous on that stream.
 */
#ifndef B200MC_H_
#define B200MC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MC_ABI_VERSION 10
#define B200MC_MAX_SCENARIOS 16

typedef struct b200mc_engine b200mc_engine_t;

typedef enum {
  B200MC_OK = 0,
  B200MC_ERR_INVALID = -1, /* bad argument (null pointer, zero size, unknown kind, ...) */
  B200MC_ERR_CUDA = -2,    /* CUDA runtime failure; message has the cudaError string */
  B200MC_ERR_COMM = -3,    /* multi-GPU exchange failure (peer mapping, oversized record block, peer rank timed out) */
  B200MC_ERR_NOMEM = -4    /* device or pinned-host allocation failed */
} b200mc_status;

/* Payoff families on the hot path. */
typedef enum {
  B200MC_EUROPEAN = 0,    /* terminal payoff;  src/pricing_models/monte_carlo.py:140-143 */
  B200MC_ASIAN_ARITH = 1, /* mean of S_t, t=1..n;  src/pricing_models/exotic_options.py:119-120 */
  B200MC_ASIAN_GEOM = 2,  /* exp(mean(log S_t));   src/pricing_models/exotic_options.py:121-122 */
  B200MC_BARRIER = 3,     /* any(S_t >= B) / any(S_t <= B), t=0..n;  exotic_options.py:201-212 */
  B200MC_LOOKBACK = 4,    /* running max / min, t=0..n;  exotic_options.py:382-399 */
  /* structured products of the same file; only b200mc_simulate_structured / b200mc_structured_from_normals take them */
  B200MC_CLIQUET = 5,     /* sum of locally clipped period returns, clipped globally;  exotic_options.py:525-552 */
  B200MC_AUTOCALLABLE = 6 /* early redemption on observation dates, coupon, knock-in put;  exotic_options.py:438-488 */
} b200mc_kind;

/* ASIAN_ARITH: always evaluate S_t = S_0 * 2^(l_t) with MUFU.EX2 instead of the multiplicative
 * small-move update the kernel picks per option when |log2 increment| <= 0.25 (mc_kernels.cuh). */
#define B200MC_FLAG_EXACT_EX2 1u
/* FP64 parity mode, EUROPEAN: use the plain-load kernel even when the bulk-async staged one applies (A/B). */
#define B200MC_FLAG_NO_BULK_COPY 2u

typedef struct {
  int32_t kind;           /* b200mc_kind */
  int32_t is_put;         /* 0 call, 1 put */
  int32_t antithetic;     /* 1: each draw also prices the mirrored path (-Z), as gbm_numpy.py:48-51.
                             Only B200MC_EUROPEAN supports it (the reference's exotics do not mirror). */
  int32_t barrier_down;   /* BARRIER: 0 = up (S_t >= B), 1 = down (S_t <= B) */
  int32_t barrier_in;     /* BARRIER: 0 = knock-out, 1 = knock-in */
  int32_t lookback_fixed; /* LOOKBACK: 0 = floating strike, 1 = fixed strike */
  uint32_t n_steps;       /* time steps per path (>= 1) */
  uint32_t flags;         /* B200MC_FLAG_* (0 = defaults) */
} b200mc_spec_t;

/* One (option, scenario) parameter set, FP64.  Scenarios of an option share its normal draws
 * (common random numbers) — the bumped re-pricings of src/greeks/unified_greeks.py:295-358. */
typedef struct {
  double S, K, T, r, sigma, q;
  double barrier; /* BARRIER only */
  double reserved;
} b200mc_params_t;

/* Terms of a structured product (dataclass fields of CliquetOption / AutocallableOption plus the schedule argument of
 * their price methods, exotic_options.py:416-427, :502-513).  One product per launch, shared by all options / scenarios.
 *   B200MC_CLIQUET      : a = local_cap, b = local_floor, c = global_cap, d = global_floor, period = n_periods
 *                         (each period spans n_steps / n_periods steps, integer division; later steps are ignored)
 *   B200MC_AUTOCALLABLE : a = autocall_barrier, b = coupon_barrier, c = coupon_rate, d = ki_barrier (all relative to S),
 *                         period = observation_freq (an observation every `period` steps, the first at step `period`) */
typedef struct {
  double a, b, c, d;
  uint32_t period;
  uint32_t reserved;
} b200mc_product_t;

/* Raw FP64 payoff moments of one (option, scenario): sum over samples of the UNDISCOUNTED payoff,
 * of its square, and the sample count (2x paths when antithetic).  The host applies exp(-rT), the
 * mean, the standard error (monte_carlo.py:145-150) and the finite-difference formulas. */
typedef struct {
  double sum, sum_sq, n;
} b200mc_moments_t;

/* Moments for the terminal-spot control variate (src/pricing_models/monte_carlo.py:154-186):
 * sums over samples of payoff, payoff^2, S_T, S_T^2 and payoff*S_T (all undiscounted), and n. */
typedef struct {
  double sum_payoff, sum_payoff_sq, sum_terminal, sum_terminal_sq, sum_payoff_terminal, n;
} b200mc_cv_moments_t;

/* Heston stochastic-volatility parameters of one option (src/pricing_models/heston.py:41-58 + the
 * arguments of price_monte_carlo, :184-195). */
typedef struct {
  double S, K, T, r, q;
  double kappa, theta, sigma_v, rho, v0;
  double reserved[2];
} b200mc_heston_params_t;

/* Jump part of a jump-diffusion (src/pricing_models/jump_diffusion.py:43-67, :274-308).
 *   model = B200MC_JUMP_MERTON : a = mu_j, b = sigma_j           (log-normal jumps)
 *   model = B200MC_JUMP_KOU    : a = p,    b = eta1, c = eta2    (double-exponential jumps) */
typedef enum { B200MC_JUMP_MERTON = 0, B200MC_JUMP_KOU = 1 } b200mc_jump_model;
#define B200MC_MODEL_PUT 1
#define B200MC_MODEL_SHARED_STREAM 2
typedef struct {
  int32_t model;
  int32_t reserved0;
  double lambda_j;
  double a, b, c;
  double reserved[3];
} b200mc_jump_params_t;

typedef struct {
  int32_t device;
  int32_t sm_count;
  int32_t cc_major, cc_minor;
  int32_t sm_clock_khz;  /* cudaDevAttrClockRate */
  int32_t mem_clock_khz;
  int64_t total_mem_bytes;
  int32_t l2_bytes;
  int32_t reserved;
  char name[64];
} b200mc_info_t;

/* Pipe-rate microbenchmarks, measured on the device the engine owns (thread-level ops / second).
 * These are the roofline denominators for the issue/XU-bound simulation kernels. */
typedef struct {
  double ffma_per_s;      /* FP32 FMA, 3-register form */
  double imad_wide_per_s; /* 32x32->64 integer multiply (the Philox multiply) */
  double lop3_per_s;      /* 3-input logic op */
  double mufu_per_s;      /* MUFU mix used by the generator (lg2, sqrt, sin, cos) */
  double mufu_ex2_per_s;
  double issue_per_s;     /* mixed independent FFMA+LOP3 stream: warp-instruction issue ceiling x 32 */
  double philox_per_s;    /* Philox4x32-10 calls / s, counter mode, nothing else */
  double normals_per_s;   /* Philox + Box-Muller normals / s, summed in registers, nothing else */
  double sm_clock_mhz_seen; /* clock64()-derived average SM clock while the probes ran */
  double reserved[3];
} b200mc_peaks_t;

/* ---- lifetime ------------------------------------------------------------------------------ */
int b200mc_abi_version(void);
int b200mc_create(b200mc_engine_t** out, int device);
void b200mc_destroy(b200mc_engine_t* eng);
/* Message of the last failed call on this engine (eng == NULL: last failed b200mc_create). */
const char* b200mc_last_error(const b200mc_engine_t* eng);
int b200mc_device_info(b200mc_engine_t* eng, b200mc_info_t* out);

/* ---- the hot path: fused Philox -> GBM -> payoff -> (sum, sum^2) ----------------------------- *
 * Replaces, per option i and scenario k:  simulate_gbm_numpy + payoff + mean/std
 * (src/simulation/gbm_numpy.py:32-53, src/pricing_models/monte_carlo.py:140-150), the batched
 * variant (src/pricing_models/monte_carlo_unified.py:321-343,622-631) and the exotic path
 * generator + payoff (src/pricing_models/exotic_options.py:54-67,116-131,198-224,380-401).
 *
 * params  : [n_opt][n_scen] (row-major), 1 <= n_scen <= B200MC_MAX_SCENARIOS
 * seed    : Philox key.  Option i draws from stream (stream_base + i); all its scenarios share it.
 * paths   : global path indices [path_begin, path_begin + n_paths) — disjoint ranges on different
 *           ranks give disjoint Philox subsequences; results depend only on the index range.
 * out     : [n_opt][n_scen] moments.
 */
int b200mc_simulate(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_params_t* params_host,
                    uint32_t n_opt, uint32_t n_scen, uint64_t seed, uint32_t stream_base,
                    uint64_t path_begin, uint64_t n_paths, b200mc_moments_t* out_host);

int b200mc_simulate_device(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_params_t* params_dev,
                           uint32_t n_opt, uint32_t n_scen, uint64_t seed, uint32_t stream_base,
                           uint64_t path_begin, uint64_t n_paths, b200mc_moments_t* out_dev,
                           void* cuda_stream);

/* ---- multi-GPU: the same launch with the all-reduce fused into its tail ------------------------------------------- *
 * rank g of G simulates global paths [g*N/G, (g+1)*N/G) of every option and the per-(option,
 * scenario) moment records are summed over the ranks.  Here that sum is the LAST thing the simulation kernel does: the
 * CTA that folds a rank's last option publishes the rank's records in its exchange block, waits for the other ranks'
 * flags and adds their records over NVLink peer memory in rank order (identical bits on every rank) - no host hop, no
 * second kernel, no library collective.  Up to 8 ranks (one NVSwitch domain), one process per GPU or several devices of
 * one process.
 *
 * Connecting (collective; a barrier must separate it from the first b200mc_simulate_allreduce):
 *   1. every rank:  b200mc_comm_export(eng, handle)      -> 64-byte CUDA IPC handle of its exchange block
 *   2. all-gather the handles with whatever the host side has (torch.distributed, MPI, a file)
 *   3. every rank:  b200mc_comm_connect(eng, rank, world, handles[world][64])
 *   one process, several devices:  b200mc_comm_connect_local(engines, n)  (rank = index; peer access instead of IPC)
 * b200mc_simulate_allreduce[_device] then behave like b200mc_simulate[_device] on this rank's path range, except that
 * `out` receives the moments of ALL ranks' paths; every rank must make the same sequence of calls.  n_paths may be 0
 * (a rank with an empty share still contributes).  A peer that does not show up within 20 s makes the call fail with
 * B200MC_ERR_COMM instead of hanging.  Unconnected engines (world <= 1) simply run b200mc_simulate. */
#define B200MC_COMM_HANDLE_BYTES 64
int b200mc_comm_export(b200mc_engine_t* eng, void* handle_out);
int b200mc_comm_connect(b200mc_engine_t* eng, int rank, int world, const void* handles);
int b200mc_comm_connect_local(b200mc_engine_t* const* engines, int n);
int b200mc_comm_disconnect(b200mc_engine_t* eng);
int b200mc_comm_world(const b200mc_engine_t* eng);
/* Collective mode: while on, EVERY fused launch of a host entry point (b200mc_simulate, _control_variate, _structured, _heston,
 * _jump_diffusion, _sobol) on a connected engine is a collective call - this rank's path / point range in, the moments of all
 * ranks' ranges out, n_paths (n_points) may be 0 - exactly like b200mc_simulate_allreduce.  Meant to be switched on around one
 * call by the sharding layer (optionslab_b200/distributed.py); off by default and after (re)connecting. */
int b200mc_comm_set_collective(b200mc_engine_t* eng, int on);
/* Bound on the in-kernel wait for a peer rank's records (default 20 000 ms); a launch that exceeds it returns B200MC_ERR_COMM. */
int b200mc_comm_set_timeout_ms(b200mc_engine_t* eng, uint32_t milliseconds);
int b200mc_simulate_allreduce(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_params_t* params_host,
                              uint32_t n_opt, uint32_t n_scen, uint64_t seed, uint32_t stream_base,
                              uint64_t path_begin, uint64_t n_paths, b200mc_moments_t* out_host);
int b200mc_simulate_allreduce_device(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_params_t* params_dev,
                                     uint32_t n_opt, uint32_t n_scen, uint64_t seed, uint32_t stream_base,
                                     uint64_t path_begin, uint64_t n_paths, b200mc_moments_t* out_dev,
                                     void* cuda_stream);

/* European payoff + the sums np.cov(discounted, terminal) needs, same launch shape as b200mc_simulate
 * (spec->kind must be B200MC_EUROPEAN).  Replaces price_with_control_variate's simulation and
 * reductions (src/pricing_models/monte_carlo.py:166-186); beta and the adjustment stay on the host. */
int b200mc_simulate_control_variate(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_params_t* params_host,
                                    uint32_t n_opt, uint32_t n_scen, uint64_t seed, uint32_t stream_base,
                                    uint64_t path_begin, uint64_t n_paths, b200mc_cv_moments_t* out_host);

/* ---- quasi-Monte Carlo: scrambled Sobol points generated on the device ------------------------- *
 * Replaces simulate_gbm_qmc (src/simulation/gbm_qmc.py:14-47), the MCMethod.QMC backend of
 * src/pricing_models/monte_carlo.py:94-97: Sobol(d = n_steps, scramble=True, seed).random(N) ->
 * norm.ppf(clip(u, 1e-10, 1 - 1e-10)) -> terminal GBM -> payoff -> moments (no mirroring).
 *
 * The point set is passed as its GF(2)-linear description, in NATURAL (binary) order of the point index:
 *     x_j(i) = shift[j] XOR (XOR over the set bits b of i) dirnums[j][b],    u_j(i) = x_j(i) * 2^-bits
 * dirnums : [spec->n_steps][32] words (entries b >= bits must be 0), shift : [spec->n_steps].
 *           For a Gray-code generator with direction numbers v_j[b] (scipy: Sobol._sv, Sobol._shift),
 *           dirnums[j][b] = v_j[b] ^ v_j[b-1] reproduces its points bit for bit.
 * points  : indices [point_begin, point_begin + n_points) of the sequence; point_begin must be a multiple
 *           of 4096 (one CTA) — ranks take disjoint ranges of the same sequence.
 * spec    : kind = B200MC_EUROPEAN, antithetic = 0.  All options / scenarios share the points (CRN).
 */
int b200mc_simulate_sobol(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_params_t* params_host,
                          uint32_t n_opt, uint32_t n_scen, const uint32_t* dirnums_host, const uint32_t* shift_host,
                          uint32_t bits, uint64_t point_begin, uint64_t n_points, b200mc_moments_t* out_host);
/* Inspection: out_host[(i - point_begin) * n_dims + j] = x_j(i), the integers the QMC kernel works from. */
int b200mc_sobol_points(b200mc_engine_t* eng, const uint32_t* dirnums_host, const uint32_t* shift_host, uint32_t n_dims,
                        uint32_t bits, uint64_t point_begin, uint64_t n_points, uint32_t* out_host);
/* Inspection: the FP32 inverse-normal values the QMC kernel derives from the given Sobol integers. */
int b200mc_sobol_normals(b200mc_engine_t* eng, const uint32_t* x_host, uint64_t n, uint32_t bits, float* out_host);

/* ---- other Euler Monte Carlo models of the reference -------------------- *
 * Heston, full-truncation Euler: replaces HestonPricer.price_monte_carlo (src/pricing_models/heston.py:184-255).
 * One Box-Muller pair per step (Z1 and the independent part of Z2).  params: [n_opt]; out: [n_opt].
 * Paths / seed / stream conventions as b200mc_simulate (option i uses stream stream_base + i).
 * is_put carries B200MC_MODEL_* bits: B200MC_MODEL_PUT prices puts; B200MC_MODEL_SHARED_STREAM makes every entry of
 * params use stream stream_base, i.e. the option axis becomes a common-random-number scenario axis (the bumped
 * re-pricings of compute_greeks_unified, src/greeks/unified_greeks.py:295-358, in one launch). */
int b200mc_simulate_heston(b200mc_engine_t* eng, const b200mc_heston_params_t* params_host, uint32_t n_opt, int is_put,
                           uint32_t n_steps, uint64_t seed, uint32_t stream_base, uint64_t path_begin, uint64_t n_paths,
                           b200mc_moments_t* out_host);
/* Merton / Kou jump diffusion: replaces MertonJumpDiffusion.price_monte_carlo and KouJumpDiffusion.price_monte_carlo
 * (src/pricing_models/jump_diffusion.py:160-225, :325-377).  The diffusion is stepped n_steps times with the
 * compensated drift; the compound-Poisson jump sum of each path is drawn once with its exact law.
 * params: [n_opt] (S, K, T, r, sigma, q); jumps: [n_opt]; out: [n_opt]. */
int b200mc_simulate_jump_diffusion(b200mc_engine_t* eng, const b200mc_params_t* params_host, const b200mc_jump_params_t* jumps_host,
                                   uint32_t n_opt, int is_put, uint32_t n_steps, uint64_t seed, uint32_t stream_base,
                                   uint64_t path_begin, uint64_t n_paths, b200mc_moments_t* out_host);
/* FP64 parity mode for those models.  Draws are STEP-major, the order the reference consumes its generator:
 *   Heston : Z  [n_steps][2][n_paths]  (heston.py:228-229: Z1, then the normal mixed into Z2)
 *   jumps  : dW [n_steps][n_paths] and J [n_steps][n_paths] = sum of the jump sizes hitting path i in step t
 *            (jump_diffusion.py:205-216 / :352-367; J may be NULL); lambda_kappa = lambda_j * E[e^Y - 1].
 * payoffs: per-path undiscounted payoffs [n_paths] (may be NULL); out: their moments. */
int b200mc_heston_from_normals(b200mc_engine_t* eng, const b200mc_heston_params_t* p, int is_put, uint32_t n_steps,
                               const double* Z_host, uint64_t n_paths, double* payoffs_host, b200mc_moments_t* out_host);
int b200mc_jump_diffusion_from_draws(b200mc_engine_t* eng, const b200mc_params_t* p, double lambda_kappa, int is_put,
                                     uint32_t n_steps, const double* dW_host, const double* J_host, uint64_t n_paths,
                                     double* payoffs_host, b200mc_moments_t* out_host);

/* ---- structured products (exotic_options.py:404-552) ------------------------------------------------------------- *
 * Same fused path as b200mc_simulate (Philox normals in registers, log-Euler steps, payoff, FP64 moments, scenarios on
 * common random numbers), with the period / observation schedule counted down in registers.
 *   B200MC_CLIQUET      : out = moments of the UNDISCOUNTED payoff max(clip(sum_p clip(R_p)), 0) * S; the host applies
 *                         exp(-rT) (exotic_options.py:552).
 *   B200MC_AUTOCALLABLE : out = moments of the DISCOUNTED payoff per unit notional - redemption at observation i pays
 *                         (1 + c*(i/n_obs)*T) * exp(-r t_i), maturity pays exp(-rT) * {1 (+ c*T above the coupon barrier),
 *                         or S_T/S after a knock-in with S_T < S}; the price is sum / n (exotic_options.py:466,486-488).
 * spec: kind = B200MC_CLIQUET or B200MC_AUTOCALLABLE, n_steps; the other fields are ignored.  K is unused. */
int b200mc_simulate_structured(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_product_t* product,
                               const b200mc_params_t* params_host, uint32_t n_opt, uint32_t n_scen, uint64_t seed,
                               uint32_t stream_base, uint64_t path_begin, uint64_t n_paths, b200mc_moments_t* out_host);
/* The same two products in FP64 on caller-supplied draws Z [n_paths][n_steps] (the reference's own normals,
 * exotic_options.py:59), statement by statement as the reference evaluates them.  payoffs (may be NULL) and out follow
 * the conventions above. */
int b200mc_structured_from_normals(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_product_t* product,
                                   const b200mc_params_t* p, const double* Z_host, uint64_t n_paths, double* payoffs_host,
                                   b200mc_moments_t* out_host);

/* ---- the simulation layer: terminal price arrays ------------------------------------------------------------------- *
 * The array the reference's simulation backends return (src/simulation/__init__.py: simulate_terminal_prices(S, T, r,
 * sigma, q, n_paths, n_steps, seed) -> ndarray): simulate_gbm_numpy / _fast / simulate_gbm_numba (gbm_numpy.py:15-83,
 * gbm_numba.py:100-129) and, for the Sobol variant, simulate_gbm_qmc / _antithetic (gbm_qmc.py:14-76).
 *   out_host[i]           = S_T of path (path_begin + i) on its draws,            i in [0, n_paths)
 *   out_host[n_paths + i] = S_T of the mirrored path (-Z) when antithetic != 0    (the reference's concatenate layout)
 * Values are the fused path's FP32 arithmetic on log2(S_T/S) scaled to FP64; K is unused.  This is the one call that
 * writes per-path data to HBM (8 or 16 bytes per path, because the caller asks for the array). */
int b200mc_terminal_prices(b200mc_engine_t* eng, const b200mc_params_t* p, uint32_t n_steps, int antithetic, uint64_t seed,
                           uint32_t stream, uint64_t path_begin, uint64_t n_paths, double* out_host);
/* Sobol points [point_begin, point_begin + n_points) of the sequence described by (dirnums, shift, bits) as in
 * b200mc_simulate_sobol; point_begin must be a multiple of 4096. */
int b200mc_terminal_prices_sobol(b200mc_engine_t* eng, const b200mc_params_t* p, uint32_t n_steps, int antithetic,
                                 const uint32_t* dirnums_host, const uint32_t* shift_host, uint32_t bits, uint64_t point_begin,
                                 uint64_t n_points, double* out_host);

/* ---- FP64 parity mode: price from caller-supplied normal draws ------------------------------- *
 * Z is row-major [n_paths][spec->n_steps] FP64 — exactly the array the reference draws at
 * gbm_numpy.py:43 / monte_carlo_unified.py:329 (one option) / exotic_options.py:59.
 * accumulate = 0: log S_T = ln S + drift*n + vol*sum(Z)            (gbm_numpy.py:46-50)
 * accumulate = 1: log S_t = ln S + running sum of (drift + vol*Z)   (monte_carlo_unified.py:333-337,
 *                 exotic_options.py:62-65); forced for path-dependent kinds.
 * payoffs : per-path undiscounted payoffs; [0,N) from +Z and, when spec->antithetic, [N,2N) from -Z
 *           (the reference's concatenate layout, gbm_numpy.py:51).  May be NULL.
 * out     : moments over all samples (fixed-order FP64 tree; deterministic).
 */
int b200mc_payoffs_from_normals(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_params_t* p,
                                int accumulate, const double* Z_host, uint64_t n_paths,
                                double* payoffs_host, b200mc_moments_t* out_host);

int b200mc_payoffs_from_normals_device(b200mc_engine_t* eng, const b200mc_spec_t* spec, const b200mc_params_t* p,
                                       int accumulate, const double* Z_dev, uint64_t n_paths,
                                       double* payoffs_dev, b200mc_moments_t* out_dev, void* cuda_stream);

/* ---- inspection of the random stream (tests / diagnostics) ----------------------------------- */
/* out_host[(p - path_begin) * n_steps + s] = the FP32 standard normal the simulation kernels use
 * for step s of global path p of (seed, stream). */
int b200mc_generate_normals(b200mc_engine_t* eng, uint64_t seed, uint32_t stream, uint64_t path_begin,
                            uint64_t n_paths, uint32_t n_steps, float* out_host);
/* Statistics of the normals of paths [path_begin, path_begin + n_paths) x n_steps of (seed, stream), gathered on the
 * device (the stream a simulation consumes, at GPU scale: 1e10 draws take tens of milliseconds):
 *   hist_z[256]      counts of z in 256 equal bins over [-6, 6) (outer bins collect the rest)
 *   hist_joint[64*64] counts of (z_cos, z_sin) - the two normals of ONE random word - on a 64 x 64 grid over [-4, 4)^2,
 *                    row = cosine-branch normal (outer cells collect the rest)
 *   tails[4]         exact counts of z > 4, z > 5, z < -4, z < -5
 *   moments[16]      0..3: sum z, z^2, z^3, z^4;  4..7: same-word sums z1 z2, z1^2 z2^2, z1 z2^3, z1^3 z2;
 *                    8..10: lag-1 sums a b, a^2 b^2, a b^3 (a = sine branch of word n, b = cosine branch of word n+1);
 *                    11: normals counted, 12: same-word pairs counted, 13: lag-1 pairs counted; 14, 15 unused */
typedef struct {
  uint64_t hist_z[256];
  uint64_t hist_joint[64 * 64];
  uint64_t tails[4];
  double moments[16];
} b200mc_rng_stats_t;
int b200mc_rng_statistics(b200mc_engine_t* eng, uint64_t seed, uint32_t stream, uint64_t path_begin, uint64_t n_paths,
                          uint32_t n_steps, b200mc_rng_stats_t* out);
/* Raw Philox4x32-10: in = n x {c0,c1,c2,c3,k0,k1}, out = n x 4 words (known-answer tests). */
int b200mc_philox_raw(b200mc_engine_t* eng, const uint32_t* ctr_key_host, uint32_t n, uint32_t* out_host);

/* ---- measurement ------------------------------------------------------------------------------ */
int b200mc_measure_peaks(b200mc_engine_t* eng, b200mc_peaks_t* out);
/* Kernels launched by this engine since creation (the caller's gpu_launches evidence). */
uint64_t b200mc_kernel_launches(const b200mc_engine_t* eng);
/* Tuning / tests: pin the tile shape the planner would otherwise choose from the problem size.  split_shift: 2^split_shift
 * adjacent lanes share each European path (0..3; < 0 = automatic); paths_per_thread: 1..32 (0 = automatic).  Prices do not
 * depend on the shape beyond FP32 summation order (~1e-7 relative). */
int b200mc_set_plan(b200mc_engine_t* eng, int split_shift, uint32_t paths_per_thread);
/* The tile shape the planner gives a fused launch on a device of `sm_count` SMs - a pure function of its arguments (no engine,
 * no device: the host logic is testable without a GPU).  A tile = one CTA of 256 threads; a thread owns paths_per_thread paths,
 * or 2^split_shift adjacent lanes share each path (European launches that would leave most SMs idle). */
int b200mc_plan_tiles(int sm_count, const b200mc_spec_t* spec, uint32_t n_opt, uint32_t n_scen, uint64_t n_paths,
                      int control_variate, uint32_t* tiles, uint32_t* paths_per_thread, uint32_t* split_shift);
/* The tile shape of the most recent b200mc_simulate* launch: tiles per option, paths per thread, split shift. */
int b200mc_last_plan(b200mc_engine_t* eng, uint32_t* tiles, uint32_t* paths_per_thread, uint32_t* split_shift);
/* When enabled, every simulation / from-normals kernel is bracketed by a CUDA event pair on the
 * stream it is launched on (a ring of 64 pairs; enabling resets it).  b200mc_kernel_timing waits for
 * the recorded kernels and returns their mean and minimum duration and how many were timed. */
int b200mc_set_kernel_timing(b200mc_engine_t* eng, int enabled);
int b200mc_kernel_timing(b200mc_engine_t* eng, float* mean_ms, float* min_ms, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* B200MC_H_ */
