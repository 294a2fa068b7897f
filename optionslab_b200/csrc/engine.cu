// libb200mc.so — C ABI (include/b200mc.h) over the sm_100a simulation kernels.
//
// One engine = one device + one stream + scratch (tile partials, staged parameters, pinned host
// mirrors).  No exceptions cross the boundary; every failure becomes a status code plus a message.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200mc.h"
#include "f64_kernels.cuh"
#include "mc_kernels.cuh"
#include "models.cuh"
#include "peaks.cuh"
#include "rng_stats.cuh"
#include "sobol.cuh"
#include "structured.cuh"

using namespace b200mc;

namespace {

std::mutex g_create_mutex;
std::string g_create_error;

struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct b200mc_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop{};
  std::string error;
  std::mutex mutex;  // one call at a time per engine (the reference's pricer objects are single-threaded too)
  DeviceBuffer partials, tickets, params_dev, moments_dev, scratch_a, scratch_b, qmc_wpart, qmc_tickets;
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  // single-option launches: parameters ride in the kernel arguments, the finishing CTA writes the records straight into
  // this mapped pinned block and then the sequence word the host is polling (no H2D / D2H copy, no stream synchronise)
  char* mapped_host = nullptr;
  char* mapped_dev = nullptr;
  unsigned long long seq = 0;
  // launches on caller-supplied streams share partials / tickets: order each enqueue after the previous one
  cudaStream_t last_stream = nullptr;
  bool last_valid = false, last_own = true;
  cudaEvent_t order_ev = nullptr;
  // fused all-reduce over peer memory (XchgArgs in mc_kernels.cuh); world <= 1: not connected
  struct Comm {
    int world = 0, rank = 0;
    char* block = nullptr;            // this rank's exchange block (cudaMalloc; exported by IPC handle)
    char* peer[kMaxRanks] = {};       // peer[rank] == block
    bool ipc_opened[kMaxRanks] = {};  // peers mapped with cudaIpcOpenMemHandle (to be closed)
    unsigned long long epoch = 0;
    unsigned long long timeout_ns = kXchgTimeoutNs;
    uint32_t* launch_ticket = nullptr;
    bool collective = false;          // b200mc_comm_set_collective: every fused launch of a host entry point exchanges
  } comm;
  int plan_split_shift = -1;  // b200mc_set_plan: -1 / 0 = automatic
  uint32_t plan_ppt = 0;
  uint32_t last_plan[3] = {0, 0, 0};  // tiles, paths per thread, split shift of the most recent fused launch
  uint64_t launches = 0;
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // scratch pair for the probes
  static constexpr int kRing = 64;
  cudaEvent_t ring0[kRing] = {}, ring1[kRing] = {};
  uint64_t timed = 0;  // kernels timed since timing was enabled
};

namespace {

int fail(b200mc_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->error = buf;
  else g_create_error = buf;
  return code;
}

#define CU_TRY(e, call)                                                                              \
  do {                                                                                               \
    cudaError_t err__ = (call);                                                                      \
    if (err__ != cudaSuccess)                                                                        \
      return fail(e, err__ == cudaErrorMemoryAllocation ? B200MC_ERR_NOMEM : B200MC_ERR_CUDA,         \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__);    \
  } while (0)

int reserve(b200mc_engine* e, DeviceBuffer& b, size_t bytes) {
  if (b.bytes >= bytes) return 0;
  if (b.ptr) {
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    CU_TRY(e, cudaFree(b.ptr));
    b.ptr = nullptr, b.bytes = 0;
  }
  const size_t want = std::max(bytes, (size_t)1 << 16);
  CU_TRY(e, cudaMalloc(&b.ptr, want));
  b.bytes = want;
  return 0;
}

int reserve_pinned(b200mc_engine* e, size_t bytes) {
  if (e->pinned_bytes >= bytes) return 0;
  if (e->pinned) {
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    CU_TRY(e, cudaFreeHost(e->pinned));
    e->pinned = nullptr, e->pinned_bytes = 0;
  }
  const size_t want = std::max(bytes, (size_t)1 << 16);
  CU_TRY(e, cudaMallocHost(&e->pinned, want));
  e->pinned_bytes = want;
  return 0;
}

int check_spec(b200mc_engine* e, const b200mc_spec_t* s) {
  if (!s) return fail(e, B200MC_ERR_INVALID, "spec is null");
  if (s->kind == B200MC_CLIQUET || s->kind == B200MC_AUTOCALLABLE)
    return fail(e, B200MC_ERR_INVALID, "structured products (kind %d) go through b200mc_simulate_structured", s->kind);
  if (s->kind < B200MC_EUROPEAN || s->kind > B200MC_LOOKBACK) return fail(e, B200MC_ERR_INVALID, "unknown payoff kind %d", s->kind);
  if (s->n_steps == 0) return fail(e, B200MC_ERR_INVALID, "n_steps must be >= 1");
  if (s->antithetic && s->kind != B200MC_EUROPEAN)
    return fail(e, B200MC_ERR_INVALID, "antithetic mirroring is only defined for the European payoff");
  return 0;
}

uint32_t pad_scenarios(uint32_t n) {
  uint32_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

constexpr size_t kMappedRecordBytes = B200MC_MAX_SCENARIOS * sizeof(b200mc_cv_moments_t);  // the largest single-option result
constexpr size_t kMappedBytes = 4096;
constexpr size_t kMappedTimeoutOffset = kMappedRecordBytes + 64;  // the exchange's timed_out word (after the sequence word)
constexpr size_t kXchgSlotBytes = 8u << 20;                       // per epoch parity: header + up to ~174k moment records

// Ticket counters start at zero and every launch leaves them at zero (finish_tile).
// Scratch of the in-kernel fold (finish_tile): tile partials + group totals, and the ticket words.
int reserve_fold(b200mc_engine* e, uint32_t n_opt, uint32_t tiles, uint32_t values_per_tile) {
  if (int rc = reserve(e, e->partials, (size_t)n_opt * fold_scratch_doubles(tiles, values_per_tile) * sizeof(double))) return rc;
  const size_t bytes = (size_t)n_opt * fold_ticket_words(tiles) * sizeof(uint32_t);
  if (e->tickets.bytes >= bytes) return 0;
  if (int rc = reserve(e, e->tickets, bytes)) return rc;
  CU_TRY(e, cudaMemset(e->tickets.ptr, 0, e->tickets.bytes));
  return 0;
}

// Kernels of consecutive calls share the partials / ticket scratch.  Calls on the engine's own stream are ordered by the
// stream; a call on a different stream than the previous one first waits for that previous call.
int order_before(b200mc_engine* e, cudaStream_t stream) {
  if (e->last_valid && e->last_stream != stream) {
    if (e->last_own) CU_TRY(e, cudaEventRecord(e->order_ev, e->stream));  // everything enqueued on our stream so far
    CU_TRY(e, cudaStreamWaitEvent(stream, e->order_ev, 0));
  }
  return 0;
}

int order_after(b200mc_engine* e, cudaStream_t stream) {
  e->last_stream = stream;
  e->last_own = stream == e->stream;
  e->last_valid = true;
  if (!e->last_own) CU_TRY(e, cudaEventRecord(e->order_ev, stream));  // a caller's stream may be gone by the next call: record now
  return 0;
}

// Is this call collective?  (explicitly - b200mc_simulate_allreduce - or because the caller switched the engine to
// collective mode around it; unconnected engines never are.)
bool exchanging(const b200mc_engine* e, bool asked) { return (asked || e->comm.collective) && e->comm.world > 1; }

// The exchange part of a launch's FoldArgs: checks, peer table, the next epoch.  Every rank must make the same calls.
int setup_exchange(b200mc_engine* e, FoldArgs& f, size_t record_bytes) {
  if (record_bytes > kXchgSlotBytes - kXchgHeaderBytes)
    return fail(e, B200MC_ERR_COMM, "%zu bytes of moment records exceed the exchange slot (%zu)", record_bytes, kXchgSlotBytes - kXchgHeaderBytes);
  if (*(volatile unsigned int*)(e->mapped_host + kMappedTimeoutOffset))
    return fail(e, B200MC_ERR_COMM, "an earlier fused all-reduce timed out waiting for a peer rank; reconnect the communicator");
  f.x.world = (uint32_t)e->comm.world;
  f.x.rank = (uint32_t)e->comm.rank;
  f.x.epoch = ++e->comm.epoch;
  for (int r = 0; r < e->comm.world; ++r) f.x.peer[r] = e->comm.peer[r];
  f.x.slot_bytes = kXchgSlotBytes;
  f.x.launch_ticket = e->comm.launch_ticket;
  f.x.timed_out = (unsigned int*)(e->mapped_dev + kMappedTimeoutOffset);
  f.x.timeout_ns = e->comm.timeout_ns;
  return 0;
}

int check_exchange_done(b200mc_engine* e, bool exchange) {  // after the launch has completed
  if (exchange && *(volatile unsigned int*)(e->mapped_host + kMappedTimeoutOffset))
    return fail(e, B200MC_ERR_COMM, "fused all-reduce timed out waiting for a peer rank");
  return 0;
}

// Tile shape.  A tile = one CTA = kBlock threads; a thread owns ppt paths of one option - or, for European launches
// too small to fill the chip that way, 2^split_shift adjacent lanes share each path (european_kernel<SPLIT>).  Large
// problems simply take ppt = kMaxPathsPerThread (the per-CTA prologue/reduction is amortised over 32 paths per thread).
// Small ones (one option, 1e4..1e6 paths) trade three things: CTA overhead (favours few fat CTAs), SM load balance
// (ceil(ctas / n_sm) CTAs on the busiest SM) and latency hiding (an SM needs ~kSaturatingCtas resident CTAs to keep
// the XU pipe fed) plus a drain tail (the last resident CTAs of an SM finish at different times and run
// under-occupied: ~0.3 of one full residency).  Every (split, ppt) is scored and the cheapest wins; the plan depends
// only on (n_opt, n_paths, n_steps, n_scen, SM count), so a given call is reproducible bit for bit.
struct TilePlan {
  uint32_t tiles = 1, ppt = 1, split_shift = 0;
};

TilePlan plan_tiles(int sm_count, int pin_split_shift, uint32_t pin_ppt, uint32_t n_opt, uint64_t n_paths, uint32_t n_steps, uint32_t ns,
                    bool path_dependent, bool may_split) {
  const int resident = ns <= 2 ? 6 : ns == 4 ? 3 : 2;    // CTAs per SM the register budgets below allow
  constexpr double kSaturatingCtas = 3.0;
  const double n_sm = (double)sm_count;
  // issued instructions per path: step loop (+ per-scenario state updates of the path-dependent kinds) + payoff epilogue
  const double per_path_loop = (11.0 + (path_dependent ? 3.5 * ns : 0.0)) * n_steps;
  const double per_path_payoff = 14.0 * ns;
  const double per_cta = 500.0 + 60.0 * ns;              // coefficient set-up + FP64 block reduction
  const uint32_t calls = (n_steps + 7u) >> 3;
  const bool wide = ns >= 8;
  // (only launches of up to 16 single-pass CTAs per SM: beyond that the general model's finer tiles balance better, measured)
  const bool mid_european = (ns == 3 || ns == 4) && !path_dependent && std::ceil((double)n_paths / kBlock) * n_opt <= 16.0 * n_sm;
  double best = 0.0;
  TilePlan plan;
  // Lane split (european_kernel<SPLIT>): measured on B200 (profiles/r02_plan_sweep.jsonl) it pays only while one thread
  // per path leaves most SMs without a CTA - 10k paths: 19.5 -> 15.4 us with 2 lanes per path; from 30k paths up the
  // unsplit launch wins.  So: the smallest split that puts a CTA on at least half of the SMs, none otherwise.
  uint32_t auto_shift = 0;
  if (may_split) {
    const double base_ctas = std::ceil((double)n_paths / kBlock) * n_opt;
    while ((1u << (auto_shift + 1)) <= (uint32_t)kMaxSplit && calls >= (4u << auto_shift) && base_ctas * (1u << auto_shift) < 0.5 * n_sm)
      ++auto_shift;
  }
  uint32_t max_shift = 0;
  if (may_split)
    while ((1u << (max_shift + 1)) <= (uint32_t)kMaxSplit && calls >= (4u << max_shift)) ++max_shift;  // >= 2 calls per lane
  const uint32_t lo = pin_split_shift >= 0 ? std::min<uint32_t>((uint32_t)pin_split_shift, max_shift) : auto_shift;
  const uint32_t hi = lo;
  for (uint32_t shift = lo; shift <= hi; ++shift) {
    const uint32_t lanes = 1u << shift;
    const uint64_t per_pass = (uint64_t)kBlock >> shift;
    const double per_path = per_path_loop / lanes + per_path_payoff + 8.0 * shift;  // + the butterfly
    for (uint32_t p = 1; p <= (uint32_t)kMaxPathsPerThread; ++p) {
      if (pin_ppt && p != pin_ppt) continue;
      const uint64_t t = (n_paths + per_pass * p - 1) / (per_pass * p);
      const uint64_t p_even = (n_paths + per_pass * t - 1) / (per_pass * t);  // spread evenly over t tiles
      if (p_even != p && !pin_ppt) continue;                                  // same tiling as a smaller p
      const double ctas = (double)t * n_opt;
      const double per_sm = std::ceil(ctas / n_sm);
      const double concurrency = std::min(per_sm, (double)resident);
      const double eff = std::min(1.0, concurrency / kSaturatingCtas);
      double cost = (per_sm + 0.3 * resident) * ((double)p * per_path + per_cta) / eff;
      if (wide && shift == 0) {
        // 8-16 scenario launches (two 8-warp CTAs per SM), measured on B200 (profiles/r02_experiments.txt #10,
        // profiles/r02_wide_plan_check.jsonl): a lone CTA runs a pass in w; each of two co-resident CTAs takes 1.68 w in the
        // European kernel (two CTAs = 1.19x the throughput of one) and 1.92 w in the path-dependent kinds (14 independent
        // update chains per thread saturate the FMA pipe from 8 warps: co-residency buys 4 %); every CTA costs ~3 us of
        // prologue + FP64 reduction + ticket.  The busiest SM works through its per_sm CTAs two at a time, a left-over one alone.
        const double pairs = std::floor(per_sm / 2.0), lone = per_sm - 2.0 * pairs;
        cost = (double)p * per_path * ((path_dependent ? 1.92 : 1.68) * pairs + lone) + per_cta * (pairs + lone);
      }
      if (mid_european && shift == 0) {
        // 3-4 scenario European launches (70 registers: three 8-warp CTAs per SM), same measurements: two co-resident CTAs take
        // 1.6 w each, three 2.45 w - the third buys nothing - so the best shapes put two CTAs on an SM
        const double triples = std::floor(per_sm / 3.0), rem = per_sm - 3.0 * triples;
        cost = (double)p * per_path * (2.45 * triples + (rem == 2.0 ? 1.6 : rem)) + per_cta * (triples + (rem > 0.0 ? 1.0 : 0.0));
      }
      if (best == 0.0 || cost < best * 0.999) best = cost, plan.ppt = p, plan.tiles = (uint32_t)t, plan.split_shift = shift;
    }
  }
  return plan;
}

TilePlan plan_tiles(const b200mc_engine* e, uint32_t n_opt, uint64_t n_paths, uint32_t n_steps, uint32_t ns, bool path_dependent,
                    bool may_split = false) {
  return plan_tiles(e->prop.multiProcessorCount, e->plan_split_shift, e->plan_ppt, n_opt, n_paths, n_steps, ns, path_dependent, may_split);
}

// __launch_bounds__ minBlocksPerSM per kernel family, picked from measurements on B200
// (profiles/r01_variants.txt): the schedulers want the three Philox chains of a superblock
// interleaved, which needs ~60-75 registers; squeezing below 48 costs 5-10%.
#ifndef B200MC_SMALL_MINB
#define B200MC_SMALL_MINB 6
#endif
#ifndef B200MC_SMALL_UNROLL
#define B200MC_SMALL_UNROLL 1
#endif
constexpr int kMinBlocksSmall = B200MC_SMALL_MINB;  // European, 1-2 scenarios (40-47 registers)
constexpr int kMinBlocksPathdep = 6;  // Asian / barrier / lookback, 1-2 scenarios (<= 40 registers)
constexpr int kMinBlocksAsian = 6;    // arithmetic Asian, 1-2 scenarios (profiles/r01_variants14*)
constexpr int kMinBlocksStructured = 5;  // cliquet / autocallable, 1-2 scenarios (47-48 registers)
#ifndef B200MC_WIDE_MINB
#define B200MC_WIDE_MINB 2
#endif
#ifndef B200MC_WIDE_UNROLL
#define B200MC_WIDE_UNROLL 1
#endif
constexpr int kMinBlocksWide = B200MC_WIDE_MINB;  // 4-16 scenarios (<= 128 registers)
constexpr int kUnrollWide = B200MC_WIDE_UNROLL;   // Philox calls in flight per thread of the 4-16 scenario European launches

template <int NS>
cudaError_t launch_european(const SimArgs& a, bool anti, bool cv, dim3 grid, cudaStream_t s) {
  constexpr int kMinBlocks = NS <= 2 ? kMinBlocksSmall : kMinBlocksWide;
  if (cv) {
    if (anti) european_kernel<NS, true, kMinBlocksWide, true><<<grid, kBlock, 0, s>>>(a);
    else european_kernel<NS, false, kMinBlocksWide, true><<<grid, kBlock, 0, s>>>(a);
  } else if (a.split_shift) {  // under-filled launches: 2^split_shift lanes per path
    if (anti) european_kernel<NS, true, kMinBlocks, false, 1, true><<<grid, kBlock, 0, s>>>(a);
    else european_kernel<NS, false, kMinBlocks, false, 1, true><<<grid, kBlock, 0, s>>>(a);
  } else {
    constexpr int kUnroll = NS <= 2 ? B200MC_SMALL_UNROLL : kUnrollWide;
    if (anti) european_kernel<NS, true, kMinBlocks, false, kUnroll><<<grid, kBlock, 0, s>>>(a);
    else european_kernel<NS, false, kMinBlocks, false, kUnroll><<<grid, kBlock, 0, s>>>(a);
  }
  return cudaGetLastError();
}

template <int KIND>
cudaError_t launch_pathdep(const SimArgs& a, uint32_t ns, dim3 grid, cudaStream_t s) {
  constexpr int kMinB = KIND == B200MC_ASIAN_ARITH ? kMinBlocksAsian : kMinBlocksPathdep;
  switch (ns) {
    case 1: pathdep_kernel<KIND, 1, kMinB><<<grid, kBlock, 0, s>>>(a); break;
    case 2: pathdep_kernel<KIND, 2, kMinB><<<grid, kBlock, 0, s>>>(a); break;
    // 4+ scenarios: the scalar Horner form (immediates instead of 10 registers of packed coefficients; same roundings,
    // so a scenario's bits do not depend on which form its launch used)
    case 4: pathdep_kernel<KIND, 4, kMinBlocksWide, 1, 4><<<grid, kBlock, 0, s>>>(a); break;
    case 8: pathdep_kernel<KIND, 8, kMinBlocksWide, 1, 4><<<grid, kBlock, 0, s>>>(a); break;
    default: pathdep_kernel<KIND, 16, kMinBlocksWide, 1, 4><<<grid, kBlock, 0, s>>>(a); break;
  }
  return cudaGetLastError();
}

// ---- structured products: argument checks and launch dispatch ---------------------------------------------------
int check_structured(b200mc_engine* e, const b200mc_spec_t* spec, const b200mc_product_t* product, uint32_t* period, uint32_t* n_events,
                     uint32_t* sim_steps) {
  if (!spec || !product) return fail(e, B200MC_ERR_INVALID, "spec / product is null");
  if (spec->kind != B200MC_CLIQUET && spec->kind != B200MC_AUTOCALLABLE)
    return fail(e, B200MC_ERR_INVALID, "kind %d is not a structured product", spec->kind);
  if (spec->n_steps == 0) return fail(e, B200MC_ERR_INVALID, "n_steps must be >= 1");
  if (product->period == 0) return fail(e, B200MC_ERR_INVALID, "n_periods / observation_freq must be >= 1");
  if (spec->kind == B200MC_CLIQUET) {
    *n_events = product->period;                 // n_periods
    *period = spec->n_steps / product->period;   // steps per period (exotic_options.py:532)
    if (*period == 0) return fail(e, B200MC_ERR_INVALID, "more cliquet periods (%u) than steps (%u): every return is 0", product->period, spec->n_steps);
    *sim_steps = *period * *n_events;
  } else {
    *period = product->period;                   // observation_freq
    *n_events = spec->n_steps / product->period; // len(range(freq, n_steps + 1, freq))
    *sim_steps = spec->n_steps;
  }
  return 0;
}

template <int KIND>
cudaError_t launch_structured(const StructuredArgs& g, uint32_t ns, dim3 grid, cudaStream_t s) {
  switch (ns) {
    case 1: structured_kernel<KIND, 1, kMinBlocksStructured><<<grid, kBlock, 0, s>>>(g); break;
    case 2: structured_kernel<KIND, 2, kMinBlocksStructured><<<grid, kBlock, 0, s>>>(g); break;
    case 4: structured_kernel<KIND, 4, kMinBlocksWide><<<grid, kBlock, 0, s>>>(g); break;
    case 8: structured_kernel<KIND, 8, kMinBlocksWide><<<grid, kBlock, 0, s>>>(g); break;
    default: structured_kernel<KIND, 16, kMinBlocksWide><<<grid, kBlock, 0, s>>>(g); break;
  }
  return cudaGetLastError();
}

// Enqueue the fused simulation on `stream`.  params_dev != null: parameters and `out` are device pointers.  params_dev ==
// null (single-option launches from the host entry points): the block rides in the kernel arguments (inline_params) and
// the finishing CTA writes the records into the mapped pinned block, then publishes `seq` for the polling host.
int enqueue_simulation(b200mc_engine* e, const b200mc_spec_t* spec, const b200mc_params_t* params_dev, const b200mc_params_t* inline_params,
                       uint32_t n_opt, uint32_t n_scen, uint64_t seed, uint32_t stream_base, uint64_t path_begin, uint64_t n_paths,
                       void* out_dev, cudaStream_t stream, bool time_it, bool cv = false, bool allreduce = false) {
  if (int rc = check_spec(e, spec)) return rc;
  if ((!params_dev && !inline_params) || !out_dev) return fail(e, B200MC_ERR_INVALID, "params/out pointer is null");
  const bool exchange = exchanging(e, allreduce);
  // a rank whose share of the paths is empty still takes part in the exchange (it contributes zero records)
  if (n_opt == 0 || (n_paths == 0 && !exchange)) return fail(e, B200MC_ERR_INVALID, "n_opt and n_paths must be >= 1");
  if (n_scen == 0 || n_scen > B200MC_MAX_SCENARIOS)
    return fail(e, B200MC_ERR_INVALID, "n_scen must be in [1, %d]", B200MC_MAX_SCENARIOS);
  if (inline_params && n_opt != 1) return fail(e, B200MC_ERR_INVALID, "inline parameters carry one option");

  if (cv && spec->kind != B200MC_EUROPEAN) return fail(e, B200MC_ERR_INVALID, "the control variate is defined for the European payoff only");
  const uint32_t ns = pad_scenarios(n_scen);
  TilePlan plan = plan_tiles(e, n_opt, std::max<uint64_t>(n_paths, 1), spec->n_steps, ns, spec->kind != B200MC_EUROPEAN, spec->kind == B200MC_EUROPEAN && !cv);
  const uint64_t ctas = (uint64_t)plan.tiles * n_opt;
  if (ctas > 0x7fffffffull) return fail(e, B200MC_ERR_INVALID, "problem too large for one launch (%llu CTAs)", (unsigned long long)ctas);
  e->last_plan[0] = plan.tiles, e->last_plan[1] = plan.ppt, e->last_plan[2] = plan.split_shift;
  if (int rc = reserve_fold(e, n_opt, plan.tiles, (cv ? 6 : 3) * ns)) return rc;  // sums + one paid-count per scenario
  if (int rc = order_before(e, stream)) return rc;

  SimArgs a{};
  a.params = params_dev;
  if (inline_params) {  // the host evaluates what thread k of every CTA's prologue (and the finishing CTA) would
    const bool negate = (spec->kind == B200MC_BARRIER && spec->barrier_down) ||
                        (spec->kind == B200MC_LOOKBACK && (spec->lookback_fixed ? spec->is_put : !spec->is_put));
    for (uint32_t k = 0; k < ns; ++k) {
      const b200mc_params_t& p = inline_params[k < n_scen ? k : n_scen - 1];
      const Coef q = make_coef(p, spec->n_steps, negate ? -1.0f : 1.0f);
      a.inl.c[k] = q.c, a.inl.d[k] = q.d, a.inl.a[k] = q.a, a.inl.kappa[k] = q.kappa, a.inl.beta[k] = q.beta, a.inl.inv_n[k] = q.inv_n;
      a.inl.spot[k] = p.S, a.inl.kappa64[k] = p.K / p.S;
    }
  }
  a.fold.partials = (double*)e->partials.ptr;
  a.fold.tickets = (uint32_t*)e->tickets.ptr;
  a.fold.out = out_dev;
  a.fold.samples = (double)n_paths * (spec->antithetic ? 2.0 : 1.0);
  a.fold.n_opt = n_opt;
  if (inline_params) {
    a.fold.done = (unsigned long long*)(e->mapped_dev + kMappedRecordBytes);
    a.fold.seq = ++e->seq;
  }
  if (exchange)
    if (int rc = setup_exchange(e, a.fold, (size_t)n_opt * n_scen * (cv ? sizeof(b200mc_cv_moments_t) : sizeof(b200mc_moments_t)))) return rc;
  a.path_begin = path_begin;
  a.n_paths = n_paths;
  a.n_opt = n_opt;
  a.n_scen = n_scen;
  a.tiles = plan.tiles;
  a.paths_per_thread = plan.ppt;
  a.split_shift = plan.split_shift;
  a.n_steps = spec->n_steps;
  a.rk = philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32));
  a.stream_base = stream_base;
  a.is_put = spec->is_put;
  a.barrier_in = spec->barrier_in;
  a.lookback_fixed = spec->lookback_fixed;
  a.force_mufu_ex2 = (spec->flags & B200MC_FLAG_EXACT_EX2) ? 1 : 0;
  // which extremum the path-dependent kernels track (they follow sgn * log2(S_t/S_0) and keep its max)
  a.sgn_negative = 0;
  if (spec->kind == B200MC_BARRIER) a.sgn_negative = spec->barrier_down ? 1 : 0;
  if (spec->kind == B200MC_LOOKBACK) a.sgn_negative = (spec->lookback_fixed ? spec->is_put : !spec->is_put) ? 1 : 0;

  const dim3 grid((unsigned)ctas);
  const int slot = (int)(e->timed % b200mc_engine::kRing);
  if (time_it) CU_TRY(e, cudaEventRecord(e->ring0[slot], stream));
  cudaError_t err = cudaSuccess;
  switch (spec->kind) {
    case B200MC_EUROPEAN:
      switch (ns) {
        case 1: err = launch_european<1>(a, spec->antithetic, cv, grid, stream); break;
        case 2: err = launch_european<2>(a, spec->antithetic, cv, grid, stream); break;
        case 4: err = launch_european<4>(a, spec->antithetic, cv, grid, stream); break;
        case 8: err = launch_european<8>(a, spec->antithetic, cv, grid, stream); break;
        default: err = launch_european<16>(a, spec->antithetic, cv, grid, stream); break;
      }
      break;
    case B200MC_ASIAN_ARITH: err = launch_pathdep<B200MC_ASIAN_ARITH>(a, ns, grid, stream); break;
    case B200MC_ASIAN_GEOM: err = launch_pathdep<B200MC_ASIAN_GEOM>(a, ns, grid, stream); break;
    case B200MC_BARRIER: err = launch_pathdep<B200MC_BARRIER>(a, ns, grid, stream); break;
    case B200MC_LOOKBACK: err = launch_pathdep<B200MC_LOOKBACK>(a, ns, grid, stream); break;
  }
  if (err != cudaSuccess) return fail(e, B200MC_ERR_CUDA, "simulation kernel launch failed: %s", cudaGetErrorString(err));
  if (time_it) {
    CU_TRY(e, cudaEventRecord(e->ring1[slot], stream));
    e->timed += 1;
  }
  e->launches += 1;
  return order_after(e, stream);
}

// Wait for the finishing CTA of launch `seq` to publish its records in the mapped block.  Polling host memory sees the
// result ~1 us after the kernel's last store; a stream synchronise costs several.  Every few thousand polls the stream is
// queried so that a failed launch surfaces as an error instead of a hang.
int wait_mapped(b200mc_engine* e, unsigned long long seq) {
  const volatile unsigned long long* flag = (const volatile unsigned long long*)(e->mapped_host + kMappedRecordBytes);
  for (uint32_t spin = 0;; ++spin) {
    if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) == seq) return 0;
    if ((spin & 0x3fffu) == 0x3fffu) {
      const cudaError_t q = cudaStreamQuery(e->stream);
      if (q == cudaSuccess) return __atomic_load_n(flag, __ATOMIC_ACQUIRE) == seq ? 0 : fail(e, B200MC_ERR_CUDA, "kernel finished without publishing its result");
      if (q != cudaErrorNotReady) return fail(e, B200MC_ERR_CUDA, "simulation kernel failed: %s", cudaGetErrorString(q));
    }
  }
}

int enqueue_from_normals(b200mc_engine* e, const b200mc_spec_t* spec, const b200mc_params_t* p, int accumulate,
                         const double* Z_dev, uint64_t n_paths, double* payoffs_dev, b200mc_moments_t* out_dev,
                         cudaStream_t stream, bool time_it) {
  if (int rc = check_spec(e, spec)) return rc;
  if (!p || !Z_dev || !out_dev) return fail(e, B200MC_ERR_INVALID, "null pointer argument");
  if (n_paths == 0) return fail(e, B200MC_ERR_INVALID, "n_paths must be >= 1");
  const uint64_t ctas = (n_paths + kF64Block - 1) / kF64Block;
  if (ctas > 0x7fffffffull) return fail(e, B200MC_ERR_INVALID, "too many paths for one launch");
  if (int rc = reserve(e, e->partials, ctas * 2 * sizeof(double))) return rc;
  if (int rc = order_before(e, stream)) return rc;
  F64Args a{};
  a.Z = Z_dev;
  a.payoffs = payoffs_dev;
  a.partials = (double*)e->partials.ptr;
  a.n_paths = n_paths;
  a.n_steps = spec->n_steps;
  a.accumulate = (accumulate || spec->kind != B200MC_EUROPEAN) ? 1 : 0;
  a.antithetic = spec->antithetic;
  a.is_put = spec->is_put;
  a.barrier_down = spec->barrier_down;
  a.barrier_in = spec->barrier_in;
  a.lookback_fixed = spec->lookback_fixed;
  a.S = p->S, a.K = p->K, a.T = p->T, a.r = p->r, a.sigma = p->sigma, a.q = p->q, a.barrier = p->barrier;
  const dim3 grid((unsigned)ctas);
  const int slot = (int)(e->timed % b200mc_engine::kRing);
  if (time_it) CU_TRY(e, cudaEventRecord(e->ring0[slot], stream));
  switch (spec->kind) {
    case B200MC_EUROPEAN:
      // bulk-async staged kernel when the rows can be copied in 16-byte units, else the plain-load kernel
      if ((spec->n_steps & 1u) == 0 && ((uintptr_t)Z_dev & 15u) == 0 && !(spec->flags & B200MC_FLAG_NO_BULK_COPY)) {
        static_assert(kTmaSmemBytes <= 227 * 1024, "tile ring exceeds shared memory");
        CU_TRY(e, cudaFuncSetAttribute(european_from_normals_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes));
        european_from_normals_tma_kernel<<<grid, kF64Block, kTmaSmemBytes, stream>>>(a);
      } else {
        from_normals_kernel<B200MC_EUROPEAN><<<grid, kF64Block, 0, stream>>>(a);
      }
      break;
    case B200MC_ASIAN_ARITH: from_normals_kernel<B200MC_ASIAN_ARITH><<<grid, kF64Block, 0, stream>>>(a); break;
    case B200MC_ASIAN_GEOM: from_normals_kernel<B200MC_ASIAN_GEOM><<<grid, kF64Block, 0, stream>>>(a); break;
    case B200MC_BARRIER: from_normals_kernel<B200MC_BARRIER><<<grid, kF64Block, 0, stream>>>(a); break;
    case B200MC_LOOKBACK: from_normals_kernel<B200MC_LOOKBACK><<<grid, kF64Block, 0, stream>>>(a); break;
  }
  CU_TRY(e, cudaGetLastError());
  if (time_it) {
    CU_TRY(e, cudaEventRecord(e->ring1[slot], stream));
    e->timed += 1;
  }
  const double samples = (double)n_paths * (spec->antithetic ? 2.0 : 1.0);
  fold_kernel<<<1, 32, 0, stream>>>((const double*)e->partials.ptr, out_dev, (uint32_t)ctas, samples);
  CU_TRY(e, cudaGetLastError());
  e->launches += 2;
  return order_after(e, stream);
}

}  // namespace

extern "C" {

int b200mc_abi_version(void) { return B200MC_ABI_VERSION; }

int b200mc_create(b200mc_engine_t** out, int device) {
  std::lock_guard<std::mutex> g(g_create_mutex);
  if (!out) return fail(nullptr, B200MC_ERR_INVALID, "out pointer is null");
  *out = nullptr;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0)
    return fail(nullptr, B200MC_ERR_CUDA, "no CUDA device available: %s", err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
  if (device < 0 || device >= count) return fail(nullptr, B200MC_ERR_INVALID, "device %d out of range [0, %d)", device, count);
  b200mc_engine* e = new (std::nothrow) b200mc_engine();
  if (!e) return fail(nullptr, B200MC_ERR_NOMEM, "host allocation failed");
  e->device = device;
  auto bail = [&](cudaError_t er, const char* what) {
    fail(nullptr, B200MC_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(er));
    delete e;
    return (int)B200MC_ERR_CUDA;
  };
  if ((err = cudaSetDevice(device)) != cudaSuccess) return bail(err, "cudaSetDevice");
  if ((err = cudaGetDeviceProperties(&e->prop, device)) != cudaSuccess) return bail(err, "cudaGetDeviceProperties");
  if (e->prop.major < 10) {
    fail(nullptr, B200MC_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device, e->prop.major, e->prop.minor);
    delete e;
    return B200MC_ERR_CUDA;
  }
  if ((err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(err, "cudaStreamCreate");
  if ((err = cudaEventCreate(&e->ev0)) != cudaSuccess) return bail(err, "cudaEventCreate");
  if ((err = cudaEventCreate(&e->ev1)) != cudaSuccess) return bail(err, "cudaEventCreate");
  for (int i = 0; i < b200mc_engine::kRing; ++i) {
    if ((err = cudaEventCreate(&e->ring0[i])) != cudaSuccess) return bail(err, "cudaEventCreate");
    if ((err = cudaEventCreate(&e->ring1[i])) != cudaSuccess) return bail(err, "cudaEventCreate");
  }
  if ((err = cudaEventCreateWithFlags(&e->order_ev, cudaEventDisableTiming)) != cudaSuccess) return bail(err, "cudaEventCreate");
  if ((err = cudaHostAlloc((void**)&e->mapped_host, kMappedBytes, cudaHostAllocMapped)) != cudaSuccess) return bail(err, "cudaHostAlloc");
  memset(e->mapped_host, 0, kMappedBytes);
  if ((err = cudaHostGetDevicePointer((void**)&e->mapped_dev, e->mapped_host, 0)) != cudaSuccess) return bail(err, "cudaHostGetDevicePointer");
  *out = e;
  return 0;
}

void b200mc_destroy(b200mc_engine_t* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (DeviceBuffer* b : {&e->partials, &e->tickets, &e->params_dev, &e->moments_dev, &e->scratch_a, &e->scratch_b, &e->qmc_wpart, &e->qmc_tickets})
    if (b->ptr) cudaFree(b->ptr);
  for (int r = 0; r < kMaxRanks; ++r)
    if (e->comm.ipc_opened[r] && e->comm.peer[r]) cudaIpcCloseMemHandle(e->comm.peer[r]);
  if (e->comm.block) cudaFree(e->comm.block);
  if (e->comm.launch_ticket) cudaFree(e->comm.launch_ticket);
  if (e->pinned) cudaFreeHost(e->pinned);
  if (e->mapped_host) cudaFreeHost(e->mapped_host);
  if (e->order_ev) cudaEventDestroy(e->order_ev);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  for (int i = 0; i < b200mc_engine::kRing; ++i) {
    if (e->ring0[i]) cudaEventDestroy(e->ring0[i]);
    if (e->ring1[i]) cudaEventDestroy(e->ring1[i]);
  }
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

const char* b200mc_last_error(const b200mc_engine_t* e) { return e ? e->error.c_str() : g_create_error.c_str(); }

int b200mc_device_info(b200mc_engine_t* e, b200mc_info_t* out) {
  if (!e || !out) return fail(e, B200MC_ERR_INVALID, "null argument");
  memset(out, 0, sizeof *out);
  out->device = e->device;
  out->sm_count = e->prop.multiProcessorCount;
  out->cc_major = e->prop.major;
  out->cc_minor = e->prop.minor;
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, e->device);
  out->sm_clock_khz = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMemoryClockRate, e->device);
  out->mem_clock_khz = v;
  out->total_mem_bytes = (int64_t)e->prop.totalGlobalMem;
  out->l2_bytes = e->prop.l2CacheSize;
  strncpy(out->name, e->prop.name, sizeof(out->name) - 1);
  return 0;
}

static int simulate_device_impl(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* params_dev, uint32_t n_opt,
                                uint32_t n_scen, uint64_t seed, uint32_t stream_base, uint64_t path_begin, uint64_t n_paths,
                                b200mc_moments_t* out_dev, void* cuda_stream, bool allreduce) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  CU_TRY(e, cudaSetDevice(e->device));
  if (!params_dev) return fail(e, B200MC_ERR_INVALID, "params pointer is null");
  return enqueue_simulation(e, spec, params_dev, nullptr, n_opt, n_scen, seed, stream_base, path_begin, n_paths, out_dev,
                            (cudaStream_t)cuda_stream, e->timing, false, allreduce);
}

int b200mc_simulate_device(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* params_dev, uint32_t n_opt,
                           uint32_t n_scen, uint64_t seed, uint32_t stream_base, uint64_t path_begin, uint64_t n_paths,
                           b200mc_moments_t* out_dev, void* cuda_stream) {
  return simulate_device_impl(e, spec, params_dev, n_opt, n_scen, seed, stream_base, path_begin, n_paths, out_dev, cuda_stream, false);
}

int b200mc_simulate_allreduce_device(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* params_dev, uint32_t n_opt,
                                     uint32_t n_scen, uint64_t seed, uint32_t stream_base, uint64_t path_begin, uint64_t n_paths,
                                     b200mc_moments_t* out_dev, void* cuda_stream) {
  return simulate_device_impl(e, spec, params_dev, n_opt, n_scen, seed, stream_base, path_begin, n_paths, out_dev, cuda_stream, true);
}

static int simulate_host(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* params_host, uint32_t n_opt,
                         uint32_t n_scen, uint64_t seed, uint32_t stream_base, uint64_t path_begin, uint64_t n_paths,
                         void* out_host, bool cv, bool allreduce = false) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (!params_host || !out_host) return fail(e, B200MC_ERR_INVALID, "params/out pointer is null");
  if (n_opt == 0 || n_scen == 0 || n_scen > B200MC_MAX_SCENARIOS) return fail(e, B200MC_ERR_INVALID, "bad n_opt / n_scen");
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t n = (size_t)n_opt * n_scen;
  const size_t in_bytes = n * sizeof(b200mc_params_t), out_bytes = n * (cv ? sizeof(b200mc_cv_moments_t) : sizeof(b200mc_moments_t));
  if (n_opt == 1) {  // one launch, nothing else: parameters in the kernel arguments, records through mapped memory
    if (int rc = enqueue_simulation(e, spec, nullptr, params_host, 1, n_scen, seed, stream_base, path_begin, n_paths, e->mapped_dev,
                                    e->stream, e->timing, cv, allreduce))
      return rc;
    if (int rc = wait_mapped(e, e->seq)) return rc;
    if (int rc = check_exchange_done(e, exchanging(e, allreduce))) return rc;
    memcpy(out_host, e->mapped_host, out_bytes);
    return 0;
  }
  if (int rc = reserve(e, e->params_dev, in_bytes)) return rc;
  if (int rc = reserve(e, e->moments_dev, out_bytes)) return rc;
  if (int rc = reserve_pinned(e, in_bytes + out_bytes)) return rc;
  char* pin_in = (char*)e->pinned;
  char* pin_out = pin_in + in_bytes;
  memcpy(pin_in, params_host, in_bytes);
  CU_TRY(e, cudaMemcpyAsync(e->params_dev.ptr, pin_in, in_bytes, cudaMemcpyHostToDevice, e->stream));
  if (int rc = enqueue_simulation(e, spec, (const b200mc_params_t*)e->params_dev.ptr, nullptr, n_opt, n_scen, seed, stream_base,
                                  path_begin, n_paths, e->moments_dev.ptr, e->stream, e->timing, cv, allreduce))
    return rc;
  CU_TRY(e, cudaMemcpyAsync(pin_out, e->moments_dev.ptr, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  if (int rc = check_exchange_done(e, exchanging(e, allreduce))) return rc;
  memcpy(out_host, pin_out, out_bytes);
  return 0;
}

int b200mc_simulate(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* params_host, uint32_t n_opt,
                    uint32_t n_scen, uint64_t seed, uint32_t stream_base, uint64_t path_begin, uint64_t n_paths,
                    b200mc_moments_t* out_host) {
  return simulate_host(e, spec, params_host, n_opt, n_scen, seed, stream_base, path_begin, n_paths, out_host, false);
}

int b200mc_simulate_allreduce(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* params_host, uint32_t n_opt,
                              uint32_t n_scen, uint64_t seed, uint32_t stream_base, uint64_t path_begin, uint64_t n_paths,
                              b200mc_moments_t* out_host) {
  return simulate_host(e, spec, params_host, n_opt, n_scen, seed, stream_base, path_begin, n_paths, out_host, false, true);
}

// ---- communicator of the fused all-reduce -------------------------------------------------------------------------
static int comm_alloc(b200mc_engine_t* e) {
  CU_TRY(e, cudaSetDevice(e->device));
  if (!e->comm.block) {
    CU_TRY(e, cudaMalloc((void**)&e->comm.block, 2 * kXchgSlotBytes));
    CU_TRY(e, cudaMalloc((void**)&e->comm.launch_ticket, 256));
    CU_TRY(e, cudaMemset(e->comm.launch_ticket, 0, 256));
  }
  return 0;
}

static int comm_reset(b200mc_engine_t* e) {  // forget the peers; flags back to zero, epochs restart at 1
  CU_TRY(e, cudaSetDevice(e->device));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  for (int r = 0; r < kMaxRanks; ++r) {
    if (e->comm.ipc_opened[r] && e->comm.peer[r]) cudaIpcCloseMemHandle(e->comm.peer[r]);
    e->comm.peer[r] = nullptr, e->comm.ipc_opened[r] = false;
  }
  e->comm.world = 0, e->comm.rank = 0, e->comm.epoch = 0, e->comm.collective = false;
  if (e->comm.block) {
    CU_TRY(e, cudaMemset(e->comm.block, 0, kXchgHeaderBytes));
    CU_TRY(e, cudaMemset(e->comm.block + kXchgSlotBytes, 0, kXchgHeaderBytes));
    CU_TRY(e, cudaMemset(e->comm.launch_ticket, 0, 256));
  }
  *(volatile unsigned int*)(e->mapped_host + kMappedTimeoutOffset) = 0;
  return 0;
}

int b200mc_comm_export(b200mc_engine_t* e, void* handle_out) {
  if (!e || !handle_out) return fail(e, B200MC_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mutex);
  static_assert(sizeof(cudaIpcMemHandle_t) == B200MC_COMM_HANDLE_BYTES, "handle size");
  if (int rc = comm_alloc(e)) return rc;
  if (int rc = comm_reset(e)) return rc;
  cudaIpcMemHandle_t h;
  const cudaError_t err = cudaIpcGetMemHandle(&h, e->comm.block);
  if (err != cudaSuccess) return fail(e, B200MC_ERR_COMM, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(err));
  memcpy(handle_out, &h, sizeof h);
  return 0;
}

int b200mc_comm_connect(b200mc_engine_t* e, int rank, int world, const void* handles) {
  if (!e || !handles) return fail(e, B200MC_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mutex);
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world)
    return fail(e, B200MC_ERR_INVALID, "rank %d / world %d out of range (at most %d ranks: one NVSwitch domain)", rank, world, kMaxRanks);
  if (!e->comm.block) return fail(e, B200MC_ERR_INVALID, "call b200mc_comm_export first");
  CU_TRY(e, cudaSetDevice(e->device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      e->comm.peer[r] = e->comm.block;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * sizeof h, sizeof h);
    void* p = nullptr;
    const cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      comm_reset(e);
      return fail(e, B200MC_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(err));
    }
    e->comm.peer[r] = (char*)p, e->comm.ipc_opened[r] = true;
  }
  e->comm.world = world, e->comm.rank = rank, e->comm.epoch = 0;
  return 0;
}

int b200mc_comm_connect_local(b200mc_engine_t* const* engines, int n) {
  if (!engines || n < 1 || n > kMaxRanks) return B200MC_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    if (!engines[i]) return B200MC_ERR_INVALID;
    std::lock_guard<std::mutex> g(engines[i]->mutex);
    if (int rc = comm_alloc(engines[i])) return rc;
    if (int rc = comm_reset(engines[i])) return rc;
  }
  for (int i = 0; i < n; ++i) {
    b200mc_engine_t* e = engines[i];
    std::lock_guard<std::mutex> g(e->mutex);
    CU_TRY(e, cudaSetDevice(e->device));
    for (int j = 0; j < n; ++j) {
      if (j != i && engines[j]->device != e->device) {
        int can = 0;
        CU_TRY(e, cudaDeviceCanAccessPeer(&can, e->device, engines[j]->device));
        if (!can) return fail(e, B200MC_ERR_COMM, "device %d cannot access device %d", e->device, engines[j]->device);
        const cudaError_t err = cudaDeviceEnablePeerAccess(engines[j]->device, 0);
        if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled)
          return fail(e, B200MC_ERR_COMM, "cudaDeviceEnablePeerAccess(%d) failed: %s", engines[j]->device, cudaGetErrorString(err));
        cudaGetLastError();
      }
      e->comm.peer[j] = engines[j]->comm.block;
    }
    e->comm.world = n, e->comm.rank = i, e->comm.epoch = 0;
  }
  return 0;
}

int b200mc_comm_disconnect(b200mc_engine_t* e) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  return comm_reset(e);
}

int b200mc_comm_world(const b200mc_engine_t* e) { return e ? e->comm.world : 0; }

int b200mc_comm_set_collective(b200mc_engine_t* e, int on) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  e->comm.collective = on != 0;
  return 0;
}

int b200mc_comm_set_timeout_ms(b200mc_engine_t* e, uint32_t milliseconds) {
  if (!e || milliseconds == 0) return fail(e, B200MC_ERR_INVALID, "timeout must be >= 1 ms");
  std::lock_guard<std::mutex> g(e->mutex);
  e->comm.timeout_ns = (unsigned long long)milliseconds * 1000000ull;
  return 0;
}

int b200mc_simulate_control_variate(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* params_host,
                                    uint32_t n_opt, uint32_t n_scen, uint64_t seed, uint32_t stream_base, uint64_t path_begin,
                                    uint64_t n_paths, b200mc_cv_moments_t* out_host) {
  return simulate_host(e, spec, params_host, n_opt, n_scen, seed, stream_base, path_begin, n_paths, out_host, true);
}

// ---- structured products ---------------------------------------------------------------------------------------
int b200mc_simulate_structured(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_product_t* product,
                               const b200mc_params_t* params_host, uint32_t n_opt, uint32_t n_scen, uint64_t seed, uint32_t stream_base,
                               uint64_t path_begin, uint64_t n_paths, b200mc_moments_t* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  uint32_t period, n_events, sim_steps;
  if (int rc = check_structured(e, spec, product, &period, &n_events, &sim_steps)) return rc;
  if (!params_host || !out_host) return fail(e, B200MC_ERR_INVALID, "params/out pointer is null");
  const bool exchange = exchanging(e, false);
  if (n_opt == 0 || (n_paths == 0 && !exchange) || n_scen == 0 || n_scen > B200MC_MAX_SCENARIOS) return fail(e, B200MC_ERR_INVALID, "bad n_opt / n_scen / n_paths");
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t n = (size_t)n_opt * n_scen;
  const size_t in_bytes = n * sizeof(b200mc_params_t), out_bytes = n * sizeof(b200mc_moments_t);
  if (int rc = reserve(e, e->params_dev, in_bytes)) return rc;
  if (int rc = reserve(e, e->moments_dev, out_bytes)) return rc;
  if (int rc = reserve_pinned(e, in_bytes + out_bytes)) return rc;
  char* pin_in = (char*)e->pinned;
  char* pin_out = pin_in + in_bytes;
  memcpy(pin_in, params_host, in_bytes);
  CU_TRY(e, cudaMemcpyAsync(e->params_dev.ptr, pin_in, in_bytes, cudaMemcpyHostToDevice, e->stream));

  const uint32_t ns = pad_scenarios(n_scen);
  const TilePlan plan = plan_tiles(e, n_opt, std::max<uint64_t>(n_paths, 1), sim_steps, ns, true);
  const uint32_t tiles = plan.tiles;
  const uint64_t ctas = (uint64_t)tiles * n_opt;
  if (ctas > 0x7fffffffull) return fail(e, B200MC_ERR_INVALID, "problem too large for one launch (%llu CTAs)", (unsigned long long)ctas);
  if (int rc = reserve_fold(e, n_opt, tiles, 3 * ns)) return rc;
  if (int rc = order_before(e, e->stream)) return rc;
  StructuredArgs a{};
  a.sim.params = (const b200mc_params_t*)e->params_dev.ptr;
  a.sim.fold.partials = (double*)e->partials.ptr;
  a.sim.fold.tickets = (uint32_t*)e->tickets.ptr;
  a.sim.fold.out = e->moments_dev.ptr;
  a.sim.fold.samples = (double)n_paths;
  a.sim.fold.n_opt = n_opt;
  if (exchange)
    if (int rc = setup_exchange(e, a.sim.fold, out_bytes)) return rc;
  a.sim.path_begin = path_begin, a.sim.n_paths = n_paths;
  a.sim.n_opt = n_opt, a.sim.n_scen = n_scen, a.sim.tiles = tiles, a.sim.paths_per_thread = plan.ppt, a.sim.n_steps = spec->n_steps;
  a.sim.rk = philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32));
  a.sim.stream_base = stream_base;
  a.a = product->a, a.b = product->b, a.c = product->c, a.d = product->d;
  a.period = period, a.n_events = n_events, a.sim_steps = sim_steps;
  const dim3 grid((unsigned)ctas);
  const int slot = (int)(e->timed % b200mc_engine::kRing);
  if (e->timing) CU_TRY(e, cudaEventRecord(e->ring0[slot], e->stream));
  const cudaError_t err = spec->kind == B200MC_CLIQUET ? launch_structured<B200MC_CLIQUET>(a, ns, grid, e->stream)
                                                      : launch_structured<B200MC_AUTOCALLABLE>(a, ns, grid, e->stream);
  if (err != cudaSuccess) return fail(e, B200MC_ERR_CUDA, "structured kernel launch failed: %s", cudaGetErrorString(err));
  if (e->timing) {
    CU_TRY(e, cudaEventRecord(e->ring1[slot], e->stream));
    e->timed += 1;
  }
  e->launches += 1;
  if (int rc = order_after(e, e->stream)) return rc;
  CU_TRY(e, cudaMemcpyAsync(pin_out, e->moments_dev.ptr, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  if (int rc = check_exchange_done(e, exchange)) return rc;
  memcpy(out_host, pin_out, out_bytes);
  return 0;
}

int b200mc_structured_from_normals(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_product_t* product, const b200mc_params_t* p,
                                   const double* Z_host, uint64_t n_paths, double* payoffs_host, b200mc_moments_t* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  uint32_t period, n_events, sim_steps;
  if (int rc = check_structured(e, spec, product, &period, &n_events, &sim_steps)) return rc;
  if (!p || !Z_host || !out_host) return fail(e, B200MC_ERR_INVALID, "null pointer argument");
  if (n_paths == 0) return fail(e, B200MC_ERR_INVALID, "n_paths must be >= 1");
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t z_bytes = (size_t)n_paths * spec->n_steps * sizeof(double);
  const uint64_t ctas = (n_paths + 127) / 128;
  if (ctas > 0x7fffffffull) return fail(e, B200MC_ERR_INVALID, "too many paths for one launch");
  if (int rc = reserve(e, e->scratch_a, z_bytes)) return rc;
  if (int rc = reserve(e, e->scratch_b, n_paths * sizeof(double))) return rc;
  if (int rc = reserve(e, e->partials, ctas * 2 * sizeof(double))) return rc;
  if (int rc = reserve(e, e->moments_dev, sizeof(b200mc_moments_t))) return rc;
  if (int rc = order_before(e, e->stream)) return rc;
  CU_TRY(e, cudaMemcpyAsync(e->scratch_a.ptr, Z_host, z_bytes, cudaMemcpyHostToDevice, e->stream));
  StructuredF64Args a{};
  a.Z = (const double*)e->scratch_a.ptr;
  a.payoffs = payoffs_host ? (double*)e->scratch_b.ptr : nullptr;
  a.partials = (double*)e->partials.ptr;
  a.n_paths = n_paths, a.n_steps = spec->n_steps, a.period = period, a.n_events = n_events;
  a.S = p->S, a.T = p->T, a.r = p->r, a.sigma = p->sigma, a.q = p->q;
  a.a = product->a, a.b = product->b, a.c = product->c, a.d = product->d;
  if (spec->kind == B200MC_CLIQUET) structured_from_normals_kernel<B200MC_CLIQUET><<<(unsigned)ctas, 128, 0, e->stream>>>(a);
  else structured_from_normals_kernel<B200MC_AUTOCALLABLE><<<(unsigned)ctas, 128, 0, e->stream>>>(a);
  CU_TRY(e, cudaGetLastError());
  fold_kernel<<<1, 32, 0, e->stream>>>((const double*)e->partials.ptr, (b200mc_moments_t*)e->moments_dev.ptr, (uint32_t)ctas, (double)n_paths);
  CU_TRY(e, cudaGetLastError());
  e->launches += 2;
  if (payoffs_host) CU_TRY(e, cudaMemcpyAsync(payoffs_host, e->scratch_b.ptr, n_paths * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaMemcpyAsync(out_host, e->moments_dev.ptr, sizeof(b200mc_moments_t), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return 0;
}

int b200mc_payoffs_from_normals_device(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* p, int accumulate,
                                       const double* Z_dev, uint64_t n_paths, double* payoffs_dev, b200mc_moments_t* out_dev,
                                       void* cuda_stream) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  CU_TRY(e, cudaSetDevice(e->device));
  return enqueue_from_normals(e, spec, p, accumulate, Z_dev, n_paths, payoffs_dev, out_dev, (cudaStream_t)cuda_stream, e->timing);
}

int b200mc_payoffs_from_normals(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* p, int accumulate,
                                const double* Z_host, uint64_t n_paths, double* payoffs_host, b200mc_moments_t* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (int rc = check_spec(e, spec)) return rc;
  if (!p || !Z_host || !out_host) return fail(e, B200MC_ERR_INVALID, "null pointer argument");
  if (n_paths == 0) return fail(e, B200MC_ERR_INVALID, "n_paths must be >= 1");
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t z_bytes = (size_t)n_paths * spec->n_steps * sizeof(double);
  const size_t n_out = (size_t)n_paths * (spec->antithetic ? 2 : 1);
  if (int rc = reserve(e, e->scratch_a, z_bytes)) return rc;
  if (int rc = reserve(e, e->scratch_b, n_out * sizeof(double))) return rc;
  if (int rc = reserve(e, e->moments_dev, sizeof(b200mc_moments_t))) return rc;
  CU_TRY(e, cudaMemcpyAsync(e->scratch_a.ptr, Z_host, z_bytes, cudaMemcpyHostToDevice, e->stream));
  if (int rc = enqueue_from_normals(e, spec, p, accumulate, (const double*)e->scratch_a.ptr, n_paths,
                                    payoffs_host ? (double*)e->scratch_b.ptr : nullptr, (b200mc_moments_t*)e->moments_dev.ptr,
                                    e->stream, e->timing))
    return rc;
  if (payoffs_host)
    CU_TRY(e, cudaMemcpyAsync(payoffs_host, e->scratch_b.ptr, n_out * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaMemcpyAsync(out_host, e->moments_dev.ptr, sizeof(b200mc_moments_t), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return 0;
}

int b200mc_generate_normals(b200mc_engine_t* e, uint64_t seed, uint32_t stream, uint64_t path_begin, uint64_t n_paths,
                            uint32_t n_steps, float* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (!out_host || n_paths == 0 || n_steps == 0) return fail(e, B200MC_ERR_INVALID, "bad argument");
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t bytes = (size_t)n_paths * n_steps * sizeof(float);
  if (int rc = reserve(e, e->scratch_a, bytes)) return rc;
  const unsigned grid = (unsigned)std::min<uint64_t>((n_paths + 255) / 256, (uint64_t)e->prop.multiProcessorCount * 32);
  normals_kernel<<<grid, 256, 0, e->stream>>>(philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32)), stream, path_begin, n_paths, n_steps,
                                              (float*)e->scratch_a.ptr);
  CU_TRY(e, cudaGetLastError());
  e->launches += 1;
  CU_TRY(e, cudaMemcpyAsync(out_host, e->scratch_a.ptr, bytes, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return 0;
}

// ---- Heston / jump-diffusion models -----------------------------------------------------------
}  // extern "C" (templates below)

// Shared host wrapper: stage `in_bytes` of parameters (two blocks), launch (the kernel folds its own tiles), copy
// [n_opt] moments back.
template <class Launch>
static int simulate_model_host(b200mc_engine_t* e, const void* block_a, size_t bytes_a, const void* block_b, size_t bytes_b,
                               uint32_t n_opt, uint32_t n_steps, uint32_t per_step_cost, uint64_t n_paths, b200mc_moments_t* out_host,
                               Launch&& launch) {
  if (!block_a || !out_host) return fail(e, B200MC_ERR_INVALID, "params/out pointer is null");
  const bool exchange = exchanging(e, false);
  if (n_opt == 0 || (n_paths == 0 && !exchange) || n_steps == 0) return fail(e, B200MC_ERR_INVALID, "n_opt, n_paths and n_steps must be >= 1");
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t in_bytes = bytes_a + bytes_b, out_bytes = (size_t)n_opt * sizeof(b200mc_moments_t);
  if (int rc = reserve(e, e->params_dev, in_bytes)) return rc;
  if (int rc = reserve(e, e->moments_dev, out_bytes)) return rc;
  if (int rc = reserve_pinned(e, in_bytes + out_bytes)) return rc;
  char* pin_in = (char*)e->pinned;
  char* pin_out = pin_in + in_bytes;
  memcpy(pin_in, block_a, bytes_a);
  if (bytes_b) memcpy(pin_in + bytes_a, block_b, bytes_b);
  CU_TRY(e, cudaMemcpyAsync(e->params_dev.ptr, pin_in, in_bytes, cudaMemcpyHostToDevice, e->stream));
  const TilePlan plan = plan_tiles(e, n_opt, std::max<uint64_t>(n_paths, 1), n_steps * per_step_cost, 1, false);
  const uint32_t tiles = plan.tiles, ppt = plan.ppt;
  const uint64_t ctas = (uint64_t)tiles * n_opt;
  if (ctas > 0x7fffffffull) return fail(e, B200MC_ERR_INVALID, "problem too large for one launch (%llu CTAs)", (unsigned long long)ctas);
  if (int rc = reserve_fold(e, n_opt, tiles, 3)) return rc;
  if (int rc = order_before(e, e->stream)) return rc;
  FoldArgs fold{};
  fold.partials = (double*)e->partials.ptr;
  fold.tickets = (uint32_t*)e->tickets.ptr;
  fold.out = e->moments_dev.ptr;
  fold.samples = (double)n_paths;
  fold.n_opt = n_opt;
  if (exchange)
    if (int rc = setup_exchange(e, fold, out_bytes)) return rc;
  const int slot = (int)(e->timed % b200mc_engine::kRing);
  if (e->timing) CU_TRY(e, cudaEventRecord(e->ring0[slot], e->stream));
  launch((const char*)e->params_dev.ptr, (const char*)e->params_dev.ptr + bytes_a, fold, tiles, ppt, dim3((unsigned)ctas));
  CU_TRY(e, cudaGetLastError());
  if (e->timing) {
    CU_TRY(e, cudaEventRecord(e->ring1[slot], e->stream));
    e->timed += 1;
  }
  e->launches += 1;
  if (int rc = order_after(e, e->stream)) return rc;
  CU_TRY(e, cudaMemcpyAsync(pin_out, e->moments_dev.ptr, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  if (int rc = check_exchange_done(e, exchange)) return rc;
  memcpy(out_host, pin_out, out_bytes);
  return 0;
}

extern "C" {

int b200mc_simulate_heston(b200mc_engine_t* e, const b200mc_heston_params_t* params_host, uint32_t n_opt, int is_put, uint32_t n_steps,
                           uint64_t seed, uint32_t stream_base, uint64_t path_begin, uint64_t n_paths, b200mc_moments_t* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (n_steps > 0x7fffffffu) return fail(e, B200MC_ERR_INVALID, "n_steps too large");
  return simulate_model_host(e, params_host, (size_t)n_opt * sizeof(b200mc_heston_params_t), nullptr, 0, n_opt, n_steps, 2, n_paths, out_host,
                             [&](const char* pa, const char*, const FoldArgs& fold, uint32_t tiles, uint32_t ppt, dim3 grid) {
                               HestonArgs a{};
                               a.params = (const b200mc_heston_params_t*)pa;
                               a.fold = fold;
                               a.path_begin = path_begin, a.n_paths = n_paths;
                               a.n_opt = n_opt, a.tiles = tiles, a.paths_per_thread = ppt, a.n_steps = n_steps;
                               a.rk = philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32)), a.stream_base = stream_base;
                               a.is_put = is_put;
                               heston_kernel<<<grid, kBlock, 0, e->stream>>>(a);
                             });
}

int b200mc_simulate_jump_diffusion(b200mc_engine_t* e, const b200mc_params_t* params_host, const b200mc_jump_params_t* jumps_host,
                                   uint32_t n_opt, int is_put, uint32_t n_steps, uint64_t seed, uint32_t stream_base,
                                   uint64_t path_begin, uint64_t n_paths, b200mc_moments_t* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (!jumps_host) return fail(e, B200MC_ERR_INVALID, "jump parameter pointer is null");
  for (uint32_t i = 0; i < n_opt; ++i) {
    const b200mc_jump_params_t& j = jumps_host[i];
    if (j.model != B200MC_JUMP_MERTON && j.model != B200MC_JUMP_KOU) return fail(e, B200MC_ERR_INVALID, "unknown jump model %d", j.model);
    if (!(j.lambda_j >= 0.0)) return fail(e, B200MC_ERR_INVALID, "lambda_j must be non-negative");
    // the jump count is inverted from p0 = exp(-lambda_j*T) upwards (models.cuh): beyond ~700 expected jumps p0 underflows
    if (params_host && !(j.lambda_j * params_host[i].T <= 700.0))
      return fail(e, B200MC_ERR_INVALID, "lambda_j * T = %g expected jumps per path is beyond the supported 700", j.lambda_j * params_host[i].T);
  }
  return simulate_model_host(e, params_host, (size_t)n_opt * sizeof(b200mc_params_t), jumps_host, (size_t)n_opt * sizeof(b200mc_jump_params_t),
                             n_opt, n_steps, 1, n_paths, out_host,
                             [&](const char* pa, const char* pb, const FoldArgs& fold, uint32_t tiles, uint32_t ppt, dim3 grid) {
                               JumpArgs a{};
                               a.params = (const b200mc_params_t*)pa;
                               a.jumps = (const b200mc_jump_params_t*)pb;
                               a.fold = fold;
                               a.path_begin = path_begin, a.n_paths = n_paths;
                               a.n_opt = n_opt, a.tiles = tiles, a.paths_per_thread = ppt, a.n_steps = n_steps;
                               a.rk = philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32)), a.stream_base = stream_base;
                               a.is_put = is_put;
                               jump_kernel<<<grid, kBlock, 0, e->stream>>>(a);
                             });
}

}  // extern "C" (template below)

// FP64 parity mode of the two models: upload step-major draws, one thread per path, fixed-order fold.
template <class Launch>
static int model_from_draws_host(b200mc_engine_t* e, const double* draws_a, size_t count_a, const double* draws_b, size_t count_b,
                                 uint64_t n_paths, double* payoffs_host, b200mc_moments_t* out_host, Launch&& launch) {
  if (!draws_a || !out_host || n_paths == 0) return fail(e, B200MC_ERR_INVALID, "null pointer or zero size");
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t bytes_a = count_a * sizeof(double), bytes_b = draws_b ? count_b * sizeof(double) : 0;
  if (int rc = reserve(e, e->scratch_a, bytes_a + bytes_b)) return rc;
  if (int rc = reserve(e, e->scratch_b, n_paths * sizeof(double))) return rc;
  if (int rc = reserve(e, e->moments_dev, sizeof(b200mc_moments_t))) return rc;
  const uint64_t ctas = (n_paths + kBlock - 1) / kBlock;
  if (int rc = reserve(e, e->partials, ctas * 2 * sizeof(double))) return rc;
  if (int rc = order_before(e, e->stream)) return rc;
  CU_TRY(e, cudaMemcpyAsync(e->scratch_a.ptr, draws_a, bytes_a, cudaMemcpyHostToDevice, e->stream));
  if (bytes_b) CU_TRY(e, cudaMemcpyAsync((char*)e->scratch_a.ptr + bytes_a, draws_b, bytes_b, cudaMemcpyHostToDevice, e->stream));
  launch((const double*)e->scratch_a.ptr, bytes_b ? (const double*)((char*)e->scratch_a.ptr + bytes_a) : nullptr,
         payoffs_host ? (double*)e->scratch_b.ptr : nullptr, (double*)e->partials.ptr, dim3((unsigned)ctas));
  CU_TRY(e, cudaGetLastError());
  fold_kernel<<<1, 32, 0, e->stream>>>((const double*)e->partials.ptr, (b200mc_moments_t*)e->moments_dev.ptr, (uint32_t)ctas, (double)n_paths);
  CU_TRY(e, cudaGetLastError());
  e->launches += 2;
  if (payoffs_host) CU_TRY(e, cudaMemcpyAsync(payoffs_host, e->scratch_b.ptr, n_paths * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaMemcpyAsync(out_host, e->moments_dev.ptr, sizeof(b200mc_moments_t), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return 0;
}

extern "C" {

int b200mc_heston_from_normals(b200mc_engine_t* e, const b200mc_heston_params_t* p, int is_put, uint32_t n_steps, const double* Z_host,
                               uint64_t n_paths, double* payoffs_host, b200mc_moments_t* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (!p || n_steps == 0) return fail(e, B200MC_ERR_INVALID, "null parameters or zero steps");
  return model_from_draws_host(e, Z_host, (size_t)n_steps * 2 * n_paths, nullptr, 0, n_paths, payoffs_host, out_host,
                               [&](const double* Z, const double*, double* pay, double* partials, dim3 grid) {
                                 HestonF64Args a{};
                                 a.Z = Z, a.payoffs = pay, a.partials = partials, a.n_paths = n_paths, a.n_steps = n_steps, a.is_put = is_put;
                                 a.S = p->S, a.K = p->K, a.T = p->T, a.r = p->r, a.q = p->q;
                                 a.kappa = p->kappa, a.theta = p->theta, a.sigma_v = p->sigma_v, a.rho = p->rho, a.v0 = p->v0;
                                 heston_from_normals_kernel<<<grid, kBlock, 0, e->stream>>>(a);
                               });
}

int b200mc_jump_diffusion_from_draws(b200mc_engine_t* e, const b200mc_params_t* p, double lambda_kappa, int is_put, uint32_t n_steps,
                                     const double* dW_host, const double* J_host, uint64_t n_paths, double* payoffs_host,
                                     b200mc_moments_t* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (!p || n_steps == 0) return fail(e, B200MC_ERR_INVALID, "null parameters or zero steps");
  return model_from_draws_host(e, dW_host, (size_t)n_steps * n_paths, J_host, (size_t)n_steps * n_paths, n_paths, payoffs_host, out_host,
                               [&](const double* dW, const double* J, double* pay, double* partials, dim3 grid) {
                                 JumpF64Args a{};
                                 a.dW = dW, a.J = J, a.payoffs = pay, a.partials = partials, a.n_paths = n_paths, a.n_steps = n_steps, a.is_put = is_put;
                                 a.S = p->S, a.K = p->K, a.T = p->T, a.r = p->r, a.sigma = p->sigma, a.q = p->q, a.lambda_kappa = lambda_kappa;
                                 jump_from_draws_kernel<<<grid, kBlock, 0, e->stream>>>(a);
                               });
}

// ---- quasi-Monte Carlo (Sobol) ---------------------------------------------------------------
static int check_sobol_table(b200mc_engine_t* e, const uint32_t* dirnums_host, const uint32_t* shift_host, uint32_t n_dims, uint32_t bits) {
  if (!dirnums_host || !shift_host) return fail(e, B200MC_ERR_INVALID, "direction-number / shift pointer is null");
  if (n_dims == 0) return fail(e, B200MC_ERR_INVALID, "n_dims must be >= 1");
  if (bits < 1 || bits > 31) return fail(e, B200MC_ERR_INVALID, "bits must be in [1, 31], got %u", bits);
  return 0;
}

// Stage [n_dims][32] direction words + [n_dims] shifts into scratch_a (pinned bounce, async on the engine stream).
static int upload_sobol_table(b200mc_engine_t* e, const uint32_t* dirnums_host, const uint32_t* shift_host, uint32_t n_dims,
                              const uint32_t** dir_dev, const uint32_t** shift_dev, size_t extra_pinned) {
  const size_t dir_bytes = (size_t)n_dims * kSobolWords * sizeof(uint32_t), shift_bytes = (size_t)n_dims * sizeof(uint32_t);
  if (int rc = reserve(e, e->scratch_a, dir_bytes + shift_bytes)) return rc;
  if (int rc = reserve_pinned(e, dir_bytes + shift_bytes + extra_pinned)) return rc;
  memcpy(e->pinned, dirnums_host, dir_bytes);
  memcpy((char*)e->pinned + dir_bytes, shift_host, shift_bytes);
  CU_TRY(e, cudaMemcpyAsync(e->scratch_a.ptr, e->pinned, dir_bytes + shift_bytes, cudaMemcpyHostToDevice, e->stream));
  *dir_dev = (const uint32_t*)e->scratch_a.ptr;
  *shift_dev = (const uint32_t*)((char*)e->scratch_a.ptr + dir_bytes);
  return 0;
}

// Shared body of b200mc_simulate_sobol and b200mc_terminal_prices_sobol (terminal_host != null: also return S_T per point).
static int sobol_run(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* params_host, uint32_t n_opt,
                     uint32_t n_scen, const uint32_t* dirnums_host, const uint32_t* shift_host, uint32_t bits,
                     uint64_t point_begin, uint64_t n_points, b200mc_moments_t* out_host, double* terminal_host, int terminal_anti) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (int rc = check_spec(e, spec)) return rc;
  if (spec->kind != B200MC_EUROPEAN || spec->antithetic)
    return fail(e, B200MC_ERR_INVALID, "the Sobol path prices the plain European payoff (gbm_qmc.py:14-47): kind must be EUROPEAN, antithetic 0");
  if (!params_host || !out_host) return fail(e, B200MC_ERR_INVALID, "params/out pointer is null");
  if (n_opt == 0 || n_scen == 0 || n_scen > B200MC_MAX_SCENARIOS) return fail(e, B200MC_ERR_INVALID, "bad n_opt / n_scen");
  if (int rc = check_sobol_table(e, dirnums_host, shift_host, spec->n_steps, bits)) return rc;
  constexpr uint64_t kAlign = 1ull << kSobolAlignShift;
  const bool exchange = exchanging(e, false) && !terminal_host;
  if (n_points == 0 && !exchange) return fail(e, B200MC_ERR_INVALID, "n_points must be >= 1");
  if (n_points != 0 && point_begin % kAlign != 0)
    return fail(e, B200MC_ERR_INVALID, "point_begin must be a multiple of %llu (one full-size CTA of Sobol points)", (unsigned long long)kAlign);
  if (point_begin + n_points > (1ull << bits))
    return fail(e, B200MC_ERR_INVALID, "points [%llu, %llu) exceed the 2^%u points of the sequence", (unsigned long long)point_begin,
                (unsigned long long)(point_begin + n_points), bits);
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t n = (size_t)n_opt * n_scen;
  const size_t in_bytes = n * sizeof(b200mc_params_t), out_bytes = n * sizeof(b200mc_moments_t);
  const uint32_t* dir_dev;
  const uint32_t* shift_dev;
  if (int rc = upload_sobol_table(e, dirnums_host, shift_host, spec->n_steps, &dir_dev, &shift_dev, in_bytes + out_bytes)) return rc;
  if (int rc = reserve(e, e->params_dev, in_bytes)) return rc;
  if (int rc = reserve(e, e->moments_dev, out_bytes)) return rc;
  const size_t table_bytes = (size_t)spec->n_steps * (kSobolWords + 1) * sizeof(uint32_t);
  char* pin_in = (char*)e->pinned + table_bytes;
  char* pin_out = pin_in + in_bytes;
  memcpy(pin_in, params_host, in_bytes);
  CU_TRY(e, cudaMemcpyAsync(e->params_dev.ptr, pin_in, in_bytes, cudaMemcpyHostToDevice, e->stream));

  const uint32_t ns = pad_scenarios(n_scen);
  // Tile shape: 2^pb points per thread (16 amortises the per-thread table work best) x dsplit CTAs per point tile, each taking a
  // share of the dimensions (whole 64-dimension chunks).  What a small point set needs is BALANCE, not occupancy: 2^20 points
  // are 256 sixteen-point tiles on 148 SMs - the busiest SM gets 2 of them against an average of 1.73 - and two half-tiles per
  // tile leave that ratio where it was (measured: 0.275 -> 0.273 ms).  Four quarter-tiles (1024 CTAs, handed out dynamically:
  // at most 7 per SM against an average of 6.9) even it out.  So: the fewest CTAs whose per-SM maximum is within 5% of the mean.
  const uint32_t chunks = (spec->n_steps + kSobolDimChunk - 1) / kSobolDimChunk;
  const double n_sm = (double)e->prop.multiProcessorCount;
  int pb = kSobolMaxPointBits;
  uint32_t dsplit = 1;
  double best_balance = -1.0;
  for (int cand_pb = kSobolMaxPointBits; cand_pb >= 0 && best_balance < 0.95; cand_pb -= 2) {
    for (uint32_t cand_ds = 1; cand_ds <= 4 && cand_ds <= chunks && best_balance < 0.95; cand_ds *= 2) {
      const double ctas_c = std::ceil((double)std::max<uint64_t>(n_points, 1) / (double)((uint64_t)kBlock << cand_pb)) * n_opt * cand_ds;
      const double balance = (ctas_c / n_sm) / std::ceil(ctas_c / n_sm);
      if (balance > best_balance + 0.05) best_balance = balance, pb = cand_pb, dsplit = cand_ds;
    }
  }
  if (terminal_host) dsplit = 1;  // the terminal-price output is written by the CTA that owns all dimensions
  const uint64_t tile_points = (uint64_t)kBlock << pb;
  const uint64_t tiles = (std::max<uint64_t>(n_points, 1) + tile_points - 1) / tile_points;
  const uint64_t ctas = tiles * n_opt * dsplit;
  if (ctas > 0x7fffffffull) return fail(e, B200MC_ERR_INVALID, "problem too large for one launch (%llu CTAs)", (unsigned long long)ctas);
  if (dsplit > 1) {
    if (int rc = reserve(e, e->qmc_wpart, tiles * n_opt * dsplit * tile_points * sizeof(float))) return rc;
    const size_t ticket_bytes = tiles * n_opt * sizeof(uint32_t);
    if (e->qmc_tickets.bytes < ticket_bytes) {
      if (int rc = reserve(e, e->qmc_tickets, ticket_bytes)) return rc;
      CU_TRY(e, cudaMemset(e->qmc_tickets.ptr, 0, e->qmc_tickets.bytes));
    }
  }
  if (int rc = reserve_fold(e, n_opt, (uint32_t)tiles, 3 * ns)) return rc;
  if (int rc = order_before(e, e->stream)) return rc;
  SobolArgs a{};
  a.params = (const b200mc_params_t*)e->params_dev.ptr;
  a.fold.partials = (double*)e->partials.ptr;
  a.fold.tickets = (uint32_t*)e->tickets.ptr;
  a.fold.out = e->moments_dev.ptr;
  a.fold.samples = (double)n_points;
  a.fold.n_opt = n_opt;
  if (exchange)
    if (int rc = setup_exchange(e, a.fold, out_bytes)) return rc;
  a.dsplit = dsplit;
  a.wpart = (float*)e->qmc_wpart.ptr;
  a.tile_tickets = (uint32_t*)e->qmc_tickets.ptr;
  a.dirnums = dir_dev;
  a.shift = shift_dev;
  a.point_begin = point_begin;
  a.n_points = n_points;
  a.n_opt = n_opt, a.n_scen = n_scen, a.tiles = (uint32_t)tiles, a.n_steps = spec->n_steps, a.bits = bits;
  a.is_put = spec->is_put;
  const size_t terminal_bytes = terminal_host ? (size_t)n_points * (terminal_anti ? 2 : 1) * sizeof(double) : 0;
  if (terminal_host) {
    if (int rc = reserve(e, e->scratch_b, terminal_bytes)) return rc;
    a.terminal_out = (double*)e->scratch_b.ptr;
    a.terminal_anti = terminal_anti;
  }
  const dim3 grid((unsigned)ctas);
  const int slot = (int)(e->timed % b200mc_engine::kRing);
  if (e->timing) CU_TRY(e, cudaEventRecord(e->ring0[slot], e->stream));
#define B200MC_QMC_LAUNCH(NS_)                                                                  \
  switch (pb) {                                                                                  \
    case 4: qmc_european_kernel<NS_, 4><<<grid, kBlock, 0, e->stream>>>(a); break;              \
    case 2: qmc_european_kernel<NS_, 2><<<grid, kBlock, 0, e->stream>>>(a); break;              \
    default: qmc_european_kernel<NS_, 0><<<grid, kBlock, 0, e->stream>>>(a); break;             \
  }
  switch (ns) {
    case 1: B200MC_QMC_LAUNCH(1) break;
    case 2: B200MC_QMC_LAUNCH(2) break;
    case 4: B200MC_QMC_LAUNCH(4) break;
    case 8: B200MC_QMC_LAUNCH(8) break;
    default: B200MC_QMC_LAUNCH(16) break;
  }
#undef B200MC_QMC_LAUNCH
  CU_TRY(e, cudaGetLastError());
  if (e->timing) {
    CU_TRY(e, cudaEventRecord(e->ring1[slot], e->stream));
    e->timed += 1;
  }
  e->launches += 1;
  if (int rc = order_after(e, e->stream)) return rc;
  CU_TRY(e, cudaMemcpyAsync(pin_out, e->moments_dev.ptr, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  if (terminal_host) CU_TRY(e, cudaMemcpyAsync(terminal_host, e->scratch_b.ptr, terminal_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  if (int rc = check_exchange_done(e, exchange)) return rc;
  memcpy(out_host, pin_out, out_bytes);
  return 0;
}

int b200mc_simulate_sobol(b200mc_engine_t* e, const b200mc_spec_t* spec, const b200mc_params_t* params_host, uint32_t n_opt,
                          uint32_t n_scen, const uint32_t* dirnums_host, const uint32_t* shift_host, uint32_t bits,
                          uint64_t point_begin, uint64_t n_points, b200mc_moments_t* out_host) {
  return sobol_run(e, spec, params_host, n_opt, n_scen, dirnums_host, shift_host, bits, point_begin, n_points, out_host, nullptr, 0);
}

int b200mc_terminal_prices_sobol(b200mc_engine_t* e, const b200mc_params_t* p, uint32_t n_steps, int antithetic,
                                 const uint32_t* dirnums_host, const uint32_t* shift_host, uint32_t bits, uint64_t point_begin,
                                 uint64_t n_points, double* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  if (!out_host) return fail(e, B200MC_ERR_INVALID, "out pointer is null");
  b200mc_spec_t spec{};
  spec.kind = B200MC_EUROPEAN, spec.n_steps = n_steps;
  b200mc_moments_t unused;
  return sobol_run(e, &spec, p, 1, 1, dirnums_host, shift_host, bits, point_begin, n_points, &unused, out_host, antithetic != 0);
}

int b200mc_terminal_prices(b200mc_engine_t* e, const b200mc_params_t* p, uint32_t n_steps, int antithetic, uint64_t seed, uint32_t stream,
                           uint64_t path_begin, uint64_t n_paths, double* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (!p || !out_host) return fail(e, B200MC_ERR_INVALID, "params/out pointer is null");
  if (n_steps == 0 || n_paths == 0) return fail(e, B200MC_ERR_INVALID, "n_steps and n_paths must be >= 1");
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t out_bytes = (size_t)n_paths * (antithetic ? 2 : 1) * sizeof(double);
  if (int rc = reserve(e, e->params_dev, sizeof(b200mc_params_t))) return rc;
  if (int rc = reserve(e, e->scratch_b, out_bytes)) return rc;
  if (int rc = reserve_pinned(e, sizeof(b200mc_params_t))) return rc;
  memcpy(e->pinned, p, sizeof(b200mc_params_t));
  CU_TRY(e, cudaMemcpyAsync(e->params_dev.ptr, e->pinned, sizeof(b200mc_params_t), cudaMemcpyHostToDevice, e->stream));
  SimArgs a{};
  a.params = (const b200mc_params_t*)e->params_dev.ptr;
  a.path_begin = path_begin, a.n_paths = n_paths, a.n_opt = 1, a.n_scen = 1, a.n_steps = n_steps;
  a.rk = philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32));
  a.stream_base = stream;
  const uint64_t want = (n_paths + kBlock - 1) / kBlock;
  const unsigned grid = (unsigned)std::min<uint64_t>(want, (uint64_t)e->prop.multiProcessorCount * 48);
  terminal_prices_kernel<<<grid, kBlock, 0, e->stream>>>(a, antithetic != 0, (double*)e->scratch_b.ptr);
  CU_TRY(e, cudaGetLastError());
  e->launches += 1;
  CU_TRY(e, cudaMemcpyAsync(out_host, e->scratch_b.ptr, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return 0;
}

int b200mc_sobol_points(b200mc_engine_t* e, const uint32_t* dirnums_host, const uint32_t* shift_host, uint32_t n_dims, uint32_t bits,
                        uint64_t point_begin, uint64_t n_points, uint32_t* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (int rc = check_sobol_table(e, dirnums_host, shift_host, n_dims, bits)) return rc;
  if (!out_host || n_points == 0 || point_begin + n_points > (1ull << bits)) return fail(e, B200MC_ERR_INVALID, "bad point range");
  CU_TRY(e, cudaSetDevice(e->device));
  const uint32_t* dir_dev;
  const uint32_t* shift_dev;
  if (int rc = upload_sobol_table(e, dirnums_host, shift_host, n_dims, &dir_dev, &shift_dev, 0)) return rc;
  const size_t bytes = (size_t)n_points * n_dims * sizeof(uint32_t);
  if (int rc = reserve(e, e->scratch_b, bytes)) return rc;
  const unsigned grid = (unsigned)std::min<uint64_t>((n_points * n_dims + 255) / 256, (uint64_t)e->prop.multiProcessorCount * 32);
  sobol_points_kernel<<<grid, 256, 0, e->stream>>>(dir_dev, shift_dev, point_begin, n_points, n_dims, (uint32_t*)e->scratch_b.ptr);
  CU_TRY(e, cudaGetLastError());
  e->launches += 1;
  CU_TRY(e, cudaMemcpyAsync(out_host, e->scratch_b.ptr, bytes, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return 0;
}

int b200mc_sobol_normals(b200mc_engine_t* e, const uint32_t* x_host, uint64_t n, uint32_t bits, float* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (!x_host || !out_host || n == 0 || bits < 1 || bits > 31) return fail(e, B200MC_ERR_INVALID, "bad argument");
  CU_TRY(e, cudaSetDevice(e->device));
  if (int rc = reserve(e, e->scratch_a, n * sizeof(uint32_t))) return rc;
  if (int rc = reserve(e, e->scratch_b, n * sizeof(float))) return rc;
  CU_TRY(e, cudaMemcpyAsync(e->scratch_a.ptr, x_host, n * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
  const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)e->prop.multiProcessorCount * 32);
  sobol_normals_kernel<<<grid, 256, 0, e->stream>>>((const uint32_t*)e->scratch_a.ptr, n, bits, (float*)e->scratch_b.ptr);
  CU_TRY(e, cudaGetLastError());
  e->launches += 1;
  CU_TRY(e, cudaMemcpyAsync(out_host, e->scratch_b.ptr, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return 0;
}

int b200mc_rng_statistics(b200mc_engine_t* e, uint64_t seed, uint32_t stream, uint64_t path_begin, uint64_t n_paths, uint32_t n_steps,
                          b200mc_rng_stats_t* out) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (!out || n_paths == 0 || n_steps == 0) return fail(e, B200MC_ERR_INVALID, "bad argument");
  static_assert(sizeof(b200mc_rng_stats_t) == (kStatsBinsZ + kStatsBinsJoint * kStatsBinsJoint + 4) * 8 + kStatsMoments * 8, "stats layout");
  CU_TRY(e, cudaSetDevice(e->device));
  if (int rc = reserve(e, e->scratch_a, sizeof(b200mc_rng_stats_t))) return rc;
  if (int rc = order_before(e, e->stream)) return rc;
  CU_TRY(e, cudaMemsetAsync(e->scratch_a.ptr, 0, sizeof(b200mc_rng_stats_t), e->stream));
  b200mc_rng_stats_t* d = (b200mc_rng_stats_t*)e->scratch_a.ptr;
  RngStatsArgs a{};
  a.rk = philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32));
  a.stream = stream, a.path_begin = path_begin, a.n_paths = n_paths, a.n_steps = n_steps;
  a.hist_z = (unsigned long long*)d->hist_z;
  a.hist_joint = (unsigned long long*)d->hist_joint;
  a.moments = d->moments;
  a.tails = (unsigned long long*)d->tails;
  const unsigned grid = (unsigned)std::min<uint64_t>((n_paths + kBlock - 1) / kBlock, (uint64_t)e->prop.multiProcessorCount * 8);
  rng_stats_kernel<<<grid, kBlock, 0, e->stream>>>(a);
  CU_TRY(e, cudaGetLastError());
  e->launches += 1;
  CU_TRY(e, cudaMemcpyAsync(out, d, sizeof(b200mc_rng_stats_t), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return order_after(e, e->stream);
}

int b200mc_philox_raw(b200mc_engine_t* e, const uint32_t* in_host, uint32_t n, uint32_t* out_host) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (!in_host || !out_host || n == 0) return fail(e, B200MC_ERR_INVALID, "bad argument");
  CU_TRY(e, cudaSetDevice(e->device));
  if (int rc = reserve(e, e->scratch_a, (size_t)n * 6 * sizeof(uint32_t))) return rc;
  if (int rc = reserve(e, e->scratch_b, (size_t)n * 4 * sizeof(uint32_t))) return rc;
  CU_TRY(e, cudaMemcpyAsync(e->scratch_a.ptr, in_host, (size_t)n * 6 * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
  philox_raw_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>((const uint32_t*)e->scratch_a.ptr, n, (uint32_t*)e->scratch_b.ptr);
  CU_TRY(e, cudaGetLastError());
  e->launches += 1;
  CU_TRY(e, cudaMemcpyAsync(out_host, e->scratch_b.ptr, (size_t)n * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return 0;
}

uint64_t b200mc_kernel_launches(const b200mc_engine_t* e) { return e ? e->launches : 0; }

int b200mc_set_plan(b200mc_engine_t* e, int split_shift, uint32_t paths_per_thread) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  if (split_shift > 3 || paths_per_thread > (uint32_t)kMaxPathsPerThread) return fail(e, B200MC_ERR_INVALID, "split_shift must be <= 3, paths_per_thread <= %d", kMaxPathsPerThread);
  e->plan_split_shift = split_shift < 0 ? -1 : split_shift;
  e->plan_ppt = paths_per_thread;
  return 0;
}

int b200mc_plan_tiles(int sm_count, const b200mc_spec_t* spec, uint32_t n_opt, uint32_t n_scen, uint64_t n_paths, int control_variate,
                      uint32_t* tiles, uint32_t* paths_per_thread, uint32_t* split_shift) {
  if (!spec || !tiles || !paths_per_thread || !split_shift || sm_count < 1 || n_opt == 0 || n_paths == 0 || n_scen == 0 ||
      n_scen > B200MC_MAX_SCENARIOS || spec->n_steps == 0)
    return B200MC_ERR_INVALID;
  const bool european = spec->kind == B200MC_EUROPEAN;
  const TilePlan plan = plan_tiles(sm_count, -1, 0, n_opt, n_paths, spec->n_steps, pad_scenarios(n_scen), !european, european && !control_variate);
  *tiles = plan.tiles, *paths_per_thread = plan.ppt, *split_shift = plan.split_shift;
  return 0;
}

int b200mc_last_plan(b200mc_engine_t* e, uint32_t* tiles, uint32_t* paths_per_thread, uint32_t* split_shift) {
  if (!e || !tiles || !paths_per_thread || !split_shift) return fail(e, B200MC_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mutex);
  *tiles = e->last_plan[0], *paths_per_thread = e->last_plan[1], *split_shift = e->last_plan[2];
  return 0;
}

int b200mc_set_kernel_timing(b200mc_engine_t* e, int enabled) {
  if (!e) return B200MC_ERR_INVALID;
  std::lock_guard<std::mutex> g(e->mutex);
  e->timing = enabled != 0;
  e->timed = 0;
  return 0;
}

int b200mc_kernel_timing(b200mc_engine_t* e, float* mean_ms, float* min_ms, int32_t* count) {
  if (!e || !mean_ms || !min_ms || !count) return fail(e, B200MC_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mutex);
  if (e->timed == 0) return fail(e, B200MC_ERR_INVALID, "no timed kernel (enable with b200mc_set_kernel_timing)");
  CU_TRY(e, cudaSetDevice(e->device));
  const int n = (int)std::min<uint64_t>(e->timed, b200mc_engine::kRing);
  double sum = 0.0;
  float best = 1e30f;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    CU_TRY(e, cudaEventSynchronize(e->ring1[i]));
    CU_TRY(e, cudaEventElapsedTime(&ms, e->ring0[i], e->ring1[i]));
    sum += ms;
    best = std::min(best, ms);
  }
  *mean_ms = (float)(sum / n);
  *min_ms = best;
  *count = n;
  return 0;
}

// ---- pipe-rate probes ---------------------------------------------------------------------------
int b200mc_measure_peaks(b200mc_engine_t* e, b200mc_peaks_t* out) {
  if (!e || !out) return fail(e, B200MC_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mutex);
  CU_TRY(e, cudaSetDevice(e->device));
  memset(out, 0, sizeof *out);
  const unsigned grid = (unsigned)e->prop.multiProcessorCount * 8, block = 256;
  const double threads = (double)grid * block;
  if (int rc = reserve(e, e->scratch_a, (size_t)grid * block * sizeof(uint64_t))) return rc;
  if (int rc = reserve(e, e->scratch_b, (size_t)grid * 2 * sizeof(long long))) return rc;
  void* buf = e->scratch_a.ptr;
  cudaStream_t s = e->stream;
  float ms = 0.f;
  auto timed = [&](auto&& launch) -> int {
    launch();  // warm-up (also pages the code in)
    CU_TRY(e, cudaGetLastError());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CU_TRY(e, cudaEventRecord(e->ev0, s));
      launch();
      CU_TRY(e, cudaEventRecord(e->ev1, s));
      CU_TRY(e, cudaEventSynchronize(e->ev1));
      CU_TRY(e, cudaEventElapsedTime(&ms, e->ev0, e->ev1));
      best = std::min(best, ms);
      e->launches += 1;
    }
    ms = best;
    return 0;
  };
  const uint32_t it = 4096;
  if (int rc = timed([&] { probe::ffma<<<grid, block, 0, s>>>(it, 1.0000001f, 1e-9f, (float*)buf); })) return rc;
  out->ffma_per_s = threads * it * 4.0 * probe::kChains / (ms * 1e-3);
  if (int rc = timed([&] { probe::imad_wide<<<grid, block, 0, s>>>(it, 0xD2511F53u, (uint64_t*)buf); })) return rc;
  out->imad_wide_per_s = threads * it * 4.0 * probe::kChains / (ms * 1e-3);
  if (int rc = timed([&] { probe::lop3<<<grid, block, 0, s>>>(it, 0x9E3779B9u, 0xBB67AE85u, (uint32_t*)buf); })) return rc;
  out->lop3_per_s = threads * it * 4.0 * probe::kChains / (ms * 1e-3);
  if (int rc = timed([&] { probe::mufu_mix<<<grid, block, 0, s>>>(it / 4, (float*)buf); })) return rc;
  out->mufu_per_s = threads * (it / 4) * 4.0 * probe::kChains / (ms * 1e-3);
  if (int rc = timed([&] { probe::mufu_ex2<<<grid, block, 0, s>>>(it / 4, (float*)buf); })) return rc;
  out->mufu_ex2_per_s = threads * (it / 4) * 4.0 * probe::kChains / (ms * 1e-3);
  if (int rc = timed([&] { probe::issue_mix<<<grid, block, 0, s>>>(it, 1.0000001f, 1e-9f, 0x9E3779B9u, (float*)buf); })) return rc;
  out->issue_per_s = threads * it * 4.0 * (probe::kChains / 2) * 3.0 / (ms * 1e-3);
  if (int rc = timed([&] { probe::philox_only<<<grid, block, 0, s>>>(it, 42u, 0u, (uint32_t*)buf); })) return rc;
  out->philox_per_s = threads * it / (ms * 1e-3);
  if (int rc = timed([&] { probe::normals_only<<<grid, block, 0, s>>>(it, philox_expand_key(42u, 0u), (float*)buf, (long long*)e->scratch_b.ptr); })) return rc;
  out->normals_per_s = threads * (double)(it * 8) / (ms * 1e-3);
  std::vector<long long> clk((size_t)grid * 2);
  CU_TRY(e, cudaMemcpy(clk.data(), e->scratch_b.ptr, clk.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  std::vector<double> mhz;
  for (unsigned i = 0; i < grid; ++i)
    if (clk[2 * i + 1] > 0) mhz.push_back((double)clk[2 * i] / (double)clk[2 * i + 1] * 1e3);
  if (!mhz.empty()) {
    std::nth_element(mhz.begin(), mhz.begin() + mhz.size() / 2, mhz.end());
    out->sm_clock_mhz_seen = mhz[mhz.size() / 2];
  }
  return 0;
}

}  // extern "C"
