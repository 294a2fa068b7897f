// Scratch experiment 4 (not product): CTA size vs end-of-CTA barrier imbalance, European NS=1 ANTI.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../optionslab_b200/csrc/mc_kernels.cuh"
using namespace b200mc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) eu(const SimArgs a) {
  constexpr int WARPS = BLOCK / 32;
  __shared__ Coef coef[1];
  __shared__ double warp_sums[WARPS][2];
  const uint32_t opt = blockIdx.x / a.tiles;
  const uint32_t tile = blockIdx.x - opt * a.tiles;
  if (threadIdx.x < 1) coef[0] = make_coef(a.params[(size_t)opt * a.n_scen], a.n_steps, 1.0f);
  __syncthreads();
  float acc[2] = {0.f, 0.f};
  const uint32_t stream = a.stream_base + opt;
  const uint64_t tile_first = (uint64_t)tile * (uint64_t)(BLOCK * a.paths_per_thread);
  for (uint32_t j = 0; j < a.paths_per_thread; ++j) {
    const uint64_t local = tile_first + (uint64_t)j * BLOCK + threadIdx.x;
    if (local >= a.n_paths) break;
    const float W = terminal_sum(a.path_begin + local, a.n_steps, stream, a.seed_lo, a.seed_hi);
    const Coef q = coef[0];
    float p = vanilla(mufu_ex2(fmaf(q.c, W, q.a)), q.kappa, false);
    acc[0] += p; acc[1] = fmaf(p, p, acc[1]);
    p = vanilla(mufu_ex2(fmaf(-q.c, W, q.a)), q.kappa, false);
    acc[0] += p; acc[1] = fmaf(p, p, acc[1]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double x0 = acc[0], x1 = acc[1];
  for (int off = 16; off > 0; off >>= 1) { x0 += __shfl_xor_sync(0xffffffffu, x0, off); x1 += __shfl_xor_sync(0xffffffffu, x1, off); }
  if (WARPS == 1) { if (lane == 0) { a.partials[(size_t)blockIdx.x * 2] = x0; a.partials[(size_t)blockIdx.x * 2 + 1] = x1; } return; }
  if (lane == 0) { warp_sums[warp][0] = x0; warp_sums[warp][1] = x1; }
  __syncthreads();
  if (threadIdx.x < 2) { double s = 0; for (int w = 0; w < WARPS; ++w) s += warp_sums[w][threadIdx.x]; a.partials[(size_t)blockIdx.x * 2 + threadIdx.x] = s; }
}

// persistent warps pulling chunks from an atomic counter (no CTA-level barrier at all)
template <int MINB>
__global__ void __launch_bounds__(256, MINB) eu_persistent(const SimArgs a, unsigned* counter, uint32_t chunks_per_opt, uint32_t ppt) {
  const int lane = threadIdx.x & 31;
  const uint32_t n_chunks = chunks_per_opt * a.n_opt;
  for (;;) {
    uint32_t chunk = 0;
    if (lane == 0) chunk = atomicAdd(counter, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk >= n_chunks) break;
    const uint32_t opt = chunk / chunks_per_opt, sub = chunk - opt * chunks_per_opt;
    Coef q;
    {
      Coef mine = make_coef(a.params[(size_t)opt * a.n_scen], a.n_steps, 1.0f);  // every lane computes (redundant, no smem)
      q = mine;
    }
    float acc[2] = {0.f, 0.f};
    const uint32_t stream = a.stream_base + opt;
    const uint64_t first = (uint64_t)sub * (32ull * ppt);
    for (uint32_t j = 0; j < ppt; ++j) {
      const uint64_t local = first + (uint64_t)j * 32 + lane;
      if (local >= a.n_paths) break;
      const float W = terminal_sum(a.path_begin + local, a.n_steps, stream, a.seed_lo, a.seed_hi);
      float p = vanilla(mufu_ex2(fmaf(q.c, W, q.a)), q.kappa, false);
      acc[0] += p; acc[1] = fmaf(p, p, acc[1]);
      p = vanilla(mufu_ex2(fmaf(-q.c, W, q.a)), q.kappa, false);
      acc[0] += p; acc[1] = fmaf(p, p, acc[1]);
    }
    double x0 = acc[0], x1 = acc[1];
    for (int off = 16; off > 0; off >>= 1) { x0 += __shfl_xor_sync(0xffffffffu, x0, off); x1 += __shfl_xor_sync(0xffffffffu, x1, off); }
    if (lane == 0) { a.partials[(size_t)chunk * 2] = x0; a.partials[(size_t)chunk * 2 + 1] = x1; }
  }
}

template <class L> float time_ms(L&& launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; }
  CK(cudaGetLastError());
  return best;
}
int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint32_t n_opt = 256, n_steps = 252; const uint64_t n_paths = 1000000;
  std::vector<b200mc_params_t> hp(n_opt);
  for (auto& p : hp) p = b200mc_params_t{100, 100, 1.0, 0.05, 0.2, 0.0, 120.0, 0};
  b200mc_params_t* dp; CK(cudaMalloc(&dp, hp.size() * sizeof(hp[0]))); CK(cudaMemcpy(dp, hp.data(), hp.size() * sizeof(hp[0]), cudaMemcpyHostToDevice));
  double* partials; CK(cudaMalloc(&partials, (size_t)256 << 20));
  unsigned* counter; CK(cudaMalloc(&counter, 4));
  const double steps = (double)n_opt * n_paths * n_steps;
  auto report = [&](const char* name, float ms) { printf("%-46s %9.3f ms  %.4e /s  (%.3f per clk per SM)\n", name, ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9)); };
  auto args = [&](int block, uint32_t ppt) {
    SimArgs a{}; a.params = dp; a.partials = partials; a.n_paths = n_paths; a.n_opt = n_opt; a.n_scen = 1;
    a.tiles = (uint32_t)((n_paths + (uint64_t)block * ppt - 1) / ((uint64_t)block * ppt));
    a.paths_per_thread = (uint32_t)((n_paths + (uint64_t)block * a.tiles - 1) / ((uint64_t)block * a.tiles)); a.n_steps = n_steps; a.seed_lo = 42; return a; };
#define EU(B, M, PPT) { SimArgs a = args(B, PPT); report("eu block=" #B " minb=" #M " ppt=" #PPT, time_ms([&] { eu<B, M><<<a.n_opt * a.tiles, B>>>(a); })); }
  EU(256, 4, 32); EU(128, 8, 32); EU(64, 16, 32); EU(32, 32, 32); EU(32, 32, 128); EU(32, 24, 64); EU(64, 12, 64); EU(128, 6, 64); EU(512, 2, 32); EU(1024, 1, 32);
#define EP(M, PPT, CPS) { SimArgs a = args(32, PPT); report("persistent minb=" #M " ppt=" #PPT " ctas/sm=" #CPS, time_ms([&] { CK(cudaMemsetAsync(counter, 0, 4)); eu_persistent<M><<<sms * CPS, 256>>>(a, counter, a.tiles, a.paths_per_thread); })); }
  EP(4, 32, 4); EP(4, 8, 4); EP(3, 32, 3); EP(5, 32, 5); EP(4, 128, 4);
  return 0;
}
