// Scratch experiment 14 (not product): the arithmetic-Asian small-move loop under different register budgets
// (__launch_bounds__ minBlocks), Philox unrolls and polynomial degrees, against the MUFU.EX2 form.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../optionslab_b200/csrc/mc_kernels.cuh"
using namespace b200mc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
template <class L> float time_ms(L&& launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; }
  CK(cudaGetLastError());
  return best;
}
int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint32_t n_steps = 252; const uint64_t n_paths = 1000000;
  const uint32_t n_opt = 64;
  std::vector<b200mc_params_t> hp(n_opt * 16);
  for (auto& p : hp) p = b200mc_params_t{100, 100, 1.0, 0.05, 0.2, 0.0, 120.0, 0};
  b200mc_params_t* dp; CK(cudaMalloc(&dp, hp.size() * sizeof(hp[0]))); CK(cudaMemcpy(dp, hp.data(), hp.size() * sizeof(hp[0]), cudaMemcpyHostToDevice));
  double* partials; CK(cudaMalloc(&partials, (size_t)n_opt * 4096 * 32 * sizeof(double)));
  const uint32_t ppt = 32;
  const uint32_t tiles = (uint32_t)((n_paths + 256ull * ppt - 1) / (256ull * ppt));
  SimArgs a{}; a.params = dp; a.partials = partials; a.path_begin = 0; a.n_paths = n_paths; a.n_opt = n_opt; a.tiles = tiles; a.n_scen = 1;
  a.paths_per_thread = (uint32_t)((n_paths + 256ull * tiles - 1) / (256ull * tiles)); a.n_steps = n_steps; a.rk = philox_expand_key(42u, 0u); a.stream_base = 0;
  const double steps = (double)n_opt * n_paths * n_steps;
  const unsigned grid = n_opt * tiles;
  auto report = [&](const char* name, float ms) { printf("%-60s %9.4f ms  %.4e /s  (%.3f per clk per SM)\n", name, ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9)); fflush(stdout); };
#define AS(M, U, D, EX) a.force_mufu_ex2 = EX; report("asian minb=" #M " unroll=" #U " deg=" #D " exact_ex2=" #EX, time_ms([&] { pathdep_kernel<B200MC_ASIAN_ARITH, 1, M, U, D><<<grid, 256>>>(a); }))
  AS(6, 1, 5, 1); AS(6, 1, 5, 0);
  AS(6, 1, -5, 0); AS(5, 1, -5, 0); AS(4, 1, -5, 0); AS(3, 1, -5, 0); AS(4, 2, -5, 0);
  AS(6, 1, -4, 0); AS(5, 1, -4, 0); AS(4, 1, -4, 0);
  AS(8, 1, 5, 0); AS(5, 1, 5, 0); AS(4, 1, 5, 0); AS(3, 1, 5, 0); AS(2, 1, 5, 0);
  AS(6, 2, 5, 0); AS(4, 2, 5, 0); AS(3, 2, 5, 0);
  AS(6, 1, 4, 0); AS(4, 1, 4, 0); AS(6, 1, 3, 0); AS(4, 1, 3, 0);
  AS(4, 1, 5, 1); AS(8, 1, 5, 1);
  a.force_mufu_ex2 = 0;
#define EU(M, U) report("european<1,anti> minb=" #M " unroll=" #U, time_ms([&] { european_kernel<1, true, M, false, U><<<grid, 256>>>(a); }))
  EU(6, 1); EU(8, 1); EU(5, 1); EU(4, 1);
#define BA(M, U) report("barrier minb=" #M " unroll=" #U, time_ms([&] { pathdep_kernel<B200MC_BARRIER, 1, M, U><<<grid, 256>>>(a); }))
  BA(6, 1); BA(8, 1); BA(4, 1);
  return 0;
}
