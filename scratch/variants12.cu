// Scratch experiment 12 (not product; historical): built against a european_kernel with an extra SMEM_ACC template parameter (rejected, see profiles/r01_variants12_smem_accumulators.txt).
// registers vs in shared memory (SMEM_ACC), large grid (128 options x 1M paths) and the C2 shape (1 option x 1M paths).
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
  std::vector<b200mc_params_t> hp(128 * 16);
  for (auto& p : hp) p = b200mc_params_t{100, 100, 1.0, 0.05, 0.2, 0.0, 120.0, 0};
  b200mc_params_t* dp; CK(cudaMalloc(&dp, hp.size() * sizeof(hp[0]))); CK(cudaMemcpy(dp, hp.data(), hp.size() * sizeof(hp[0]), cudaMemcpyHostToDevice));
  double* partials; CK(cudaMalloc(&partials, (size_t)128 * 4096 * 32 * sizeof(double)));
  for (uint32_t n_opt : {128u, 1u}) {
    for (uint32_t ppt : {32u, 9u, 4u, 2u, 1u}) {
      if (n_opt == 128 && ppt != 32) continue;
      const uint32_t tiles = (uint32_t)((n_paths + 256ull * ppt - 1) / (256ull * ppt));
      SimArgs a{}; a.params = dp; a.partials = partials; a.path_begin = 0; a.n_paths = n_paths; a.n_opt = n_opt; a.tiles = tiles;
      a.paths_per_thread = (uint32_t)((n_paths + 256ull * tiles - 1) / (256ull * tiles)); a.n_steps = n_steps; a.seed_lo = 42; a.seed_hi = 0; a.stream_base = 0;
      const double steps = (double)n_opt * n_paths * n_steps;
      const unsigned grid = n_opt * tiles;
      auto report = [&](const char* name, float ms) { printf("n_opt=%3u ppt=%2u ctas=%6u %-40s %9.4f ms  %.4e /s  (%.3f per clk per SM)\n", n_opt, a.paths_per_thread, grid, name, ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9)); };
#define EU(N, SCEN, M, SM) a.n_scen = SCEN; report("european<" #N "> scen=" #SCEN " minb=" #M " smem_acc=" #SM, time_ms([&] { european_kernel<N, true, M, false, 1, SM><<<grid, 256>>>(a); }))
      EU(16, 14, 2, false); EU(16, 14, 4, true); EU(16, 14, 5, true); EU(16, 14, 6, true);
      EU(8, 8, 2, false); EU(8, 8, 5, true); EU(8, 8, 6, true);
      EU(4, 4, 2, false); EU(4, 4, 6, true);
      EU(1, 1, 6, false);
    }
  }
  return 0;
}
