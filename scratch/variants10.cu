// Scratch experiment 10 (not product): the PRODUCTION kernels (one-word-per-pair layout) under different __launch_bounds__ minBlocks / UNROLL.
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
  for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; }
  CK(cudaGetLastError());
  return best;
}
int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint32_t n_opt = 128, n_steps = 252; const uint64_t n_paths = 1000000;
  std::vector<b200mc_params_t> hp(n_opt * 16);
  for (auto& p : hp) p = b200mc_params_t{100, 100, 1.0, 0.05, 0.2, 0.0, 120.0, 0};
  b200mc_params_t* dp; CK(cudaMalloc(&dp, hp.size() * sizeof(hp[0]))); CK(cudaMemcpy(dp, hp.data(), hp.size() * sizeof(hp[0]), cudaMemcpyHostToDevice));
  const uint32_t ppt = 32, tiles = (uint32_t)((n_paths + 256ull * ppt - 1) / (256ull * ppt));
  double* partials; CK(cudaMalloc(&partials, (size_t)n_opt * tiles * 32 * sizeof(double)));
  SimArgs a{}; a.params = dp; a.partials = partials; a.path_begin = 0; a.n_paths = n_paths; a.n_opt = n_opt; a.n_scen = 1; a.tiles = tiles;
  a.paths_per_thread = (uint32_t)((n_paths + 256ull * tiles - 1) / (256ull * tiles)); a.n_steps = n_steps; a.seed_lo = 42; a.seed_hi = 0; a.stream_base = 0;
  const double steps = (double)n_opt * n_paths * n_steps;
  auto report = [&](const char* name, float ms) { printf("%-46s %9.3f ms  %.4e /s  (%.3f per clk per SM)\n", name, ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9)); };
  const unsigned grid = n_opt * tiles;
#define EU(M, U) report("european<1,anti> minb=" #M " unroll=" #U, time_ms([&] { european_kernel<1, true, M, false, U><<<grid, 256>>>(a); }))
#define PD(K, M, U) report("pathdep<" #K ",1> minb=" #M " unroll=" #U, time_ms([&] { pathdep_kernel<K, 1, M, U><<<grid, 256>>>(a); }))
  EU(1, 1); EU(3, 1); EU(4, 1); EU(5, 1); EU(6, 1); EU(8, 1); EU(1, 2); EU(4, 2); EU(6, 2);
  PD(B200MC_ASIAN_ARITH, 1, 1); PD(B200MC_ASIAN_ARITH, 3, 1); PD(B200MC_ASIAN_ARITH, 4, 1); PD(B200MC_ASIAN_ARITH, 5, 1); PD(B200MC_ASIAN_ARITH, 6, 1); PD(B200MC_ASIAN_ARITH, 8, 1);
  PD(B200MC_ASIAN_ARITH, 1, 2); PD(B200MC_ASIAN_ARITH, 4, 2); PD(B200MC_ASIAN_ARITH, 6, 2);
  PD(B200MC_ASIAN_GEOM, 1, 1); PD(B200MC_ASIAN_GEOM, 4, 1); PD(B200MC_ASIAN_GEOM, 6, 1); PD(B200MC_ASIAN_GEOM, 4, 2);
  PD(B200MC_BARRIER, 1, 1); PD(B200MC_BARRIER, 3, 1); PD(B200MC_BARRIER, 4, 1); PD(B200MC_BARRIER, 5, 1); PD(B200MC_BARRIER, 6, 1); PD(B200MC_BARRIER, 8, 1);
  PD(B200MC_BARRIER, 1, 2); PD(B200MC_BARRIER, 4, 2); PD(B200MC_BARRIER, 6, 2);
  a.n_scen = 2;
#define EUN(N, M, U) report("european<" #N ",anti> minb=" #M " unroll=" #U, time_ms([&] { european_kernel<N, true, M, false, U><<<grid, 256>>>(a); }))
#define PDN(K, N, M, U) report("pathdep<" #K "," #N "> minb=" #M " unroll=" #U, time_ms([&] { pathdep_kernel<K, N, M, U><<<grid, 256>>>(a); }))
  EUN(2, 4, 1); EUN(2, 6, 1); PDN(B200MC_ASIAN_ARITH, 2, 3, 1); PDN(B200MC_ASIAN_ARITH, 2, 4, 1); PDN(B200MC_ASIAN_ARITH, 2, 6, 1); PDN(B200MC_BARRIER, 2, 3, 1); PDN(B200MC_BARRIER, 2, 4, 1); PDN(B200MC_BARRIER, 2, 6, 1);
  a.n_scen = 4;
  EUN(4, 2, 1); EUN(4, 4, 1); PDN(B200MC_ASIAN_ARITH, 4, 2, 1); PDN(B200MC_ASIAN_ARITH, 4, 3, 1); PDN(B200MC_ASIAN_ARITH, 4, 4, 1); PDN(B200MC_BARRIER, 4, 2, 1); PDN(B200MC_BARRIER, 4, 3, 1); PDN(B200MC_BARRIER, 4, 4, 1);
  a.n_scen = 8;
  EUN(8, 2, 1); EUN(8, 3, 1); PDN(B200MC_ASIAN_ARITH, 8, 1, 1); PDN(B200MC_ASIAN_ARITH, 8, 2, 1); PDN(B200MC_ASIAN_ARITH, 8, 3, 1); PDN(B200MC_BARRIER, 8, 1, 1); PDN(B200MC_BARRIER, 8, 2, 1); PDN(B200MC_BARRIER, 8, 3, 1);
  a.n_scen = 14;
  EUN(16, 1, 1); EUN(16, 2, 1); PDN(B200MC_ASIAN_ARITH, 16, 1, 1); PDN(B200MC_ASIAN_ARITH, 16, 2, 1); PDN(B200MC_BARRIER, 16, 1, 1); PDN(B200MC_BARRIER, 16, 2, 1);
  return 0;
}
