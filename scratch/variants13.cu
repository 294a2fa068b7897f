// Scratch experiment 13 (not product; historical): built against a sobol.cuh that selected the inverse-normal variant with -DB200MC_INV_MODE=0/1/2
// 2^20 points x 252 dimensions, random direction words (timing only).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../optionslab_b200/csrc/sobol.cuh"
using namespace b200mc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint32_t d = 252, bits = 30; const uint64_t n = 1u << 20;
  std::vector<uint32_t> tab(d * 32, 0), sh(d);
  uint32_t s = 12345; auto rnd = [&] { s = s * 1664525u + 1013904223u; return (s >> 2) & ((1u << bits) - 1); };
  for (uint32_t j = 0; j < d; ++j) { for (uint32_t b = 0; b < bits; ++b) tab[j * 32 + b] = rnd(); sh[j] = rnd(); }
  uint32_t *dt, *ds; CK(cudaMalloc(&dt, tab.size() * 4)); CK(cudaMalloc(&ds, sh.size() * 4));
  CK(cudaMemcpy(dt, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ds, sh.data(), sh.size() * 4, cudaMemcpyHostToDevice));
  b200mc_params_t hp{100, 100, 1.0, 0.05, 0.2, 0.0, 0, 0}; b200mc_params_t* dp; CK(cudaMalloc(&dp, sizeof hp)); CK(cudaMemcpy(dp, &hp, sizeof hp, cudaMemcpyHostToDevice));
  const uint32_t tiles = (uint32_t)(n / (256 * kSobolPoints));
  double* partials; CK(cudaMalloc(&partials, (size_t)tiles * 2 * sizeof(double)));
  SobolArgs a{}; a.params = dp; a.partials = partials; a.dirnums = dt; a.shift = ds; a.point_begin = 0; a.n_points = n; a.n_opt = 1; a.n_scen = 1; a.tiles = tiles; a.n_steps = d; a.bits = bits;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  qmc_european_kernel<1><<<tiles, 256>>>(a); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) { CK(cudaEventRecord(e0)); qmc_european_kernel<1><<<tiles, 256>>>(a); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; }
  CK(cudaGetLastError());
  std::vector<double> hpart(tiles * 2); CK(cudaMemcpy(hpart.data(), partials, hpart.size() * 8, cudaMemcpyDeviceToHost));
  double sum = 0; for (uint32_t t = 0; t < tiles; ++t) sum += hpart[2 * t];
  printf("INV_MODE=%d  %8.4f ms  %.4e point-dims/s (%.3f per clk per SM)  mean payoff/S %.6f\n", B200MC_INV_MODE, best, (double)n * d / (best * 1e-3),
         (double)n * d / (best * 1e-3) / (sms * 1.965e9), sum / n);
  return 0;
}
