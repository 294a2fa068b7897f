// Scratch experiment 7: 6 normals per Philox call (3 radii + 3 angles, one angle from spare low bits) vs 16 per 3 calls.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../optionslab_b200/csrc/mc_kernels.cuh"
using namespace b200mc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <class F> __device__ __forceinline__ void hex_call(const u32x4& x, F&& f) {
  const uint32_t a2 = __byte_perm(x.x, x.y, 0x0040) & 0xffffu;  // byte0 of w0, byte0 of w1
  f(box_muller(x.x, (x.w & 0xffffu) | 0x4b000000u));
  f(box_muller(x.y, (x.w >> 16) | 0x4b000000u));
  f(box_muller(x.z, a2 | 0x4b000000u));
}
// MODE 0: production layout (for_each_pair, 252 steps) ; MODE 1: hex layout, UNROLL calls per iteration
template <int MODE, int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) k(uint32_t ppt, uint32_t n_steps, uint32_t k0, uint32_t k1, float* out) {
  const uint64_t base = ((uint64_t)blockIdx.x * ppt) * 256 + threadIdx.x;
  float acc = 0.f;
  for (uint32_t j = 0; j < ppt; ++j) {
    const uint64_t path = base + (uint64_t)j * 256;
    float W = 0.f;
    auto cons = [&](const NormalPair& p) { W = fmaf(p.rad, p.cs, W); W = fmaf(p.rad, p.sn, W); };
    if (MODE == 0) {
      for_each_pair(path, n_steps, 0u, k0, k1, [&](const NormalPair& p, int) { cons(p); });
    } else {
      const uint32_t calls = n_steps / 6;
      for (uint32_t c = 0; c < calls; c += UNROLL) {
        u32x4 x[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) x[u] = draw4(path, c + u, 0u, k0, k1);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) hex_call(x[u], cons);
      }
    }
    acc += W;
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
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
  const uint32_t ppt = 32, grid = sms * 8 * 16;
  float* out; CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float)));
#define RUN(M, U, B, NS) { const double steps = (double)grid * 256 * ppt * NS; float ms = time_ms([&] { k<M, U, B><<<grid, 256>>>(ppt, NS, 42u, 0u, out); }); \
    printf("mode=%d unroll=%d minb=%d steps=%d  %9.3f ms  %.4e /s  (%.3f per clk per SM)\n", M, U, B, NS, ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9)); }
  RUN(0, 1, 4, 240); RUN(0, 1, 4, 252);
  RUN(1, 1, 4, 240); RUN(1, 2, 4, 240); RUN(1, 3, 4, 252); RUN(1, 4, 4, 240); RUN(1, 2, 1, 240); RUN(1, 3, 1, 252); RUN(1, 4, 1, 240); RUN(1, 6, 1, 252); RUN(1, 2, 6, 240);
  return 0;
}
