// Scratch experiment (not product): throughput of European-kernel variants on one B200.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../optionslab_b200/csrc/mc_kernels.cuh"
using namespace b200mc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// generic: each thread simulates `ppt` paths of n_steps, sums W (so nothing is optimised away)
template <int ROUNDS, int ILP, int MINB>
__global__ void __launch_bounds__(256, MINB) k_terminal(uint32_t ppt, uint32_t n_steps, uint32_t k0, uint32_t k1, float* out) {
  const uint64_t base = ((uint64_t)blockIdx.x * ppt) * 256 + threadIdx.x;
  float acc = 0.f;
  for (uint32_t j = 0; j < ppt; j += ILP) {
    float W[ILP];
    uint64_t path[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { W[i] = 0.f; path[i] = base + (uint64_t)(j + i) * 256; }
    const uint32_t full = n_steps >> 4;
    for (uint32_t sb = 0; sb < full; ++sb) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        const u32x4 a = philox4x32<ROUNDS>((uint32_t)path[i], 3 * sb, (uint32_t)(path[i] >> 32), 0u, k0, k1);
        const u32x4 b = philox4x32<ROUNDS>((uint32_t)path[i], 3 * sb + 1, (uint32_t)(path[i] >> 32), 0u, k0, k1);
        const u32x4 c = philox4x32<ROUNDS>((uint32_t)path[i], 3 * sb + 2, (uint32_t)(path[i] >> 32), 0u, k0, k1);
        NormalPair A, B;
#define ACC(P) W[i] = fmaf(P.rad, P.cs, W[i]); W[i] = fmaf(P.rad, P.sn, W[i]);
        box_muller_quad(a.x, a.y, a.z, A, B); ACC(A) ACC(B)
        box_muller_quad(a.w, b.x, b.y, A, B); ACC(A) ACC(B)
        box_muller_quad(b.z, b.w, c.x, A, B); ACC(A) ACC(B)
        box_muller_quad(c.y, c.z, c.w, A, B); ACC(A) ACC(B)
      }
    }
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc += mufu_ex2(W[i] * 0.01f);
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}

// Box-Muller only (words from a cheap xorshift): the XU-bound ceiling of this instruction mix
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_boxmuller_only(uint32_t iters, float* out) {
  uint32_t x = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
  float W = 0.f;
  for (uint32_t it = 0; it < iters; ++it) {
    uint32_t w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; w[i] = x; }
    NormalPair A, B;
    box_muller_quad(w[0], w[1], w[2], A, B);
    W = fmaf(A.rad, A.cs, W); W = fmaf(A.rad, A.sn, W); W = fmaf(B.rad, B.cs, W); W = fmaf(B.rad, B.sn, W);
  }
  out[blockIdx.x * 256 + threadIdx.x] = W;
}

// Philox only, production counter layout (16 IMAD.WIDE per call)
template <int ROUNDS>
__global__ void __launch_bounds__(256) k_philox_only(uint32_t iters, uint32_t k0, uint32_t k1, uint32_t* out) {
  const uint32_t gid = blockIdx.x * 256 + threadIdx.x;
  uint32_t s = 0;
  for (uint32_t it = 0; it < iters; ++it) {
    const u32x4 x = philox4x32<ROUNDS>(gid, it, 0u, 7u, k0, k1);
    s ^= x.x ^ x.y ^ x.z ^ x.w;
  }
  out[gid] = s;
}

template <class L>
float time_ms(L&& launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint32_t n_steps = 256, ppt = 32;
  const uint32_t grid = sms * 8 * 16;  // 16 waves at 8 CTAs/SM
  float* out; CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float) * 2));
  const double steps = (double)grid * 256 * ppt * n_steps;
  auto report = [&](const char* name, float ms, double work) {
    printf("%-44s %9.3f ms  %.4e /s  (%.3f per clk per SM @1.965GHz)\n", name, ms, work / (ms * 1e-3), work / (ms * 1e-3) / (sms * 1.965e9));
  };
#define RUN_T(R, I, M) report("terminal rounds=" #R " ilp=" #I " minblocks=" #M, time_ms([&] { k_terminal<R, I, M><<<grid, 256>>>(ppt, n_steps, 42u, 0u, out); }), steps)
  RUN_T(10, 1, 1); RUN_T(10, 1, 6); RUN_T(10, 1, 8); RUN_T(10, 2, 1); RUN_T(10, 2, 4); RUN_T(10, 2, 6); RUN_T(10, 4, 1);
  RUN_T(7, 1, 1); RUN_T(7, 2, 1);
  const uint32_t it = 16384;
  report("box-muller only (quads, xorshift words) mb=1", time_ms([&] { k_boxmuller_only<1><<<sms * 8, 256>>>(it, out); }), (double)sms * 8 * 256 * it * 4);
  report("box-muller only mb=8", time_ms([&] { k_boxmuller_only<8><<<sms * 8, 256>>>(it, out); }), (double)sms * 8 * 256 * it * 4);
  report("philox10 only (calls)", time_ms([&] { k_philox_only<10><<<sms * 8, 256>>>(it, 42u, 0u, (uint32_t*)out); }), (double)sms * 8 * 256 * it);
  report("philox7 only (calls)", time_ms([&] { k_philox_only<7><<<sms * 8, 256>>>(it, 42u, 0u, (uint32_t*)out); }), (double)sms * 8 * 256 * it);
  return 0;
}
