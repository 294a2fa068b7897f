// Scratch experiment 11: which instruction class steals XU throughput?  A loop of 4 MUFU (lg2, sqrt, sin, cos as in
// Box-Muller) per iteration plus NI IMAD.WIDE, NL LOP3, NF FFMA of filler, all independent of the MUFU chain.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ float m_lg2(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float m_sqrt(float x) { float y; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float m_ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float m_rsq(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MODE 0: lg2,sqrt,ex2,rsq (no FMUL.RZ companions)  MODE 1: lg2, sqrt, sin, cos
template <int NI, int NL, int NF, int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) k(uint32_t iters, float seed, float* out) {
  float a = seed + threadIdx.x * 1e-3f, b = a + 0.5f, c = a + 0.25f, d = a + 0.125f;
  uint32_t p = threadIdx.x * 2654435761u + 1u, q = p ^ 0x9e3779b9u;
  uint32_t l0 = p, l1 = q;
  float f0 = a, f1 = b;
  for (uint32_t i = 0; i < iters; ++i) {
    if (MODE == 0) { a = m_lg2(a); b = m_sqrt(b); c = m_ex2(c); d = m_rsq(d); }
    else { a = m_lg2(a); b = m_sqrt(b); c = __sinf(c); d = __cosf(d); }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      if (j & 1) { uint64_t v = (uint64_t)q * 0xCD9E8D57u; q = (uint32_t)(v >> 32) ^ (uint32_t)v; }
      else { uint64_t w = (uint64_t)p * 0xD2511F53u; p = (uint32_t)(w >> 32) ^ (uint32_t)w; }  // IMAD.WIDE + LOP3
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) { l0 = (l0 ^ 0x5bd1e995u) + (l1 >> 3); l1 ^= l0; }  // 3 ALU ops per j
#pragma unroll
    for (int j = 0; j < NF; ++j) { f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 0.9999f, f0); }  // 2 FFMA per j
  }
  out[blockIdx.x * 256 + threadIdx.x] = a + b + c + d + (float)(p ^ q ^ l0 ^ l1) + f0 + f1;
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
  const uint32_t grid = sms * 8 * 4, iters = 20000;
  float* out; CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float)));
  const double mufu = (double)grid * 256 * iters * 4;
#define RUN(NI, NL, NF, MODE, MINB) { float ms = time_ms([&] { k<NI, NL, NF, MODE, MINB><<<grid, 256>>>(iters, 1.5f, out); }); \
    printf("imad.wide=%d lop-triples=%d ffma-pairs=%d mode=%d minb=%d  %8.3f ms  MUFU/clk/SM = %.3f\n", NI, NL, NF, MODE, MINB, ms, mufu / (ms * 1e-3) / (sms * 1.965e9)); }
  RUN(0, 0, 0, 0, 6); RUN(0, 0, 0, 1, 6);
  RUN(2, 0, 0, 0, 6); RUN(4, 0, 0, 0, 6); RUN(4, 0, 0, 1, 6); RUN(6, 0, 0, 0, 6);
  RUN(0, 2, 0, 0, 6); RUN(0, 4, 0, 0, 6); RUN(0, 4, 0, 1, 6);
  RUN(0, 0, 2, 0, 6); RUN(0, 0, 4, 0, 6); RUN(0, 0, 4, 1, 6);
  RUN(4, 2, 2, 1, 6); RUN(4, 2, 2, 1, 8); RUN(4, 2, 2, 1, 4); RUN(4, 2, 2, 0, 6);
  RUN(2, 2, 2, 1, 6); RUN(3, 2, 2, 1, 6);
  return 0;
}
