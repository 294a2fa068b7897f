// Scratch experiment 17 (not product): which FP32 instruction forms can execute on fmalite?  One kernel per form,
// 8 independent chains per thread; read sm__inst_executed_pipe_fmaheavy / fmalite with ncu, time without it.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// FORM: 0 FFMA r,r,imm   1 FFMA r,r,r (3 distinct)   2 FFMA r,r,r with c == a (fma(a,b,a))   3 FMUL r,r   4 FMUL r,imm
//       5 FADD r,r       6 FADD r,imm                7 FFMA r,imm,r (x*imm + y)              8 FFMA r,r,r + IMAD.WIDE mix
template <int FORM>
__global__ void __launch_bounds__(256, 6) forms(uint32_t iters, float* out, float seed) {
  float f[8], g[8];
  uint32_t w[4];
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = 1.0f + 1e-6f * (float)(gid + i), g[i] = seed + 1e-7f * (float)(gid * 3 + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = gid * 2654435761u + i;
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      float& a = f[k & 7];
      const float b = g[k & 7], c = g[(k + 3) & 7];
      if (FORM == 0) a = fmaf(a, b, 1e-7f);
      if (FORM == 1) a = fmaf(a, b, c);
      if (FORM == 2) a = fmaf(a, b, a);
      if (FORM == 3) a = a * b;
      if (FORM == 4) a = a * 1.0000001f;
      if (FORM == 5) a = a + b;
      if (FORM == 6) a = a + 1e-7f;
      if (FORM == 7) a = fmaf(a, 1.0000001f, b);
      if (FORM == 8) {
        a = fmaf(a, b, c);
        if (k < 16) { const uint64_t p = (uint64_t)w[k & 3] * 0xD2511F53ull; w[k & 3] = (uint32_t)(p >> 32) ^ (uint32_t)p ^ it; }
      }
      if (FORM == 9) {
        a = fmaf(a, b, 1e-7f);
        if (k < 16) { const uint64_t p = (uint64_t)w[k & 3] * 0xD2511F53ull; w[k & 3] = (uint32_t)(p >> 32) ^ (uint32_t)p ^ it; }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += (float)w[i];
  out[gid] = s;
}

template <class L> float time_ms(L&& launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 2; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; }
  CK(cudaGetLastError());
  return best;
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const unsigned grid = sms * 6 * 8; const uint32_t iters = 5000;
  float* out; CK(cudaMalloc(&out, (size_t)grid * 256 * 4));
  auto report = [&](const char* name, float ms) {
    const double warp_iters_per_smsp = (double)grid * 8 / (sms * 4) * iters;
    printf("%-44s %8.3f ms  %7.2f cycles per 64 ops per warp\n", name, ms, ms * 1e-3 * 1.965e9 / warp_iters_per_smsp);
    fflush(stdout);
  };
#define RUN(F, NAME) report(NAME, time_ms([&] { forms<F><<<grid, 256>>>(iters, out, 0.999f); }))
  RUN(0, "FFMA r,r,imm"); RUN(1, "FFMA r,r,r"); RUN(2, "FFMA a,b,a"); RUN(3, "FMUL r,r"); RUN(4, "FMUL r,imm"); RUN(5, "FADD r,r");
  RUN(6, "FADD r,imm"); RUN(7, "FFMA r,imm,r"); RUN(8, "64 FFMA r,r,r + 16 IMAD.WIDE"); RUN(9, "64 FFMA r,r,imm + 16 IMAD.WIDE");
  return 0;
}
