// Scratch experiment 15 (not product): does IMAD.WIDE.U32 block the issue port for its whole pipe occupancy?
// Per loop iteration each thread runs NW mul.wide.u32 (4 independent chains, hi^lo feeds the next multiply, like a
// Philox round) plus NF FFMA (8 independent chains) plus NL LOP3 (4 chains) plus NM MUFU.EX2 (4 chains).
// If the pipes overlap, time ~ max(pipe times); if IMAD.WIDE holds the dispatch port, time ~ 5*NW + NF + NL + NM.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NW, int NF, int NL, int NM, int MODE>
__global__ void __launch_bounds__(256, 6) mix(uint32_t iters, uint32_t seed, float* out) {
  uint32_t w[4]; float f[8]; uint32_t l[4]; float m[4];
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = gid * 2654435761u + i + seed, l[i] = gid ^ (i * 0x9e3779b9u), m[i] = 1e-3f * (float)(gid & 255) + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = 1.0f + 1e-6f * (float)(gid + i);
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < (NW > NF ? (NW > NL ? NW : NL) : (NF > NL ? NF : NL)); ++k) {
      if (k < NW) {
        if (MODE == 0) { const uint64_t p = (uint64_t)w[k & 3] * 0xD2511F53ull; w[k & 3] = (uint32_t)(p >> 32) ^ (uint32_t)p ^ it; }
        if (MODE == 1) { const uint32_t hi = __umulhi(w[k & 3], 0xD2511F53u); w[k & 3] = hi ^ it; }               // IMAD.HI only
        if (MODE == 2) { w[k & 3] = (w[k & 3] * 0xD2511F53u) ^ it; }                                            // IMAD (lo) only
      }
      if (k < NF) f[k & 7] = fmaf(f[k & 7], 1.0000001f, 1e-7f);
      if (k < NL) l[k & 3] = (l[k & 3] ^ l[(k + 1) & 3]) | it;
      if (k < NM) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[k & 3]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += (float)w[i] + (float)l[i] + m[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
  out[gid] = s;
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
  const unsigned grid = sms * 6 * 8; const uint32_t iters = 20000;
  float* out; CK(cudaMalloc(&out, (size_t)grid * 256 * 4));
  // cycles per iteration per SMSP-resident warp-instruction stream: 12 warps per SMSP share one issue port
  auto report = [&](const char* name, float ms, int nw, int nf, int nl, int nm) {
    const double warp_iters_per_smsp = (double)grid * 8 / (sms * 4) * iters;   // warp-iterations each SMSP executes
    const double cyc = ms * 1e-3 * 1.965e9 / warp_iters_per_smsp;              // SMSP cycles per warp-iteration
    printf("%-34s NW=%2d NF=%2d NL=%2d NM=%2d  %8.3f ms  %7.2f cycles/warp-iter  (instr %3d, 5*NW+rest %3d)\n", name, nw, nf, nl, nm, ms, cyc,
           nw + nf + nl + nm, 5 * nw + nf + nl + nm);
    fflush(stdout);
  };
#define RUN(NW, NF, NL, NM, MODE) report(MODE == 0 ? "mul.wide" : MODE == 1 ? "mul.hi" : "mul.lo", time_ms([&] { mix<NW, NF, NL, NM, MODE><<<grid, 256>>>(iters, 1u, out); }), NW, NF, NL, NM)
  RUN(16, 0, 0, 0, 0); RUN(0, 32, 0, 0, 0); RUN(0, 64, 0, 0, 0); RUN(0, 0, 32, 0, 0); RUN(0, 0, 0, 16, 0);
  RUN(16, 16, 0, 0, 0); RUN(16, 32, 0, 0, 0); RUN(16, 64, 0, 0, 0);
  RUN(16, 0, 16, 0, 0); RUN(16, 0, 32, 0, 0);
  RUN(16, 0, 0, 16, 0); RUN(16, 32, 0, 16, 0); RUN(16, 32, 32, 16, 0); RUN(16, 64, 32, 16, 0);
  RUN(0, 32, 32, 16, 0); RUN(0, 64, 32, 16, 0);
  RUN(16, 0, 0, 0, 1); RUN(16, 32, 0, 0, 1); RUN(16, 0, 0, 16, 1); RUN(16, 32, 32, 16, 1);
  RUN(16, 0, 0, 0, 2); RUN(16, 32, 0, 0, 2); RUN(16, 0, 0, 16, 2); RUN(16, 32, 32, 16, 2);
  RUN(32, 0, 0, 0, 2); RUN(32, 32, 32, 16, 2);
  return 0;
}
