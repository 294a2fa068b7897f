// Scratch experiment 6: can IMAD.WIDE, LOP3, FFMA and MUFU streams overlap?  (dispatch-port hypothesis)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
// per loop iteration: NW imad.wide, NL lop3, NF ffma, NM mufu(ex2) - all on independent chains
template <int NW, int NL, int NF, int NM>
__global__ void __launch_bounds__(256) mix(uint32_t iters, uint32_t m, float fa, float* out) {
  uint64_t w[NW > 0 ? NW : 1]; uint32_t l[NL > 0 ? NL : 1]; float f[NF > 0 ? NF : 1]; float x[NM > 0 ? NM : 1];
#pragma unroll
  for (int i = 0; i < NW; ++i) w[i] = threadIdx.x * 2654435761u + i;
#pragma unroll
  for (int i = 0; i < NL; ++i) l[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = (float)(threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < NM; ++i) x[i] = 0.001f * (threadIdx.x + i);
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < NW) { uint32_t lo = (uint32_t)w[i]; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(lo), "r"(m)); }
        if (i < NL) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(m), "r"(m + 1));
        if (i < NM) asm volatile("neg.f32 %0, %0; ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        if (i < NF) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fa));
        if (i + 8 < NL) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i + 8]) : "r"(m), "r"(m + 1));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NW; ++i) s += (float)(w[i] >> 32) + (float)(uint32_t)w[i];
#pragma unroll
  for (int i = 0; i < NL; ++i) s += (float)l[i];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
#pragma unroll
  for (int i = 0; i < NM; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
  float* out; CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 4));
  const uint32_t it = 8192;
#define RUN(NW, NL, NF, NM) { float ms = time_ms([&] { mix<NW, NL, NF, NM><<<sms * 8, 256>>>(it, 0xD2511F53u, 1.0001f, out); }); \
    double clk = ms * 1e-3 * 1.965e9; double per = (double)8 * 256 * it * 2 / clk; /* thread-iterations(u) per clk per SM */ \
    printf("W=%d L=%2d F=%d M=%d : %.2f cyc per SMSP per warp-iteration | per clk per SM: wide %.1f lop %.1f ffma %.1f mufu %.1f total-instr %.1f\n", NW, NL, NF, NM, \
      32.0 * 4 / per / 4 * 1.0, per * NW, per * NL, per * NF, per * NM, per * (NW + NL + NF + 2 * NM)); }
  RUN(4, 0, 0, 0); RUN(0, 8, 0, 0); RUN(0, 0, 8, 0); RUN(0, 0, 0, 4);
  RUN(4, 4, 0, 0); RUN(4, 8, 0, 0); RUN(4, 12, 0, 0); RUN(4, 4, 4, 0); RUN(4, 8, 4, 0); RUN(4, 8, 8, 0);
  RUN(3, 0, 0, 2); RUN(3, 4, 0, 2); RUN(3, 5, 2, 2); RUN(3, 5, 3, 2); RUN(6, 9, 5, 4); RUN(3, 5, 3, 1); RUN(2, 4, 3, 2);
  return 0;
}
